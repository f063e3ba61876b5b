#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/pair4; mkdir -p $O
QQQ_B200_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_pair.log 2>&1; echo "rc=$?" >> $O/pytest_pair.log
echo "--- pair=0" > $O/time.log
timeout 300 python probes/time_ours.py sweep 2>&1 | grep -v "M=    1 \|M=   16 " >> $O/time.log
timeout 200 python probes/time_ours.py llama 2>&1 | grep "M= 1024" >> $O/time.log
echo "--- pair=1" >> $O/time.log
QQQ_B200_PAIR=1 timeout 300 python probes/time_ours.py sweep 2>&1 | grep -v "M=    1 \|M=   16 " >> $O/time.log
QQQ_B200_PAIR=1 timeout 200 python probes/time_ours.py llama 2>&1 | grep "M= 1024" >> $O/time.log
for cfg in "1024 -1 4096 4096" "1024 -1 8192 21760"; do QQQ_B200_PAIR=1 QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 120 python probes/trace_timeline.py $cfg >> $O/trace.log 2>&1; done
echo done > $O/done.txt
