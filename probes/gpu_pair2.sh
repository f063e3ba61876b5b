#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/pair2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "rc=$?" >> $O/pytest_default.log
QQQ_B200_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_pair.log 2>&1; echo "rc=$?" >> $O/pytest_pair.log
QQQ_B200_PAIR=1 timeout 150 python probes/stress_eager.py 20 4 8 > $O/stress_pair.log 2>&1; echo "rc=$?" >> $O/stress_pair.log
echo "--- pair=0" > $O/time.log
timeout 300 python probes/time_ours.py sweep 2>&1 | grep -v "M=    1 \|M=   16 " >> $O/time.log
timeout 200 python probes/time_ours.py llama 2>&1 | grep "M= 1024" >> $O/time.log
echo "--- pair=1" >> $O/time.log
QQQ_B200_PAIR=1 timeout 300 python probes/time_ours.py sweep 2>&1 | grep -v "M=    1 \|M=   16 " >> $O/time.log
QQQ_B200_PAIR=1 timeout 200 python probes/time_ours.py llama 2>&1 | grep "M= 1024" >> $O/time.log
echo done > $O/done.txt
