#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/b14; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 150 python probes/stress_eager.py 20 4 8 > $O/stress.log 2>&1; echo "rc=$?" >> $O/stress.log
timeout 400 python probes/time_ours.py sweep > $O/time_sweep.log 2>&1
timeout 300 python probes/time_ours.py llama > $O/time_llama.log 2>&1
echo "forced stream-K (QQQ_B200_SPLIT=1)" > $O/time_llama_split.log
QQQ_B200_SPLIT=1 timeout 300 python probes/time_ours.py llama >> $O/time_llama_split.log 2>&1
for shape in "1024 4096 12288" "1024 4096 22016"; do
  set -- $shape
  echo -n "model | " >> $O/merged.log; timeout 100 python probes/time_ours.py one $1 $2 $3 -1 2>&1 | tail -1 >> $O/merged.log
  echo -n "split | " >> $O/merged.log; QQQ_B200_SPLIT=1 timeout 100 python probes/time_ours.py one $1 $2 $3 -1 2>&1 | tail -1 >> $O/merged.log
done
for cfg in "1024 -1 4096 4096" "1024 -1 8192 21760"; do QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 120 python probes/trace_timeline.py $cfg >> $O/trace.log 2>&1; done
timeout 600 python bench.py --no-cpu > $O/bench.json 2> $O/bench.err
echo done > $O/done.txt
