#!/bin/bash
# round 2, call N (1 GPU): final-candidate validation: full suite, bench line, planner picks at mid M, decode whole-tile check
cd "$(dirname "$0")/.."
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
T="timeout 100 python probes/time_ours.py one"
for cfg in "128 4096 4096 -1" "256 4096 4096 -1" "200 4096 4096 -1" "256 4096 11008 -1" "128 8192 21760 -1" "256 8192 21760 -1" "512 8192 21760 -1" "256 4096 1024 -1"; do
  echo "--- planner: $cfg" >> $O/time.log; $T $cfg >> $O/time.log 2>&1
  echo "--- ntok256 split0: $cfg" >> $O/time.log; QQQ_B200_NTOK=256 QQQ_B200_SPLIT=0 $T $cfg >> $O/time.log 2>&1
  echo "--- ntok64 split0: $cfg" >> $O/time.log; QQQ_B200_NTOK=64 QQQ_B200_SPLIT=0 $T $cfg >> $O/time.log 2>&1
done
for cfg in "32 4096 4096 128" "32 4096 1024 128" "32 4096 14336 128" "32 14336 4096 128" "16 4096 4096 -1" "1 4096 4096 -1" "64 4096 4096 -1"; do
  echo "--- planner: $cfg" >> $O/time_decode.log; $T $cfg >> $O/time_decode.log 2>&1
  echo "--- split0: $cfg" >> $O/time_decode.log; QQQ_B200_SPLIT=0 $T $cfg >> $O/time_decode.log 2>&1
  echo "--- split0 ksub2: $cfg" >> $O/time_decode.log; QQQ_B200_SPLIT=0 QQQ_B200_KSUB=2 $T $cfg >> $O/time_decode.log 2>&1
done
echo done > $O/done.txt
