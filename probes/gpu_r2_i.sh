#!/bin/bash
# round 2, call I (1 GPU): full GPU suite (bias epilogue, compact variant, forced planner modes), decode with/without the compact variant
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
T="timeout 100 python probes/time_ours.py one"
for cfg in "32 4096 4096 128" "32 4096 1024 128" "32 4096 6144 128" "32 4096 14336 128" "32 4096 28672 128" "32 14336 4096 128" "32 4096 4096 -1" "1 8192 21760 -1" "16 8192 21760 -1" "16 8192 21760 128" "64 8192 21760 -1" "1 4096 4096 -1" "32 128 128 -1"; do
  for c in 0 1; do
    echo "--- compact=$c: $cfg" >> $O/time_decode.log; QQQ_B200_COMPACT=$c $T $cfg >> $O/time_decode.log 2>&1
  done
done
for c in 0 1; do
  QQQ_B200_COMPACT=$c timeout 600 python bench.py --no-full --no-cpu --no-sweep > $O/bench_compact$c.json 2> $O/bench_compact$c.err
done
python probes/build_variant.py probes/libqqq_b200_trace.so -DQQQ_TRACE -DQQQ_TRACE_CTA=5 > $O/build.log 2>&1
for cfg in "32 128 4096 4096" "32 128 4096 28672" "16 -1 8192 21760"; do
  QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 100 python probes/trace_timeline.py $cfg >> $O/traces.log 2>&1
done
timeout 900 python bench.py --no-full > $O/bench.json 2> $O/bench.err
echo done > $O/done.txt
