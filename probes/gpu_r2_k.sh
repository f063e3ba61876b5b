#!/bin/bash
# round 2, call K (1 GPU): full suite on the final kernels, headline with / without graph branches, ncu evidence
cd "$(dirname "$0")/.."
O=gpurun_out/r2k; mkdir -p $O
nvidia-smi > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --no-full > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --no-full --no-cpu --no-sweep --no-overlap > $O/bench_nooverlap.json 2> $O/bench_nooverlap.err
T="timeout 100 python probes/time_ours.py one"
for cfg in "1 8192 21760 -1" "16 8192 21760 -1" "128 8192 21760 -1" "1024 4096 4096 -1" "32 4096 4096 128"; do $T $cfg >> $O/time.log 2>&1; done
# ncu: launch list of the bench command (shares, cold-cache) and full captures of the three Llama-2-7B GEMM shapes + the sweep shape
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 330 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu --no-merged --no-decode --no-full > $O/bench_under_ncu.log 2>&1
for cfg in "1024 4096 4096 -1" "1024 4096 11008 -1" "1024 11008 4096 -1" "1024 8192 21760 -1" "1024 8192 21760 128" "16 8192 21760 -1"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:qqq_gemm -s 6 -c 1 -o $O/ncu_$tag python probes/time_ours.py one $cfg > $O/ncu_$tag.log 2>&1
done
echo done > $O/done.txt
