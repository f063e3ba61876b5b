#!/bin/bash
# GPU battery 2 (dev tooling): TMA-store epilogue + compact split-K scratch; logs under gpurun_out/b2/
cd "$(dirname "$0")/.."
O=gpurun_out/b2; mkdir -p $O
timeout 60 ./probes/probe_pipes > $O/probe_pipes.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 400 python probes/time_ours.py sweep > $O/time_sweep.log 2>&1
timeout 300 python probes/time_ours.py llama > $O/time_llama.log 2>&1
for cfg in "1024 -1 2" "1024 -1 3" "4096 -1 3" "1024 128 3" "128 -1 3" "256 -1 3" "16 -1 2" "16 128 2"; do
  set -- $cfg
  echo -n "groups=$3 | " >> $O/groups.log
  QQQ_B200_GROUPS=$3 timeout 120 python probes/time_ours.py one $1 8192 21760 $2 2>&1 | tail -1 >> $O/groups.log
done
for shape in "4096 4096" "4096 11008" "11008 4096"; do
  set -- $shape
  for nt in 256 192 128; do
    echo -n "ntok=$nt | " >> $O/ntok.log
    QQQ_B200_NTOK=$nt timeout 120 python probes/time_ours.py one 1024 $1 $2 -1 2>&1 | tail -1 >> $O/ntok.log
  done
done
for cfg in "1024 -1" "128 -1" "16 -1" "1024 128"; do
  set -- $cfg
  QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 120 python probes/trace_timeline.py $1 $2 >> $O/trace.log 2>&1
done
QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 120 python probes/trace_timeline.py 1024 -1 4096 4096 >> $O/trace.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qqq_gemm_kernel|act_quant_kernel" --launch-skip 1344 -c 896 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-sweep --no-cpu --no-merged > $O/bench_under_ncu.log 2>&1
echo done > $O/done.txt
