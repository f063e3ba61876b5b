#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/b8; mkdir -p $O
run() { echo "== $1" >> $O/stress.log; shift; env "$@" timeout 150 python probes/stress_eager.py 40 0 >> $O/stress.log 2>&1; echo "rc=$?" >> $O/stress.log; }
run "fixed default M=4096" X=1
run "fixed default M=4096 again" X=1
run "fixed default M=2048" STRESS_M=2048
run "fixed NST=4" QQQ_B200_NST=4
echo "== full stress with chains" >> $O/stress.log; timeout 300 python probes/stress_eager.py 10 10 8 >> $O/stress.log 2>&1; echo "rc=$?" >> $O/stress.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 400 python probes/time_ours.py sweep > $O/time_sweep.log 2>&1
timeout 300 python probes/time_ours.py llama > $O/time_llama.log 2>&1
for cfg in "1024 8192 21760" "4096 8192 21760" "1024 4096 4096" "1024 4096 11008" "1024 11008 4096"; do
  set -- $cfg
  for nst in 3 4 5; do
    echo -n "NST=$nst | " >> $O/nst.log
    QQQ_B200_NST=$nst timeout 120 python probes/time_ours.py one $1 $2 $3 -1 2>&1 | tail -1 >> $O/nst.log
  done
done
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qqq_gemm_kernel|act_quant_kernel" --launch-skip 1344 -c 896 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-sweep --no-cpu --no-merged > $O/bench_under_ncu.log 2>&1
echo done > $O/done.txt
