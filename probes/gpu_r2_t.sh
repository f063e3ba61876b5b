#!/bin/bash
# round 2, call T (1 GPU, the last GPU-minutes): ncu launch list of the bench command on the final HEAD (shares of a step).
# As run in this round it had no kernel filter and `-s 500` fell inside the model set-up (the zero-mean synthetic weights
# take more torch launches than before), so its capture holds no product kernel and was not kept; profiles/r02/call_k holds
# the launch list of the same command.  The filter below is the fix.
cd "$(dirname "$0")/.."
O=gpurun_out/r2t; mkdir -p $O
timeout 95 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'qqq_gemm|act_quant' -s 352 -c 352 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu --no-merged --no-decode --no-full > $O/bench_under_ncu.log 2>&1
echo "rc=$?" > $O/done.txt
