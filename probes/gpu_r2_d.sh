#!/bin/bash
# round 2, call D (2 GPUs): first hardware run of the fused exchange (scatter GEMM + reduce/quant/gather)
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
timeout 400 python -m pytest tests/test_zz_tp_scatter_gpu.py -m gpu -x -q > $O/pytest_tp_scatter.log 2>&1; echo "rc=$?" >> $O/pytest_tp_scatter.log
for mode in scatter nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 --tp-mode $mode > $O/bench_tp2_$mode.json 2> $O/bench_tp2_$mode.err
done
echo done > $O/done.txt
