#!/bin/bash
# round 2, call P (N GPUs): final multi-GPU validation: tests (2 GPUs), pieces, the bench line
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r2p_$N; mkdir -p $O
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_zz_tp_scatter_gpu.py tests/test_zz_tp_fused_gpu.py -m gpu -x -q > $O/pytest_tp.log 2>&1; echo "rc=$?" >> $O/pytest_tp.log
fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 probes/tp_pieces.py > $O/tp_pieces.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_tp$N.json 2> $O/bench_tp$N.err
echo done > $O/done.txt
