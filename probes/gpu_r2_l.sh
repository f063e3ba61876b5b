#!/bin/bash
# round 2, call L (4 GPUs): the fused exchange at world 4 — pieces, then the bench line the driver will ask for
cd "$(dirname "$0")/.."
N=${1:-4}
O=gpurun_out/r2l_$N; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 probes/tp_pieces.py > $O/tp_pieces.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_tp$N.json 2> $O/bench_tp$N.err
echo done > $O/done.txt
