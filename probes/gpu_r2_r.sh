#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-4}
O=gpurun_out/r2r_$N; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_tp$N.json 2> $O/bench_tp$N.err
if [ "$N" = "4" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_tp2.json 2> $O/bench_tp2.err
fi
echo done > $O/done.txt
