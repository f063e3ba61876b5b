#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2j; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 probes/tp_pieces.py > $O/tp_pieces.log 2>&1
timeout 300 python -m pytest tests/test_zz_tp_scatter_gpu.py -m gpu -x -q > $O/pytest_tp.log 2>&1; echo "rc=$?" >> $O/pytest_tp.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_tp2.json 2> $O/bench_tp2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-overlap --no-70b --no-tp-sweep > $O/bench_tp2_nooverlap.json 2> $O/bench_tp2_nooverlap.err
echo done > $O/done.txt
