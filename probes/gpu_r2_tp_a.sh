#!/bin/bash
# round 2, call A (2 GPUs): capabilities probe + first hardware run of the fused / exact TP paths
cd "$(dirname "$0")/.."
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 probes/probe_symm.py > $O/probe_symm.log 2>&1; echo "rc=$?" >> $O/probe_symm.log
QQQ_B200_MULTI_GPU_TESTS=1 timeout 300 python -m pytest tests/test_zz_tp_fused_gpu.py -m gpu -x -q > $O/pytest_tp_fused.log 2>&1; echo "rc=$?" >> $O/pytest_tp_fused.log
for extra in "" "--fused-allreduce"; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 $extra > $O/bench_tp2$extra.json 2> $O/bench_tp2$extra.err
done
echo done > $O/done.txt
