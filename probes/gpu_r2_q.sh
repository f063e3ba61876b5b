#!/bin/bash
# round 2, call Q (1 GPU): the record run of the final tree
cd "$(dirname "$0")/.."
O=gpurun_out/r2q; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo done > $O/done.txt
