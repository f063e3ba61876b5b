#!/bin/bash
# round 2, call O (1 GPU): final-candidate validation after the planner's decode / small-tile changes and the pipelined e2e path
cd "$(dirname "$0")/.."
O=gpurun_out/r2o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
T="timeout 100 python probes/time_ours.py one"
for cfg in "128 4096 4096 -1" "256 4096 4096 -1" "256 4096 1024 -1" "1024 4096 512 -1" "1024 4096 1408 -1" "32 4096 4096 128" "32 4096 1024 128" "32 4096 14336 128" "32 14336 4096 128" "32 4096 6144 128" "32 4096 28672 128" "16 4096 4096 -1" "1 4096 4096 -1" "16 8192 21760 -1"; do $T $cfg >> $O/time.log 2>&1; done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo done > $O/done.txt
