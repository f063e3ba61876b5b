#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
timeout 300 python probes/diag_flaky.py 1024 4096 1024 128 150 > $O/diag_default.log 2>&1
QQQ_B200_SPLIT=0 timeout 300 python probes/diag_flaky.py 1024 4096 1024 128 100 > $O/diag_split0.log 2>&1
timeout 300 python probes/diag_flaky.py 1024 4096 1024 -1 100 > $O/diag_pc.log 2>&1
QQQ_B200_PDL=0 timeout 300 python probes/diag_flaky.py 1024 4096 1024 128 100 > $O/diag_nopdl.log 2>&1
echo done > $O/done.txt
