#!/bin/bash
# Pair-mode record (dev tooling): auto policy vs QQQ_B200_PAIR=0 on the two many-wave sweep shapes.
cd "$(dirname "$0")/.."
O=gpurun_out/last; mkdir -p $O
( echo "--- pair=auto"; timeout 26 python - <<'PY'
import sys; sys.argv=["x"]; sys.path.insert(0,"probes")
import time_ours as t
for gs in (-1,128):
    for M in (1024,4096): t.run(M,8192,21760,gs)
PY
echo "--- pair=0"; QQQ_B200_PAIR=0 timeout 26 python - <<'PY'
import sys; sys.argv=["x"]; sys.path.insert(0,"probes")
import time_ours as t
for gs in (-1,128):
    for M in (1024,4096): t.run(M,8192,21760,gs)
PY
) > $O/pair_mode.log 2>&1
cat $O/pair_mode.log
