#!/bin/bash
# Try tiling overrides (dev tooling): each line = "M gs NTOK KSUB NST"
cd "$(dirname "$0")/.."
while read -r M gs NTOK KSUB NST; do
  [ -z "$M" ] && continue
  echo -n "cfg ntok=$NTOK ksub=$KSUB nst=$NST | "
  QQQ_B200_NTOK=$NTOK QQQ_B200_KSUB=$KSUB QQQ_B200_NST=$NST timeout 120 python probes/time_ours.py one $M 8192 21760 $gs 2>&1 | tail -1
done <<CFG
1024 -1 256 1 4
1024 -1 256 1 3
1024 -1 128 2 4
1024 -1 128 2 3
1024 -1 128 1 6
1024 -1 192 1 5
1024 -1 64 2 6
4096 -1 256 1 4
4096 -1 128 2 4
4096 -1 192 1 5
1024 128 256 1 4
1024 128 128 2 4
256 -1 256 1 4
256 -1 128 2 4
256 -1 64 2 6
128 -1 128 2 4
128 -1 128 1 6
128 -1 64 2 6
128 -1 32 4 6
64 -1 64 2 6
64 -1 32 4 6
64 -1 64 4 3
16 -1 16 4 6
16 -1 16 2 6
16 -1 16 4 3
16 128 16 4 6
CFG
