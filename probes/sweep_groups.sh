#!/bin/bash
cd "$(dirname "$0")/.."
while read -r M gs G; do
  [ -z "$M" ] && continue
  echo -n "groups=$G | "
  QQQ_B200_GROUPS=$G timeout 120 python probes/time_ours.py one $M 8192 21760 $gs 2>&1 | tail -1
done <<CFG
16 -1 2
16 -1 3
16 -1 4
16 128 3
16 128 4
128 -1 2
128 -1 3
128 128 3
128 128 4
1024 -1 2
1024 -1 3
1024 128 3
1024 128 4
4096 -1 2
4096 128 3
CFG
