#!/bin/bash
cd "$(dirname "$0")/.."
while read -r M gs NTOK KSUB; do
  [ -z "$M" ] && continue
  echo -n "cfg ntok=$NTOK ksub=$KSUB | "
  QQQ_B200_NTOK=$NTOK QQQ_B200_KSUB=$KSUB timeout 120 python probes/time_ours.py one $M 8192 21760 $gs 2>&1 | tail -1
done <<CFG
1024 -1 256 1
1024 -1 192 1
1024 -1 192 2
1024 -1 160 1
1024 -1 160 2
1024 -1 128 2
4096 -1 256 1
4096 -1 192 1
4096 -1 192 2
4096 -1 160 2
1024 128 192 1
1024 128 192 2
CFG
