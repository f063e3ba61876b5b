#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/pair1; mkdir -p $O
for cfg in "32 1024 256 -1" "64 512 256 -1" "256 1024 512 -1" "300 768 640 128" "1024 4096 4096 -1" "128 8192 2048 -1"; do
  echo "== $cfg" >> $O/check.log
  QQQ_B200_PAIR=1 timeout 90 python probes/pair_check.py $cfg >> $O/check.log 2>&1; echo "rc=$?" >> $O/check.log
done
echo done > $O/done.txt
