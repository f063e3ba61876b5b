#!/bin/bash
# Runs every probe mode in its own process with a timeout (a hang in one mode must not take the others down).
cd "$(dirname "$0")"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for m in 0 1 2 3 4 5; do
  for n in 64 16 8 256; do
    timeout 60 ./probe_umma_i8 $m $n 2>&1 | grep -v "^device" ; echo "  (mode $m N=$n rc=${PIPESTATUS[0]})"
  done
done
for m in 10 11; do
  for n in 256 128 64 16; do
    timeout 120 ./probe_umma_i8 $m $n 2>&1 | grep BENCH
  done
done
