#!/bin/bash
# round 2, call M (1 GPU): mid-size M (64..512): remainder-first split (SPLIT=2), pairs, unpack groups
cd "$(dirname "$0")/.."
O=gpurun_out/r2m; mkdir -p $O
T="timeout 100 python probes/time_ours.py one"
for cfg in "64 8192 21760 -1" "128 8192 21760 -1" "128 8192 21760 128" "256 8192 21760 -1" "256 8192 21760 128" "512 8192 21760 -1" "128 4096 4096 -1" "128 4096 11008 -1"; do
  for v in "SPLIT=-1 PAIR=-1 GROUPS=0" "SPLIT=2 PAIR=-1 GROUPS=0" "SPLIT=0 PAIR=-1 GROUPS=0" "SPLIT=2 PAIR=1 GROUPS=0" "SPLIT=1 PAIR=1 GROUPS=0" "SPLIT=2 PAIR=-1 GROUPS=3" "SPLIT=-1 PAIR=-1 GROUPS=3" "SPLIT=2 PAIR=-1 NTOK=64" "SPLIT=2 PAIR=-1 NTOK=128"; do
    set -- $v
    echo "--- $v: $cfg" >> $O/time.log; env QQQ_B200_${1} QQQ_B200_${2} QQQ_B200_${3} $T $cfg >> $O/time.log 2>&1
  done
done
python probes/build_variant.py probes/libqqq_b200_trace.so -DQQQ_TRACE -DQQQ_TRACE_CTA=5 > $O/build.log 2>&1
for v in "SPLIT=-1" "SPLIT=2"; do
  echo "##### $v" >> $O/traces.log
  env QQQ_B200_$v QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 100 python probes/trace_timeline.py 128 -1 8192 21760 >> $O/traces.log 2>&1
done
echo done > $O/done.txt
