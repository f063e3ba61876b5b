#!/bin/bash
# round 2, call F (1 GPU): full GPU suite on the new planner / round-robin schedule; what bounds a k-block (trace markers:
# tokens vs weights); ncu of the Llama-7B GEMMs; planner choices against forced alternatives
cd "$(dirname "$0")/.."
O=gpurun_out/r2f; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
python probes/build_variant.py probes/libqqq_b200_trace.so -DQQQ_TRACE -DQQQ_TRACE_CTA=5 > $O/build.log 2>&1
TR="timeout 100 python probes/trace_timeline.py"
for cfg in "1024 -1 4096 4096" "1024 -1 4096 11008" "1024 -1 8192 21760"; do
  for v in "NTOK=256 PAIR=0" "NTOK=256 PAIR=1" "NTOK=208 PAIR=0" "NTOK=128 PAIR=0"; do
    set -- $v
    echo "##### $v : $cfg" >> $O/traces.log
    env QQQ_B200_${1} QQQ_B200_${2} QQQ_B200_SPLIT=0 QQQ_B200_LIB=probes/libqqq_b200_trace.so $TR $cfg >> $O/traces.log 2>&1
  done
done
T="timeout 100 python probes/time_ours.py one"
for cfg in "1024 4096 4096 -1" "1024 4096 11008 -1" "1024 11008 4096 -1" "1024 8192 21760 -1" "1024 8192 21760 128" "4096 8192 21760 -1" "128 8192 21760 -1" "256 8192 21760 -1" "512 8192 21760 -1" "2048 4096 4096 -1" "1024 4096 2048 -1" "1024 2048 4096 -1" "1024 4096 5504 -1" "1024 5504 4096 -1" "1024 4096 512 -1" "1024 512 4096 -1" "1024 8192 3584 -1" "1024 3584 8192 -1" "1024 8192 1024 -1" "1024 1024 8192 -1"; do
  echo "--- planner: $cfg" >> $O/time.log;  $T $cfg >> $O/time.log 2>&1
  for v in "NTOK=256 PAIR=0 SPLIT=0" "NTOK=256 PAIR=1 SPLIT=0" "NTOK=256 PAIR=0 SPLIT=1" "NTOK=208 PAIR=0 SPLIT=0" "NTOK=208 PAIR=0 SPLIT=1" "NTOK=128 PAIR=0 SPLIT=0" "NTOK=128 PAIR=0 SPLIT=1"; do
    set -- $v
    echo "--- $v: $cfg" >> $O/time.log; env QQQ_B200_${1} QQQ_B200_${2} QQQ_B200_${3} $T $cfg >> $O/time.log 2>&1
  done
done
for cfg in "1024 4096 4096 -1" "1024 4096 11008 -1" "1024 8192 21760 -1"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:qqq_gemm -s 6 -c 1 -o $O/ncu_$tag python probes/time_ours.py one $cfg > $O/ncu_$tag.log 2>&1
done
echo done > $O/done.txt
