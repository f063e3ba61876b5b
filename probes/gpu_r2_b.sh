#!/bin/bash
# round 2, call B (1 GPU): baseline on this round's box, drain variants (208-token tiles, drain helpers), floor of small shards, traces
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
python probes/build_variant.py probes/libqqq_b200_dbuf208.so -DQQQ_DBUF_MAX_TOK=208 > $O/build.log 2>&1
python probes/build_variant.py probes/libqqq_b200_helpers.so -DQQQ_DRAIN_HELPERS >> $O/build.log 2>&1
python probes/build_variant.py probes/libqqq_b200_trace.so -DQQQ_TRACE -DQQQ_TRACE_CTA=5 >> $O/build.log 2>&1
T="timeout 100 python probes/time_ours.py one"
for cfg in "1024 4096 4096 -1" "1024 4096 11008 -1" "1024 11008 4096 -1" "1024 8192 21760 -1" "1024 8192 21760 128" "4096 8192 21760 -1"; do
  echo "--- default: $cfg" >> $O/time_variants.log;  $T $cfg >> $O/time_variants.log 2>&1
  echo "--- ntok128: $cfg" >> $O/time_variants.log;  QQQ_B200_NTOK=128 $T $cfg >> $O/time_variants.log 2>&1
  echo "--- ntok192: $cfg" >> $O/time_variants.log;  QQQ_B200_NTOK=192 $T $cfg >> $O/time_variants.log 2>&1
  echo "--- dbuf208: $cfg" >> $O/time_variants.log;  QQQ_B200_NTOK=208 QQQ_B200_LIB=probes/libqqq_b200_dbuf208.so $T $cfg >> $O/time_variants.log 2>&1
  for split in -1 0 1; do
    echo "--- helpers split=$split: $cfg" >> $O/time_variants.log; QQQ_B200_SPLIT=$split QQQ_B200_LIB=probes/libqqq_b200_helpers.so $T $cfg >> $O/time_variants.log 2>&1
    echo "--- default split=$split: $cfg" >> $O/time_variants.log; QQQ_B200_SPLIT=$split $T $cfg >> $O/time_variants.log 2>&1
  done
done
QQQ_B200_LIB=probes/libqqq_b200_helpers.so timeout 400 python -m pytest tests/test_gemm_parity.py tests/test_qlinear_gpu.py -m gpu -x -q > $O/pytest_helpers.log 2>&1; echo "rc=$?" >> $O/pytest_helpers.log
QQQ_B200_NTOK=208 QQQ_B200_LIB=probes/libqqq_b200_dbuf208.so timeout 400 python -m pytest tests/test_gemm_parity.py -m gpu -x -q > $O/pytest_dbuf208.log 2>&1; echo "rc=$?" >> $O/pytest_dbuf208.log
# floor of the small tensor-parallel shards and of the decode GEMMs
for cfg in "1024 4096 2048 -1" "1024 2048 4096 -1" "1024 4096 512 -1" "1024 512 4096 -1" "1024 4096 1408 -1" "1024 1408 4096 -1" "1024 4096 1536 -1" "1024 128 128 -1" \
           "32 4096 4096 128" "32 4096 1024 128" "32 4096 6144 128" "32 4096 14336 128" "32 4096 28672 128" "32 14336 4096 128" "32 4096 4096 -1" "32 128 128 -1"; do
  $T $cfg >> $O/time_small.log 2>&1
  QQQ_B200_LIB=probes/libqqq_b200_helpers.so $T $cfg >> $O/time_small_helpers.log 2>&1
done
for cfg in "1024 -1 4096 4096" "1024 -1 4096 512" "1024 -1 512 4096" "32 128 4096 4096" "32 128 4096 1024" "32 -1 4096 4096" "128 -1 8192 21760"; do
  QQQ_B200_LIB=probes/libqqq_b200_trace.so timeout 100 python probes/trace_timeline.py $cfg >> $O/traces.log 2>&1
done
timeout 600 python bench.py --no-full --no-cpu > $O/bench.json 2> $O/bench.err
echo done > $O/done.txt
