#!/bin/bash
# First GPU call of the next round (dev tooling): confirms everything written after round 1's GPU budget was spent.
#   1 GPU :  gpurun --timeout 900 -- bash probes/gpu_round2_first.sh
#   2 GPUs:  gpurun --gpus 2 --timeout 900 -- bash probes/gpu_round2_first.sh tp
cd "$(dirname "$0")/.."
O=gpurun_out/r2_first; mkdir -p $O
if [ "$1" = "tp" ]; then
  # fused GEMM + all-reduce (multimem.red epilogue): parity first, then the TP=2 bench with and without it
  QQQ_B200_MULTI_GPU_TESTS=1 timeout 400 python -m pytest tests/test_zz_tp_fused_gpu.py -m gpu -x -q > $O/pytest_tp_fused.log 2>&1; echo "rc=$?" >> $O/pytest_tp_fused.log
  for extra in "" "--fused-allreduce"; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 5 --warmup 3 $extra > $O/bench_tp2$extra.json 2> $O/bench_tp2$extra.err
  done
else
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
  timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
  QQQ_B200_PAIR=1 timeout 600 python -m pytest tests/test_gemm_parity.py tests/test_qlinear_gpu.py -m gpu -x -q > $O/pytest_pair_forced.log 2>&1; echo "rc=$?" >> $O/pytest_pair_forced.log
  # pair mode where it is still off by policy: single token tile, 64-256 tokens (trace one CTA pair at M=128)
  for p in 0 1; do
    echo "--- pair=$p" >> $O/time_pair_midM.log
    QQQ_B200_PAIR=$p timeout 200 python - >> $O/time_pair_midM.log 2>&1 <<'PY'
import sys; sys.argv=["x"]; sys.path.insert(0,"probes")
import time_ours as t
for gs in (-1,128):
    for M in (64,128,256,512): t.run(M,8192,21760,gs)
PY
  done
  # 208-token double-buffered tiles (drain overlapped; 5 x 208 covers 1024) against the default 256-token tiles
  python probes/build_variant.py probes/libqqq_b200_dbuf208.so -DQQQ_DBUF_MAX_TOK=208 > $O/build_dbuf208.log 2>&1
  for cfg in "1024 8192 21760 -1" "1024 4096 4096 -1" "1024 4096 11008 -1" "1024 11008 4096 -1" "1024 8192 21760 128" "4096 8192 21760 -1"; do
    echo "--- default: $cfg" >> $O/time_dbuf208.log
    timeout 100 python probes/time_ours.py one $cfg >> $O/time_dbuf208.log 2>&1
    echo "--- dbuf208: $cfg" >> $O/time_dbuf208.log
    QQQ_B200_NTOK=208 QQQ_B200_LIB=probes/libqqq_b200_dbuf208.so timeout 100 python probes/time_ours.py one $cfg >> $O/time_dbuf208.log 2>&1
  done
  QQQ_B200_NTOK=208 QQQ_B200_LIB=probes/libqqq_b200_dbuf208.so timeout 600 python -m pytest tests/test_gemm_parity.py -m gpu -x -q > $O/pytest_dbuf208.log 2>&1; echo "rc=$?" >> $O/pytest_dbuf208.log
  # drain helpers (unpack warps take a share of the accumulator drain of whole tiles): parity first, then timing,
  # with the planner's own split decision and with whole tiles forced
  python probes/build_variant.py probes/libqqq_b200_helpers.so -DQQQ_DRAIN_HELPERS > $O/build_helpers.log 2>&1
  QQQ_B200_LIB=probes/libqqq_b200_helpers.so timeout 600 python -m pytest tests/test_gemm_parity.py tests/test_qlinear_gpu.py -m gpu -x -q > $O/pytest_helpers.log 2>&1; echo "rc=$?" >> $O/pytest_helpers.log
  for cfg in "1024 8192 21760 -1" "1024 4096 4096 -1" "1024 4096 11008 -1" "1024 11008 4096 -1" "4096 8192 21760 -1" "128 8192 21760 -1" "1024 8192 21760 128"; do
    for split in -1 0; do   # -1: the planner decides (same as unset); 0: whole tiles only
      echo "--- default split=$split: $cfg" >> $O/time_helpers.log
      QQQ_B200_SPLIT=$split timeout 100 python probes/time_ours.py one $cfg >> $O/time_helpers.log 2>&1
      echo "--- helpers split=$split: $cfg" >> $O/time_helpers.log
      QQQ_B200_SPLIT=$split QQQ_B200_LIB=probes/libqqq_b200_helpers.so timeout 100 python probes/time_ours.py one $cfg >> $O/time_helpers.log 2>&1
    done
  done
fi
echo done > $O/done.txt
