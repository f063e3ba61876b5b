#!/bin/bash
# round 2, call S (1 GPU, the last 3 GPU-minutes of the round): closing evidence for DESIGN §8 — what the decode GEMMs and the
# activation quantisation cost on their own (CUDA events) and one ncu --set full capture of each
cd "$(dirname "$0")/.."
O=gpurun_out/r2s; mkdir -p $O
timeout 70 python probes/final_evidence.py time > $O/time.log 2>&1
timeout 95 ncu --set full --clock-control none --import-source on -k regex:'act_quant|qqq_gemm' -s 7 -c 7 -f -o $O/ncu_final \
    python probes/final_evidence.py ncu > $O/ncu.log 2>&1
echo done > $O/done.txt
