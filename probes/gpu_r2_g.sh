#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
for v in "0 1184" "1 1184" "1 148" "3 148" "11 148" "27 148" "1 296" "0 148"; do
  set -- $v
  echo "== TPDBG=$1 TPGRID=$2" >> $O/tp_pieces.log
  QQQ_B200_TPDBG=$1 QQQ_B200_TPGRID=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 probes/tp_pieces.py 2>&1 | grep world >> $O/tp_pieces.log
done
echo done > $O/done.txt
