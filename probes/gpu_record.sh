#!/bin/bash
# Record battery (dev tooling): the numbers and profiler evidence that go into profiles/.  usage: gpu_record.sh TAG
cd "$(dirname "$0")/.."
O=gpurun_out/$1; mkdir -p $O/ncu
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
[ -z "$SKIP_REF" ] && timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qqq_gemm_kernel|act_quant_kernel" --launch-skip 1344 -c 896 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-sweep --no-cpu --no-merged --no-decode --no-full > $O/bench_under_ncu.log 2>&1
M="dram__bytes_read.sum,dram__bytes_write.sum"
for cfg in ${NCU_CFGS:-"16 -1 8192 21760" "1024 -1 8192 21760" "16 128 8192 21760" "1024 128 8192 21760" "1024 -1 4096 4096" "1024 -1 4096 11008" "1024 -1 11008 4096"}; do
  set -- $cfg
  name=m$1_g$2_k$3_n$4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:qqq_gemm_kernel --launch-skip 3 -c 1 -o $O/ncu/$name -f python probes/run_one.py $1 $2 5 $3 $4 > $O/ncu/$name.log 2>&1
  python probes/ncuget.py $O/ncu/$name.ncu-rep gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' 'lts__throughput.avg.pct_of_peak_sustained_elapsed' 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed' 'l1tex__m_xbar2l1tex_read_bytes.sum' 'sm__throughput.avg.pct_of_peak_sustained_elapsed' 'smsp__inst_executed.sum' 'sm__cycles_elapsed.max' 'launch__grid_size' 'launch__registers_per_thread' 'launch__shared_mem_per_block_dynamic' 'sm__inst_executed_pipe_tensor*' > $O/ncu/$name.txt 2>&1
done
name=actquant_m1024_k4096
timeout 200 ncu --set full --clock-control none -k regex:act_quant_kernel --launch-skip 2 -c 1 -o $O/ncu/$name -f python -c "
import torch, sys; sys.path.insert(0,'.'); import qqq_b200
x=torch.randn(1024,4096,device='cuda').half()
for i in range(4): qqq_b200.dynamic_quant(x)
torch.cuda.synchronize()" > $O/ncu/$name.log 2>&1
python probes/ncuget.py $O/ncu/$name.ncu-rep gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed' 'sm__throughput.avg.pct_of_peak_sustained_elapsed' 'launch__grid_size' 'launch__registers_per_thread' 'sm__warps_active.avg.pct_of_peak_sustained_active' > $O/ncu/$name.txt 2>&1
# keep only two full reports (64 MiB cap on gpurun_out): the headline shapes
ls -la $O/ncu > $O/ncu_files.txt
for f in $O/ncu/*.ncu-rep; do case $f in *m16_g-1_k8192*|*m1024_g-1_k4096_n4096*) ;; *) rm -f $f;; esac; done
echo done > $O/done.txt
