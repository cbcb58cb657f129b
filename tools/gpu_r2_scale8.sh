#!/bin/bash
# round 2, 8-GPU visit (gpurun --gpus 8): the driver's launch line at N = 8, 4, 2 (full contract line: kernel value, e2e
# through the drop-in call against the box's link ceiling with all ranks copying at once, extras incl. config 5 =
# 32M x 52 over 8 GPUs and the optional all-gather), and the two-device equality test
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_scale_box.txt; nproc >> gpurun_out/r2_scale_box.txt; (numactl -H 2>/dev/null || lscpu | grep -i numa) >> gpurun_out/r2_scale_box.txt
for N in 8 2; do
  timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench $N rc=$?"
  cut -c1-400 gpurun_out/r2_bench_${N}gpu.json
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_devices" 2>&1 | tail -2
