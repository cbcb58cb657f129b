#!/bin/bash
# 8-GPU scaling visit (gpurun --gpus 8): the driver's launch line at N = 8 and N = 4, headline workload and the
# 52-joint shard of BASELINE config 5 (32M x 52 over 8 GPUs = 4M x 52 per GPU).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_box.txt
for N in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench $N rc=$?"
  cut -c1-260 gpurun_out/bench_${N}gpu.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 \
    bench.py --gpus 8 --steps 20 --warmup 3 --workload fk_4m_x_52 --no-cpu-baseline > gpurun_out/bench_8gpu_52.json 2>> gpurun_out/bench_8gpu.err; echo "bench52 rc=$?"
cut -c1-260 gpurun_out/bench_8gpu_52.json
