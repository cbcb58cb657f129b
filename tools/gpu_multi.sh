#!/bin/bash
# Multi-GPU sanity visit (run under gpurun --gpus N): the driver's launch line for N ranks, then quick GPU tests.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_box.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench $N rc=$?"
cat gpurun_out/bench_${N}gpu.json | cut -c1-900; tail -3 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 20 --warmup 3 --workload fk_4m_x_52 --no-cpu-baseline > gpurun_out/bench_${N}gpu_52.json 2>> gpurun_out/bench_${N}gpu.err; echo "bench52 $N rc=$?"
cat gpurun_out/bench_${N}gpu_52.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_${N}gpu_ref.json 2>> gpurun_out/bench_${N}gpu.err; echo "ref $N rc=$?"
cut -c1-300 gpurun_out/bench_${N}gpu_ref.json
timeout 900 python -m pytest tests -m gpu -q -k "fk_quat or from_root_positions or mirror or sharding" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_multi.log
