#!/bin/bash
# One GPU-box visit: environment facts, smoke, GPU parity tests, bench (both arms), ncu launch list,
# ncu full capture of the fk kernel.
set -u
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv
  echo "host cores: $(nproc)"; free -g | head -2; lscpu | grep -E "Model name|^CPU\(s\)"
} > gpurun_out/box.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err
timeout 600 python bench.py > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_ -s 3 -c 1 -f -o gpurun_out/prof_fk \
    python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_ -s 3 -c 1 -f -o gpurun_out/prof_fk_52 \
    python bench.py --kernel-only --steps 3 --warmup 3 --workload fk_4m_x_52 > gpurun_out/ncu_full_52.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_quat -s 3 -c 1 -f -o gpurun_out/prof_fkq \
    python bench.py --kernel-only --steps 3 --warmup 3 --op fk_quat > gpurun_out/ncu_full_fkq.log 2>&1
python tools/bench_elementwise.py > gpurun_out/elementwise.jsonl 2> gpurun_out/elementwise.err
bash tools/gpu_ops.sh > /dev/null 2>&1
ls -la gpurun_out | head -30
