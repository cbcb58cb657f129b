#!/bin/bash
# Development loop on the GPU box: parity tests, then kernel-only timings of the fk configs.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in fk_1m_x_22 fk_4m_x_65 fk_4m_x_52; do
  for c in 4 8; do
    PMB_FK_CHUNK=$c timeout 300 python bench.py --kernel-only --steps 50 --warmup 5 --workload $wl 2>&1 | tail -1
  done
done | tee gpurun_out/quick.jsonl
