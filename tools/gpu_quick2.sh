#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
rm -f gpurun_out/r2_ops_fk.jsonl
for wl in fk_1m_x_22 fk_2m_x_16 fk_2m_x_24 fk_2m_x_32 fk_2m_x_40 fk_4m_x_52 fk_4m_x_65; do
  timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --op fk --workload $wl >> gpurun_out/r2_ops_fk.jsonl
done
cut -c1-300 gpurun_out/r2_ops_fk.jsonl
