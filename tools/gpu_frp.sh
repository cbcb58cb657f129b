#!/bin/bash
# from_root_positions: parity tests + error statistics + kernel-only timing at the three bench shapes, tile kernel vs thread-per-frame kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ik.py -x -q -m gpu 2>&1 | tail -6
timeout 600 python tests/dev/frp_error_stats.py 3001 1000 517 20000 6000 4000 2>/dev/null | cut -c1-330 | tee gpurun_out/r2_frp_error_stats.jsonl
rm -f gpurun_out/frp_now.jsonl
export PMB_EXPERIMENT=1
for cfg in "PMB_FRP_TILE=0" "PMB_FRP_TILE=1" "PMB_FRP_TILE=1 PMB_FRP_FAST=0" "PMB_FRP_TILE=0 PMB_FRP_FAST=0"; do
  for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
    env $cfg timeout 300 python bench.py --kernel-only --steps 20 --warmup 3 --op from_root_positions --workload $wl | sed "s/^{/{\"cfg\": \"$cfg\", /" >> gpurun_out/frp_now.jsonl
  done
done
cut -c1-330 gpurun_out/frp_now.jsonl
