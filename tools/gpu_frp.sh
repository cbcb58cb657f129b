#!/bin/bash
# from_root_positions: blocks per SM (= shared-memory carve-out, what is left is L1) sweep.
set -u
mkdir -p gpurun_out
run() { # workload, knob value or "-"
  if [ "$2" = "-" ]; then unset PMB_FRP_BLOCKS_PER_SM; else export PMB_FRP_BLOCKS_PER_SM=$2; fi
  timeout 120 python bench.py --kernel-only --op from_root_positions --steps 10 --warmup 3 --workload $1 2>/dev/null | tail -1 | sed "s/^{/{\"blocks_per_sm\": \"$2\", /"
}
{
  for b in 2 3 4 6 9; do run fk_4m_x_65 $b; done
  for b in - 3 4 6; do run fk_4m_x_52 $b; done
  for b in - 4 6; do run fk_1m_x_22 $b; done
} | tee gpurun_out/frp_sweep.jsonl
