#!/bin/bash
# kernel-only timings of every op of the path on the three skeletons
OPS=${OPS:-"fk to_dq from_dq round_trip fk_quat from_root_positions mirror_all"}
for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
  for op in $OPS; do
    timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --workload $wl --op $op 2>&1 | tail -1
  done
done | tee gpurun_out/ops.jsonl
