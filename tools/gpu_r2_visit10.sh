#!/bin/bash
# round 2, visit 10: ncu --set full of from_root_positions (the shipping thread-per-frame kernel) at 4M x 65 and 1M x 22
set -u
mkdir -p gpurun_out
for wl in fk_4m_x_65 fk_1m_x_22; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:from_root_positions -s 2 -c 1 -f \
     -o gpurun_out/r2_prof_frp_${wl} python bench.py --kernel-only --steps 3 --warmup 3 --op from_root_positions --workload $wl > gpurun_out/r2_ncu_frp_${wl}.log 2>&1
  tail -2 gpurun_out/r2_ncu_frp_${wl}.log
done
