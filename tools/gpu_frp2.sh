#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ik.py -x -q -m gpu 2>&1 | tail -30 | cut -c1-300
export PMB_EXPERIMENT=1 PMB_FRP_TILE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:from_root_positions_tile -s 2 -c 1 -f \
     -o gpurun_out/r2_prof_frp_tile_4m_x_65 python bench.py --kernel-only --steps 3 --warmup 3 --op from_root_positions --workload fk_4m_x_65 > gpurun_out/r2_ncu_frp_tile.log 2>&1
tail -2 gpurun_out/r2_ncu_frp_tile.log
