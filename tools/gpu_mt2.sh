#!/bin/bash
# fk at 52 / 65 joints: matrix track kernel against the row track kernel in BOTH harnesses on one box
set -u
mkdir -p gpurun_out
export PMB_EXPERIMENT=1
for rep in 1 2 3; do
for mt in 0 1; do
  for wl in fk_4m_x_52 fk_4m_x_65; do
    PMB_FK_MTRACKS=$mt timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --op fk --workload $wl | cut -c1-200
  done
done
done
unset PMB_EXPERIMENT
sed -i 's/--workloads [a-z0-9_,]*/--workloads fk_4m_x_52/' tools/gpu_mt.sh
bash tools/gpu_mt.sh 2>&1 | tail -4
