#!/bin/bash
set -u
mkdir -p gpurun_out
export PMB_EXPERIMENT=1 PMB_FK_MTRACKS=1
for wl in fk_4m_x_65 fk_1m_x_22; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_mtracks -s 3 -c 1 -f \
   -o gpurun_out/r2_prof_mt_${wl} python bench.py --kernel-only --steps 3 --warmup 3 --op fk --workload $wl > gpurun_out/r2_ncu_mt_${wl}.log 2>&1
tail -1 gpurun_out/r2_ncu_mt_${wl}.log
done
