#!/bin/bash
# ncu --set full of the kernels that ship for the headline shape (1M x 22): fk (row teams), fk_quat (3 x 10 lanes), mirror (fused)
set -u
mkdir -p gpurun_out
for spec in "fk fk_rows r2_prof_fk_1m_x_22" "fk_quat qtracks r2_prof_fkq_1m_x_22" "mirror_all qtracks r2_prof_mirror_1m_x_22"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$3 \
     python bench.py --kernel-only --steps 3 --warmup 3 --op $1 > gpurun_out/$3.log 2>&1
  tail -1 gpurun_out/$3.log
done
