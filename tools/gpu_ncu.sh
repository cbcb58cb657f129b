#!/bin/bash
# ncu full capture of the fk kernel for one workload; env PMB_FK_CHUNK etc. pass through.
set -u
mkdir -p gpurun_out
WL=${1:-fk_1m_x_22}; OUT=${2:-prof_fk}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_chain -s 3 -c 1 -f -o gpurun_out/$OUT \
    python bench.py --kernel-only --steps 3 --warmup 3 --workload $WL > gpurun_out/ncu_$OUT.log 2>&1
tail -2 gpurun_out/ncu_$OUT.log
