#!/bin/bash
# kernel-only sweep over env knobs: lines of "VAR=val ... -> json"
set -u
mkdir -p gpurun_out
WL=${WL:-fk_1m_x_22}; OP=${OP:-fk}
while read -r envs; do
  [ -z "$envs" ] && continue
  echo -n "$envs -> "
  env $envs timeout 300 python bench.py --kernel-only --steps 50 --warmup 5 --workload $WL --op $OP 2>&1 | tail -1
done | tee -a gpurun_out/sweep.txt
