#!/bin/bash
# from_root_dual_quat: elements per block tile (= shared memory per block = blocks per SM) sweep.
set -u
mkdir -p gpurun_out
for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
  for e in 4096 3072 2304 1536; do
    PMB_FRDQ_ELEMS=$e timeout 120 python bench.py --kernel-only --op from_dq --steps 20 --warmup 3 --workload $wl 2>/dev/null | tail -1 | sed "s/^{/{\"elems\": $e, /"
  done
done | tee gpurun_out/frdq_sweep.jsonl
