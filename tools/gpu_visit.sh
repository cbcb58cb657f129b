#!/bin/bash
# General GPU visit: whole GPU suite, fk_quat timings, other-op timings.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_fk.py --steps 30 --op fk_quat > gpurun_out/sweep_fkq.jsonl 2> gpurun_out/sweep_fkq.err <<'KNOBS'
-
PMB_FKQ_MATRIX=1
PMB_FKQ_GROUP=8
PMB_FKQ_GROUP=16
PMB_FKQ_GROUP=24
PMB_FKQ_GROUP=32
PMB_FKQ_GROUP=512
PMB_FKQ_BLOCKS_PER_SM=1
KNOBS
echo "sweep rc=$?"; cut -c1-300 gpurun_out/sweep_fkq.jsonl; tail -3 gpurun_out/sweep_fkq.err
for wl in fk_1m_x_22 fk_4m_x_65; do for op in from_root_positions mirror_all; do
  timeout 300 python bench.py --kernel-only --steps 20 --warmup 3 --workload $wl --op $op 2>&1 | tail -1
done; done | tee gpurun_out/ops_ik.jsonl
