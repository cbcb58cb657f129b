#!/bin/bash
# GPU visit: lane-kernel parity, knob sweep over all fk kernels, element-wise timings.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lane_kernel or unroll or to_matrix or primitives or mirror" > gpurun_out/pytest_lanes.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_lanes.log
tail -15 gpurun_out/pytest_lanes.log
timeout 900 python tools/sweep_fk.py --steps 30 --workloads fk_1m_x_22,fk_2m_x_32,fk_2m_x_40,fk_4m_x_52,fk_4m_x_65 > gpurun_out/sweep_lanes.jsonl 2> gpurun_out/sweep_lanes.err <<'KNOBS'
-
PMB_FK_LANES=1
PMB_FK_LANES=1 PMB_FK_FR=8
PMB_FK_LANES=1 PMB_FK_WARPS=2
PMB_FK_LANES=1 PMB_FK_WARPS=1
PMB_FK_LANES=1 PMB_FK_FR=8 PMB_FK_WARPS=2
PMB_FK_LANES=1 PMB_FK_BLOCKS_PER_SM=1
PMB_FK_LANES=1 PMB_FK_BLOCKS_PER_SM=2
PMB_FK_LANES=1 PMB_FK_BLOCKS_PER_SM=3
KNOBS
echo "sweep rc=$?"; cut -c1-300 gpurun_out/sweep_lanes.jsonl; tail -3 gpurun_out/sweep_lanes.err
python tools/bench_elementwise.py > gpurun_out/elementwise.jsonl 2> gpurun_out/elementwise.err; grep -E "to_matrix|unroll|mirror|normalize\"" gpurun_out/elementwise.jsonl
