#!/bin/bash
# General GPU visit: whole GPU suite, launch-policy calibration sweep, bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_fk.py --steps 30 --workloads fk_1m_x_22,fk_2m_x_32,fk_2m_x_40,fk_4m_x_52,fk_4m_x_65 > gpurun_out/sweep_policy.jsonl 2> gpurun_out/sweep_policy.err <<'KNOBS'
-
PMB_FK_ROWS=0
PMB_FK_ROWS=1 PMB_FK_STAGES=2
PMB_FK_ROWS=1 PMB_FK_STAGES=3
PMB_FK_ROWS=1 PMB_FK_STAGES=4
KNOBS
echo "sweep rc=$?"; cut -c1-330 gpurun_out/sweep_policy.jsonl; tail -3 gpurun_out/sweep_policy.err
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-1500
