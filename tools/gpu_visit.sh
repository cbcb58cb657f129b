#!/bin/bash
# GPU visit: whole GPU suite, auto-policy check on every calibration size, element-wise timings.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_fk.py --steps 30 --workloads fk_2m_x_16,fk_1m_x_22,fk_2m_x_24,fk_2m_x_32,fk_2m_x_40,fk_4m_x_52,fk_4m_x_65 > gpurun_out/sweep_policy.jsonl 2> gpurun_out/sweep_policy.err <<'KNOBS'
-
PMB_FK_LANES=1
PMB_FK_LANES=0 PMB_FK_ROWS=0
PMB_FK_LANES=0 PMB_FK_ROWS=0 PMB_FK_GROUP=0
PMB_FK_LANES=0 PMB_FK_ROWS=1
KNOBS
echo "sweep rc=$?"; cut -c1-300 gpurun_out/sweep_policy.jsonl; tail -3 gpurun_out/sweep_policy.err
