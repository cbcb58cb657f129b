#!/bin/bash
# GPU visit for the row-team fk kernel: parity of the forced variants, knob sweep on the three bench
# skeletons, ncu captures, then the whole GPU suite.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/box.txt 2>&1
echo "host cores: $(nproc)" >> gpurun_out/box.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "row_team" > gpurun_out/pytest_rows.log 2>&1; echo "pytest rows rc=$?" | tee -a gpurun_out/pytest_rows.log
tail -5 gpurun_out/pytest_rows.log
timeout 900 python tools/sweep_fk.py --steps 30 > gpurun_out/sweep_rows.jsonl 2> gpurun_out/sweep_rows.err <<'KNOBS'
PMB_FK_ROWS=0
PMB_FK_ROWS=1 PMB_FK_STAGES=2
PMB_FK_ROWS=1 PMB_FK_STAGES=3
PMB_FK_ROWS=1 PMB_FK_STAGES=4
PMB_FK_ROWS=1 PMB_FK_STAGES=2 PMB_FK_BLOCKS_PER_SM=4
PMB_FK_ROWS=1 PMB_FK_STAGES=2 PMB_FK_BLOCKS_PER_SM=3
PMB_FK_ROWS=1 PMB_FK_STAGES=2 PMB_FK_BLOCKS_PER_SM=2
PMB_FK_ROWS=1 PMB_FK_STAGES=2 PMB_FK_BLOCKS_PER_SM=1
PMB_FK_ROWS=1 PMB_FK_STAGES=4 PMB_FK_BLOCKS_PER_SM=1
KNOBS
echo "sweep rc=$?"; cat gpurun_out/sweep_rows.jsonl; tail -3 gpurun_out/sweep_rows.err
for wl in fk_1m_x_22 fk_4m_x_65; do
  PMB_FK_ROWS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_rows -s 3 -c 1 -f \
      -o gpurun_out/prof_rows_$wl python bench.py --kernel-only --steps 3 --warmup 3 --workload $wl > gpurun_out/ncu_rows_$wl.log 2>&1
  echo "ncu $wl rc=$?"
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
ls -la gpurun_out | head -30
