#!/bin/bash
# Experiment visit: build-time variants of the row walk (experiments/variants/lib_*.so) through the knob sweep,
# an ncu capture of a lone team per SM, then the GPU tests of the extended rotation surface.
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep_variants.jsonl
for lib in experiments/variants/lib_*.so; do
  echo "{\"lib\": \"$lib\"}" >> gpurun_out/sweep_variants.jsonl
  PMB_LIB_PATH=$PWD/$lib timeout 600 python tools/sweep_fk.py --steps 30 >> gpurun_out/sweep_variants.jsonl 2>> gpurun_out/sweep_variants.err <<'KNOBS'
PMB_FK_ROWS=1 PMB_FK_STAGES=2 PMB_FK_BLOCKS_PER_SM=4
PMB_FK_ROWS=1 PMB_FK_STAGES=3
PMB_FK_ROWS=1 PMB_FK_STAGES=4
PMB_FK_ROWS=1 PMB_FK_STAGES=3 PMB_FK_BLOCKS_PER_SM=1
PMB_FK_ROWS=1 PMB_FK_STAGES=3 PMB_FK_BLOCKS_PER_SM=2
KNOBS
done
tail -3 gpurun_out/sweep_variants.err
PMB_FK_ROWS=1 PMB_FK_STAGES=3 PMB_FK_BLOCKS_PER_SM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_rows -s 3 -c 1 -f \
    -o gpurun_out/prof_rows1_fk_4m_x_65 python bench.py --kernel-only --steps 3 --warmup 3 --workload fk_4m_x_65 > gpurun_out/ncu_rows1.log 2>&1
echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_gpu_quat_ext.py -m gpu -q > gpurun_out/pytest_ext.log 2>&1; echo "pytest ext rc=$?" | tee -a gpurun_out/pytest_ext.log
tail -30 gpurun_out/pytest_ext.log
