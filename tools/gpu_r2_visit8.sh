#!/bin/bash
# round 2, visit 8: quaternion track kernel (to_root_dual_quat, fk_quat) -- parity, then a sweep against the chain kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "track_kernel or every_variant or random_trees" 2>&1 | tail -15 > gpurun_out/r2_pytest_qt.log; cat gpurun_out/r2_pytest_qt.log
for op in to_dq fk_quat; do
  timeout 600 python tools/sweep_fk.py --op $op --steps 20 < tools/knobs_qt.txt > gpurun_out/r2_sweep_qt_$op.jsonl 2> gpurun_out/r2_sweep_qt_$op.err
  cat gpurun_out/r2_sweep_qt_$op.jsonl; tail -3 gpurun_out/r2_sweep_qt_$op.err
done
