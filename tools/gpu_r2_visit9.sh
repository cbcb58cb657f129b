#!/bin/bash
# round 2, visit 9: quaternion track kernel as the default of to_root_dual_quat / fk_quat -- full GPU suite, sanitizer,
# ncu --set full of both modes at 4M x 65, per-op kernel-only lines
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
bash tools/gpu_sanitize.sh > gpurun_out/r2_sanitize_summary.txt 2>&1; cat gpurun_out/r2_sanitize_summary.txt
for op in to_dq fk_quat; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:qtracks -s 3 -c 1 -f \
     -o gpurun_out/r2_prof_qt_${op}_4m_x_65 python bench.py --kernel-only --steps 3 --warmup 3 --op $op --workload fk_4m_x_65 > gpurun_out/r2_ncu_qt_${op}.log 2>&1
  tail -2 gpurun_out/r2_ncu_qt_${op}.log
done
rm -f gpurun_out/r2_ops_kernel_only.jsonl
for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
  for op in fk to_dq from_dq fk_quat from_root_positions mirror_all; do
    timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --op $op --workload $wl >> gpurun_out/r2_ops_kernel_only.jsonl 2>> gpurun_out/r2_ops_kernel_only.err
  done
done
cat gpurun_out/r2_ops_kernel_only.jsonl
