#!/bin/bash
# round 2, visit 13 (final state): whole GPU suite, both bench arms with default flags, launch list, per-op kernel timings,
# alternating A/B of the two fk track kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_default.json 2> gpurun_out/r2_bench_ref_default.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
grep -o 'pmb::[a-z_]*' gpurun_out/r2_launches.csv | sort | uniq -c
rm -f gpurun_out/r2_ops_kernel_only.jsonl
for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
  for op in fk to_dq from_dq round_trip fk_quat from_root_positions mirror_all; do
    timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --op $op --workload $wl >> gpurun_out/r2_ops_kernel_only.jsonl 2>> gpurun_out/r2_ops_kernel_only.err
  done
done
cut -c1-260 gpurun_out/r2_ops_kernel_only.jsonl
timeout 600 python tools/bench_elementwise.py > gpurun_out/r2_elementwise.jsonl 2> gpurun_out/r2_elementwise.err; grep -E "normalize|unroll" gpurun_out/r2_elementwise.jsonl
bash tools/gpu_mt.sh 2>&1 | tail -9
