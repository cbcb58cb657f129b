#!/bin/bash
# round 2, visit 11: refresh of the judged artefacts after the quaternion track kernel / from_root_positions changes:
# whole GPU suite, both bench arms, ncu --set full of to_root_dual_quat 1M x 22 (traffic.json) and of from_root_positions
# 4M x 65, the launch list of one bench run, element-wise and per-op kernel timings
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qtracks -s 3 -c 1 -f \
   -o gpurun_out/r2_prof_to_dq_1m_x_22 python bench.py --kernel-only --steps 3 --warmup 3 --op to_dq > gpurun_out/r2_ncu_to_dq.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:from_root_positions -s 2 -c 1 -f \
   -o gpurun_out/r2_prof_frp_paired_4m_x_65 python bench.py --kernel-only --steps 3 --warmup 3 --op from_root_positions --workload fk_4m_x_65 > gpurun_out/r2_ncu_frp_paired.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
grep -o 'pmb::[a-z_]*' gpurun_out/r2_launches.csv | sort | uniq -c
rm -f gpurun_out/r2_ops_kernel_only.jsonl
for wl in fk_1m_x_22 fk_4m_x_52 fk_4m_x_65; do
  for op in fk to_dq from_dq round_trip fk_quat from_root_positions mirror_all; do
    timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --op $op --workload $wl >> gpurun_out/r2_ops_kernel_only.jsonl 2>> gpurun_out/r2_ops_kernel_only.err
  done
done
cut -c1-300 gpurun_out/r2_ops_kernel_only.jsonl
timeout 600 python tools/bench_elementwise.py > gpurun_out/r2_elementwise.jsonl 2> gpurun_out/r2_elementwise.err; cut -c1-200 gpurun_out/r2_elementwise.jsonl
