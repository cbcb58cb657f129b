#!/bin/bash
# round 2, visit 7: compute-sanitizer over every kernel family incl. the track kernel and the host pipeline, per-call host
# latency, from_root_positions error statistics at the test sizes, fk_quat sign statistics, ncu of the dq kernels, launch list
set -u
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh > gpurun_out/r2_sanitize_summary.txt 2>&1; cat gpurun_out/r2_sanitize_summary.txt
timeout 300 python tests/dev/latency_small.py > gpurun_out/r2_latency_small.jsonl 2> gpurun_out/r2_latency_small.err; cat gpurun_out/r2_latency_small.jsonl
timeout 600 python tests/dev/frp_error_stats.py 3001 1000 517 20000 6000 4000 > gpurun_out/r2_frp_error_stats.jsonl 2> gpurun_out/r2_frp_error_stats.err; cat gpurun_out/r2_frp_error_stats.jsonl
timeout 600 python tests/dev/fkq_sign_stats.py > gpurun_out/r2_fkq_sign_stats.jsonl 2> gpurun_out/r2_fkq_sign_stats.err; cat gpurun_out/r2_fkq_sign_stats.jsonl; tail -3 gpurun_out/r2_fkq_sign_stats.err
for op in to_dq from_dq; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:root_dq -s 3 -c 1 -f \
     -o gpurun_out/r2_prof_${op}_1m_x_22 python bench.py --kernel-only --steps 3 --warmup 3 --op $op > gpurun_out/r2_ncu_${op}.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
grep -o 'pmb::[a-z_]*' gpurun_out/r2_launches.csv | sort | uniq -c
