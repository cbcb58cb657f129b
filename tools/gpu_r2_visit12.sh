#!/bin/bash
# round 2, visit 12: matrix track kernel as the default of fk above 30 joints -- whole GPU suite, sanitizer, bench, ncu of the
# kernel at 4M x 65 / 4M x 52 (traffic.json), per-op timings from a cold start
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
bash tools/gpu_sanitize.sh > gpurun_out/r2_sanitize_summary.txt 2>&1; cat gpurun_out/r2_sanitize_summary.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_default.err
for wl in fk_4m_x_65 fk_4m_x_52; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_mtracks -s 3 -c 1 -f \
     -o gpurun_out/r2_prof_mt_${wl} python bench.py --kernel-only --steps 3 --warmup 3 --op fk --workload $wl > gpurun_out/r2_ncu_mt_${wl}.log 2>&1
done
rm -f gpurun_out/r2_ops_fk.jsonl
for wl in fk_1m_x_22 fk_2m_x_24 fk_2m_x_40 fk_4m_x_52 fk_4m_x_65; do
  timeout 300 python bench.py --kernel-only --steps 30 --warmup 5 --op fk --workload $wl >> gpurun_out/r2_ops_fk.jsonl
done
cut -c1-300 gpurun_out/r2_ops_fk.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
grep -o 'pmb::[a-z_]*' gpurun_out/r2_launches.csv | sort | uniq -c
