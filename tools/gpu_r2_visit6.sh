#!/bin/bash
# round 2, visit 6: the new default policy (track kernel for large skeletons): whole GPU suite, bench both arms,
# from_root_positions error statistics, per-op kernel timings, ncu full captures of the shipping fk kernels + launch list
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/r2_bench.err
timeout 600 python tests/dev/frp_error_stats.py > gpurun_out/r2_frp_error_stats.jsonl 2> gpurun_out/r2_frp_error_stats.err
cat gpurun_out/r2_frp_error_stats.jsonl
OPS="fk to_dq from_dq fk_quat from_root_positions mirror_all" bash tools/gpu_ops.sh > /dev/null 2>&1; cp gpurun_out/ops.jsonl gpurun_out/r2_ops_kernel_only.jsonl
cut -c1-200 gpurun_out/r2_ops_kernel_only.jsonl
for wl in fk_4m_x_52 fk_4m_x_65; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_ -s 3 -c 1 -f \
     -o gpurun_out/r2_prof_${wl} python bench.py --kernel-only --steps 3 --warmup 3 --workload $wl > gpurun_out/r2_ncu_${wl}.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
ls -la gpurun_out | grep r2_ | head -40
