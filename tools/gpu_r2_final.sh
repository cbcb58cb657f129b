#!/bin/bash
# round 2, final verification exactly as the driver runs things: build check, smoke(), pytest -m gpu, both bench arms with
# default flags (wall time of each), compute-sanitizer
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
/usr/bin/time -f "reference arm wall %e s" timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_default.json 2> gpurun_out/r2_bench_ref_default.err; tail -1 gpurun_out/r2_bench_ref_default.err; cut -c1-300 gpurun_out/r2_bench_ref_default.json
/usr/bin/time -f "our arm wall %e s" timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -1 gpurun_out/r2_bench_default.err; cut -c1-400 gpurun_out/r2_bench_default.json
bash tools/gpu_sanitize.sh > gpurun_out/r2_sanitize_summary.txt 2>&1; cat gpurun_out/r2_sanitize_summary.txt
