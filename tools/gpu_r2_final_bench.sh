#!/bin/bash
# both bench arms with the driver's default flags, wall time of each
set -u
mkdir -p gpurun_out
S=$SECONDS; timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_default.json 2> gpurun_out/r2_bench_ref_default.err; echo "reference arm rc=$? wall $((SECONDS-S)) s"; cut -c1-300 gpurun_out/r2_bench_ref_default.json
S=$SECONDS; timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "our arm rc=$? wall $((SECONDS-S)) s"; cut -c1-400 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
