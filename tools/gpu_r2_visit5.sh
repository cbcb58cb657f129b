#!/bin/bash
# round 2, visit 5: track kernel with merged table / L2 prefetch / deeper ring; the new bench line; the whole GPU suite
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "track or knobs or bvh" > gpurun_out/r2_pytest_tracks.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_tracks.log
tail -5 gpurun_out/r2_pytest_tracks.log
timeout 600 python tools/sweep_fk.py --steps 30 --workloads fk_4m_x_52,fk_4m_x_65,fk_2m_x_40,fk_1m_x_22 < tools/knobs_tracks.txt > gpurun_out/r2_sweep_tracks.jsonl 2> gpurun_out/r2_sweep_tracks.err
cut -c1-250 gpurun_out/r2_sweep_tracks.jsonl
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
