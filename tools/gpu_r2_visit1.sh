#!/bin/bash
# round 2, visit 1: FFMA2 probe, the whole GPU suite after the api split, sweep of the new track kernel
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/r2_box.txt 2>&1
echo "host cores: $(nproc)" >> gpurun_out/r2_box.txt
timeout 60 ./experiments/ffma2_probe > gpurun_out/r2_ffma2_probe.jsonl 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "track or knobs" > gpurun_out/r2_pytest_tracks.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_tracks.log
tail -5 gpurun_out/r2_pytest_tracks.log
timeout 600 python tools/sweep_fk.py --steps 30 < tools/knobs_tracks.txt > gpurun_out/r2_sweep_tracks.jsonl 2> gpurun_out/r2_sweep_tracks.err
cut -c1-260 gpurun_out/r2_sweep_tracks.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
