#!/bin/bash
# round 2, visit 2: box-fed track kernel: parity, sweep, ncu of the best candidates
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "track or knobs" > gpurun_out/r2_pytest_tracks.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_tracks.log
tail -5 gpurun_out/r2_pytest_tracks.log
timeout 600 python tools/sweep_fk.py --steps 30 --workloads fk_1m_x_22,fk_4m_x_52,fk_4m_x_65,fk_2m_x_40 < tools/knobs_tracks.txt > gpurun_out/r2_sweep_tracks.jsonl 2> gpurun_out/r2_sweep_tracks.err
cut -c1-250 gpurun_out/r2_sweep_tracks.jsonl
for wl in fk_4m_x_52 fk_4m_x_65; do
  PMB_EXPERIMENT=1 PMB_FK_TRACKS=1 PMB_FK_UL=2 PMB_FK_U=1 PMB_FK_NB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk_tracks -s 3 -c 1 -f \
     -o gpurun_out/r2_prof_tracks_${wl} python bench.py --kernel-only --steps 3 --warmup 3 --workload $wl > gpurun_out/r2_ncu_tracks_${wl}.log 2>&1
done
ls -la gpurun_out | grep r2_
