#!/bin/bash
# fk matrix track kernel: parity, then an alternating A/B against the shipping policy (five rounds, 30 steps each)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matrix_track" 2>&1 | tail -3 | cut -c1-250
timeout 900 python tools/sweep_fk.py --steps 30 --workloads fk_2m_x_24,fk_2m_x_40,fk_4m_x_52,fk_4m_x_65 < tools/knobs_mt.txt > gpurun_out/r2_sweep_mtracks.jsonl 2> gpurun_out/r2_sweep_mtracks.err
python - <<'PY'
import json, statistics as st
rows=[json.loads(l) for l in open('gpurun_out/r2_sweep_mtracks.jsonl') if l.startswith('{')]
for wl in dict.fromkeys(r['workload'] for r in rows):
    for kn in ("PMB_FK_MTRACKS=0", "PMB_FK_MTRACKS=1"):
        ms=[r['ms'] for r in rows if r['workload']==wl and r['knobs']==[kn]]
        v=[r['variant'] for r in rows if r['workload']==wl and r['knobs']==[kn]][0]
        print(wl, kn, 'min %.4f median %.4f max %.4f' % (min(ms), st.median(ms), max(ms)), v[:70])
PY
tail -3 gpurun_out/r2_sweep_mtracks.err
