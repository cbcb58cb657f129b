#!/bin/bash
# quaternion track kernel: lane shapes 4 x 8 against 3 x 10 -- parity, then an alternating A/B on the three bench shapes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ik.py -x -q -m gpu -k "quat_every_variant or dual_quat_track or mirror_fused or quaternion_track" 2>&1 | tail -6 | cut -c1-250
for op in to_dq fk_quat; do
  timeout 600 python tools/sweep_fk.py --op $op --steps 30 < tools/knobs_qt.txt > gpurun_out/r2_sweep_qt_shape_$op.jsonl 2> gpurun_out/r2_sweep_qt_shape_$op.err
  python - "$op" <<'PY'
import json, sys, statistics as st
op=sys.argv[1]
rows=[json.loads(l) for l in open(f'gpurun_out/r2_sweep_qt_shape_{op}.jsonl') if l.startswith('{')]
for wl in dict.fromkeys(r['workload'] for r in rows):
    for kn in ("PMB_QT_SHAPE=1", "PMB_QT_SHAPE=2"):
        ms=[r['ms'] for r in rows if r['workload']==wl and r['knobs']==[kn]]
        v=[r['variant'] for r in rows if r['workload']==wl and r['knobs']==[kn]][0]
        print(op, wl, kn, 'min %.4f median %.4f max %.4f' % (min(ms), st.median(ms), max(ms)), v[:80])
PY
  tail -2 gpurun_out/r2_sweep_qt_shape_$op.err
done
