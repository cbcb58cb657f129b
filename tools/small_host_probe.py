#!/usr/bin/env python
"""Drop-in call on small HOST batches (BASELINE config 1 and around): single-shot path against the chunked host pipeline."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from pymotion_b200.ops import skeleton as sk
from pymotion_b200.topologies import parents_of, synth_numpy

par = parents_of("body22")
def best(fn, reps=30):
    for _ in range(5): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, sorted(ts)[len(ts)//2] * 1e3
for n in (100, 1000, 4000, 16000, 64000):
    rot, gp, off = synth_numpy(n, par, seed=0)
    single = lambda: sk.fk(rot, gp, off, par)
    piped = lambda: sk.fk_host(rot, gp, off, par)
    old = sk._HOST_PATH_MIN_BYTES
    sk._HOST_PATH_MIN_BYTES = 1 << 60
    a = best(single)
    sk._HOST_PATH_MIN_BYTES = old
    b = best(piped)
    p1, r1 = single(); p2, r2 = piped()
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    c = best(lambda: sk.to_root_dual_quat(rot, gp, par, off))
    d = best(lambda: sk.from_root_dual_quat(dq, par))
    print(json.dumps({"frames": n, "to_root_dual_quat_ms_best_median": [round(x, 4) for x in c],
                      "from_root_dual_quat_ms_best_median": [round(x, 4) for x in d]}), flush=True)
    print(json.dumps({"frames": n, "bytes": n * (64 * 22 + 12), "single_shot_ms_best_median": [round(x, 4) for x in a],
                      "host_pipeline_ms_best_median": [round(x, 4) for x in b], "equal": bool(np.array_equal(p1, p2) and np.array_equal(r1, r2))}), flush=True)
