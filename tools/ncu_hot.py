#!/usr/bin/env python
"""Hot spots of an ncu report (source page): stall-reason totals and the top instructions by samples.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; out = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name": break
    if len(r) == len(hdr): out.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in out)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in out) for h in stalls}
print("instructions", len(out), "samples", tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]: print(f"  {k:24s}{v:8d} {v/tot:.3f}")
for n in sorted(range(len(out)), key=lambda n: -int(out[n][ix["# Samples"]]))[:topn]:
    r = out[n]
    st = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
    t3 = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{n:5d} {r[1].strip()[:58]:58s} ex={int(r[ix['Instructions Executed']]):9d} smp={int(r[ix['# Samples']]):6d} wf={r[ix['L1 Wavefronts Shared']]:>9s}/{r[ix['L1 Wavefronts Shared Ideal']]:>9s} {t3}")
ops = collections.Counter()
for r in out:
    p = r[1].strip().split(); op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]])
te = sum(ops.values()); print("executed", te)
print("  " + "  ".join(f"{op}:{c/te:.3f}" for op, c in ops.most_common(14)))
