#!/usr/bin/env python
"""Does the time of the fk track kernels at 4M x 52 depend on WHERE the arrays sit?  Allocates a pad of varying size before the
workload's tensors (fresh process state each time via empty_cache) and times both kernels on the same arrays."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from pymotion_b200 import _lib
from pymotion_b200.topologies import parents_of, synth_torch

lib = _lib.load()
dev = torch.device("cuda", 0)
par = parents_of(sys.argv[1] if len(sys.argv) > 1 else "smplh52")
J, F = len(par), int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
st = torch.cuda.current_stream(dev).cuda_stream
for pad_mb in (0, 1, 3, 17, 64, 100, 513, 1024, 2049, 0):
    torch.cuda.empty_cache()
    pad = torch.empty(pad_mb * (1 << 20) + (4096 if pad_mb else 0), dtype=torch.uint8, device=dev) if pad_mb else None
    rot, gpos, off = synth_torch(F, par, dev, seed=1234)
    pos = torch.empty((F, J, 3), device=dev)
    rotm = torch.empty((F, J, 9), device=dev)
    res = {"pad_mb": pad_mb, "rot": hex(rot.data_ptr()), "pos": hex(pos.data_ptr()), "rotm": hex(rotm.data_ptr())}
    for mt in ("0", "1"):
        os.environ["PMB_EXPERIMENT"] = "1"; os.environ["PMB_FK_MTRACKS"] = mt
        lib.pmb_reload_knobs()
        call = lambda: lib.pmb_fk_f32(rot.data_ptr(), gpos.data_ptr(), 3, off.data_ptr(), 0, par.ctypes.data, F, J, pos.data_ptr(), rotm.data_ptr(), st)
        for _ in range(5): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        res["mtracks_ms" if mt == "1" else "tracks_ms"] = round(e0.elapsed_time(e1) / 20, 4)
    print(json.dumps(res), flush=True)
    del rot, gpos, off, pos, rotm, pad
