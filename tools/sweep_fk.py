#!/usr/bin/env python
"""Kernel-only sweep of fk launch variants in ONE process (PMB_EXPERIMENT=1 + pmb_reload_knobs() between
variants).  Each line: workload, knobs, variant picked, ms, algorithmic GB/s, max |diff| vs the first
knob set of the workload.

    python tools/sweep_fk.py [--workloads fk_1m_x_22,fk_4m_x_52,fk_4m_x_65] [--steps 30] < knobs.txt

knobs.txt: one whitespace-separated list of VAR=value per line ("-" = no knobs).
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from bench import WORKLOADS, op_bytes_per_pose  # noqa: E402
from pymotion_b200 import _lib  # noqa: E402
from pymotion_b200.topologies import parents_of, synth_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="fk_1m_x_22,fk_4m_x_52,fk_4m_x_65")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--op", default="fk", choices=["fk", "fk_quat", "to_dq"])
    args = ap.parse_args()
    knob_sets = [ln.split() for ln in sys.stdin.read().splitlines() if ln.strip() and not ln.startswith("#")]
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for wl in args.workloads.split(","):
        topo, frames = WORKLOADS[wl]
        par = parents_of(topo)
        J = len(par)
        rot, gpos, off = synth_torch(frames, par, dev, seed=1234)
        pos = torch.empty((frames, J, 3), device=dev, dtype=torch.float32)
        wide = {"fk": 9, "fk_quat": 4, "to_dq": 8}[args.op]
        out = torch.empty((frames, J, wide), device=dev, dtype=torch.float32)
        st = torch.cuda.current_stream(dev).cuda_stream
        import numpy as np

        off0 = np.zeros(3, dtype=np.float32)

        def step():
            if args.op == "fk":
                return lib.pmb_fk_f32(rot.data_ptr(), gpos.data_ptr(), 3, off.data_ptr(), 0, par.ctypes.data, frames, J,
                                      pos.data_ptr(), out.data_ptr(), st)
            if args.op == "fk_quat":
                return lib.pmb_fk_quat_f32(rot.data_ptr(), gpos.data_ptr(), 3, off.data_ptr(), 0, par.ctypes.data, frames,
                                           J, pos.data_ptr(), out.data_ptr(), st)
            return lib.pmb_to_root_dual_quat_f32(rot.data_ptr(), gpos.data_ptr(), 3, par.ctypes.data, off.data_ptr(),
                                                 off0.ctypes.data, frames, J, out.data_ptr(), st)

        ref_pos = ref_out = None
        for ks in knob_sets:
            for k in [k for k in os.environ if k.startswith("PMB_") and k != "PMB_LIB_PATH"]:
                os.environ.pop(k)
            for kv in ks:
                if kv != "-":
                    k, v = kv.split("=")
                    os.environ[k] = v
            os.environ["PMB_EXPERIMENT"] = "1"
            lib.pmb_reload_knobs()
            pos.zero_(), out.zero_()
            rc = step()
            if rc != 0:
                print(json.dumps({"workload": wl, "knobs": ks, "error": lib.pmb_last_error().decode()}), flush=True)
                continue
            torch.cuda.synchronize()
            sl = slice(0, frames, max(1, frames // 65536))  # strided sample of frames incl. every tile position
            tail = slice(frames - 4096, frames)
            if ref_pos is None:
                ref_pos = (pos[sl].clone(), pos[tail].clone())
                ref_out = (out[sl].clone(), out[tail].clone())
                diff = 0.0
            else:
                diff = max(float((pos[sl] - ref_pos[0]).abs().max()), float((pos[tail] - ref_pos[1]).abs().max()),
                           float((out[sl] - ref_out[0]).abs().max()), float((out[tail] - ref_out[1]).abs().max()))
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            variant = lib.pmb_last_variant().decode() if hasattr(lib, "pmb_last_variant") else ""
            print(json.dumps({"workload": wl, "op": args.op, "knobs": ks, "variant": variant, "ms": round(ms, 5),
                              "GBps": round(op_bytes_per_pose(args.op, J) * frames / ms / 1e6, 1),
                              "max_diff_vs_first": diff}), flush=True)
        del rot, gpos, off, pos, out, ref_pos, ref_out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
