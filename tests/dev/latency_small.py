#!/usr/bin/env python
"""Small-batch latency of fk (BASELINE config 1: 1000 x 22): per-call cost through the Python drop-in and through the
C ABI, synchronous (call + wait) and pipelined (back-to-back launches), beside the NumPy oracle port on the host."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pymotion_oracle as orc  # noqa: E402
from pymotion_b200 import _lib  # noqa: E402
from pymotion_b200.ops import skeleton as sk  # noqa: E402
from pymotion_b200.topologies import parents_of, synth_torch  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    par = parents_of("body22")
    for frames in (1, 1000, 100_000):
        rot, gp, off = synth_torch(frames, par, dev, seed=1)
        pos = torch.empty((frames, 22, 3), device=dev)
        rotm = torch.empty((frames, 22, 3, 3), device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream

        def c_call():
            return lib.pmb_fk_f32(rot.data_ptr(), gp.data_ptr(), 3, off.data_ptr(), 0, par.ctypes.data, frames, 22,
                                  pos.data_ptr(), rotm.data_ptr(), st)

        def timeit(fn, n, sync_each):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
                if sync_each:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n * 1e6

        n = 2000
        out = {"frames": frames, "joints": 22,
               "c_abi_us_sync": round(timeit(c_call, n, True), 2), "c_abi_us_pipelined": round(timeit(c_call, n, False), 2),
               "python_us_sync": round(timeit(lambda: sk.fk(rot, gp, off, par), n, True), 2),
               "python_us_pipelined": round(timeit(lambda: sk.fk(rot, gp, off, par), n, False), 2)}
        if frames == 1000:  # the other two ops of the path and a large skeleton (track kernel: schedule cache, TMA map cache)
            dq = torch.empty((frames, 22, 8), device=dev)
            rots = torch.empty((frames, 22, 4), device=dev)
            off0 = np.zeros(3, dtype=np.float32)
            out["to_dq_c_abi_us_pipelined"] = round(timeit(lambda: lib.pmb_to_root_dual_quat_f32(
                rot.data_ptr(), gp.data_ptr(), 3, par.ctypes.data, off.data_ptr(), off0.ctypes.data, frames, 22, dq.data_ptr(), st), n, False), 2)
            out["from_dq_c_abi_us_pipelined"] = round(timeit(lambda: lib.pmb_from_root_dual_quat_f32(
                dq.data_ptr(), par.ctypes.data, frames, 22, pos.data_ptr(), rots.data_ptr(), st), n, False), 2)
            par65 = parents_of("deep65")
            rot65, gp65, off65 = synth_torch(frames, par65, dev, seed=2)
            pos65 = torch.empty((frames, 65, 3), device=dev)
            rotm65 = torch.empty((frames, 65, 3, 3), device=dev)
            out["fk65_c_abi_us_pipelined"] = round(timeit(lambda: lib.pmb_fk_f32(
                rot65.data_ptr(), gp65.data_ptr(), 3, off65.data_ptr(), 0, par65.ctypes.data, frames, 65, pos65.data_ptr(),
                rotm65.data_ptr(), st), n, False), 2)
        if frames <= 1000:
            r, g, o = rot.cpu().numpy(), gp.cpu().numpy(), off.cpu().numpy()
            t0 = time.perf_counter()
            for _ in range(50):
                orc.fk(r, g, o, par)
            out["numpy_port_us"] = round((time.perf_counter() - t0) / 50 * 1e6, 1)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
