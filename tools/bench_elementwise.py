#!/usr/bin/env python
"""Kernel-only timings of the element-wise primitives and the non-chain skeleton ops through the C ABI:
algorithmic GB/s (bytes of the operands, each moved once) per kernel.  One JSON line per op.

    python tools/bench_elementwise.py [--n 22000000] [--steps 20]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from pymotion_b200 import _lib  # noqa: E402
from pymotion_b200.topologies import parents_of  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=22_000_000)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    n = args.n
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(7)
    q0 = torch.nn.functional.normalize(torch.randn((n, 4), device=dev, generator=g), dim=-1)
    q1 = torch.nn.functional.normalize(torch.randn((n, 4), device=dev, generator=g), dim=-1)
    v = torch.randn((n, 3), device=dev, generator=g)
    v2 = torch.randn((n, 3), device=dev, generator=g)
    m = torch.empty((n, 9), device=dev)
    dq = torch.empty((n, 8), device=dev)
    o4 = torch.empty((n, 4), device=dev)
    o3 = torch.empty((n, 3), device=dev)
    o1 = torch.empty((n,), device=dev)
    tt = torch.rand((n, 1), device=dev, generator=g)
    codes = torch.full((1,), 5, device=dev, dtype=torch.uint8)
    flags = torch.zeros(3, device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: t.data_ptr()  # noqa: E731
    lib.pmb_quat_to_matrix_f32(p(q0), p(m), n, st)
    lib.pmb_dq_from_rotation_translation_f32(p(q0), p(v), p(dq), n, st)
    J = 22
    F = n // J
    par = parents_of("body22")
    work = torch.empty(int(lib.pmb_unroll_workspace_bytes(F, J)), device=dev, dtype=torch.uint8)

    ops = {  # name: (callable, bytes per element)
        "quat.mul": (lambda: lib.pmb_quat_mul_f32(p(q0), p(q1), p(o4), n, st), 48),
        "quat.mul_vec": (lambda: lib.pmb_quat_mul_vec_f32(p(q0), p(v), p(o3), n, st), 40),
        "quat.length": (lambda: lib.pmb_quat_length_f32(p(q0), p(o1), n, st), 20),
        "quat.normalize": (lambda: lib.pmb_quat_normalize_f32(p(q0), 1e-8, p(o4), n, st), 32),
        "quat.conjugate": (lambda: lib.pmb_quat_conjugate_f32(p(q0), p(o4), n, st), 32),
        "quat.to_matrix": (lambda: lib.pmb_quat_to_matrix_f32(p(q0), p(m), n, st), 52),
        "quat.from_matrix": (lambda: lib.pmb_quat_from_matrix_f32(p(m), p(o4), n, st), 52),
        "dual_quat.from_rotation_translation": (lambda: lib.pmb_dq_from_rotation_translation_f32(p(q0), p(v), p(dq), n, st), 60),
        "dual_quat.from_translation": (lambda: lib.pmb_dq_from_translation_f32(p(v), p(dq), n, st), 44),
        "dual_quat.to_rotation_translation": (lambda: lib.pmb_dq_to_rotation_translation_f32(p(dq), p(o4), p(o3), n, st), 60),
        "quat.from_angle_axis": (lambda: lib.pmb_quat_from_angle_axis_f32(p(o1), p(v), p(o4), n, st), 32),
        "quat.from_euler": (lambda: lib.pmb_quat_from_euler_f32(p(v), p(codes), 0, p(o4), n, st), 28),
        "quat.to_euler": (lambda: lib.pmb_quat_to_euler_f32(p(q0), p(codes), 0, p(o3), n, st), 28),
        "quat.to_angle_axis": (lambda: lib.pmb_quat_to_angle_axis_f32(p(q0), p(o1), p(o3), n, st), 32),
        "quat.slerp": (lambda: lib.pmb_quat_slerp_f32(p(q0), p(q1), p(tt), 1, 1, p(o4), n, st), 52),
        "quat.from_to": (lambda: lib.pmb_quat_from_to_f32(p(v), p(v2), 1, p(o4), n, st), 40),
        "quat.unroll [F x 22]": (lambda: lib.pmb_unroll_f32(p(q0), 4, F, J, p(o4), p(work), work.numel(), st), 32),
        "dual_quat.normalize": (lambda: lib.pmb_dq_normalize_f32(p(dq), p(m), n, p(flags), st) if False else lib.pmb_dq_normalize_f32(p(dq), p(dq), n, p(flags), st), 64),
        "ortho6d.from_quat": (lambda: lib.pmb_ortho6d_from_quat_f32(p(q0), p(dq), n, st), 40),
        "ortho6d.to_matrix": (lambda: lib.pmb_ortho6d_to_matrix_f32(p(dq), p(m), n, st), 60),
        "ortho6d.to_quat": (lambda: lib.pmb_ortho6d_to_quat_f32(p(dq), p(o4), n, st), 40),
        "vector.normalize [n x 3]": (lambda: lib.pmb_vec_normalize_f32(p(v), 1e-8, p(o3), n, 3, st), 24),
        "center_of_mass [F x 22]": (lambda: lib.pmb_center_of_mass_f32(p(v), p(tt), 0, F, J, p(o3), st), 12 + 12.0 / J),
        "from_global_rotations [F x 22]": (lambda: lib.pmb_from_global_rotations_f32(p(q0), par.ctypes.data, F, J, p(o4), st), 32),
        "mirror_to_local [F x 22]": (lambda: lib.pmb_mirror_to_local_f32(p(q0), par.ctypes.data, None, 0, F, J, p(o4), st), 32),
    }
    for name, (fn, bpe) in ops.items():
        rc = fn()
        if rc != 0:
            print(json.dumps({"op": name, "error": lib.pmb_last_error().decode()}), flush=True)
            continue
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"op": name, "n": n, "ms": round(ms, 4), "GBps": round(bpe * n / ms / 1e6, 1), "bytes_per_element": bpe}), flush=True)


if __name__ == "__main__":
    main()
