#!/usr/bin/env python
"""Benchmark of the fk hot path (BASELINE.json metric: skeleton poses/sec, 22 joints).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (`fk`) over one resident batch.  At N GPUs
every rank runs the same per-GPU batch on its own frames (frame-axis shard, no
data-path collective): weak scaling, value = all frames processed / max-over-ranks
device time.  Prints ONE JSON line on rank 0.

  value      kernel throughput with inputs resident in HBM (CUDA events on the launch stream)
  roofline   algorithmic bytes (64*J + 12 per pose) / kernel time vs MEASURED_PEAKS.json
  e2e        same metric through the public host-buffer API (pymotion_b200.ops.skeleton.fk_host):
             pinned host inputs -> H2D -> kernel -> D2H -> pinned host outputs, all inside the timing
  cpu_baseline  the NumPy oracle port of the reference path timed on this box's host (rank 0, N=1)

`--impl reference` times the reference's CPU algorithm instead (NumPy oracle port --
the reference is pure Python/NumPy and is not shipped to the GPU box -- on all host
threads it can use, frames split across a thread pool).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "skeleton poses/sec (22 joints)"  # BASELINE.json; other workloads carry their own joint count


def metric_of(n_joints: int) -> str:
    return f"skeleton poses/sec ({n_joints} joints)"
UNIT = "poses/s"

WORKLOADS = {
    # name: (topology, frames per GPU)   -- BASELINE.json configs[1], [3], [4]
    "fk_1m_x_22": ("body22", 1_000_000),
    "fk_4m_x_65": ("deep65", 4_000_000),
    "fk_4m_x_52": ("smplh52", 4_000_000),
    # launch-heuristic calibration sizes (not BASELINE configs)
    "fk_2m_x_32": ("body32", 2_000_000),
    "fk_2m_x_40": ("body40", 2_000_000),
    "fk_2m_x_16": ("body16", 2_000_000),
    "fk_2m_x_24": ("body24", 2_000_000),
}
# --kernel-only development workloads for the other ops of the path (BASELINE.json configs[2])
DEV_OPS = ("fk", "to_dq", "from_dq", "round_trip", "fk_quat", "from_root_positions", "mirror_all")


def fk_bytes_per_pose(n_joints: int) -> int:
    return 64 * n_joints + 12  # read 16J + 12, write 12J + 36J


def op_bytes_per_pose(op: str, n_joints: int) -> int:
    return {"fk": 64 * n_joints + 12, "to_dq": 48 * n_joints + 12, "from_dq": 60 * n_joints,
            "round_trip": 108 * n_joints + 12, "fk_quat": 44 * n_joints + 12,
            # 8f rank 2 ops: positions in, rotations out; mirror = rotations in, rotations out (its two kernels
            # move 76 J + 12: the fk_quat positions and the global quaternions in between are not algorithmic)
            "from_root_positions": 28 * n_joints, "mirror_all": 32 * n_joints}[op]


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(workload: str):
    """dram bytes per launch from the committed `ncu --set full` capture, if any."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Polls NVML (same counters as the nvidia-smi clocks line of B200_PROFILING.md)
    every ~2 ms while the timed regions run."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (t, sm_mhz, reasons_bitmask, power_w)
        self.stop_flag = threading.Event()
        self.windows = []
        self.ok = False
        self.sm_max = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                try:
                    power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    power = None
                self.samples.append((time.perf_counter(), mhz, reasons, power))
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        nv = self.nv
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        note = "sampled inside the timed regions"
        if len(inside) < 3:
            inside, note = self.samples, "timed region shorter than the sampling period: all samples of the run"
        mhz = sorted(s[1] for s in inside)
        mask = 0
        for s in inside:
            mask |= s[2]
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
            "applications_clocks_setting": getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2),
        }
        reasons = [k for k, bit in names.items() if mask & bit]
        powers = [s[3] for s in inside if s[3] is not None]
        return {
            "sm_mhz": mhz[len(mhz) // 2],
            "sm_max_mhz": self.sm_max,
            "reasons": reasons,
            "samples": len(inside),
            "power_w_max": max(powers) if powers else None,
            "note": note,
        }


# --------------------------------------------------------------------------- reference arm
def numpy_port_threaded(rot, gpos, offsets, parents, n_threads: int):
    """The reference algorithm (NumPy oracle port) over frame shards on a thread
    pool: NumPy releases the GIL inside its ufunc / matmul loops."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from oracle import pymotion_oracle as orc

    bounds = np.linspace(0, rot.shape[0], n_threads + 1).astype(int)

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        if hi > lo:
            orc.fk(rot[lo:hi], gpos[lo:hi], offsets, parents)

    with ThreadPoolExecutor(n_threads) as pool:
        list(pool.map(work, range(n_threads)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    import numpy as np

    from pymotion_b200.topologies import parents_of, synth_numpy

    topo, _ = WORKLOADS[args.workload]
    par = parents_of(topo)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    sample = 2000 * threads  # frames per step: keeps a step at ~0.1-0.2 s per thread
    rot, gpos, off = synth_numpy(sample, par, seed=0)
    for _ in range(max(1, min(args.warmup, 3))):
        numpy_port_threaded(rot, gpos, off, par, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        numpy_port_threaded(rot, gpos, off, par, threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference",
        "metric": metric_of(len(par)),
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32 in / f64 chain (NumPy promotion of the reference)",
        "data": "synthetic",
        "config": {"workload": args.workload, "sample_frames_per_step": sample, "n_joints": len(par),
                   "note": "CPU arm: does not use GPUs; same number for every --gpus"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} frames x {len(par)} joints per step, NumPy oracle port of "
                                   f"pymotion.ops.skeleton.fk over {threads} threads (host has {cores} cores)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from pymotion_b200 import _lib
    from pymotion_b200.ops import skeleton as sk
    from pymotion_b200.topologies import parents_of, synth_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()  # raises if the CUDA library is missing: no fallback

    topo, frames = WORKLOADS[args.workload]
    par = parents_of(topo)
    n_joints = len(par)
    rot, gpos, off = synth_torch(frames, par, dev, seed=1234 + rank)
    pos = torch.empty((frames, n_joints, 3), device=dev, dtype=torch.float32)
    rotm = torch.empty((frames, n_joints, 3, 3), device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream(dev)

    def step():
        _lib.check(lib.pmb_fk_f32(rot.data_ptr(), gpos.data_ptr(), 3, off.data_ptr(), 0, par.ctypes.data, frames,
                                  n_joints, pos.data_ptr(), rotm.data_ptr(), stream.cuda_stream))

    if args.kernel_only and args.op != "fk":
        if args.op == "from_root_positions":
            step()  # needs `rotm` alive for this one launch
            torch.cuda.synchronize(dev)
        del rotm
        dq = torch.empty((frames, n_joints, 8), device=dev, dtype=torch.float32)
        rots = torch.empty((frames, n_joints, 4), device=dev, dtype=torch.float32)
        off0 = np.zeros(3, dtype=np.float32)
        st = stream.cuda_stream

        def to_dq():
            _lib.check(lib.pmb_to_root_dual_quat_f32(rot.data_ptr(), gpos.data_ptr(), 3, par.ctypes.data, off.data_ptr(),
                                                     off0.ctypes.data, frames, n_joints, dq.data_ptr(), st))

        def from_dq():
            _lib.check(lib.pmb_from_root_dual_quat_f32(dq.data_ptr(), par.ctypes.data, frames, n_joints, pos.data_ptr(),
                                                       rots.data_ptr(), st))

        def fk_quat():
            _lib.check(lib.pmb_fk_quat_f32(rot.data_ptr(), gpos.data_ptr(), 3, off.data_ptr(), 0, par.ctypes.data, frames,
                                           n_joints, pos.data_ptr(), rots.data_ptr(), st))

        def from_root_positions():
            _lib.check(lib.pmb_from_root_positions_f32(pos.data_ptr(), par.ctypes.data, off.data_ptr(), frames, n_joints,
                                                       rots.data_ptr(), st))

        def mirror_all():  # device part of mirror(mode="all"): fk_quat (rotations only) -> flip -> local
            _lib.check(lib.pmb_fk_quat_f32(rot.data_ptr(), gpos.data_ptr(), 0, off.data_ptr(), 0, par.ctypes.data, frames,
                                           n_joints, None, rots.data_ptr(), st))
            _lib.check(lib.pmb_mirror_to_local_f32(rots.data_ptr(), par.ctypes.data, None, 0, frames, n_joints,
                                                   dq.data_ptr(), st))

        to_dq()
        step = {"to_dq": to_dq, "from_dq": from_dq, "fk_quat": fk_quat, "from_root_positions": from_root_positions,
                "mirror_all": mirror_all, "round_trip": lambda: (to_dq(), from_dq())}[args.op]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    sampler.windows.append((w0, time.perf_counter()))
    barrier()
    ms = ev0.elapsed_time(ev1)
    variant = lib.pmb_last_variant().decode()  # what the timed launches actually ran
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * frames * args.steps / (ms_max * 1e-3)

    if args.kernel_only:
        gather = None
        if args.all_gather and world > 1:
            # the OPTIONAL exchange of SURVEY 8e: every rank ends up with the positions of all frames (NCCL all-gather
            # of equal frame blocks).  Timed on its own, never part of poses/s.
            from pymotion_b200 import sharding

            for _ in range(2):
                full = sharding.all_gather_frames(pos, frames * world)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            g0.record(stream)
            for _ in range(reps):
                full = sharding.all_gather_frames(pos, frames * world)
            g1.record(stream)
            torch.cuda.synchronize(dev)
            tg = torch.tensor([g0.elapsed_time(g1) / reps], device=dev, dtype=torch.float64)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(full[rank * frames:(rank + 1) * frames], pos))
            shard_bytes = pos.numel() * 4
            gather = {"ms": float(tg.item()), "shard_bytes": shard_bytes, "received_bytes_per_rank": shard_bytes * (world - 1),
                      "GBps_received_per_rank": shard_bytes * (world - 1) / (float(tg.item()) * 1e-3) / 1e9,
                      "own_block_intact": same, "includes": "NCCL all_gather_into_tensor into one [F, J, 3] array (equal shards: no extra copy)"}
            del full
        if rank == 0:
            bytes_per_launch = op_bytes_per_pose(args.op, n_joints) * frames
            kernel_ms = ms_max / args.steps
            line = {"workload": args.workload, "op": args.op, "n_gpus": world, "ms_per_step": kernel_ms, "value": value,
                    "GBps": bytes_per_launch / (kernel_ms * 1e-3) / 1e9, "env_chunk": os.environ.get("PMB_FK_CHUNK")}
            if gather:
                line["all_gather_positions"] = gather
            print(json.dumps(line), flush=True)
        sampler.stop_flag.set()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public host-buffer API (pinned host memory both ways)
    e2e_steps = max(1, min(args.steps, 5))
    h_rot = torch.empty((frames, n_joints, 4), dtype=torch.float32, pin_memory=True).copy_(rot)
    h_gpos = torch.empty((frames, 3), dtype=torch.float32, pin_memory=True).copy_(gpos)
    h_off = off.cpu()
    h_pos = torch.empty((frames, n_joints, 3), dtype=torch.float32, pin_memory=True)
    h_rotm = torch.empty((frames, n_joints, 3, 3), dtype=torch.float32, pin_memory=True)
    sk.fk_host(h_rot, h_gpos, h_off, par, out=(h_pos, h_rotm))  # warm-up: allocates the staging workspace
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        sk.fk_host(h_rot, h_gpos, h_off, par, out=(h_pos, h_rotm))  # returns when the outputs are in host memory
    w1 = time.perf_counter()
    sampler.windows.append((w0, w1))
    te = torch.tensor([w1 - w0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * frames * e2e_steps / float(te.item())
    h2d = h_rot.numel() * 4 + h_gpos.numel() * 4 + h_off.numel() * 4
    d2h = h_pos.numel() * 4 + h_rotm.numel() * 4
    # the e2e result must be the same answer as the resident path
    same = bool(torch.equal(h_pos[:4096], pos[:4096].cpu()) and torch.equal(h_rotm[-4096:], rotm[-4096:].cpu()))
    lib.pmb_release_workspace()

    sampler.stop_flag.set()
    sampler.join(timeout=1.0)
    clocks = sampler.summary()

    # ---- CPU baseline: the NumPy oracle port, single thread, bounded sample (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pymotion_oracle as orc
        from pymotion_b200.topologies import synth_numpy

        sample = 400_000 if n_joints <= 22 else 100_000
        c_rot, c_gpos, c_off = synth_numpy(sample, par, seed=0)
        t0 = time.perf_counter()
        c_pos, c_rotm = orc.fk(c_rot, c_gpos, c_off, par)
        dt = time.perf_counter() - t0
        # and use it as the checker it is: the GPU answer on the same sample
        g_pos, g_rotm = sk.fk(torch.from_numpy(c_rot[:50_000]).to(dev), torch.from_numpy(c_gpos[:50_000]).to(dev),
                              torch.from_numpy(c_off).to(dev), par)
        err_p = float(np.abs(g_pos.cpu().numpy() - c_pos[:50_000]).max())
        err_r = float(np.abs(g_rotm.cpu().numpy() - c_rotm[:50_000]).max())
        cpu = {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{sample} frames x {n_joints} joints, one pass of the NumPy oracle port of "
                         f"pymotion.ops.skeleton.fk ({dt:.1f} s, 1 thread of {os.cpu_count()} host cores)",
               "max_abs_err_gpu_vs_cpu": {"positions": err_p, "rotmats": err_r}}

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_launch = fk_bytes_per_pose(n_joints) * frames
        kernel_ms = ms_max / args.steps
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": metric_of(n_joints),
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": kernel_ms,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": args.workload,
                "frames_per_gpu": frames,
                "n_joints": n_joints,
                "topology": topo,
                "sharding": f"frame axis, {world} shard(s), no data-path collective",
                "l2": f"working set {bytes_per_launch / 1e9:.2f} GB per launch >> 126 MB L2: no flush needed",
            },
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload),
                "peak_source": peak_src,
                "frac_of_nominal_8000": achieved / 8000.0,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel": "pmb::" + variant,
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "pymotion_b200.ops.skeleton.fk_host (pmb_fk_f32_host)",
                    "matches_resident_path": same},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fk_1m_x_22", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="development: skip the e2e and CPU legs")
    ap.add_argument("--op", default="fk", choices=DEV_OPS, help="with --kernel-only: which op of the path to time")
    ap.add_argument("--all-gather", action="store_true",
                    help="with --kernel-only under torchrun: also time the optional all-gather of positions")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
