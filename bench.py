#!/usr/bin/env python
"""Benchmark of the fk hot path (BASELINE.json metric: skeleton poses/sec, 22 joints).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (`fk`) over one resident batch.  At N GPUs every rank runs the same per-GPU
batch on its own frames (frame-axis shard, no data-path collective): weak scaling, value = all frames processed /
max-over-ranks device time.  Prints ONE JSON line on rank 0.

  value         kernel throughput of BASELINE config 2 (fk, 1M frames x 22 joints per GPU) with inputs resident in
                HBM, CUDA events on the launch stream
  roofline      algorithmic bytes (64 J + 12 per pose) / kernel time vs MEASURED_PEAKS.json
  extra         the other BASELINE configs measured the same way in the same run: dual-quaternion round trip
                1M x 22 (config 3), fk 4M x 65 (config 4), fk 4M x 52 per GPU (config 5: at --gpus 8 this IS
                32M x 52) and, for N > 1, the OPTIONAL all-gather of positions timed on its own
  e2e           same metric through the drop-in call the reference's users make -- pymotion_b200.ops.skeleton.fk
                (host arrays in) -> host arrays out -- with the H2D / D2H copies inside the timing: page-locked
                host tensors (headline `e2e.value`), plain pageable NumPy arrays, and fk_quat; next to the
                measured H2D+D2H ceiling of this box for the same byte mix
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, staged by __graft_entry__.build()) timed on this box's host:
                BASELINE config 1 exactly (1000 x 22, best of 20, fk / to_root_dual_quat / from_root_dual_quat)
                and a bounded single-thread sample of the 1M x 22 workload

`--impl reference` times the reference's own CPU implementation of the path on all host threads: the real
`pymotion.ops.skeleton.fk` from oracle/_ref (the NumPy oracle port only if the staged reference is missing), one
full 1M x 22 batch per step, frames split over a thread pool in cache-sized blocks (NumPy releases the GIL).
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
UNIT = "poses/s"
NOMINAL_HBM_GBS = 8000.0  # the peak north_star's 70 % target refers to


def metric_of(n_joints: int) -> str:
    return f"skeleton poses/sec ({n_joints} joints)"


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
REF_BLOCK_FRAMES = 2048  # frames per reference call inside a thread: the working set of one call stays in cache


def fk_bytes_per_pose(n_joints: int) -> int:
    return 64 * n_joints + 12  # read 16J + 12, write 12J + 36J


def op_bytes_per_pose(op: str, n_joints: int) -> int:
    return {"fk": 64 * n_joints + 12, "to_dq": 48 * n_joints + 12, "from_dq": 60 * n_joints,
            "round_trip": 108 * n_joints + 12, "fk_quat": 44 * n_joints + 12,
            # 8f rank 2 ops: positions in, rotations out; mirror = rotations in, rotations out
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


def shared_config(workload: str, world: int) -> dict:
    """The `config` object: identical in both arms (the reference arm runs the same workload on the host)."""
    from pymotion_b200.topologies import parents_of

    topo, frames = WORKLOADS[workload]
    n_joints = len(parents_of(topo))
    return {
        "workload": workload,
        "frames_per_gpu": frames,
        "n_joints": n_joints,
        "topology": topo,
        "sharding": f"frame axis, {world} shard(s), no data-path collective",
        "l2": f"working set {fk_bytes_per_pose(n_joints) * frames / 1e9:.2f} GB per launch >> 126 MB L2: no flush needed",
    }


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

    def summary(self, windows=None):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        nv = self.nv
        windows = self.windows if windows is None else windows
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        note = "sampled inside the timed regions"
        if len(inside) < 3:
            lo = min(a for a, _ in windows) - 0.05 if windows else 0.0
            hi = max(b for _, b in windows) + 0.05 if windows else float("inf")
            near = [s for s in self.samples if lo <= s[0] <= hi]
            inside = near if len(near) >= 3 else self.samples
            note = "timed region shorter than the sampling period: samples within 50 ms of it" if inside is near else \
                "timed region shorter than the sampling period: all samples of the run"
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
def reference_modules():
    """(fk, to_root_dual_quat, from_root_dual_quat, kind): the real reference from oracle/_ref when it has been
    staged (`kind: "reference"`), else the NumPy oracle port (`kind: "port"`)."""
    from oracle import fetch_ref

    mods = fetch_ref.import_reference()
    if mods is not None:
        sk = mods[0]
        return sk.fk, sk.to_root_dual_quat, sk.from_root_dual_quat, "reference"
    from oracle import pymotion_oracle as orc

    return orc.fk, orc.to_root_dual_quat, orc.from_root_dual_quat, "port"


def reference_threaded(fk, rot, gpos, offsets, parents, n_threads: int):
    """One pass of the reference's fk over the whole batch: contiguous frame shards on a thread pool, every
    thread calling the reference on cache-sized blocks of its shard (NumPy releases the GIL in its loops)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    bounds = np.linspace(0, rot.shape[0], n_threads + 1).astype(int)

    def work(i):
        for lo in range(bounds[i], bounds[i + 1], REF_BLOCK_FRAMES):
            hi = min(lo + REF_BLOCK_FRAMES, bounds[i + 1])
            fk(rot[lo:hi], gpos[lo:hi], offsets, parents)

    with ThreadPoolExecutor(n_threads) as pool:
        list(pool.map(work, range(n_threads)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    from pymotion_b200.topologies import parents_of, synth_numpy

    topo, frames = WORKLOADS[args.workload]
    par = parents_of(topo)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    fk, _, _, kind = reference_modules()
    frames = min(frames, args.ref_frames) if args.ref_frames else frames
    rot, gpos, off = synth_numpy(frames, par, seed=0)
    for _ in range(max(1, min(args.warmup, 2))):
        reference_threaded(fk, rot, gpos, off, par, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_threaded(fk, rot, gpos, off, par, threads)
    dt = time.perf_counter() - t0
    value = frames * args.steps / dt
    what = "unmodified pymotion.ops.skeleton.fk (oracle/_ref)" if kind == "reference" else "NumPy oracle port of pymotion.ops.skeleton.fk"
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
        "config": shared_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{frames} frames x {len(par)} joints per step (one GPU's batch; the CPU arm does not use "
                                   f"GPUs, same number for every --gpus), {what} over {threads} threads in blocks of "
                                   f"{REF_BLOCK_FRAMES} frames (host has {cores} cores)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(par, n_joints, dev, sk):
    """Rank 0, N = 1: the reference on this box's host.  BASELINE config 1 exactly, plus a bounded single-thread
    sample of the bench workload (the way a user of the reference calls it: one call, one thread), used as the
    checker it is on the way."""
    import numpy as np
    import torch

    from pymotion_b200.topologies import parents_of, synth_numpy

    fk, to_dq, from_dq, kind = reference_modules()
    # ---- 1000 x 22, default_rng(0), 3 warm-ups, best / median of 20
    par1 = parents_of("body22")
    rot1, gpos1, off1 = synth_numpy(1000, par1, seed=0)
    dq1 = to_dq(rot1, gpos1, par1, off1)

    def best_of(fn, reps=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return {"best_ms": 1e3 * ts[0], "median_ms": 1e3 * ts[len(ts) // 2], "poses_per_s_best": 1000 / ts[0]}

    config1 = {"shape": "1000 frames x 22 joints, fp32 inputs, np.random.default_rng(0), 3 warm-ups + best of 20",
               "fk": best_of(lambda: fk(rot1, gpos1, off1, par1)),
               "to_root_dual_quat": best_of(lambda: to_dq(rot1, gpos1, par1, off1)),
               "from_root_dual_quat": best_of(lambda: from_dq(dq1, par1))}
    # the same three calls through the drop-in, NumPy in -> NumPy out (tiny batch: launch + copy latency)
    dq1_f32 = dq1.astype(np.float32)  # the reference promotes to float64; the drop-in computes (and is fed) float32
    ours1 = {"fk": best_of(lambda: sk.fk(rot1, gpos1, off1, par1)),
             "to_root_dual_quat": best_of(lambda: sk.to_root_dual_quat(rot1, gpos1, par1, off1)),
             "from_root_dual_quat": best_of(lambda: sk.from_root_dual_quat(dq1_f32, par1))}
    config1["drop_in_same_calls"] = ours1
    # ---- bounded sample of the bench workload, one call, one thread
    sample = 250_000 if n_joints <= 22 else 60_000
    c_rot, c_gpos, c_off = synth_numpy(sample, par, seed=0)
    t0 = time.perf_counter()
    c_pos, c_rotm = fk(c_rot, c_gpos, c_off, par)
    dt = time.perf_counter() - t0
    n_chk = 50_000
    g_pos, g_rotm = sk.fk(torch.from_numpy(c_rot[:n_chk]).to(dev), torch.from_numpy(c_gpos[:n_chk]).to(dev),
                          torch.from_numpy(c_off).to(dev), par)
    err_p = float(np.abs(g_pos.cpu().numpy() - c_pos[:n_chk]).max())
    err_r = float(np.abs(g_rotm.cpu().numpy() - c_rotm[:n_chk]).max())
    what = "unmodified pymotion.ops.skeleton.fk from oracle/_ref" if kind == "reference" else "NumPy oracle port"
    return {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{sample} frames x {n_joints} joints of the bench workload, ONE call of the {what} "
                      f"({dt:.1f} s, 1 thread of {os.cpu_count()} host cores; NumPy runs this path single-threaded)",
            "max_abs_err_gpu_vs_cpu": {"positions": err_p, "rotmats": err_r, "frames": n_chk},
            "config1": config1}


# --------------------------------------------------------------------------- our arm
class Timer:
    """CUDA-event timing on the launch stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, stream, world, sampler):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.dev, self.stream, self.world, self.sampler = torch, dist, dev, stream, world, sampler

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, step, steps: int, warmup: int):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        ev0.record(self.stream)
        for _ in range(steps):
            step()
        ev1.record(self.stream)
        torch.cuda.synchronize(self.dev)
        window = (w0, time.perf_counter())
        self.sampler.windows.append(window)
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1)) / steps, window


def roofline_of(bytes_per_launch: int, kernel_ms: float, variant: str, traffic_key: str):
    peak, peak_src = measured_peak()
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(traffic_key), "peak_source": peak_src,
            "frac_of_nominal_8000": achieved / NOMINAL_HBM_GBS, "algorithmic_bytes_per_launch": bytes_per_launch,
            "kernel": variant}


def link_ceiling(dev, h2d_bytes: int, d2h_bytes: int, timer: Timer):
    """What this box's host link gives the byte mix of one e2e step: page-locked buffers, plain cudaMemcpyAsync in both
    directions at once on two streams (all ranks at the same time, like the e2e legs)."""
    import torch

    cap = 256 << 20
    scale = min(1.0, cap / max(h2d_bytes, d2h_bytes))
    nh, nd = max(1 << 20, int(h2d_bytes * scale)) // 4, max(1 << 20, int(d2h_bytes * scale)) // 4
    h_in = torch.empty(nh, dtype=torch.float32, pin_memory=True).fill_(1.0)
    h_out = torch.empty(nd, dtype=torch.float32, pin_memory=True)
    d_in, d_out = torch.empty(nh, device=dev), torch.ones(nd, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def both(k_in=1, k_out=1):
        with torch.cuda.stream(s_in):
            for _ in range(k_in):
                d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            for _ in range(k_out):
                h_out.copy_(d_out, non_blocking=True)
        s_in.synchronize(), s_out.synchronize()

    def rate(fn, nbytes, reps=3):
        fn()
        timer.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = timer.max_over_ranks(time.perf_counter() - t0)
        return nbytes * reps / dt / 1e9

    out = {"h2d_alone_GBps": rate(lambda: both(1, 0), 4 * nh), "d2h_alone_GBps": rate(lambda: both(0, 1), 4 * nd)}
    t_mix = (4 * nh + 4 * nd) / rate(lambda: both(1, 1), 4 * nh + 4 * nd) / 1e9  # seconds for the scaled mix, both directions at once
    out["concurrent_mix_GBps"] = (4 * nh + 4 * nd) / t_mix / 1e9
    out["step_seconds_at_ceiling"] = t_mix / scale
    out["how"] = (f"{4 * nh >> 20} MiB H2D + {4 * nd >> 20} MiB D2H of page-locked memory, cudaMemcpyAsync on two streams, "
                  f"per rank, all {timer.world} rank(s) at once")
    return out


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
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    timer = Timer(dev, stream, world, sampler)
    launches = 0  # kernels of ours launched inside timed regions

    class Batch:
        """One resident synthetic batch of a workload and the C-ABI calls on it."""

        def __init__(self, workload):
            self.workload = workload
            self.topo, self.frames = WORKLOADS[workload]
            self.par = parents_of(self.topo)
            self.J = len(self.par)
            self.rot, self.gpos, self.off = synth_torch(self.frames, self.par, dev, seed=1234 + rank)
            self.pos = torch.empty((self.frames, self.J, 3), device=dev, dtype=torch.float32)
            self.off0 = np.zeros(3, dtype=np.float32)

        def alloc(self, *names):
            shapes = {"rotm": (self.frames, self.J, 3, 3), "dq": (self.frames, self.J, 8), "rots": (self.frames, self.J, 4)}
            for n in names:
                setattr(self, n, torch.empty(shapes[n], device=dev, dtype=torch.float32))

        def fk(self):
            _lib.check(lib.pmb_fk_f32(self.rot.data_ptr(), self.gpos.data_ptr(), 3, self.off.data_ptr(), 0, self.par.ctypes.data,
                                      self.frames, self.J, self.pos.data_ptr(), self.rotm.data_ptr(), st))

        def to_dq(self):
            _lib.check(lib.pmb_to_root_dual_quat_f32(self.rot.data_ptr(), self.gpos.data_ptr(), 3, self.par.ctypes.data,
                                                     self.off.data_ptr(), self.off0.ctypes.data, self.frames, self.J,
                                                     self.dq.data_ptr(), st))

        def from_dq(self):
            _lib.check(lib.pmb_from_root_dual_quat_f32(self.dq.data_ptr(), self.par.ctypes.data, self.frames, self.J,
                                                       self.pos.data_ptr(), self.rots.data_ptr(), st))

        def fk_quat(self):
            _lib.check(lib.pmb_fk_quat_f32(self.rot.data_ptr(), self.gpos.data_ptr(), 3, self.off.data_ptr(), 0,
                                           self.par.ctypes.data, self.frames, self.J, self.pos.data_ptr(), self.rots.data_ptr(), st))

        def from_root_positions(self):
            _lib.check(lib.pmb_from_root_positions_f32(self.pos.data_ptr(), self.par.ctypes.data, self.off.data_ptr(), self.frames,
                                                       self.J, self.rots.data_ptr(), st))

        def mirror_all(self):  # device part of mirror(mode="all"): fk (rotations only) -> flip -> local, one fused launch
            _lib.check(lib.pmb_mirror_local_f32(self.rot.data_ptr(), self.par.ctypes.data, None, 0, self.frames, self.J,
                                                self.rots.data_ptr(), self.dq.data_ptr(), st))

    # ------------------------------------------------------------------ development mode: one op, kernel only
    if args.kernel_only:
        b = Batch(args.workload)
        if args.op == "fk":
            b.alloc("rotm")
            step = b.fk
        else:
            if args.op == "from_root_positions":
                b.alloc("rotm")
                b.fk()
                torch.cuda.synchronize(dev)
                del b.rotm
            b.alloc("dq", "rots")
            b.to_dq()
            step = {"to_dq": b.to_dq, "from_dq": b.from_dq, "fk_quat": b.fk_quat, "from_root_positions": b.from_root_positions,
                    "mirror_all": b.mirror_all, "round_trip": lambda: (b.to_dq(), b.from_dq())}[args.op]
        kernel_ms, _ = timer.run(step, args.steps, args.warmup)
        variant = lib.pmb_last_variant().decode()
        gather = all_gather_leg(b, timer, world, rank, dev, stream) if (args.all_gather and world > 1) else None
        if rank == 0:
            bytes_per_launch = op_bytes_per_pose(args.op, b.J) * b.frames
            line = {"workload": args.workload, "op": args.op, "n_gpus": world, "ms_per_step": kernel_ms,
                    "value": world * b.frames / (kernel_ms * 1e-3), "GBps": bytes_per_launch / (kernel_ms * 1e-3) / 1e9,
                    "variant": variant}
            if gather:
                line["all_gather_positions"] = gather
            print(json.dumps(line), flush=True)
        sampler.stop_flag.set()
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ headline: config 2 (or --workload)
    main = Batch(args.workload)
    main.alloc("rotm")
    kernel_ms, main_window = timer.run(main.fk, args.steps, args.warmup)
    launches += args.steps
    variant = lib.pmb_last_variant().decode()  # what the timed launches actually ran
    value = world * main.frames / (kernel_ms * 1e-3)
    n_joints, frames, par = main.J, main.frames, main.par

    # ------------------------------------------------------------------ end to end through the drop-in call
    e2e = e2e_legs(main, sk, lib, timer, world, dev, sampler, args)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(par, n_joints, dev, sk)
    del main
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the other BASELINE configs, same run
    extra = []
    if not args.no_extra:
        x_steps = max(3, min(args.steps, 20))

        def record(name, baseline_config, b, op, step, n_kernels):
            nonlocal launches
            ms, window = timer.run(step, x_steps, 3)
            launches += x_steps * n_kernels
            bytes_per_launch = op_bytes_per_pose(op, b.J) * b.frames
            extra.append({"name": name, "baseline_config": baseline_config, "op": op, "frames_per_gpu": b.frames, "n_joints": b.J,
                          "steps": x_steps, "ms_per_step": ms, "value": world * b.frames / (ms * 1e-3), "unit": UNIT,
                          "roofline": roofline_of(bytes_per_launch, ms, "pmb::" + lib.pmb_last_variant().decode(), name),
                          "clocks": sampler.summary([window])})

        b = Batch("fk_1m_x_22")
        b.alloc("dq", "rots")
        record("round_trip_1m_x_22", "configs[2]: to_root_dual_quat + from_root_dual_quat, 1M x 22", b, "round_trip",
               lambda: (b.to_dq(), b.from_dq()), 2)
        record("to_root_dual_quat_1m_x_22", "configs[2], first half", b, "to_dq", b.to_dq, 1)
        record("from_root_dual_quat_1m_x_22", "configs[2], second half", b, "from_dq", b.from_dq, 1)
        # SURVEY 8f rank 2 on the headline shape: positions -> local rotations, and mirror's rotation step (one fused launch)
        b.alloc("rotm")
        b.fk()
        del b.rotm
        record("from_root_positions_1m_x_22", "SURVEY 8f rank 2: from_root_positions on the fk positions of configs[1]", b,
               "from_root_positions", b.from_root_positions, 1)
        record("mirror_rotations_1m_x_22", "SURVEY 8f rank 2: mirror(mode='all') rotation step (fk -> flip -> local, fused)", b,
               "mirror_all", b.mirror_all, 1)
        del b
        torch.cuda.empty_cache()
        for name, cfg in (("fk_4m_x_65", "configs[3]: fk 4M x 65 deep hierarchy"),
                          ("fk_4m_x_52", f"configs[4]: fk 32M x 52 over 8 GPUs = 4M x 52 per GPU (this run: {world} GPU(s), "
                                         f"{4 * world}M frames)")):
            b = Batch(name)
            b.alloc("rotm")
            record(name, cfg, b, "fk", b.fk, 1)
            if name == "fk_4m_x_65":  # the other two walks at the deep hierarchy (quaternion track kernel), same inputs
                del b.rotm
                b.alloc("dq", "rots")
                record("to_root_dual_quat_4m_x_65", "configs[3] shape, to_root_dual_quat", b, "to_dq", b.to_dq, 1)
                record("fk_quat_4m_x_65", "configs[3] shape, fk emitting global quaternions (SURVEY 8f rank 1)", b, "fk_quat", b.fk_quat, 1)
                record("from_root_positions_4m_x_65", "configs[3] shape, from_root_positions on the fk_quat positions (SURVEY 8f rank 2)", b,
                       "from_root_positions", b.from_root_positions, 1)
                record("mirror_rotations_4m_x_65", "configs[3] shape, mirror(mode='all') rotation step (SURVEY 8f rank 2)", b,
                       "mirror_all", b.mirror_all, 1)
            if name == "fk_4m_x_52" and world > 1:
                extra.append({"name": "all_gather_positions_4m_x_52", "baseline_config": "configs[4], OPTIONAL exchange, never part of poses/s",
                              **all_gather_leg(b, timer, world, rank, dev, stream)})
            del b
            torch.cuda.empty_cache()

    sampler.stop_flag.set()
    sampler.join(timeout=1.0)

    if rank == 0:
        bytes_per_launch = fk_bytes_per_pose(n_joints) * frames
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
            "config": shared_config(args.workload, world),
            "roofline": roofline_of(bytes_per_launch, kernel_ms, "pmb::" + variant, args.workload),
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches + e2e.get("gpu_launches", 0),
            "clocks": sampler.summary(),
            "extra": extra,
        }
        line["clocks"]["headline_region"] = sampler.summary([main_window])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def all_gather_leg(b, timer, world, rank, dev, stream):
    """The OPTIONAL exchange of SURVEY 8e: every rank ends up with the positions of all frames (NCCL all-gather of
    equal frame blocks).  Timed on its own, never part of poses/s."""
    import torch

    from pymotion_b200 import sharding

    frames = b.frames
    for _ in range(2):
        full = sharding.all_gather_frames(b.pos, frames * world)
    timer.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    g0.record(stream)
    for _ in range(reps):
        full = sharding.all_gather_frames(b.pos, frames * world)
    g1.record(stream)
    torch.cuda.synchronize(dev)
    ms = timer.max_over_ranks(g0.elapsed_time(g1) / reps)
    same = bool(torch.equal(full[rank * frames:(rank + 1) * frames], b.pos))
    shard_bytes = b.pos.numel() * 4
    del full
    return {"ms": ms, "shard_bytes": shard_bytes, "received_bytes_per_rank": shard_bytes * (world - 1),
            "GBps_received_per_rank": shard_bytes * (world - 1) / (ms * 1e-3) / 1e9, "own_block_intact": same,
            "includes": "NCCL all_gather_into_tensor into one [F, J, 3] array (equal shards: no extra copy)"}


def e2e_legs(main, sk, lib, timer, world, dev, sampler, args):
    """The same metric through the call a user of the reference makes -- skeleton.fk(host arrays) -> host arrays --
    with every step's H2D and D2H inside the timing (wall clock around the blocking calls, max over ranks)."""
    import numpy as np
    import torch

    frames, n_joints, par = main.frames, main.J, main.par
    steps = max(1, min(args.steps, 5))
    h_rot = torch.empty((frames, n_joints, 4), dtype=torch.float32, pin_memory=True).copy_(main.rot)
    h_gpos = torch.empty((frames, 3), dtype=torch.float32, pin_memory=True).copy_(main.gpos)
    h_off = main.off.cpu()
    h2d = h_rot.numel() * 4 + h_gpos.numel() * 4 + h_off.numel() * 4
    d2h_fk = frames * n_joints * 48
    d2h_fkq = frames * n_joints * 28

    def timed(call):
        # warm-up to the steady state of a loop `pos, rotm = sk.fk(...)`: the staging workspace exists and TWO sets of
        # page-locked result blocks are in torch's host allocator (the previous result is alive while the next is made)
        first = call()
        second = call()
        del first, second
        timer.barrier()
        w0 = time.perf_counter()
        for _ in range(steps):
            out = call()  # returns when the outputs are in host memory
        w1 = time.perf_counter()
        sampler.windows.append((w0, w1))
        return timer.max_over_ranks(w1 - w0) / steps, out

    ceiling = link_ceiling(dev, h2d, d2h_fk, timer)
    # (1) page-locked host tensors in -> page-locked host tensors out, the reference's signature
    t_pin, (p_pos, p_rotm) = timed(lambda: sk.fk(h_rot, h_gpos, h_off, par))
    same = bool(torch.equal(p_pos[:4096], main.pos[:4096].cpu()) and torch.equal(p_rotm[-4096:], main.rotm[-4096:].cpu()))
    del p_pos, p_rotm
    # (2) plain pageable NumPy arrays in -> NumPy arrays out (what `import pymotion_b200.ops.skeleton as sk` users pass)
    n_rot, n_gpos, n_off = np.array(h_rot.numpy()), np.array(h_gpos.numpy()), np.array(h_off.numpy())
    t_np, (n_pos, n_rotm) = timed(lambda: sk.fk(n_rot, n_gpos, n_off, par))
    same_np = bool(isinstance(n_pos, np.ndarray) and np.array_equal(n_rotm[-4096:], main.rotm[-4096:].cpu().numpy()))
    del n_pos, n_rotm
    # (3) fk_quat: global quaternions back instead of matrices (28 J instead of 48 J bytes per frame over PCIe)
    t_q, (q_pos, q_rot) = timed(lambda: sk.fk_quat(h_rot, h_gpos, h_off, par))
    del q_pos, q_rot
    # (4) BASELINE configs[2] called the reference's way: to_root_dual_quat then from_root_dual_quat on pageable NumPy arrays
    def round_trip():
        dq = sk.to_root_dual_quat(n_rot, n_gpos, par, n_off)
        return sk.from_root_dual_quat(dq, par)
    t_rt, (rt_t, rt_r) = timed(round_trip)
    same_rt = bool(np.abs(rt_r[-4096:] - n_rot[-4096:]).max() < 1e-5)
    del rt_t, rt_r
    lib.pmb_release_workspace()
    chunks = -(-frames // max(1, ((48 << 20) // (64 * n_joints + 12) + 31) // 32 * 32))
    return {
        "value": world * frames / t_pin, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_fk, "steps": steps,
        "api": "pymotion_b200.ops.skeleton.fk(rot, global_pos, offsets, parents) on page-locked CPU tensors -> CPU tensors "
               "(the reference's signature, ops/skeleton.py:16; chunked pipeline pmb_fk_f32_host underneath)",
        "ms_per_step": 1e3 * t_pin, "matches_resident_path": same,
        "frac_of_link_ceiling": ceiling["step_seconds_at_ceiling"] / t_pin,
        "link_ceiling": ceiling,
        "pageable_numpy": {"value": world * frames / t_np, "ms_per_step": 1e3 * t_np, "matches_resident_path": same_np,
                           "api": "skeleton.fk(pageable NumPy arrays) -> NumPy arrays (results in page-locked memory; inputs staged "
                                  "through the library's pinned ring by host threads)",
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_fk},
        "fk_quat": {"value": world * frames / t_q, "ms_per_step": 1e3 * t_q, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_fkq,
                    "api": "skeleton.fk_quat(page-locked CPU tensors): global quaternions instead of rotation matrices"},
        "round_trip_numpy": {"value": world * frames / t_rt, "ms_per_step": 1e3 * t_rt, "recovers_rotations": same_rt,
                             "api": "skeleton.to_root_dual_quat(pageable NumPy) -> skeleton.from_root_dual_quat(...) -> NumPy (configs[2] "
                                    "through the drop-in calls; both through the host pipeline)",
                             "h2d_bytes_per_step": h2d + frames * n_joints * 32, "d2h_bytes_per_step": frames * n_joints * 60},
        "gpu_launches": 5 * (steps + 2) * chunks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fk_1m_x_22", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (extra records)")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: cap the frames per step (tests)")
    ap.add_argument("--kernel-only", action="store_true", help="development: one op, kernel time only")
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
