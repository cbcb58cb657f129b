// Host-side launcher of the quaternion track kernel (qtracks_kernel.cuh), shared by the dual-quaternion and the fk
// translation units: looks up the whole-skeleton four-track schedule, picks the block shape that puts the most warps
// (= tiles of 8 frames in flight) on an SM and launches.  Nothing here computes on the CPU.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>

#include "host_common.h"
#include "qtracks_kernel.cuh"

namespace pmbh {

struct QtShape {
    int warps = 0, blocks = 0, smem = 0;  // warps per block, blocks per SM, dynamic shared memory per block
};
// Every block carries its own copy of the item table and costs 1 KB of system shared memory.
inline QtShape qt_shape(int mode, int n_joints, int n_items, int warps_cap, const DeviceProps &dp, int n_tracks = pmb::kQtTracks,
                        int tile_frames = pmb::kQtFrames) {
    QtShape best;
    for (int blocks = 1; blocks <= 4; ++blocks)
        for (int warps = 16; warps >= 1; --warps) {
            const int smem = pmb::qt_geom(mode, warps, n_joints, n_items, n_tracks, tile_frames).block_bytes;
            if (smem > dp.smem_optin || blocks * (smem + 1024) > dp.smem_sm) continue;
            if (blocks * warps > warps_cap) continue;
            if (blocks * warps > best.blocks * best.warps) best = {warps, blocks, smem};
            break;  // fewer warps per block only lowers the product for this block count
        }
    return best;
}

template <int MODE, int NT, int FQ>
bool launch_qtracks_shape(const float *rot, const float *gpos, long long gstride, const float *offsets, long long n_frames, int n_joints,
                          float *out_q, float *out_p, cudaStream_t stream, const DeviceProps &dp, bool forced, int &rc,
                          const pmb::TrackProgram &tp, int n_steps, const typename pmb::QtMirrorArg<MODE>::type &mir) {
    const QtShape sh = qt_shape(MODE, n_joints, n_steps * NT, knob(K_QT_WARPS_PER_SM, 32), dp, NT, FQ);
    if (sh.warps == 0) return false;
    if (!forced && (2 * n_joints < n_steps * NT || sh.warps * sh.blocks < 4)) return false;
    constexpr bool kCanPipe = MODE != pmb::kQtDq;
    const bool pipe = kCanPipe && knob(K_QT_PIPE, n_joints > 30 ? 1 : 0) != 0;
    auto kernel = pipe ? pmb::qtracks_kernel<MODE, kCanPipe, NT, FQ> : pmb::qtracks_kernel<MODE, false, NT, FQ>;
    int per_sm = 0;
    if ((rc = kernel_fit(kernel, dp, sh.warps * 32, sh.smem, per_sm))) return true;
    if (per_sm < 1) return false;
    per_sm = std::min(per_sm, sh.blocks);
    const long long tiles = (n_frames + FQ - 1) / FQ;
    const long long blocks = std::min<long long>((tiles + sh.warps - 1) / sh.warps, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("qtracks_kernel<MODE=%d,PIPE=%d,%dx%d> steps=%d grid=%lld x %d warps (%d warps/SM) smem=%d", MODE, int(pipe), NT, FQ, n_steps,
                 blocks, sh.warps, per_sm * sh.warps, sh.smem);
    kernel<<<static_cast<unsigned>(blocks), sh.warps * 32, sh.smem, stream>>>(reinterpret_cast<const float4 *>(rot), gpos, gstride, offsets,
                                                                            reinterpret_cast<float4 *>(out_q), out_p, n_frames, n_joints,
                                                                            n_steps, knob(K_QT_DYNAMIC, 1), tp, mir);
    rc = PMB_OK;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "qtracks_kernel launch");
    return true;
}

// The lane shape and its schedule: 3 tracks x 10 frames for fk_quat where three tracks need no more steps than four, else 4 x 8
// (PMB_QT_SHAPE = 1 / 2 forces 4 x 8 / 3 x 10).  n_steps = 0: no schedule fits.
inline int qt_pick_shape(int mode, const int64_t *parents_host, int n_joints, const pmb::TrackProgram *&tp, int &n_steps, bool &three) {
    int steps4 = 0, steps3 = 0, rc;
    if ((rc = track_program(parents_host, n_joints, 3, 0, tp, steps3))) return rc;
    if ((rc = track_program(parents_host, n_joints, 4, 0, tp, steps4))) return rc;  // tp: the four-track program
    const int shape = knob(K_QT_SHAPE, 0);
    // measured at 1M x 22 (8 steps either way; profiles/r2_sweep_qt_shape_*.jsonl): fk_quat 0.1729 -> 0.1533 ms with 3 x 10 (16
    // warps per SM, improving with every warp); to_root_dual_quat 0.1783 -> 0.1871 and mirror 0.1764 -> 0.1953 with the 15 warps
    // that fit (dual quaternions do reach 0.1687 at 12 warps, but 0.1826 at 14: not an optimum to ship): only fk_quat takes 3 x 10
    three = shape == 2 || (shape == 0 && mode == pmb::kQtFkQuat && steps3 > 0 && (steps4 == 0 || steps3 <= steps4));
    n_steps = three ? steps3 : steps4;
    if (three && steps3 > 0) return track_program(parents_host, n_joints, 3, 0, tp, steps3);  // a cache hit: tp = the three-track program
    return PMB_OK;
}

// Returns false if the kernel does not apply (schedule or stage does not fit, or -- unless forced -- a topology the
// level schedule fills less than half, or a skeleton so large that fewer than four warps fit an SM: the
// thread-per-frame chain kernels take those); rc carries the status when it returns true.
// Measured on B200 (profiles/r2_sweep_qt_*.jsonl), chain kernel -> this kernel:
//     to_root_dual_quat  1M x 22 0.186 -> 0.178 ms,  4M x 52 2.00 -> 1.70,  4M x 65 2.60 -> 2.20
//     fk_quat            1M x 22 0.186 -> 0.172 ms,  4M x 52 1.96 -> 1.50,  4M x 65 3.17 -> 1.91
// PIPE (next step's table entry / quaternion fetched a step early): pays where the normalisation of the local quaternion
// leaves the critical path with it and the walk is long (fk_quat, 52 / 65 joints: 4 - 15 %), costs 3 - 5 % elsewhere.
// Lane shape: 3 tracks x 10 frames where three tracks need no more steps than four (the 22-joint body: 8 steps either way),
// else 4 x 8; PMB_QT_SHAPE = 1 / 2 forces 4 x 8 / 3 x 10.
template <int MODE>
bool launch_qtracks(const float *rot, const float *gpos, long long gstride, const float *offsets, const int64_t *parents_host,
                    long long n_frames, int n_joints, float *out_q, float *out_p, cudaStream_t stream, const DeviceProps &dp,
                    bool forced, int &rc, const typename pmb::QtMirrorArg<MODE>::type &mir = typename pmb::QtMirrorArg<MODE>::type()) {
    const pmb::TrackProgram *tp = nullptr;
    int n_steps = 0;
    bool three = false;
    if ((rc = qt_pick_shape(MODE, parents_host, n_joints, tp, n_steps, three))) return true;
    if (n_steps == 0) return false;
    if (three)
        return launch_qtracks_shape<MODE, 3, 10>(rot, gpos, gstride, offsets, n_frames, n_joints, out_q, out_p, stream, dp, forced, rc, *tp, n_steps, mir);
    return launch_qtracks_shape<MODE, 4, 8>(rot, gpos, gstride, offsets, n_frames, n_joints, out_q, out_p, stream, dp, forced, rc, *tp, n_steps, mir);
}

}  // namespace pmbh
