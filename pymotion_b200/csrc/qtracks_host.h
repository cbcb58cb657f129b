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
inline QtShape qt_shape(int mode, int n_joints, int n_items, int warps_cap, const DeviceProps &dp) {
    QtShape best;
    for (int blocks = 1; blocks <= 4; ++blocks)
        for (int warps = 16; warps >= 1; --warps) {
            const int smem = pmb::qt_geom(mode, warps, n_joints, n_items).block_bytes;
            if (smem > dp.smem_optin || blocks * (smem + 1024) > dp.smem_sm) continue;
            if (blocks * warps > warps_cap) continue;
            if (blocks * warps > best.blocks * best.warps) best = {warps, blocks, smem};
            break;  // fewer warps per block only lowers the product for this block count
        }
    return best;
}

// Returns false if the kernel does not apply (schedule or stage does not fit, or -- unless forced -- a topology the
// four-track schedule fills less than half, or a skeleton so large that fewer than four warps fit an SM: the
// thread-per-frame chain kernels take those); rc carries the status when it returns true.
// Measured on B200 (profiles/r2_sweep_qt_*.jsonl), chain kernel -> this kernel:
//     to_root_dual_quat  1M x 22 0.186 -> 0.178 ms,  4M x 52 2.00 -> 1.70,  4M x 65 2.60 -> 2.20
//     fk_quat            1M x 22 0.186 -> 0.172 ms,  4M x 52 1.96 -> 1.50,  4M x 65 3.17 -> 1.91
// PIPE (next step's table entry / quaternion fetched a step early): pays where the normalisation of the local quaternion
// leaves the critical path with it and the walk is long (fk_quat, 52 / 65 joints: 4 - 15 %), costs 3 - 5 % elsewhere.
template <int MODE>
bool launch_qtracks(const float *rot, const float *gpos, long long gstride, const float *offsets, const int64_t *parents_host,
                    long long n_frames, int n_joints, float *out_q, float *out_p, cudaStream_t stream, const DeviceProps &dp,
                    bool forced, int &rc, const typename pmb::QtMirrorArg<MODE>::type &mir = typename pmb::QtMirrorArg<MODE>::type()) {
    const pmb::TrackProgram *tp = nullptr;
    int n_steps = 0;
    if ((rc = track_program(parents_host, n_joints, pmb::kQtTracks, 0, tp, n_steps))) return true;
    if (n_steps == 0) return false;
    const QtShape sh = qt_shape(MODE, n_joints, n_steps * pmb::kQtTracks, knob(K_QT_WARPS_PER_SM, 32), dp);
    if (sh.warps == 0) return false;
    if (!forced && (2 * n_joints < n_steps * pmb::kQtTracks || sh.warps * sh.blocks < 4)) return false;
    constexpr bool kCanPipe = MODE != pmb::kQtDq;
    const bool pipe = kCanPipe && knob(K_QT_PIPE, n_joints > 30 ? 1 : 0) != 0;
    auto kernel = pipe ? pmb::qtracks_kernel<MODE, kCanPipe> : pmb::qtracks_kernel<MODE, false>;
    int per_sm = 0;
    if ((rc = kernel_fit(kernel, dp, sh.warps * 32, sh.smem, per_sm))) return true;
    if (per_sm < 1) return false;
    per_sm = std::min(per_sm, sh.blocks);
    const long long tiles = (n_frames + pmb::kQtFrames - 1) / pmb::kQtFrames;
    const long long blocks = std::min<long long>((tiles + sh.warps - 1) / sh.warps, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("qtracks_kernel<MODE=%d,PIPE=%d> steps=%d grid=%lld x %d warps (%d warps/SM) smem=%d", MODE, int(pipe), n_steps, blocks, sh.warps,
                 per_sm * sh.warps, sh.smem);
    kernel<<<static_cast<unsigned>(blocks), sh.warps * 32, sh.smem, stream>>>(reinterpret_cast<const float4 *>(rot), gpos, gstride, offsets,
                                                                            reinterpret_cast<float4 *>(out_q), out_p, n_frames, n_joints,
                                                                            n_steps, knob(K_QT_DYNAMIC, 1), *tp, mir);
    rc = PMB_OK;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "qtracks_kernel launch");
    return true;
}

}  // namespace pmbh
