// libpymotion_b200.so, dual-quaternion translation unit: to_root_dual_quat, from_root_dual_quat, from_global_rotations.
// Host side only validates, looks up the per-topology program, picks a launch configuration and launches.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "dq_kernels.cuh"
#include "host_common.h"
#include "qtracks_host.h"

using namespace pmbh;

extern "C" {

int pmb_to_root_dual_quat_f32(const float *rotations, const float *global_pos, int64_t gpos_frame_stride,
                              const int64_t *parents_host, const float *offsets, const float *offsets_host0,
                              int64_t n_frames, int32_t n_joints, float *dq, void *stream) {
    if (!rotations || !global_pos || !offsets || !dq) return fail(PMB_ERR_NULL, "to_root_dual_quat: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (gpos_frame_stride != 0 && gpos_frame_stride != 3) return fail(PMB_ERR_SHAPE, "gpos_frame_stride must be 0 or 3");
    if (!aligned16(rotations) || !aligned16(dq)) return fail(PMB_ERR_ALIGN, "rotations and dq must be 16-byte aligned");
    if (offsets_host0 && (offsets_host0[0] != 0.f || offsets_host0[1] != 0.f || offsets_host0[2] != 0.f))
        return fail(PMB_ERR_ROOT_OFFSET, "offsets[0] must be zero (ops/skeleton.py:227)");
    const pmb::JointProgram *prog_p = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, true, prog_p, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    if (n_frames > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames must be below 2^31 per call");
    // The quaternion track kernel (qtracks_kernel.cuh) unless it does not apply; PMB_DQ_TRACKS = 0 / 1 forces, and forcing
    // a flush group of the chain kernel (PMB_DQ_GROUP / PMB_DQ_BLOCKS_PER_SM) selects the chain kernel.
    const int tracks = knob(K_DQ_TRACKS, (knob_set(K_DQ_GROUP) || knob_set(K_DQ_BLOCKS_PER_SM)) ? 0 : -1);
    if (tracks != 0) {
        int trc = PMB_OK;
        if (launch_qtracks<pmb::kQtDq>(rotations, global_pos, gpos_frame_stride, offsets, parents_host, n_frames, n_joints, dq, nullptr,
                                       static_cast<cudaStream_t>(stream), dp, tracks == 1, trc))
            return trc;
    }
    // Joints per flush.  Dual quaternions are whole 32-byte sectors, so partial flushes cost DRAM little and
    // occupancy matters more than for fk (measured, 1M x 22: whole rows / 4 warps per SM 0.239 ms, 8 joints /
    // 12 warps 0.214 ms, 16 joints / 8 warps 0.186 ms): take the largest group that still lets TWO 4-warp
    // blocks share an SM.
    constexpr int WARPS = 4;
    const int budget = (dp.smem_optin - 2048) / 2;
    int group = n_joints;
    if (pmb::dq_geom(group, WARPS, n_joints, n_slots).block_bytes > budget) {
        group = ((n_joints + 7) / 8) * 8;
        while (group > 8 && pmb::dq_geom(group, WARPS, n_joints, n_slots).block_bytes > budget) group -= 8;
    }
    if (knob_set(K_DQ_GROUP)) {
        const int v = knob(K_DQ_GROUP, 0);
        if (v >= 8 && v % 8 == 0 && v < n_joints) group = v;
    }
    const int smem = pmb::dq_geom(group, WARPS, n_joints, n_slots).block_bytes;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", n_slots);
    auto kernel = pmb::to_root_dq_kernel<WARPS>;
    int per_sm = 0;
    if ((rc = kernel_fit(kernel, dp, WARPS * 32, smem, per_sm))) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, rotations, n_frames, n_joints, pmb::kChunk))) return rc;
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "to_root_dual_quat kernel does not fit on an SM");
    per_sm = std::max(1, std::min(per_sm, knob(K_DQ_BLOCKS_PER_SM, per_sm)));
    const long long tiles = (n_frames + 31) / 32;
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    auto magic_of = [](int d) { return static_cast<uint32_t>((1ULL << 32) / static_cast<uint32_t>(d)) + 1u; };
    const int tail = n_joints % group ? n_joints % group : group;
    note_variant("to_root_dq_kernel<WARPS=%d> group=%d grid=%lld smem=%d", WARPS, group, blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        tm, global_pos, gpos_frame_stride, offsets, reinterpret_cast<float4 *>(dq), n_frames, n_joints, n_slots, group,
        magic_of(2 * group), magic_of(2 * tail), *prog_p);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_from_root_dual_quat_f32(const float *dq, const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                                float *translations, float *rotations, void *stream) {
    if (!dq || !translations || !rotations) return fail(PMB_ERR_NULL, "from_root_dual_quat: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (!aligned16(dq) || !aligned16(rotations) || !aligned16(translations))
        return fail(PMB_ERR_ALIGN, "dq, rotations and translations must be 16-byte aligned");
    const pmb::JointProgram *prog_p = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, true, prog_p, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    constexpr int THREADS = 256;
    // Elements (frame, joint) per block tile.  The translations stage costs 12 bytes of shared memory per element and
    // every thread has one 32-byte element in flight, so the tile size sets the bytes in flight per SM: 2304 elements
    // (27 KB) lets the 8 blocks of 256 threads the SM can hold all be resident.  Measured against the 4096 of the
    // first half (profiles/r1_sweep_frdq_tile.jsonl): 4M x 52 2.249 -> 1.888 ms, 4M x 65 2.756 -> 2.319 ms, 1M x 22 unchanged.
    const int fb = tile_frames(n_joints, knob(K_FRDQ_ELEMS, 2304));
    const int smem = ((fb * n_joints * 12 + 15) & ~15) + ((n_joints * 2 + 15) & ~15);
    auto kernel = pmb::from_root_dq_kernel<THREADS>;
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    int per_sm_unused = 0;
    if ((rc = kernel_fit(kernel, dp, THREADS, smem, per_sm_unused))) return rc;
    const long long blocks = (n_frames + fb - 1) / fb;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    const uint32_t magic = magic_small(n_joints);
    note_variant("from_root_dq_kernel<%d> tile=%d frames grid=%lld smem=%d", THREADS, fb, blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(dq), translations, reinterpret_cast<float4 *>(rotations), n_frames, n_joints,
        fb, magic, *prog_p);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_from_global_rotations_f32(const float *global_quats, const int64_t *parents_host, int64_t n_frames,
                                  int32_t n_joints, float *local_quats, void *stream) {
    if (!global_quats || !local_quats) return fail(PMB_ERR_NULL, "from_global_rotations: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (!aligned16(global_quats) || !aligned16(local_quats))
        return fail(PMB_ERR_ALIGN, "quaternion arrays must be 16-byte aligned");
    const pmb::JointProgram *prog_p = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, false, prog_p, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    constexpr int THREADS = 256;
    const int fb = tile_frames(n_joints, 4096);
    const int smem = (n_joints * 2 + 15) & ~15;
    const long long blocks = (n_frames + fb - 1) / fb;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    const uint32_t magic = magic_small(n_joints);
    pmb::from_global_rotations_kernel<THREADS><<<static_cast<unsigned>(blocks), THREADS, smem,
                                                 static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(global_quats), reinterpret_cast<float4 *>(local_quats), n_frames, n_joints, fb,
        magic, *prog_p);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

}  // extern "C"
