// libpymotion_b200.so, element-wise translation unit: the quat / dual_quat / ortho6d primitives, unroll, center_of_mass,
// interpolate_positions, vector.normalize, and the fk consumers from_root_positions / mirror (which share
// the from_to device code of rotations_ext.cuh).
// Host side only validates, looks up the per-topology program, picks a launch configuration and launches.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "elementwise.cuh"
#include "ik_kernels.cuh"
#include "misc_ops.cuh"
#include "rotations_ext.cuh"
#include "host_common.h"

using namespace pmbh;

extern "C" {

// ---- element-wise ---------------------------------------------------------------
int pmb_quat_mul_f32(const float *q0, const float *q1, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q0, q1, out);
    PMB_NEED16(q0); PMB_NEED16(q1); PMB_NEED16(out);
    pmb::quat_mul_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q0, (const float4 *)q1, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_mul_vec_f32(const float *q, const float *v, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, v, out);
    PMB_NEED16(q);
    pmb::quat_mul_vec_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, v, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_length_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q);
    pmb::quat_length_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_normalize_f32(const float *q, float eps, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q); PMB_NEED16(out);
    pmb::quat_normalize_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, eps, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_conjugate_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q); PMB_NEED16(out);
    pmb::quat_conjugate_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_matrix_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q);
    PMB_NEED16(out);
    pmb::quat_to_matrix_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_matrix_f32(const float *m, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, m, out);
    PMB_NEED16(out);
    pmb::quat_from_matrix_kernel<<<grid_, 256, 0, st_>>>(m, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_from_rotation_translation_f32(const float *rotations, const float *translations, float *dq, int64_t n,
                                         void *stream) {
    PMB_EW_PROLOGUE(n, rotations, translations, dq);
    PMB_NEED16(rotations); PMB_NEED16(dq);
    if (aligned32(dq)) pmb::dq_from_rt_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)rotations, translations, (float4 *)dq, n);
    else pmb::dq_from_rt_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)rotations, translations, (float4 *)dq, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_from_translation_f32(const float *translations, float *dq, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, translations, dq);
    PMB_NEED16(dq);
    if (aligned32(dq)) pmb::dq_from_t_kernel<true><<<grid_, 256, 0, st_>>>(translations, (float4 *)dq, n);
    else pmb::dq_from_t_kernel<false><<<grid_, 256, 0, st_>>>(translations, (float4 *)dq, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_to_rotation_translation_f32(const float *dq, float *rotations, float *translations, int64_t n,
                                       void *stream) {
    PMB_EW_PROLOGUE(n, dq, rotations, translations);
    PMB_NEED16(dq); PMB_NEED16(rotations);
    if (aligned32(dq)) pmb::dq_to_rt_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)rotations, translations, n);
    else pmb::dq_to_rt_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)rotations, translations, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- the rest of the quaternion / dual-quaternion surface (rotations_ext.cuh) ------------------
int pmb_quat_from_angle_axis_f32(const float *angle, const float *axis, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, angle, axis, out);
    PMB_NEED16(out);
    pmb::quat_from_angle_axis_kernel<<<grid_, 256, 0, st_>>>(angle, axis, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_scaled_angle_axis_f32(const float *scaled_axis, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, scaled_axis, out);
    PMB_NEED16(out);
    pmb::quat_from_scaled_angle_axis_kernel<<<grid_, 256, 0, st_>>>(scaled_axis, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_euler_f32(const float *euler, const uint8_t *order_codes, int64_t order_stride, float *out, int64_t n,
                            void *stream) {
    PMB_EW_PROLOGUE(n, euler, order_codes, out);
    PMB_NEED16(out);
    if (order_stride != 0 && order_stride != 1) return fail(PMB_ERR_SHAPE, "%s: order_stride must be 0 or 1", __func__);
    pmb::quat_from_euler_kernel<<<grid_, 256, 0, st_>>>(euler, order_codes, order_stride, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_euler_f32(const float *q, const uint8_t *order_codes, int64_t order_stride, float *out, int64_t n,
                          void *stream) {
    PMB_EW_PROLOGUE(n, q, order_codes, out);
    PMB_NEED16(q);
    if (order_stride != 0 && order_stride != 1) return fail(PMB_ERR_SHAPE, "%s: order_stride must be 0 or 1", __func__);
    pmb::quat_to_euler_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, order_codes, order_stride, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_angle_axis_f32(const float *q, float *angle, float *axis, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, angle, axis);
    PMB_NEED16(q);
    pmb::quat_to_angle_axis_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, angle, axis, 0, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_scaled_angle_axis_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q);
    pmb::quat_to_angle_axis_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, nullptr, out, 1, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_slerp_f32(const float *q0, const float *q1, const float *t, int64_t t_stride, int32_t shortest, float *out,
                       int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q0, q1, t, out);
    PMB_NEED16(q0); PMB_NEED16(q1); PMB_NEED16(out);
    if (t_stride != 0 && t_stride != 1) return fail(PMB_ERR_SHAPE, "%s: t_stride must be 0 or 1", __func__);
    pmb::quat_slerp_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q0, (const float4 *)q1, t, t_stride, shortest,
                                                   (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_to_f32(const float *v1, const float *v2, int32_t normalize_input, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, v1, v2, out);
    PMB_NEED16(out);
    pmb::quat_from_to_kernel<<<grid_, 256, 0, st_>>>(v1, v2, nullptr, normalize_input, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_to_axis_f32(const float *v1, const float *v2, const float *rot_axis, int32_t normalize_input, float *out,
                              int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, v1, v2, rot_axis, out);
    PMB_NEED16(out);
    pmb::quat_from_to_kernel<<<grid_, 256, 0, st_>>>(v1, v2, rot_axis, normalize_input, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int64_t pmb_unroll_workspace_bytes(int64_t n_steps, int64_t n_cols) {
    if (n_steps < 0 || n_cols < 0) return 0;
    const int64_t chunks = (n_steps + pmb::kUnrollChunk - 1) / pmb::kUnrollChunk;
    return n_steps * n_cols + chunks * n_cols + 16;
}
int pmb_unroll_f32(const float *x, int32_t width, int64_t n_steps, int64_t n_cols, float *out, void *workspace,
                   int64_t workspace_bytes, void *stream) {
    if (width != 4 && width != 8) return fail(PMB_ERR_SHAPE, "%s: width must be 4 (quaternions) or 8 (dual quaternions)", __func__);
    if (n_steps < 0 || n_cols < 0) return fail(PMB_ERR_SHAPE, "%s: negative size", __func__);
    if (n_steps == 0 || n_cols == 0) return PMB_OK;
    if (!x || !out || !workspace) return fail(PMB_ERR_NULL, "%s: NULL array pointer", __func__);
    PMB_NEED16(x); PMB_NEED16(out);
    if (workspace_bytes < pmb_unroll_workspace_bytes(n_steps, n_cols))
        return fail(PMB_ERR_SHAPE, "%s: workspace too small (%lld < %lld bytes)", __func__,
                    static_cast<long long>(workspace_bytes), static_cast<long long>(pmb_unroll_workspace_bytes(n_steps, n_cols)));
    DeviceProps dp;
    int rc = device_props(dp);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t chunks = (n_steps + pmb::kUnrollChunk - 1) / pmb::kUnrollChunk;
    if (n_cols > 0x7FFFFFFFLL || (chunks * n_cols + 127) / 128 > 0x7FFFFFFFLL)
        return fail(PMB_ERR_SHAPE, "%s: array too large for one launch", __func__);
    uint8_t *local = static_cast<uint8_t *>(workspace), *agg = local + n_steps * n_cols;
    const int w4 = width / 4;
    const long long local_threads = chunks * n_cols;
    pmb::unroll_local_kernel<<<static_cast<unsigned>((local_threads + 127) / 128), 128, 0, st>>>(
        reinterpret_cast<const float4 *>(x), w4, n_steps, n_cols, chunks, local, agg);
    PMB_CUDA(cudaGetLastError());
    pmb::unroll_chunks_kernel<<<static_cast<unsigned>(n_cols), 256, 0, st>>>(agg, chunks, n_cols);
    PMB_CUDA(cudaGetLastError());
    if (n_cols <= 512 && knob(K_UNROLL_CHUNK_APPLY, 1))
        pmb::unroll_apply_chunk_kernel<<<static_cast<unsigned>(chunks), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(x), w4, n_steps, static_cast<int>(n_cols), magic_small(static_cast<int>(n_cols)),
            local, agg, reinterpret_cast<float4 *>(out));
    else
        pmb::unroll_apply_kernel<<<ew_grid(n_steps * n_cols, 256, dp), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(x), w4, n_steps, n_cols, local, agg, reinterpret_cast<float4 *>(out));
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// io/bvh.py:352-359: the numeric part of BVH.get_data, one fused three-pass scan (rotations_ext.cuh)
int pmb_bvh_rotations_to_quat_f32(const float *euler_deg, const uint8_t *order_codes, int64_t n_frames, int64_t n_joints, float *out,
                                  void *workspace, int64_t workspace_bytes, void *stream) {
    if (n_frames < 0 || n_joints < 0) return fail(PMB_ERR_SHAPE, "%s: negative size", __func__);
    if (n_frames == 0 || n_joints == 0) return PMB_OK;
    if (!euler_deg || !order_codes || !out || !workspace) return fail(PMB_ERR_NULL, "%s: NULL array pointer", __func__);
    PMB_NEED16(out);
    if (workspace_bytes < pmb_unroll_workspace_bytes(n_frames, n_joints))
        return fail(PMB_ERR_SHAPE, "%s: workspace too small (%lld < %lld bytes)", __func__, static_cast<long long>(workspace_bytes),
                    static_cast<long long>(pmb_unroll_workspace_bytes(n_frames, n_joints)));
    DeviceProps dp;
    int rc = device_props(dp);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t chunks = (n_frames + pmb::kUnrollChunk - 1) / pmb::kUnrollChunk;
    if (n_joints > 0x7FFFFFFFLL || (chunks * n_joints + 127) / 128 > 0x7FFFFFFFLL)
        return fail(PMB_ERR_SHAPE, "%s: array too large for one launch", __func__);
    uint8_t *local = static_cast<uint8_t *>(workspace), *agg = local + n_frames * n_joints;
    const long long local_threads = chunks * n_joints;
    pmb::bvh_local_kernel<<<static_cast<unsigned>((local_threads + 127) / 128), 128, 0, st>>>(euler_deg, order_codes, n_frames, n_joints,
                                                                                              chunks, local, agg);
    PMB_CUDA(cudaGetLastError());
    pmb::unroll_chunks_kernel<<<static_cast<unsigned>(n_joints), 256, 0, st>>>(agg, chunks, n_joints);
    PMB_CUDA(cudaGetLastError());
    pmb::bvh_apply_kernel<<<ew_grid(n_frames * n_joints, 256, dp), 256, 0, st>>>(euler_deg, order_codes, n_frames, n_joints, local, agg,
                                                                                reinterpret_cast<float4 *>(out));
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_dq_is_unit_f32(const float *dq, float atol, int64_t n, int32_t *flags3, void *stream) {
    if (!flags3) return fail(PMB_ERR_NULL, "%s: flags3 is NULL", __func__);
    PMB_CUDA(cudaMemsetAsync(flags3, 0, 3 * sizeof(int32_t), static_cast<cudaStream_t>(stream)));
    PMB_EW_PROLOGUE(n, dq);
    PMB_NEED16(dq);
    pmb::dq_is_unit_kernel<<<grid_, 256, 0, st_>>>((const float4 *)dq, atol, flags3, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_normalize_f32(const float *dq, float *out, int64_t n, int32_t *flags3, void *stream) {
    if (!flags3) return fail(PMB_ERR_NULL, "%s: flags3 is NULL", __func__);
    PMB_CUDA(cudaMemsetAsync(flags3, 0, 3 * sizeof(int32_t), static_cast<cudaStream_t>(stream)));
    PMB_EW_PROLOGUE(n, dq, out);
    PMB_NEED16(dq); PMB_NEED16(out);
    // pass 1: one read, scaled values written, whole-array is_unit verdict reduced into flags3; pass 2 returns at once when
    // the verdict is "unit", else projects the output in place (rotations_ext.cuh)
    if (aligned32(dq) && aligned32(out)) {
        pmb::dq_normalize_scale_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)out, flags3, n);
        PMB_CUDA(cudaGetLastError());
        pmb::dq_normalize_project_kernel<true><<<grid_, 256, 0, st_>>>((float4 *)out, flags3, n);
    } else {
        pmb::dq_normalize_scale_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)out, flags3, n);
        PMB_CUDA(cudaGetLastError());
        pmb::dq_normalize_project_kernel<false><<<grid_, 256, 0, st_>>>((float4 *)out, flags3, n);
    }
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- ortho6d, center_of_mass, interpolate_positions, vector.normalize, and the fk consumers from_root_positions / mirror (which share
// the from_to device code of rotations_ext.cuh) (misc_ops.cuh) ---------
int pmb_ortho6d_from_matrix_f32(const float *rotmats, float *ortho6d, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, rotmats, ortho6d);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    pmb::ortho6d_from_matrix_kernel<<<grid_, 256, 0, st_>>>(rotmats, ortho6d, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_ortho6d_from_quat_f32(const float *q, float *ortho6d, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, ortho6d);
    PMB_NEED16(q);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    pmb::ortho6d_from_quat_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, ortho6d, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_ortho6d_to_matrix_f32(const float *ortho6d, float *rotmats, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, ortho6d, rotmats);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    PMB_NEED16(rotmats);
    pmb::ortho6d_to_matrix_kernel<<<grid_, 256, 0, st_>>>(ortho6d, rotmats, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_ortho6d_to_quat_f32(const float *ortho6d, float *q, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, ortho6d, q);
    PMB_NEED16(q);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    pmb::ortho6d_to_quat_kernel<<<grid_, 256, 0, st_>>>(ortho6d, (float4 *)q, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_center_of_mass_f32(const float *joints, const float *weights, int64_t weights_frame_stride, int64_t n_frames,
                           int32_t n_joints, float *out, void *stream) {
    if (n_joints < 1) return fail(PMB_ERR_SHAPE, "%s: n_joints < 1", __func__);
    if (weights_frame_stride != 0 && weights_frame_stride != n_joints)
        return fail(PMB_ERR_SHAPE, "%s: weights_frame_stride must be 0 or n_joints", __func__);
    PMB_EW_PROLOGUE(n_frames, joints, weights, out);
    pmb::center_of_mass_kernel<<<ew_grid(3 * n_frames, 256, dp_), 256, 0, st_>>>(joints, weights, weights_frame_stride, out,
                                                                              n_frames, n_joints);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_interpolate_positions_f32(const double *sample_times, const double *original_times, const float *positions,
                                  int64_t outer, int64_t n_original, int64_t n_samples, int64_t inner, float *out,
                                  int32_t *idx_workspace, float *weight_workspace, void *stream) {
    if (n_original < 2) return fail(PMB_ERR_SHAPE, "%s: at least two original times are needed", __func__);
    if (outer < 0 || inner < 0 || n_samples < 0) return fail(PMB_ERR_SHAPE, "%s: negative size", __func__);
    if (n_original > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "%s: n_original must be below 2^31", __func__);
    if (outer == 0 || inner == 0 || n_samples == 0) return PMB_OK;
    PMB_EW_PROLOGUE(n_samples, sample_times, original_times, positions, out, idx_workspace, weight_workspace);
    pmb::interp_coeff_kernel<<<grid_, 256, 0, st_>>>(sample_times, original_times, n_samples, n_original, idx_workspace,
                                                     weight_workspace);
    PMB_CUDA(cudaGetLastError());
    pmb::interp_apply_kernel<<<ew_grid(outer * n_samples * inner, 256, dp_), 256, 0, st_>>>(
        positions, idx_workspace, weight_workspace, out, outer, n_original, n_samples, inner);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_vec_normalize_f32(const float *v, float eps, float *out, int64_t n, int32_t k, void *stream) {
    if (k < 1) return fail(PMB_ERR_SHAPE, "%s: k < 1", __func__);
    PMB_EW_PROLOGUE(n, v, out);
    if (k == 3 && aligned16(v) && aligned16(out) && n >= 4 && knob(K_VEC3_X4, 1)) {
        const int64_t n4 = n / 4;
        pmb::vec3_normalize_x4_kernel<<<ew_grid(n4, 256, dp_), 256, 0, st_>>>((const float4 *)v, eps, (float4 *)out, n4);
        PMB_CUDA(cudaGetLastError());
        if (n % 4) pmb::vec_normalize_kernel<<<1, 32, 0, st_>>>(v + 12 * n4, eps, out + 12 * n4, n % 4, 3);
    } else {
        pmb::vec_normalize_kernel<<<grid_, 256, 0, st_>>>(v, eps, out, n, k);
    }
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- fk consumers: from_root_positions, mirror (ik_kernels.cuh) ------------------------------
int pmb_from_root_positions_f32(const float *positions, const int64_t *parents_host, const float *offsets,
                                int64_t n_frames, int32_t n_joints, float *rotations, void *stream) {
    if (!positions || !offsets || !rotations) return fail(PMB_ERR_NULL, "from_root_positions: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    PMB_NEED16(rotations);
    const pmb::JointProgram *prog_p = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, false, prog_p, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    // children in index order, like the reference's `children` lists (skeleton.py:122-126)
    pmb::ChildTable kids;
    {
        int count[PMB_MAX_JOINTS + 1] = {0};
        for (int i = 1; i < n_joints; ++i) ++count[parents_host[i]];
        kids.start[0] = 0;
        for (int j = 0; j < n_joints; ++j) kids.start[j + 1] = static_cast<uint16_t>(kids.start[j] + count[j]);
        int fill[PMB_MAX_JOINTS] = {0};
        for (int i = 1; i < n_joints; ++i) {
            const int p = static_cast<int>(parents_host[i]);
            kids.child[kids.start[p] + fill[p]++] = static_cast<uint16_t>(i);
        }
    }
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    const bool fast = knob(K_FRP_FAST, 1) != 0;
    constexpr int THREADS = 128;
    const int smem = n_joints * 16 + n_slots * THREADS * 16;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", n_slots);
    // aligned 16- / 8-byte point loads only where every lane of a warp takes the same case (J a multiple of 4: +5 % at 4M x 52);
    // with per-lane cases the divergence costs more than the sector look-ups save (1M x 22: 0.182 -> 0.208 ms, 4M x 65: 3.06 -> 3.86)
    const bool wide = knob(K_FRP_WIDE, n_joints % 4 == 0 ? 1 : 0) != 0;
    auto kernel = fast ? (wide ? pmb::from_root_positions_kernel<THREADS, true, true> : pmb::from_root_positions_kernel<THREADS, true, false>)
                       : pmb::from_root_positions_kernel<THREADS, false, false>;
    int per_sm_unused = 0;
    if ((rc = kernel_fit(kernel, dp, THREADS, smem, per_sm_unused))) return rc;
    // Thread = frame reads its positions row 12 bytes at a time, so the op lives on L1 hits, and L1 is what the
    // shared-memory carve-out leaves of the SM's 256 KB.  Left to the driver, the carve-out is sized for the nine
    // blocks the registers allow: with many live branch slots (2 KB each per block) that takes nearly all of it
    // (measured at 4M x 65, 10 slots: 46 % L1 hits, 7.5 x the input re-read from L2, 10.4 ms).  Ask for what six
    // blocks need instead; the occupancy follows the carve-out.  Measured (profiles/r1_sweep_frp_carveout.jsonl):
    // 4M x 65 10.42 -> 4.79 ms (4.95 for 3 .. 4 blocks, 8.2 for 2), 4M x 52 3.84 -> 3.66 ms, 1M x 22 0.366 -> 0.360 ms.
    {
        const int target = std::max(1, std::min(9, knob(K_FRP_BLOCKS_PER_SM, 6)));
        const int want = target * (smem + 1024);
        const int pct = std::max(1, std::min(100, (want * 100 + dp.smem_optin - 1) / dp.smem_optin));
        PMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        note_variant("from_root_positions_kernel<%d,FAST=%d,WIDE=%d> smem=%d carveout=%d%% (for %d blocks/SM)", THREADS, int(fast), int(fast && wide), smem, pct, target);
    }
    const long long blocks = (n_frames + THREADS - 1) / THREADS;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    kernel<<<static_cast<unsigned>(blocks), THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        positions, offsets, reinterpret_cast<float4 *>(rotations), n_frames, n_joints, n_slots, *prog_p, kids);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_mirror_to_local_f32(const float *global_quats, const int64_t *parents_host, const int64_t *joints_mapping_host,
                            int32_t mirror_axis, int64_t n_frames, int32_t n_joints, float *local_quats, void *stream) {
    if (!global_quats || !local_quats) return fail(PMB_ERR_NULL, "mirror_to_local: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (mirror_axis < 0 || mirror_axis > 2) return fail(PMB_ERR_SHAPE, "mirror_axis must be 0 (X), 1 (Y) or 2 (Z)");
    PMB_NEED16(global_quats); PMB_NEED16(local_quats);
    const pmb::JointProgram *prog_p = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, false, prog_p, n_slots);
    if (rc) return rc;
    pmb::JointMap jm;
    for (int j = 0; j < n_joints; ++j) {
        const int64_t m = joints_mapping_host ? joints_mapping_host[j] : j;
        if (m < 0 || m >= n_joints) return fail(PMB_ERR_SHAPE, "joints_mapping[%d] = %lld outside [0, %d)", j, static_cast<long long>(m), n_joints);
        jm.map[j] = static_cast<uint16_t>(m);
    }
    if (n_frames == 0) return PMB_OK;
    constexpr int THREADS = 256;
    const int fb = tile_frames(n_joints, 4096);
    const long long blocks = (n_frames + fb - 1) / fb;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    const uint32_t magic = magic_small(n_joints);
    // the two vector components that change sign (skeleton.py:307-315): X -> (y, z), Y -> (x, z), Z -> (x, y)
    const float fx = mirror_axis == 0 ? 1.f : -1.f, fy = mirror_axis == 1 ? 1.f : -1.f, fz = mirror_axis == 2 ? 1.f : -1.f;
    pmb::mirror_to_local_kernel<THREADS><<<static_cast<unsigned>(blocks), THREADS, (4 * n_joints + 15) & ~15, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(global_quats), reinterpret_cast<float4 *>(local_quats), n_frames, n_joints, fb,
        magic, fx, fy, fz, *prog_p, jm);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_vec_mirror_f32(const float *v, int32_t axis, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, v, out);
    if (axis < 0 || axis > 2) return fail(PMB_ERR_SHAPE, "%s: axis must be 0, 1 or 2", __func__);
    pmb::vec_mirror_kernel<<<ew_grid(3 * n, 256, dp_), 256, 0, st_>>>(v, out, axis, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_root_center_f32(const float *positions, float *out, int64_t n_frames, int32_t n_joints, void *stream) {
    if (n_joints < 1) return fail(PMB_ERR_SHAPE, "%s: n_joints < 1", __func__);
    PMB_EW_PROLOGUE(n_frames, positions, out);
    pmb::root_center_kernel<<<ew_grid(n_frames * n_joints * 3, 256, dp_), 256, 0, st_>>>(positions, out, n_frames, n_joints);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

}  // extern "C"
