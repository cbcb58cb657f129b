/*
 * pymotion_b200 -- C ABI of the B200-native (sm_100a) forward-kinematics /
 * root dual-quaternion engine.  Drop-in boundary for the hot path of
 * UPC-ViRVIG/pymotion v0.2.3:
 *
 *   pymotion/ops/skeleton.py        fk :16, from_root_dual_quat :173, to_root_dual_quat :207,
 *                                   from_global_rotations :64, from_root_positions :96, mirror :247
 *   pymotion/rotations/quat.py      mul :337, mul_vec :320, length :364, inverse :379,
 *                                   conjugate :396, normalize :411, to_matrix :276, from_matrix :85
 *   pymotion/rotations/dual_quat.py from_rotation_translation :12, from_translation :39,
 *                                   to_rotation_translation :62
 * and, off the hot path, the rest of those two modules: quat from/to angle-axis,
 * scaled angle-axis, Euler; unroll, slerp, from_to, from_to_axis; dual_quat normalize, is_unit, unroll.
 *
 * The reference has no FFI layer of its own (it is pure NumPy / PyTorch); these
 * are the entry points a ctypes / cffi stub inside those modules would bind
 * (INTEGRATION.md shows the stub).  Conventions:
 *
 *  - plain pointers and sizes only; no torch / CUDA types in any signature;
 *  - every array pointer is a DEVICE pointer on the current CUDA device unless
 *    the parameter name ends in `_host`;
 *  - arrays are dense row-major float32: quaternion = 4 floats (w,x,y,z),
 *    dual quaternion = 8 floats (real wxyz | dual wxyz), rotation matrix = 9
 *    floats m[r][c], vector = 3 floats.  There are no float64 entry points (north_star: fp32,
 *    HBM-bound); the Python layer rounds float64 callers to float32 with a one-time PrecisionWarning;
 *  - the caller owns every buffer; the library never allocates user-visible
 *    memory, never frees and never writes an input;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    calls are asynchronous on it and never synchronise;
 *  - return value 0 = PMB_OK, negative = pmb_status; pmb_last_error() returns a
 *    thread-local message for the last failing call on this thread;
 *  - `parents_host` is a HOST array of n_joints int64.  parents[0] is ignored
 *    (0 and -1 both accepted, like the reference); parents[i] must satisfy
 *    0 <= parents[i] < i for i >= 1 (BVH depth-first order, io/bvh.py:77-85),
 *    otherwise PMB_ERR_TOPOLOGY (the reference silently returns a non-FK result
 *    for such tables; see DESIGN.md "deliberate deviations");
 *  - a `*_frame_stride` is the distance in ELEMENTS between consecutive frames
 *    of that operand; 0 means one row shared by every frame (NumPy broadcasting).
 */
#ifndef PYMOTION_B200_H_
#define PYMOTION_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMB_VERSION 100          /* 0.1.0 */
#define PMB_MAX_JOINTS 512       /* joint programs travel as kernel parameters */

typedef enum pmb_status {
    PMB_OK = 0,
    PMB_ERR_NULL = -1,           /* a required pointer is NULL */
    PMB_ERR_SHAPE = -2,          /* n_frames < 0, n_joints < 1 or > PMB_MAX_JOINTS, bad stride */
    PMB_ERR_ALIGN = -3,          /* quaternion / dual-quaternion pointer not 16-byte aligned */
    PMB_ERR_TOPOLOGY = -4,       /* parents[i] outside [0, i) for some i >= 1 */
    PMB_ERR_CUDA = -5,           /* CUDA runtime / launch error (message has the cudaError string) */
    PMB_ERR_ROOT_OFFSET = -6     /* to_root_dual_quat: offsets[0] != 0 (AssertionError in the reference, skeleton.py:227) */
} pmb_status;

int pmb_version(void);
const char *pmb_last_error(void);
const char *pmb_status_string(int status);
/* Diagnostics: which kernel variant (template arguments, grid, shared memory) the last fk /
 * to_root_dual_quat launch on this thread picked.  Used by bench.py and the variant tests. */
const char *pmb_last_variant(void);

/* Number of SMs / name of the current device (diagnostics, used by bench.py). */
int pmb_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len);

/* ---- skeleton ops ------------------------------------------------------- */

/* ops/skeleton.py:16-61  fk(rot, global_pos, offsets, parents) -> (positions, rotmats)
 *   rot          [n_frames][n_joints][4]   local rotations, normalised inside (quat.py:411)
 *   global_pos   [n_frames][3]             (gpos_frame_stride = 3) or one shared row (0)
 *   offsets      [n_joints][3]             (offsets_frame_stride = 0) or per frame (3*n_joints)
 *   positions    [n_frames][n_joints][3]   out
 *   rotmats      [n_frames][n_joints][3][3] out
 * offsets[0] is ignored: the root translation is global_pos (skeleton.py:49). */
int pmb_fk_f32(const float *rot, const float *global_pos, int64_t gpos_frame_stride,
               const float *offsets, int64_t offsets_frame_stride, const int64_t *parents_host,
               int64_t n_frames, int32_t n_joints, float *positions, float *rotmats, void *stream);

/* Variant that emits global QUATERNIONS instead of matrices (SURVEY 8f rank 1:
 * fk -> quat.from_matrix fused away; 44J instead of 64J bytes per pose).
 *   global_rots  [n_frames][n_joints][4]   out; equals quat.from_matrix(rotmats) up to sign
 * With shared offsets, `positions` may be NULL: rotations only (what mirror needs; no translation chain). */
int pmb_fk_quat_f32(const float *rot, const float *global_pos, int64_t gpos_frame_stride,
                    const float *offsets, int64_t offsets_frame_stride, const int64_t *parents_host,
                    int64_t n_frames, int32_t n_joints, float *positions, float *global_rots, void *stream);

/* ops/skeleton.py:207-244  to_root_dual_quat(rotations, global_pos, parents, offsets) -> dq
 *   offsets      [n_joints][3] shared; offsets_host0 = its first row read by the caller
 *                (host, 3 floats) so the reference's `offsets[0] == 0` assert can be
 *                enforced without a device sync; pass NULL to skip the check.
 *   dq           [n_frames][n_joints][8]   out */
int pmb_to_root_dual_quat_f32(const float *rotations, const float *global_pos, int64_t gpos_frame_stride,
                              const int64_t *parents_host, const float *offsets, const float *offsets_host0,
                              int64_t n_frames, int32_t n_joints, float *dq, void *stream);

/* ops/skeleton.py:173-204  from_root_dual_quat(dq, parents) -> (translations, rotations) */
int pmb_from_root_dual_quat_f32(const float *dq, const int64_t *parents_host, int64_t n_frames,
                                int32_t n_joints, float *translations, float *rotations, void *stream);

/* ops/skeleton.py:64-93  from_global_rotations(global_quats, parents) -> local_quats */
int pmb_from_global_rotations_f32(const float *global_quats, const int64_t *parents_host, int64_t n_frames,
                                  int32_t n_joints, float *local_quats, void *stream);

/* ---- host-buffer (end-to-end) entry points ------------------------------- */

/* fk on HOST buffers (the reference's own calling convention, ops/skeleton.py:16: arrays in
 * host memory in, arrays in host memory out).  The frame axis is cut into chunks of
 * `chunk_frames` (0 = library default, ~48 MB of traffic per chunk) and H2D copy / kernel /
 * D2H copy are pipelined over two internal streams.  Page-locked buffers are DMA'd directly;
 * pageable buffers (plain malloc / NumPy memory) go through a page-locked staging ring filled
 * and drained by a few library-owned host threads.  Workspaces (device + pinned staging) are
 * per device and per concurrent caller, owned by the library, freed by pmb_release_workspace().
 * Blocks until the outputs are in host memory; on error no copy is left in flight.
 * pmb_fk_quat_f32_host returns global quaternions (16 J instead of 36 J bytes per frame back
 * over PCIe) -- what every in-repo consumer of fk computes next (ops/skeleton.py:322, :140). */
int pmb_fk_f32_host(const float *rot_host, const float *global_pos_host, const float *offsets_host,
                    const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                    float *positions_host, float *rotmats_host, int64_t chunk_frames);
int pmb_fk_quat_f32_host(const float *rot_host, const float *global_pos_host, const float *offsets_host,
                         const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                         float *positions_host, float *global_rots_host, int64_t chunk_frames);

/* The dual-quaternion pair on HOST arrays, same pipeline: ops/skeleton.py:207 `to_root_dual_quat(rotations, global_pos, parents,
 * offsets) -> dq` and :173 `from_root_dual_quat(dq, parents) -> (translations, rotations)` called the reference's way, NumPy in /
 * NumPy out.  offsets_host[0] must be zero (PMB_ERR_ROOT_OFFSET, the reference's assert at :227). */
int pmb_to_root_dual_quat_f32_host(const float *rotations_host, const float *global_pos_host, const int64_t *parents_host,
                                   const float *offsets_host, int64_t n_frames, int32_t n_joints, float *dq_host, int64_t chunk_frames);
int pmb_from_root_dual_quat_f32_host(const float *dq_host, const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                                     float *translations_host, float *rotations_host, int64_t chunk_frames);
void pmb_release_workspace(void);

/* ---- element-wise quaternion primitives (n = number of quaternions) ----- */
int pmb_quat_mul_f32(const float *q0, const float *q1, float *out, int64_t n, void *stream);          /* quat.py:337 */
int pmb_quat_mul_vec_f32(const float *q, const float *v, float *out, int64_t n, void *stream);        /* quat.py:320 */
int pmb_quat_length_f32(const float *q, float *out, int64_t n, void *stream);                         /* quat.py:364 */
int pmb_quat_normalize_f32(const float *q, float eps, float *out, int64_t n, void *stream);           /* quat.py:411 */
int pmb_quat_conjugate_f32(const float *q, float *out, int64_t n, void *stream);                      /* quat.py:396 (= inverse :379) */
int pmb_quat_to_matrix_f32(const float *q, float *out, int64_t n, void *stream);                      /* quat.py:276 */
int pmb_quat_from_matrix_f32(const float *m, float *out, int64_t n, void *stream);                    /* quat.py:85 */

/* ---- element-wise dual-quaternion primitives ---------------------------- */
int pmb_dq_from_rotation_translation_f32(const float *rotations, const float *translations, float *dq,
                                         int64_t n, void *stream);                                    /* dual_quat.py:12 */
int pmb_dq_from_translation_f32(const float *translations, float *dq, int64_t n, void *stream);       /* dual_quat.py:39 */
int pmb_dq_to_rotation_translation_f32(const float *dq, float *rotations, float *translations,
                                       int64_t n, void *stream);                                      /* dual_quat.py:62 */

/* ---- the rest of the quat / dual_quat surface -------- */
/* n = number of elements; angle [n], axis / vectors [n][3], euler [n][3], quaternions [n][4]. */
int pmb_quat_from_angle_axis_f32(const float *angle, const float *axis, float *out, int64_t n, void *stream);      /* quat.py:24 */
int pmb_quat_from_scaled_angle_axis_f32(const float *scaled_axis, float *out, int64_t n, void *stream);           /* quat.py:6 */
/* Euler orders ('x'|'y'|'z' per slot, quat.py:43-82, :159-227) travel as ONE byte per element,
 * o0 + 3*o1 + 9*o2 with o = 0|1|2 for x|y|z (device array); order_stride 1 = per element, 0 = one shared order. */
int pmb_quat_from_euler_f32(const float *euler, const uint8_t *order_codes, int64_t order_stride, float *out,
                            int64_t n, void *stream);                                                             /* quat.py:43 */
int pmb_quat_to_euler_f32(const float *q, const uint8_t *order_codes, int64_t order_stride, float *out,
                          int64_t n, void *stream);                                                               /* quat.py:159 */
int pmb_quat_to_angle_axis_f32(const float *q, float *angle, float *axis, int64_t n, void *stream);               /* quat.py:247 */
int pmb_quat_to_scaled_angle_axis_f32(const float *q, float *out, int64_t n, void *stream);                       /* quat.py:230 */
/* t: device array, t_stride 1 = one value per element, 0 = one shared value.  shortest != 0: flip q1 when q0.q1 < 0. */
int pmb_quat_slerp_f32(const float *q0, const float *q1, const float *t, int64_t t_stride, int32_t shortest,
                       float *out, int64_t n, void *stream);                                                      /* quat.py:465 */
int pmb_quat_from_to_f32(const float *v1, const float *v2, int32_t normalize_input, float *out, int64_t n,
                         void *stream);                                                                           /* quat.py:504 */
int pmb_quat_from_to_axis_f32(const float *v1, const float *v2, const float *rot_axis, int32_t normalize_input,
                              float *out, int64_t n, void *stream);                                               /* quat.py:579 */
/* quat.unroll (quat.py:426, width 4) / dual_quat.unroll (dual_quat.py:139, width 8) along the LEADING axis of
 * x [n_steps][n_cols][width]: entry t is negated iff its dot product (first 4 numbers) with the already
 * unrolled entry t-1 is < 0.  Chunked scan; `workspace` is a device scratch buffer of at least
 * pmb_unroll_workspace_bytes(n_steps, n_cols) bytes owned by the caller.  out may not alias x. */
int64_t pmb_unroll_workspace_bytes(int64_t n_steps, int64_t n_cols);
int pmb_unroll_f32(const float *x, int32_t width, int64_t n_steps, int64_t n_cols, float *out, void *workspace,
                   int64_t workspace_bytes, void *stream);
/* dual_quat.is_unit (dual_quat.py:118): ONE verdict for the whole array, returned as three device flags the
 * caller reads after synchronising: flags3[0] = some real part is not ~0, [1] = some |real|^2 is not ~1,
 * [2] = some |real.dual| > atol;  unit  <=>  !flags3[0] || (!flags3[1] && !flags3[2]). */
int pmb_dq_is_unit_f32(const float *dq, float atol, int64_t n, int32_t *flags3, void *stream);
/* dual_quat.normalize (dual_quat.py:86): divide by |real|; if the WHOLE scaled array is not unit, project the
 * real direction out of every dual part.  flags3: device scratch (3 int32), no host synchronisation. */
int pmb_dq_normalize_f32(const float *dq, float *out, int64_t n, int32_t *flags3, void *stream);

/* ---- fk consumers ------------------------------------ */
/* ops/skeleton.py:96-170  from_root_positions(positions, parents, offsets) -> rotations
 *   positions [n_frames][n_joints][3] root-centred joint positions, offsets [n_joints][3] (device),
 *   rotations [n_frames][n_joints][4] out.  One tree walk per frame instead of the reference's O(J) fk passes. */
int pmb_from_root_positions_f32(const float *positions, const int64_t *parents_host, const float *offsets,
                                int64_t n_frames, int32_t n_joints, float *rotations, void *stream);
/* Second half of mirror / _true_mirror (ops/skeleton.py:324-331, :410-416): global quaternions re-indexed by
 * joints_mapping_host (NULL = identity; 'symmetry' mode), the two vector components for `mirror_axis`
 * (0 = X, 1 = Y, 2 = Z) negated, then back to local space as from_global_rotations does. */
int pmb_mirror_to_local_f32(const float *global_quats, const int64_t *parents_host, const int64_t *joints_mapping_host,
                            int32_t mirror_axis, int64_t n_frames, int32_t n_joints, float *local_quats, void *stream);

/* The whole device side of mirror's rotation step in ONE call: fk (rotations only, root at the origin) -> sign convention of
 * quat.from_matrix -> re-index by joints_mapping (NULL = identity), flip two vector components -> back to local space
 * (ops/skeleton.py:322-331 `mirror`, :410-416 `_true_mirror`).  Equivalent to pmb_fk_quat_f32 (positions = NULL) followed by
 * pmb_mirror_to_local_f32, but where the quaternion track kernel applies the global quaternions never leave the SM (32 J
 * instead of 64 J bytes per pose).  global_quats_scratch [n_frames][n_joints][4] is only used by skeletons that take the
 * two-kernel path (pmb_mirror_local_needs_scratch returns 1: sparse level schedules, very large skeletons) and may be NULL
 * otherwise.  local_quats is not modified. */
int pmb_mirror_local_needs_scratch(const int64_t *parents_host, int32_t n_joints);
int pmb_mirror_local_f32(const float *local_quats, const int64_t *parents_host, const int64_t *joints_mapping_host, int32_t mirror_axis,
                         int64_t n_frames, int32_t n_joints, float *global_quats_scratch, float *mirrored_local_quats, void *stream);
/* out = v with component `axis` negated, v [n][3] (translations / offsets / end sites, skeleton.py:405-408). */
int pmb_vec_mirror_f32(const float *v, int32_t axis, float *out, int64_t n, void *stream);
/* out[f][j] = positions[f][j] - positions[f][0] (mirror mode 'positions', skeleton.py:338). */
int pmb_root_center_f32(const float *positions, float *out, int64_t n_frames, int32_t n_joints, void *stream);

/* ---- the numeric modules either side of fk (SURVEY 8f rank 3 tail, rank 4) -------------------- */
/* rotations/ortho6d.py: ortho6d [n][3][2] = the first two columns of the rotation matrix (8-byte aligned). */
int pmb_ortho6d_from_matrix_f32(const float *rotmats, float *ortho6d, int64_t n, void *stream);   /* ortho6d.py:31 */
int pmb_ortho6d_from_quat_f32(const float *q, float *ortho6d, int64_t n, void *stream);           /* ortho6d.py:14 */
int pmb_ortho6d_to_matrix_f32(const float *ortho6d, float *rotmats, int64_t n, void *stream);     /* ortho6d.py:67 */
int pmb_ortho6d_to_quat_f32(const float *ortho6d, float *q, int64_t n, void *stream);             /* ortho6d.py:50 */
/* ops/center_of_mass.py:52  out[f][:] = sum_j joints[f][j][:] * weights[j];  weights_frame_stride 0 = one row of
 * weights shared by every frame, n_joints = weights per frame. */
int pmb_center_of_mass_f32(const float *joints, const float *weights, int64_t weights_frame_stride, int64_t n_frames,
                           int32_t n_joints, float *out, void *stream);
/* ops/time.py:4  linear interpolation along the time axis of positions [outer][n_original][inner] ->
 * out [outer][n_samples][inner].  Times are DEVICE float64 arrays (sorted original_times); idx_workspace
 * (n_samples int32) and weight_workspace (n_samples float) are device scratch owned by the caller. */
int pmb_interpolate_positions_f32(const double *sample_times, const double *original_times, const float *positions,
                                  int64_t outer, int64_t n_original, int64_t n_samples, int64_t inner, float *out,
                                  int32_t *idx_workspace, float *weight_workspace, void *stream);
/* ops/vector.py:4  v / (|v| + eps) over rows of length k. */
int pmb_vec_normalize_f32(const float *v, float eps, float *out, int64_t n, int32_t k, void *stream);

/* ---- the numeric part of BVH.get_data (io/bvh.py:332-365) ----------------- */

/* rots = quat.normalize(quat.unroll(quat.from_euler(np.radians(rotations), rot_order tiled over
 * the frames), axis=0))  (io/bvh.py:352-359) as ONE fused three-pass scan: the un-unrolled
 * quaternions are never stored.
 *   euler_deg    [n_frames][n_joints][3]   Euler angles in DEGREES, as read from the file
 *   order_codes  [n_joints]                one byte per joint, o0 + 3*o1 + 9*o2 with 0|1|2 = 'x'|'y'|'z'
 *   out          [n_frames][n_joints][4]   unit quaternions, sign-continuous along the frame axis
 *   workspace    pmb_unroll_workspace_bytes(n_frames, n_joints) bytes of device memory
 * BVH text parsing stays on the host (out of scope). */
int pmb_bvh_rotations_to_quat_f32(const float *euler_deg, const uint8_t *order_codes, int64_t n_frames,
                                  int64_t n_joints, float *out, void *workspace, int64_t workspace_bytes,
                                  void *stream);

/* ---- introspection of the host-side joint program (tests) ---- */

/* Builds the per-joint program the chain kernels execute for `parents_host`:
 * codes_out[i] bits 0-7 = slot the parent transform is fetched from (0xFF: the
 * previous joint's registers), bits 8-15 = slot this joint is saved to (0xFF:
 * none), bits 16-30 = parent index.  Returns the number of slots (>= 0) or a
 * negative pmb_status. */
int pmb_build_joint_program(const int64_t *parents_host, int32_t n_joints, uint32_t *codes_out);

/* Builds the schedule the fk track kernel executes for `parents_host` with `n_tracks` (1..8)
 * independent joints per step: codes_out[step * n_tracks + track], bits 0-9 joint, bits 10-19
 * parent, bit 20 = the parent is the same track's previous item (kept in registers), bit 21 =
 * no-op.  Every joint appears exactly once, after its parent's step.  `window` = 8: joints are
 * scheduled box by box (8 consecutive joints, what the kernel holds of the input at a time) and
 * window_first_out (n_windows + 1 entries, may be NULL) receives the first step of every window
 * and the total; `window` = 0: the whole skeleton at once (the optimum for n_tracks machines).
 * Returns the number of steps (> 0) or a negative pmb_status (PMB_ERR_SHAPE if codes_capacity
 * is too small). */
int pmb_build_track_schedule(const int64_t *parents_host, int32_t n_joints, int32_t n_tracks, int32_t window,
                             uint32_t *codes_out, int32_t codes_capacity, int32_t *window_first_out);

/* Kernel-variant knobs (PMB_FK_*, PMB_DQ_*, ...) are honoured only when PMB_EXPERIMENT=1 is in
 * the environment, and are read once; this re-reads them (sweeps and the forced-variant tests
 * switch variants inside one process).  Not for production use, not thread-safe against
 * concurrent launches. */
void pmb_reload_knobs(void);

#ifdef __cplusplus
}
#endif
#endif /* PYMOTION_B200_H_ */
