// The fk consumers of the reference:
//   ops/skeleton.py  from_root_positions :96-170   positions -> local rotations (analytic two-bone alignment)
//                    mirror :247-344, _true_mirror :347-418   (their device-side pieces)
#pragma once
#include "common.cuh"
#include "dq_kernels.cuh"
#include "rotations_ext.cuh"

namespace pmb {

// Children of every joint in CSR form (index order), built on the host from parents[].
struct ChildTable {
    uint16_t start[PMB_MAX_JOINTS + 1];
    uint16_t child[PMB_MAX_JOINTS];
};

// ---------------------------------------------------------------------------------------------------
// from_root_positions (ops/skeleton.py:96-170).  The reference sets the joints one after the other and
// re-runs a whole fk before each alignment (one pass per joint with children plus one per extra child,
// O(J) fk passes).  Everything an alignment of joint j reads from those passes is (a) the global rotation of
// j under the rotations chosen so far -- the parent's final global rotation times j's current local
// rotation -- and (b) rest directions pos[c] - pos[j] rotated back into j's frame, which are the offsets of
// the children.  So ONE walk down the tree per frame does the same job: thread = frame, the parent's global
// quaternion travels in registers / shared-memory slots exactly like the other chain kernels.
//   first child c (:137-146):  rot_j = from_to(off_c, G^-1 (P_c - P_j)),   G = global rotation of j's parent
//   every further child g (:148-166), with G_j = G (x) rot_j:
//       roll = from_to_axis(off_g, G_j^-1 (P_g - P_j), G_j^-1 normalize(P_c - P_j));   rot_j = rot_j (x) roll
// Joints without children keep the identity.  Global rotations are composed from NORMALISED local rotations,
// as the reference's fk does (from_to / from_to_axis results are unit only up to their eps terms).
// Conditioning: the further children of a joint are almost aligned once the first one is, so the roll corrections
// are SMALL angles: the half-angle sine comes from |a x b| / (2 w) (q_from_to_axis_stable), not from the reference's
// sqrt((1 - dot) / 2), which cancels in fp32.  What remains against the float64 reference are its np.isclose snaps to
// the identity (|angle| < 4.5e-3: a frame on the other side of the threshold differs by up to 2.2e-3) and the sign of
// a roll whose axis is perpendicular to the correction; the reference's own float32 torch twin differs from its
// NumPy path by up to 2.3e-3 for the same reasons.
// ---------------------------------------------------------------------------------------------------
// (A tile-staged variant -- one warp per block, positions and rotations through shared memory with coalesced
// 16-byte accesses -- was measured SLOWER: 0.54 ms against 0.36 ms at 1M x 22.  The op is bound by the IEEE
// square roots and divisions of from_to / from_to_axis, not by its strided global accesses.)
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
from_root_positions_kernel(const float *__restrict__ pos, const float *__restrict__ offsets, float4 *__restrict__ rots,
                           long long n_frames, int n_joints, int n_slots, const __grid_constant__ JointProgram prog,
                           const __grid_constant__ ChildTable kids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tab = reinterpret_cast<float4 *>(smem_raw);                 // [J] rest directions: offsets / (|offsets| + 1e-8)
    float4 *slots = tab + n_joints;                                      // [n_slots][THREADS] global quaternions
    for (int j = threadIdx.x; j < n_joints; j += THREADS) {
        // from_to / from_to_axis normalise their first argument (quat.py:541, :616): the same value for every frame
        const Vec3<float> d = v_normalize(Vec3<float>{offsets[3 * j], offsets[3 * j + 1], offsets[3 * j + 2]}, 1e-8f);
        tab[j] = make_float4(d.x, d.y, d.z, 0.f);
    }
    __syncthreads();
    const long long f = blockIdx.x * static_cast<long long>(THREADS) + threadIdx.x;
    if (f >= n_frames) return;
    const float *P = pos + f * n_joints * 3;
    float4 *R = rots + f * n_joints;
    auto point = [&](int j) { return Vec3<float>{__ldg(P + 3 * j), __ldg(P + 3 * j + 1), __ldg(P + 3 * j + 2)}; };

    Quat<float> cur{1.f, 0.f, 0.f, 0.f};  // global rotation of the previous joint
    for (int j = 0; j < n_joints; ++j) {
        const uint32_t code = prog.code[j];
        Quat<float> G{1.f, 0.f, 0.f, 0.f};  // global rotation of j's parent (the root's parent is the world)
        if (j > 0) {
            const uint32_t src = prog_src(code);
            if (src == kSrcReg) {
                G = cur;
            } else {
                const float4 s = slots[src * THREADS + threadIdx.x];
                G = {s.x, s.y, s.z, s.w};
            }
        }
        Quat<float> rot{1.f, 0.f, 0.f, 0.f};
        const int k0 = kids.start[j], k1 = kids.start[j + 1];
        if (k1 > k0) {
            const Vec3<float> pj = point(j);
            const int c = kids.child[k0];
            const Vec3<float> pc = point(c);
            const Vec3<float> to_c{pc.x - pj.x, pc.y - pj.y, pc.z - pj.z};
            const float4 oc = tab[c];
            rot = q_from_to_stable(Vec3<float>{oc.x, oc.y, oc.z}, q_rotate(q_conj(G), to_c));
            for (int k = k0 + 1; k < k1; ++k) {
                const int g = kids.child[k];
                const Quat<float> inv = q_conj(q_mul(G, q_normalize(rot, 1e-8f)));  // fk normalises local rotations (quat.py:411)
                const Vec3<float> pg = point(g);
                const float4 og = tab[g];
                const Vec3<float> pred = q_rotate(inv, Vec3<float>{pg.x - pj.x, pg.y - pj.y, pg.z - pj.z});
                const Vec3<float> axis = q_rotate(inv, v_normalize(to_c, 1e-8f));
                rot = q_mul(rot, q_from_to_axis_stable(Vec3<float>{og.x, og.y, og.z}, pred, axis));
            }
        }
        R[j] = make_float4(rot.w, rot.x, rot.y, rot.z);
        cur = q_mul(G, q_normalize(rot, 1e-8f));
        const uint32_t sv = prog_save(code);
        if (sv != kNoSave) slots[sv * THREADS + threadIdx.x] = make_float4(cur.w, cur.x, cur.y, cur.z);
    }
}

// ---------------------------------------------------------------------------------------------------
// mirror / _true_mirror, second half (ops/skeleton.py:324-331, :410-416): global quaternions are re-indexed by
// the joints mapping ('symmetry' mode; identity otherwise), two vector components change sign (which two
// depends on the mirror axis: X -> (y, z), Y -> (x, z), Z -> (x, y)), and the result goes back to local space
// like from_global_rotations (:64-93).  Thread per (frame, joint).
// ---------------------------------------------------------------------------------------------------
struct JointMap {
    uint16_t map[PMB_MAX_JOINTS];
};
__device__ __forceinline__ Quat<float> q_flip(const float4 &a, float fx, float fy, float fz) {
    return {a.x, fx * a.y, fy * a.z, fz * a.w};
}
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
mirror_to_local_kernel(const float4 *__restrict__ gq, float4 *__restrict__ lq, long long n_frames, int n_joints,
                       int frames_per_block, uint32_t magic, float fx, float fy, float fz,
                       const __grid_constant__ JointProgram prog, const __grid_constant__ JointMap jm) {
    // per-thread joint indices: the tables go to shared memory first (a divergent index into the constant bank
    // serialises: measured 2.3 TB/s against 6.4 TB/s for from_global_rotations, which stages its table)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    short *src = reinterpret_cast<short *>(smem_raw);   // [J] source joint of joint j
    short *psrc = src + n_joints;                        // [J] source joint of j's parent
    for (int j = threadIdx.x; j < n_joints; j += THREADS) {
        src[j] = static_cast<short>(jm.map[j]);
        psrc[j] = static_cast<short>(jm.map[prog_parent(prog.code[j])]);
    }
    __syncthreads();
    const long long fbase = static_cast<long long>(blockIdx.x) * frames_per_block;
    const int nf = static_cast<int>(min(static_cast<long long>(frames_per_block), n_frames - fbase));
    const int n_el = nf * n_joints;
    const float4 *gt = gq + fbase * n_joints;
    float4 *lt = lq + fbase * n_joints;
    for (int i = threadIdx.x; i < n_el; i += THREADS) {
        const int fl = div_small(i, magic);
        const int j = i - fl * n_joints;
        const int row = i - j;
        Quat<float> r = q_flip(__ldg(gt + row + src[j]), fx, fy, fz);
        if (j > 0) r = q_mul(q_conj(q_flip(__ldg(gt + row + psrc[j]), fx, fy, fz)), r);
        __stcs(lt + i, make_float4(r.w, r.x, r.y, r.z));
    }
}

// v[..., axis] = -v[..., axis] for [n][3] vectors (global translations, offsets, end sites)
__global__ void vec_mirror_kernel(const float *v, float *o, int axis, long long n) {
    PMB_GRID_STRIDE(i, 3 * n) {
        const float x = __ldcs(v + i);
        o[i] = (i % 3 == axis) ? -x : x;
    }
}
// p[f][j] - p[f][0]  (mirror 'positions' mode, ops/skeleton.py:338)
__global__ void root_center_kernel(const float *p, float *o, long long n_frames, int n_joints) {
    const long long n = n_frames * n_joints * 3;
    PMB_GRID_STRIDE(i, n) {
        const long long f = i / (3 * n_joints);
        const int c = static_cast<int>(i % 3);
        o[i] = __ldcs(p + i) - __ldg(p + f * 3 * n_joints + c);
    }
}

}  // namespace pmb
