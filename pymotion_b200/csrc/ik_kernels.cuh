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
// ---- the same alignments with approximate square roots / reciprocals -----------------------------------------------
// sqrt.approx and rcp.approx are good to about 1 ulp (against 0.5 ulp for the IEEE versions) and cost one MUFU
// instruction each instead of ~10 instructions with a slow path; the stable half-angle formulas keep those relative
// errors relative (nothing here subtracts nearly equal quantities).  Error statistics against the float64 reference are
// the same as with IEEE arithmetic at every quantile (profiles/r2_frp_error_stats.jsonl).  (A shorter chain from two
// levels of rsqrt.approx without the q / (|q| + eps) of the intermediate rotations was measured too: 10 % faster, but
// median / p99 errors 1.2 - 1.5 x larger -- not taken.)
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ Vec3<float> v_normalize_approx(const Vec3<float> &v, float eps) {
    const float inv = rcp_approx(sqrt_approx(v.x * v.x + v.y * v.y + v.z * v.z) + eps);
    return {v.x * inv, v.y * inv, v.z * inv};
}
__device__ __forceinline__ Quat<float> q_normalize_approx(const Quat<float> &q, float eps) {
    const float inv = rcp_approx(sqrt_approx(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z) + eps);
    return {q.w * inv, q.x * inv, q.y * inv, q.z * inv};
}
__device__ __forceinline__ void half_angle_approx(float dot, float ncr, float &w, float &s) {
    const float big = sqrt_approx((1.f + fabsf(dot)) * 0.5f);
    const float small = ncr * rcp_approx(big + big);
    w = dot >= 0.f ? big : small;
    s = dot >= 0.f ? small : big;
}
template <bool FAST>
__device__ __forceinline__ Quat<float> ik_from_to(const Vec3<float> &a, Vec3<float> b) {
    if (!FAST) return q_from_to_stable(a, b);
    b = v_normalize_approx(b, 1e-8f);
    const Vec3<float> cr = cross3(a, b);
    const float dot = dot3_np(a, b);
    const float ncr = sqrt_approx(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
    float w, s;
    half_angle_approx(dot, ncr, w, s);
    const float k = s * rcp_approx(ncr + 1e-8f);  // normalize(cross) * s
    Quat<float> r{w, cr.x * k, cr.y * k, cr.z * k};
    if (np_isclose(dot, 1.f)) r = {1.f, 0.f, 0.f, 0.f};
    if (np_isclose(dot, -1.f)) {
        const Vec3<float> e = np_isclose(fabsf(a.x), 1.f) ? Vec3<float>{0.f, 1.f, 0.f} : Vec3<float>{1.f, 0.f, 0.f};
        const Vec3<float> h = v_normalize(cross3(a, e), 1e-8f);
        r = {0.f, h.x, h.y, h.z};
    }
    return r;
}
template <bool FAST>
__device__ __forceinline__ Quat<float> ik_from_to_axis(const Vec3<float> &a, Vec3<float> b, const Vec3<float> &u) {
    if (!FAST) return q_from_to_axis_stable(a, b, u);
    b = v_normalize_approx(b, 1e-8f);
    const Vec3<float> cr = cross3(a, b);
    const float dot = dot3_np(a, b);
    const float ncr = sqrt_approx(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
    float w, s;
    half_angle_approx(dot, ncr, w, s);
    const float side = dot3_np(cr, u);
    s *= side > 0.f ? 1.f : (side < 0.f ? -1.f : side);  // np.sign (keeps 0 and nan)
    Quat<float> r{w, u.x * s, u.y * s, u.z * s};
    if (np_isclose(dot, 1.f)) r = {1.f, 0.f, 0.f, 0.f};
    if (np_isclose(dot, -1.f)) r = {0.f, u.x, u.y, u.z};
    return r;
}
template <bool FAST>
__device__ __forceinline__ Quat<float> ik_q_normalize(const Quat<float> &q) {
    return FAST ? q_normalize_approx(q, 1e-8f) : q_normalize(q, 1e-8f);
}
template <bool FAST>
__device__ __forceinline__ Vec3<float> ik_v_normalize(const Vec3<float> &v) {
    return FAST ? v_normalize_approx(v, 1e-8f) : v_normalize(v, 1e-8f);
}

// The walk of one frame, shared by the two kernels below.  point(j) returns P_j, load_slot / save_slot move the global
// quaternion of a branch joint, emit(j, rot) receives the local rotation of joint j (joints in index order).
template <bool FAST, class Point, class LoadSlot, class SaveSlot, class Emit>
__device__ __forceinline__ void frp_walk(int n_joints, const float4 *tab, const JointProgram &prog, const ChildTable &kids,
                                         Point point, LoadSlot load_slot, SaveSlot save_slot, Emit emit) {
    Quat<float> cur{1.f, 0.f, 0.f, 0.f};  // global rotation of the previous joint
    Vec3<float> pc_prev{0.f, 0.f, 0.f};   // point of the previous joint's first child: the next joint's own point on a chain
    int c_prev = -1;
    for (int j = 0; j < n_joints; ++j) {
        const uint32_t code = prog.code[j];
        Quat<float> G{1.f, 0.f, 0.f, 0.f};  // global rotation of j's parent (the root's parent is the world)
        if (j > 0) {
            const uint32_t src = prog_src(code);
            G = src == kSrcReg ? cur : load_slot(src);
        }
        Quat<float> rot{1.f, 0.f, 0.f, 0.f};
        const int k0 = kids.start[j], k1 = kids.start[j + 1];
        if (k1 > k0) {
            const Vec3<float> pj = j == c_prev ? pc_prev : point(j);
            const int c = kids.child[k0];
            const Vec3<float> pc = point(c);
            pc_prev = pc, c_prev = c;
            const Vec3<float> to_c{pc.x - pj.x, pc.y - pj.y, pc.z - pj.z};
            const float4 oc = tab[c];
            rot = ik_from_to<FAST>(Vec3<float>{oc.x, oc.y, oc.z}, q_rotate(q_conj(G), to_c));
            for (int k = k0 + 1; k < k1; ++k) {
                const int g = kids.child[k];
                const Quat<float> inv = q_conj(q_mul(G, ik_q_normalize<FAST>(rot)));  // fk normalises local rotations (quat.py:411)
                const Vec3<float> pg = point(g);
                const float4 og = tab[g];
                const Vec3<float> pred = q_rotate(inv, Vec3<float>{pg.x - pj.x, pg.y - pj.y, pg.z - pj.z});
                const Vec3<float> axis = q_rotate(inv, ik_v_normalize<FAST>(to_c));
                rot = q_mul(rot, ik_from_to_axis<FAST>(Vec3<float>{og.x, og.y, og.z}, pred, axis));
            }
        }
        emit(j, rot);
        cur = q_mul(G, ik_q_normalize<FAST>(rot));
        const uint32_t sv = prog_save(code);
        if (sv != kNoSave) save_slot(sv, cur);
    }
}

// Rest directions: from_to / from_to_axis normalise their first argument (quat.py:541, :616) -- the same value for every frame.
__device__ __forceinline__ void frp_fill_rest_directions(float4 *tab, const float *offsets, int n_joints) {
    for (int j = threadIdx.x; j < n_joints; j += blockDim.x) {
        const Vec3<float> d = v_normalize(Vec3<float>{offsets[3 * j], offsets[3 * j + 1], offsets[3 * j + 2]}, 1e-8f);
        tab[j] = make_float4(d.x, d.y, d.z, 0.f);
    }
}

// Thread = frame straight from / to global memory: 12-byte reads 12 J bytes apart, 16-byte stores 16 J bytes apart.  Kept
// for skeletons whose tile does not fit the tile kernel's shared memory and for unaligned position arrays.
// ncu at 4M x 65 (profiles/r2_frp_fk_4m_x_65_ncu_summary.txt): 1.37 G sector look-ups in L1 at a 56 % hit rate, 573 M sectors
// read from L2 for an input of 97.5 M (each point is touched twice, as a child and as a joint, and 24 warps x 25 KB of
// rows do not stay in L1), 260 M partial-sector stores; 174 G sectors/s through L2 -- the sector rate, not the arithmetic,
// bounds it (approximate square roots / reciprocals change nothing: 4.85 -> 4.82 ms).
template <int THREADS, bool FAST, bool WIDE>
__global__ void __launch_bounds__(THREADS)
from_root_positions_kernel(const float *__restrict__ pos, const float *__restrict__ offsets, float4 *__restrict__ rots,
                           long long n_frames, int n_joints, int n_slots, const __grid_constant__ JointProgram prog,
                           const __grid_constant__ ChildTable kids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tab = reinterpret_cast<float4 *>(smem_raw);                 // [J] rest directions: offsets / (|offsets| + 1e-8)
    float4 *slots = tab + n_joints;                                      // [n_slots][THREADS] global quaternions
    frp_fill_rest_directions(tab, offsets, n_joints);
    __syncthreads();
    const long long f = blockIdx.x * static_cast<long long>(THREADS) + threadIdx.x;
    if (f >= n_frames) return;
    const float *P = pos + f * n_joints * 3;
    float4 *R = rots + f * n_joints;
    const bool last_frame = f == n_frames - 1;
    // Rotations leave as whole 32-byte sectors where the row allows it: joints (j - 1, j) share a sector iff the float4
    // index f J + j is odd, so the rotation of an even-index joint waits in registers for its neighbour (one 256-bit store
    // instead of two half-sector ones: 260 M -> 130 M store sectors at 4M x 65).
    // this frame's row starts on a sector boundary (the array itself may start mid-sector: it is 16-byte aligned)
    const bool pair_at_odd = ((f * n_joints + static_cast<long long>((reinterpret_cast<uintptr_t>(rots) >> 4) & 1u)) & 1LL) == 0;
    Quat<float> held{1.f, 0.f, 0.f, 0.f};
    frp_walk<FAST>(
        n_joints, tab, prog, kids,
        [&](int j) {
            // A point is 12 bytes at a 4-byte aligned address; three scalar loads make a warp look up 96 sectors in L1 for 32
            // points (l1tex throughput 74 %, the busiest unit of the kernel).  WIDE: aligned 16- / 8-byte loads that cover the
            // point -- one 16-byte load when it starts 0 or 4 bytes into a 16-byte unit, an 8-byte plus a 4-byte load otherwise.
            // The case is the same for every lane iff J is a multiple of 4, and only then it pays (the host decides).  (The
            // 16-byte load at offset 0 reads 4 bytes past the point: not for the last point of the array.)
            const float *p = P + 3 * j;
            const unsigned w = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p) >> 2) & 3u;
            if (WIDE && w == 0 && !(last_frame && j == n_joints - 1)) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
                return Vec3<float>{v.x, v.y, v.z};
            }
            if (WIDE && w == 1) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p - 1));
                return Vec3<float>{v.y, v.z, v.w};
            }
            if (WIDE && w == 2) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
                return Vec3<float>{v.x, v.y, __ldg(p + 2)};
            }
            if (WIDE && w == 3) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(p + 1));
                return Vec3<float>{__ldg(p), v.x, v.y};
            }
            return Vec3<float>{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        },
        [&](uint32_t s) { const float4 v = slots[s * THREADS + threadIdx.x]; return Quat<float>{v.x, v.y, v.z, v.w}; },
        [&](uint32_t s, const Quat<float> &q) { slots[s * THREADS + threadIdx.x] = make_float4(q.w, q.x, q.y, q.z); },
        [&](int j, const Quat<float> &r) {
            const bool second = ((j & 1) != 0) == pair_at_odd;  // j closes the sector that j - 1 opened
            if (second && j > 0) {
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(R + j - 1), "f"(held.w), "f"(held.x), "f"(held.y),
                             "f"(held.z), "f"(r.w), "f"(r.x), "f"(r.y), "f"(r.z) : "memory");
            } else if (second || j == n_joints - 1) {  // the first joint of a row that starts mid-sector / the last one of a row that ends there
                R[j] = make_float4(r.w, r.x, r.y, r.z);
            }
            held = r;
        });
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
