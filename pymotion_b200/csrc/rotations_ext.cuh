// The rest of the quaternion / dual-quaternion surface of the reference:
//   pymotion/rotations/quat.py       from_scaled_angle_axis :6, from_angle_axis :24, from_euler :43,
//                                    to_euler :159, to_scaled_angle_axis :230, to_angle_axis :247,
//                                    unroll :426, slerp :465, from_to :504, from_to_axis :579
//   pymotion/rotations/dual_quat.py  normalize :86, is_unit :118, unroll :139
// Element-wise kernels: one thread per element, grid-stride, 16-byte accesses for quaternions.  `unroll` is
// the one op with a dependency ALONG the frame axis: a segmented prefix product of signs, done as a
// three-kernel chunked scan over a 2-bit state.  All arithmetic is fp32 with correctly rounded sqrt /
// division and the accurate libdevice sin / cos / acos / atan2 (no fast-math): these feed tolerance tests at
// 1e-5 against the NumPy reference.
#pragma once
#include "common.cuh"
#include "dq_kernels.cuh"
#include "elementwise.cuh"

namespace pmb {

// np.isclose(x, target) with the NumPy defaults rtol = 1e-5, atol = 1e-8 (evaluated in fp32 for fp32 input)
__device__ __forceinline__ bool np_isclose(float x, float target) {
    return fabsf(x - target) <= 1e-8f + 1e-5f * fabsf(target);
}
// products and sums rounded separately, like NumPy's element-wise multiply followed by np.sum
__device__ __forceinline__ float dot3_np(const Vec3<float> &a, const Vec3<float> &b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float dot4_np(const Quat<float> &a, const Quat<float> &b) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.w, b.w), __fmul_rn(a.x, b.x)), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ Vec3<float> cross3(const Vec3<float> &a, const Vec3<float> &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// ops/vector.py:4-19 and quat.py:411-423 applied to a 3-vector: v / (|v| + eps)
__device__ __forceinline__ Vec3<float> v_normalize(const Vec3<float> &v, float eps) {
    const float d = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z) + eps;
    return {v.x / d, v.y / d, v.z / d};
}

// quat.py:24-40
__device__ __forceinline__ Quat<float> q_from_angle_axis(float angle, const Vec3<float> &axis) {
    float s, c;
    sincosf(angle * 0.5f, &s, &c);
    return {c, s * axis.x, s * axis.y, s * axis.z};
}
__global__ void quat_from_angle_axis_kernel(const float *angle, const float *axis, float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) stq(o, i, q_from_angle_axis(__ldcs(angle + i), ldv(axis, i)));
}
// quat.py:6-21: the null vector gives 0 / 0 = nan in the axis, as in the reference
__global__ void quat_from_scaled_angle_axis_kernel(const float *sa, float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const Vec3<float> v = ldv(sa, i);
        const float angle = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
        stq(o, i, q_from_angle_axis(angle, Vec3<float>{v.x / angle, v.y / angle, v.z / angle}));
    }
}

// Euler orders travel as one byte per element: o0 + 3 * o1 + 9 * o2, o = 0 | 1 | 2 for 'x' | 'y' | 'z'.
__device__ __forceinline__ void order_unpack(uint8_t code, int &o0, int &o1, int &o2) {
    o0 = code % 3, o1 = (code / 3) % 3, o2 = code / 9;
}
__device__ __forceinline__ Quat<float> q_about_axis(float angle, int ax) {
    float s, c;
    sincosf(angle * 0.5f, &s, &c);
    return {c, ax == 0 ? s : 0.f, ax == 1 ? s : 0.f, ax == 2 ? s : 0.f};
}
// quat.py:43-82: q(e0 about o0) (x) q(e1 about o1) (x) q(e2 about o2)
__global__ void quat_from_euler_kernel(const float *euler, const uint8_t *order, long long order_stride, float4 *o,
                                       long long n) {
    PMB_GRID_STRIDE(i, n) {
        int o0, o1, o2;
        order_unpack(order[i * order_stride], o0, o1, o2);
        const Vec3<float> e = ldv(euler, i);
        stq(o, i, q_mul(q_about_axis(e.x, o0), q_mul(q_about_axis(e.y, o1), q_about_axis(e.z, o2))));
    }
}
__device__ __forceinline__ float np_mod_2pi(float x) {
    constexpr float kTwoPi = 6.283185307179586f;
    float r = fmodf(x, kTwoPi);
    if (r < 0.f) r += kTwoPi;
    return r;
}
__device__ __forceinline__ float q_component(const Quat<float> &q, int k) {  // k = 0 | 1 | 2 -> x | y | z
    return k == 0 ? q.x : (k == 1 ? q.y : q.z);
}
// quat.py:159-227
__global__ void quat_to_euler_kernel(const float4 *q, const uint8_t *order, long long order_stride, float *o,
                                     long long n) {
    PMB_GRID_STRIDE(idx, n) {
        int o0, o1, o2;
        order_unpack(order[idx * order_stride], o0, o1, o2);
        const int i = o2, j = o1, k = o0;                      // :191-193
        const float sign = static_cast<float>((i - j) * (j - k) * (k - i) / 2);  // +-1 (+-2 / 2, exact)
        const Quat<float> a4 = ldq(q, idx);
        const float qi = q_component(a4, i), qj = q_component(a4, j), qk = q_component(a4, k);
        const float a = a4.w - qj, b = qi + qk * sign, c = qj + a4.w, d = qk * sign - qi;
        const float second = 2.f * atan2f(hypotf(c, d), hypotf(a, b)) - 1.5707963267948966f;
        const float half_sum = atan2f(b, a), half_diff = atan2f(d, c);
        o[3 * idx] = np_mod_2pi((half_sum + half_diff) * sign);
        o[3 * idx + 1] = np_mod_2pi(second);
        o[3 * idx + 2] = np_mod_2pi(half_sum - half_diff);
    }
}

// quat.py:247-273 (scaled = false: angle [n], axis [n][3]) and :230-244 (scaled = true: angle * axis into `axis`)
__global__ void quat_to_angle_axis_kernel(const float4 *q, float *angle, float *axis, int scaled, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const Quat<float> a = ldq(q, i);
        const float ang = 2.f * acosf(fminf(fmaxf(a.w, -1.f), 1.f));
        const float s = sqrtf(fminf(fmaxf(__fsub_rn(1.f, __fmul_rn(a.w, a.w)), 0.f), 1.f));  // two roundings, like NumPy: 1 - w*w is ill-conditioned near |w| = 1
        Vec3<float> ax{0.f, 0.f, 0.f};
        if (s > 1e-8f) ax = {a.x / s, a.y / s, a.z / s};
        if (scaled) {
            stv(axis, i, Vec3<float>{ang * ax.x, ang * ax.y, ang * ax.z});
        } else {
            angle[i] = ang;
            stv(axis, i, ax);
        }
    }
}

// quat.py:465-501; t is one value per element (t_stride = 1) or one shared value (t_stride = 0)
__global__ void quat_slerp_kernel(const float4 *q0, const float4 *q1, const float *t, long long t_stride, int shortest,
                                  float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const Quat<float> a = ldq(q0, i);
        Quat<float> b = ldq(q1, i);
        float dot = dot4_np(a, b);
        if (shortest && dot < 0.f) b = {-b.w, -b.x, -b.y, -b.z}, dot = -dot;
        dot = fminf(fmaxf(dot, -1.f), 1.f);
        const float theta = acosf(dot) * t[i * t_stride];
        Quat<float> c{b.w - a.w * dot, b.x - a.x * dot, b.y - a.y * dot, b.z - a.z * dot};
        // :499: 1e-6 joins every COMPONENT before the norm
        const float w1 = c.w + 1e-6f, x1 = c.x + 1e-6f, y1 = c.y + 1e-6f, z1 = c.z + 1e-6f;
        const float nrm = sqrtf(w1 * w1 + x1 * x1 + y1 * y1 + z1 * z1);
        c = {c.w / nrm, c.x / nrm, c.y / nrm, c.z / nrm};
        float sn, cs;
        sincosf(theta, &sn, &cs);
        stq(o, i, Quat<float>{cs * a.w + sn * c.w, cs * a.x + sn * c.x, cs * a.y + sn * c.y, cs * a.z + sn * c.z});
    }
}

// quat.py:504-576: rotation taking direction v1 onto v2 (half-angle quaternion about normalize(v1 x v2));
// identity where np.isclose(dot, 1); where np.isclose(dot, -1) a half turn about normalize(v1 x e), e = y if
// v1 lies along x, else x (:552-562).
__device__ __forceinline__ Quat<float> q_from_to(Vec3<float> a, Vec3<float> b, bool normalize_input) {
    if (normalize_input) a = v_normalize(a, 1e-8f), b = v_normalize(b, 1e-8f);
    const Vec3<float> cr = cross3(a, b);
    const float dot = dot3_np(a, b);
    const float w = sqrtf((1.f + dot) * 0.5f), s = sqrtf((1.f - dot) * 0.5f);
    const Vec3<float> u = v_normalize(cr, 1e-8f);
    Quat<float> r{w, u.x * s, u.y * s, u.z * s};
    if (np_isclose(dot, 1.f)) r = {1.f, 0.f, 0.f, 0.f};
    if (np_isclose(dot, -1.f)) {
        const Vec3<float> e = np_isclose(fabsf(a.x), 1.f) ? Vec3<float>{0.f, 1.f, 0.f} : Vec3<float>{1.f, 0.f, 0.f};
        const Vec3<float> h = v_normalize(cross3(a, e), 1e-8f);
        r = {0.f, h.x, h.y, h.z};
    }
    return r;
}
// quat.py:579-650: the same half angle about the GIVEN axis u, signed by sign((v1 x v2) . u); identity where
// np.isclose(dot, 1); (0, u) where np.isclose(dot, -1).
__device__ __forceinline__ Quat<float> q_from_to_axis(Vec3<float> a, Vec3<float> b, const Vec3<float> &u, bool normalize_input) {
    if (normalize_input) a = v_normalize(a, 1e-8f), b = v_normalize(b, 1e-8f);
    const Vec3<float> cr = cross3(a, b);
    const float dot = dot3_np(a, b);
    const float w = sqrtf((1.f + dot) * 0.5f);
    float s = sqrtf((1.f - dot) * 0.5f);
    const float side = dot3_np(cr, u);
    s *= side > 0.f ? 1.f : (side < 0.f ? -1.f : side);  // np.sign (keeps 0 and nan)
    Quat<float> r{w, u.x * s, u.y * s, u.z * s};
    if (np_isclose(dot, 1.f)) r = {1.f, 0.f, 0.f, 0.f};
    if (np_isclose(dot, -1.f)) r = {0.f, u.x, u.y, u.z};
    return r;
}
// The same two rotations for from_root_positions (ik_kernels.cuh), where the angle is usually SMALL (a rigid skeleton's
// further children are almost aligned once the first one is) and the result feeds a chain of products: the reference's
// float64 sqrt((1 - dot) / 2) loses everything to cancellation in fp32 there (1 - dot ~ 1e-6 carries 6 % of rounding).
// sin(theta) = |a x b| = 2 s w gives the small one of (w, s) from the well-conditioned one: same values as the
// reference's formulas, conditioned for fp32.  `a` is already normalised (a rest offset, normalised once per block).
__device__ __forceinline__ void half_angle_stable(float dot, float ncr, float &w, float &s) {
    if (dot >= 0.f) {
        w = sqrtf((1.f + dot) * 0.5f);
        s = ncr / (w + w);
    } else {
        s = sqrtf((1.f - dot) * 0.5f);
        w = ncr / (s + s);
    }
}
__device__ __forceinline__ Quat<float> q_from_to_stable(const Vec3<float> &a, Vec3<float> b) {
    b = v_normalize(b, 1e-8f);
    const Vec3<float> cr = cross3(a, b);
    const float dot = dot3_np(a, b);
    const float ncr = sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
    float w, s;
    half_angle_stable(dot, ncr, w, s);
    const float k = s / (ncr + 1e-8f);  // normalize(cross) * s
    Quat<float> r{w, cr.x * k, cr.y * k, cr.z * k};
    if (np_isclose(dot, 1.f)) r = {1.f, 0.f, 0.f, 0.f};
    if (np_isclose(dot, -1.f)) {
        const Vec3<float> e = np_isclose(fabsf(a.x), 1.f) ? Vec3<float>{0.f, 1.f, 0.f} : Vec3<float>{1.f, 0.f, 0.f};
        const Vec3<float> h = v_normalize(cross3(a, e), 1e-8f);
        r = {0.f, h.x, h.y, h.z};
    }
    return r;
}
__device__ __forceinline__ Quat<float> q_from_to_axis_stable(const Vec3<float> &a, Vec3<float> b, const Vec3<float> &u) {
    b = v_normalize(b, 1e-8f);
    const Vec3<float> cr = cross3(a, b);
    const float dot = dot3_np(a, b);
    const float ncr = sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
    float w, s;
    half_angle_stable(dot, ncr, w, s);
    const float side = dot3_np(cr, u);
    s *= side > 0.f ? 1.f : (side < 0.f ? -1.f : side);  // np.sign (keeps 0 and nan)
    Quat<float> r{w, u.x * s, u.y * s, u.z * s};
    if (np_isclose(dot, 1.f)) r = {1.f, 0.f, 0.f, 0.f};
    if (np_isclose(dot, -1.f)) r = {0.f, u.x, u.y, u.z};
    return r;
}
__global__ void quat_from_to_kernel(const float *v1, const float *v2, const float *axis, int normalize_input, float4 *o,
                                    long long n) {
    PMB_GRID_STRIDE(i, n) {
        const Vec3<float> a = ldv(v1, i), b = ldv(v2, i);
        stq(o, i, axis == nullptr ? q_from_to(a, b, normalize_input != 0)
                                  : q_from_to_axis(a, b, ldv(axis, i), normalize_input != 0));
    }
}

// ---- dual_quat.normalize / is_unit -------------------------------------------------------------------
// flags[0] = some real part is not ~0, flags[1] = some |real|^2 is not ~1, flags[2] = some real.dual is not ~0
// (dual_quat.py:118-136 reduces over the WHOLE array).  flags must be zero on entry.
__device__ __forceinline__ void unit_flags(const Quat<float> &r, const Quat<float> &d, float atol, int *flags) {
    const float n2 = dot4_np(r, r);
    const bool f0 = !np_isclose(n2, 0.f), f1 = !np_isclose(n2, 1.f), f2 = !(fabsf(dot4_np(r, d)) <= atol);
    // one atomic per warp and flag
    if (__any_sync(__activemask(), f0) && f0) atomicOr(flags + 0, 1);
    if (__any_sync(__activemask(), f1) && f1) atomicOr(flags + 1, 1);
    if (__any_sync(__activemask(), f2) && f2) atomicOr(flags + 2, 1);
}
__global__ void dq_is_unit_kernel(const float4 *dq, float atol, int *flags, long long n) {
    PMB_GRID_STRIDE(i, n) unit_flags(ldq(dq, 2 * i), ldq(dq, 2 * i + 1), atol, flags);
}
// dual_quat.normalize (dual_quat.py:86-115).  The reference scales both parts by 1 / |real|, asks whether the scaled ARRAY
// AS A WHOLE is unit (is_unit, atol 1e-3) and only if it is not removes the real direction from every dual part.
// Pass 1 reads the input once, writes the scaled values (one square root and one division per element, within an ulp of
// the reference's quotients) and reduces the whole-array verdict on the way: for r / |r| the squared norm is 1 to rounding
// unless |r|^2 is 0 / inf / nan (so "all real parts ~ 0" can never hold and flags[0] is always raised), and real . dual of
// the scaled pair is (r . d) / |r|^2.  Pass 2 is launched behind it, reads the three flags and returns at once when the
// verdict is "unit" -- the case of dual quaternions that were built from rotations and translations -- so that case
// costs the algorithmic 64 bytes per element; otherwise it projects IN PLACE on the output (k = rn . dn equals
// (r . d) / |r|^2), 128 bytes per element in total.  (Round 1 read the input twice, 96 bytes per element in both cases.)
template <bool A32>
__global__ void dq_normalize_scale_kernel(const float4 *dq, float4 *o, int *flags, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const F8 x = ld_dq<A32>(dq, i);
        const Quat<float> r{x.lo.x, x.lo.y, x.lo.z, x.lo.w}, d{x.hi.x, x.hi.y, x.hi.z, x.hi.w};
        const float n2 = dot4_np(r, r);
        const bool f1 = !(n2 > 0.f && n2 < 3.0e38f);
        const bool f2 = !(fabsf(dot4_np(r, d) / n2) <= 1e-3f);
        if (i == 0) atomicOr(flags + 0, 1);
        if (__any_sync(__activemask(), f1) && f1) atomicOr(flags + 1, 1);
        if (__any_sync(__activemask(), f2) && f2) atomicOr(flags + 2, 1);
        const float inv = 1.f / sqrtf(r.w * r.w + r.x * r.x + r.y * r.y + r.z * r.z);
        st_dq<A32>(o, i, make_float4(r.w * inv, r.x * inv, r.y * inv, r.z * inv), make_float4(d.w * inv, d.x * inv, d.y * inv, d.z * inv));
    }
}
template <bool A32>
__global__ void dq_normalize_project_kernel(float4 *o, const int *flags, long long n) {
    if (flags[0] == 0 || (flags[1] == 0 && flags[2] == 0)) return;  // the scaled array is unit: nothing to do
    PMB_GRID_STRIDE(i, n) {
        const F8 x = ld_dq<A32>(o, i);
        const Quat<float> rn{x.lo.x, x.lo.y, x.lo.z, x.lo.w}, dn{x.hi.x, x.hi.y, x.hi.z, x.hi.w};
        const float k = dot4_np(rn, dn);
        st_dq<A32>(o, i, x.lo, make_float4(dn.w - rn.w * k, dn.x - rn.x * k, dn.y - rn.y * k, dn.z - rn.z * k));
    }
}

// ---- unroll (quat.py:426-462, dual_quat.py:139-167) ---------------------------------------------------
// x is [T][M][W4] float4 (W4 = 1 quaternions, 2 dual quaternions; the decision uses the first float4).
// Entry t is negated iff its dot product with the UNROLLED entry t-1 is < 0.  With d_t the dot product of the
// ORIGINAL entries, the sign s_t obeys s_t = s_{t-1} * sgn(d_t) when d_t != 0 and s_t = +1 when d_t is 0 or
// nan (no flip): a prefix product with resets.  State = 2 bits {neg, reset seen}; combine(p, c) =
// {c.reset ? c.neg : p.neg ^ c.neg, p.reset | c.reset} is associative, so the scan is chunked:
//   unroll_local   per (chunk of kUnrollChunk steps, column): running state of every step -> local[T][M],
//                  state of the whole chunk -> agg[chunks][M]
//   unroll_chunks  per column: exclusive scan of agg over the chunks (in place)
//   unroll_apply   per (step, column): combine(agg[chunk], local[t]) decides the flip
constexpr int kUnrollChunk = 128;
__device__ __forceinline__ uint8_t unroll_combine(uint8_t p, uint8_t c) {
    const uint8_t neg = (c & 2) ? (c & 1) : ((p ^ c) & 1);
    return neg | ((p | c) & 2);
}
__global__ void unroll_local_kernel(const float4 *__restrict__ x, int w4, long long n_steps, long long n_cols,
                                    long long n_chunks, uint8_t *__restrict__ local, uint8_t *__restrict__ agg) {
    // one thread per (chunk, column), columns fastest: neighbouring threads read neighbouring quaternions.  The steps
    // of a chunk are fetched eight at a time, so a thread has 128 bytes in flight instead of 16 (the chain over the
    // steps is cheap; what this kernel waits for is memory).
    constexpr int B = 8;
    const long long id = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (id >= n_chunks * n_cols) return;
    const long long c = id / n_cols, m = id - c * n_cols, t0 = c * kUnrollChunk;
    const long long t1 = min(t0 + kUnrollChunk, n_steps);
    float4 prev = t0 > 0 ? __ldg(x + ((t0 - 1) * n_cols + m) * w4) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint8_t state = 0;
    for (long long tb = t0; tb < t1; tb += B) {
        float4 buf[B];
#pragma unroll
        for (int k = 0; k < B; ++k)
            if (tb + k < t1) buf[k] = __ldg(x + ((tb + k) * n_cols + m) * w4);
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const long long t = tb + k;
            if (t < t1) {
                const float4 cur = buf[k];
                uint8_t el;
                if (t == 0) {
                    el = 2;  // the first entry keeps its cover
                } else {
                    // np.sum(r[i] * r[i - 1], axis=-1): separate roundings
                    const float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(cur.x, prev.x), __fmul_rn(cur.y, prev.y)), __fmul_rn(cur.z, prev.z)),
                                              __fmul_rn(cur.w, prev.w));
                    el = d < 0.f ? 1 : (d > 0.f ? 0 : 2);
                }
                state = unroll_combine(state, el);
                local[t * n_cols + m] = state;
                prev = cur;
            }
        }
    }
    agg[c * n_cols + m] = state;
}
// one block per column: every thread folds a contiguous run of chunks, the block scans the 256 partial
// states in shared memory, then every thread rewrites its run with the exclusive prefixes
__global__ void __launch_bounds__(256) unroll_chunks_kernel(uint8_t *agg, long long n_chunks, long long n_cols) {
    __shared__ uint8_t part[256];
    const long long m = blockIdx.x;
    const long long per = (n_chunks + 255) / 256;
    const long long c0 = min(threadIdx.x * per, n_chunks), c1 = min(c0 + per, n_chunks);
    uint8_t run = 0;
    for (long long c = c0; c < c1; ++c) run = unroll_combine(run, agg[c * n_cols + m]);
    part[threadIdx.x] = run;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {  // inclusive Hillis-Steele scan with the (associative, non-commutative) combine
        const uint8_t left = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        if (threadIdx.x >= d) part[threadIdx.x] = unroll_combine(left, part[threadIdx.x]);
        __syncthreads();
    }
    run = threadIdx.x > 0 ? part[threadIdx.x - 1] : 0;  // state before this thread's run
    for (long long c = c0; c < c1; ++c) {
        const uint8_t a = agg[c * n_cols + m];
        agg[c * n_cols + m] = run;  // exclusive: the state BEFORE the chunk
        run = unroll_combine(run, a);
    }
}
__global__ void unroll_apply_kernel(const float4 *__restrict__ x, int w4, long long n_steps, long long n_cols,
                                    const uint8_t *__restrict__ local, const uint8_t *__restrict__ agg, float4 *__restrict__ o) {
    const long long n = n_steps * n_cols;
    PMB_GRID_STRIDE(i, n) {
        const long long t = i / n_cols, m = i - t * n_cols;
        const uint8_t s = unroll_combine(agg[(t / kUnrollChunk) * n_cols + m], local[i]);
        const float f = (s & 1) ? -1.f : 1.f;
        for (int k = 0; k < w4; ++k) {
            const float4 v = __ldcs(x + i * w4 + k);
            __stcs(o + i * w4 + k, make_float4(f * v.x, f * v.y, f * v.z, f * v.w));
        }
    }
}

// The same for n_cols <= 512: one block per chunk, so the column of an element comes from a 16-bit multiply-high
// (div_small, dq_kernels.cuh) instead of two 64-bit divisions per element.
__global__ void __launch_bounds__(256)
unroll_apply_chunk_kernel(const float4 *__restrict__ x, int w4, long long n_steps, int n_cols, uint32_t magic,
                          const uint8_t *__restrict__ local, const uint8_t *__restrict__ agg, float4 *__restrict__ o) {
    const long long c = blockIdx.x, t0 = c * kUnrollChunk;
    const int steps = static_cast<int>(min(static_cast<long long>(kUnrollChunk), n_steps - t0));
    const int n_el = steps * n_cols;  // <= 128 * 512 = 2^16
    const long long base = t0 * n_cols;
    const uint8_t *a = agg + c * n_cols;
    for (int il = threadIdx.x; il < n_el; il += 256) {
        const int m = il - div_small(il, magic) * n_cols;
        const long long i = base + il;
        const uint8_t s = unroll_combine(a[m], local[i]);
        const float f = (s & 1) ? -1.f : 1.f;
        for (int k = 0; k < w4; ++k) {
            const float4 v = __ldcs(x + i * w4 + k);
            __stcs(o + i * w4 + k, make_float4(f * v.x, f * v.y, f * v.z, f * v.w));
        }
    }
}

// ---- BVH.get_data producer chain, fused (io/bvh.py:352-359) -----------------------------------------------------
// rots = quat.normalize(quat.unroll(quat.from_euler(np.radians(rotations), order tiled over the frames), axis=0))
// as the same chunked scan as `unroll`, with the quaternion of an entry COMPUTED from its Euler angles where it is
// needed (never stored un-unrolled) and the final normalisation applied by the pass that writes:
//   bvh_local   per (chunk of kUnrollChunk frames, joint): quaternions from Euler angles on the fly, running sign state
//               -> local[T][J], state of the chunk -> agg[chunks][J]
//   unroll_chunks (above) exclusive scan of agg over the chunks
//   bvh_apply   per (frame, joint): quaternion again, sign from combine(agg, local), q / (|q| + 1e-8), store
// 12 + 12 bytes of angles read, 16 written, 2 of state per entry: 42 bytes against 110 for the three separate ops
// (from_euler 28, unroll 50, normalize 32); 28 is the algorithmic minimum.
__device__ __forceinline__ Quat<float> q_from_euler_deg(const float *euler_deg, long long i, int o0, int o1, int o2) {
    constexpr float kRad = 0.017453292519943295f;  // np.radians: x * pi / 180
    const Vec3<float> e = ldv(euler_deg, i);
    return q_mul(q_about_axis(e.x * kRad, o0), q_mul(q_about_axis(e.y * kRad, o1), q_about_axis(e.z * kRad, o2)));
}
__global__ void bvh_local_kernel(const float *__restrict__ euler_deg, const uint8_t *__restrict__ order, long long n_steps,
                                 long long n_cols, long long n_chunks, uint8_t *__restrict__ local, uint8_t *__restrict__ agg) {
    const long long id = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (id >= n_chunks * n_cols) return;
    const long long c = id / n_cols, m = id - c * n_cols, t0 = c * kUnrollChunk;
    const long long t1 = min(t0 + kUnrollChunk, n_steps);
    int o0, o1, o2;
    order_unpack(order[m], o0, o1, o2);
    Quat<float> prev = t0 > 0 ? q_from_euler_deg(euler_deg, (t0 - 1) * n_cols + m, o0, o1, o2) : Quat<float>{0.f, 0.f, 0.f, 0.f};
    uint8_t state = 0;
    for (long long t = t0; t < t1; ++t) {
        const Quat<float> cur = q_from_euler_deg(euler_deg, t * n_cols + m, o0, o1, o2);
        uint8_t el = 2;  // the first entry keeps its cover
        if (t > 0) {
            const float d = dot4_np(cur, prev);  // np.sum(r[i] * r[i - 1], axis=-1): separate roundings
            el = d < 0.f ? 1 : (d > 0.f ? 0 : 2);
        }
        state = unroll_combine(state, el);
        local[t * n_cols + m] = state;
        prev = cur;
    }
    agg[c * n_cols + m] = state;
}
__global__ void bvh_apply_kernel(const float *__restrict__ euler_deg, const uint8_t *__restrict__ order, long long n_steps,
                                 long long n_cols, const uint8_t *__restrict__ local, const uint8_t *__restrict__ agg,
                                 float4 *__restrict__ o) {
    const long long n = n_steps * n_cols;
    PMB_GRID_STRIDE(i, n) {
        const long long t = i / n_cols, m = i - t * n_cols;
        int o0, o1, o2;
        order_unpack(order[m], o0, o1, o2);
        const uint8_t s = unroll_combine(agg[(t / kUnrollChunk) * n_cols + m], local[i]);
        Quat<float> q = q_from_euler_deg(euler_deg, i, o0, o1, o2);
        if (s & 1) q = {-q.w, -q.x, -q.y, -q.z};
        stq(o, i, q_normalize(q, 1e-8f));
    }
}

}  // namespace pmb
