// Shared device helpers: quaternion / rigid-transform algebra in registers and
// the status plumbing of the C ABI.  All arithmetic follows the formulas (and
// the term order) of pymotion/rotations/quat.py so that fp32 results differ
// from the reference's only by FMA contraction and the final rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pymotion_b200.h"

namespace pmb {

constexpr int kWarp = 32;
constexpr uint32_t kSrcReg = 0xFFu;   // joint program: parent transform is the previous joint (kept in registers)
constexpr uint32_t kNoSave = 0xFFu;   // joint program: this joint is not saved to a slot

// Per-joint program word (see pmb_build_joint_program in pymotion_b200.h).
struct JointProgram {
    uint32_t code[PMB_MAX_JOINTS];
};

__host__ __device__ __forceinline__ uint32_t prog_src(uint32_t c) { return c & 0xFFu; }
__host__ __device__ __forceinline__ uint32_t prog_save(uint32_t c) { return (c >> 8) & 0xFFu; }
__host__ __device__ __forceinline__ uint32_t prog_parent(uint32_t c) { return (c >> 16) & 0x7FFFu; }

// ---- streaming loads / stores ------------------------------------------------
// Inputs are read once: read-only path, no L1 allocation for pure streams.
template <typename V>
__device__ __forceinline__ V ldg_stream(const V *p) {
    return __ldcs(p);
}
template <typename V>
__device__ __forceinline__ void stg_stream(V *p, const V &v) {
    __stcs(p, v);
}

// ---- 256-bit global accesses (Blackwell: LDG.E.ENL2.256 / STG.E.ENL2.256) ----------------------------
// A dual quaternion is 32 bytes = one DRAM sector.  Written as two 16-byte stores by the same thread every
// sector arrives at L2 in two halves from two different instructions (measured: the dual-quaternion writers ran
// at 2.6 .. 2.9 TB/s of write traffic against 4.6 TB/s for the bulk-stored fk output); one 32-byte access moves
// the sector whole.  The address must be 32-byte aligned.
struct F8 {
    float4 lo, hi;
};
__device__ __forceinline__ F8 ldg256(const void *p) {
    F8 v;
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg256(void *p, const float4 &lo, const float4 &hi) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z),
                 "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                 : "memory");
}
// dual quaternion i of a [n][8] array: one 32-byte access when the array is 32-byte aligned (A32), else two float4
template <bool A32>
__device__ __forceinline__ F8 ld_dq(const float4 *dq, long long i) {
    if (A32) return ldg256(dq + 2 * i);
    return F8{__ldcs(dq + 2 * i), __ldcs(dq + 2 * i + 1)};
}
template <bool A32>
__device__ __forceinline__ void st_dq(float4 *dq, long long i, const float4 &lo, const float4 &hi) {
    if (A32) {
        stg256(dq + 2 * i, lo, hi);
    } else {
        __stcs(dq + 2 * i, lo);
        __stcs(dq + 2 * i + 1, hi);
    }
}

// ---- quaternion algebra, (w,x,y,z) ------------------------------------------
template <typename T>
struct Quat {
    T w, x, y, z;
};
template <typename T>
struct Vec3 {
    T x, y, z;
};

template <typename T>
__device__ __forceinline__ T t_sqrt(T v);
template <>
__device__ __forceinline__ float t_sqrt<float>(float v) { return sqrtf(v); }
template <>
__device__ __forceinline__ double t_sqrt<double>(double v) { return sqrt(v); }

// quat.py:364-376
template <typename T>
__device__ __forceinline__ T q_length(const Quat<T> &q) {
    return t_sqrt<T>(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
}

// quat.py:411-423: q / (|q| + eps); eps joins the NORM, so a zero quaternion maps to zero
// (and to the identity matrix in to_matrix).
template <typename T>
__device__ __forceinline__ Quat<T> q_normalize(const Quat<T> &q, T eps) {
    const T inv = T(1) / (q_length(q) + eps);
    return {q.w * inv, q.x * inv, q.y * inv, q.z * inv};
}

// Same value as q_normalize to ~3e-7 relative with two bare MUFU ops (no IEEE slow paths, no denormal
// rescaling): |q| = n2 * rsqrt(n2), 1 / (|q| + eps) through the approximate reciprocal.  eps still joins
// the NORM; squared norms below FLT_MIN (|q| < 1.1e-19) count as zero, where the reference's
// q / (|q| + 1e-8) is below 1.1e-11, i.e. the same identity matrix.
__device__ __forceinline__ Quat<float> q_normalize_fast(const Quat<float> &q, float eps) {
    const float n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    float r, inv;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n2));
    const float n = n2 >= 1.17549435e-38f ? n2 * r : 0.f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(n + eps));
    return {q.w * inv, q.x * inv, q.y * inv, q.z * inv};
}

// quat.py:337-361, same term order.
template <typename T>
__device__ __forceinline__ Quat<T> q_mul(const Quat<T> &a, const Quat<T> &b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + b.w * a.x + a.y * b.z - a.z * b.y,
            a.w * b.y + b.w * a.y + a.z * b.x - a.x * b.z,
            a.w * b.z + b.w * a.z + a.x * b.y - a.y * b.x};
}

template <typename T>
__device__ __forceinline__ Quat<T> q_conj(const Quat<T> &q) {
    return {q.w, -q.x, -q.y, -q.z};
}

// quat.py:320-334 (+ _fast_cross :653-674): t = 2 (u x v); v + w t + u x t.
template <typename T>
__device__ __forceinline__ Vec3<T> q_rotate(const Quat<T> &q, const Vec3<T> &v) {
    const T tx = T(2) * (q.y * v.z - q.z * v.y);
    const T ty = T(2) * (q.z * v.x - q.x * v.z);
    const T tz = T(2) * (q.x * v.y - q.y * v.x);
    return {v.x + q.w * tx + (q.y * tz - q.z * ty),
            v.y + q.w * ty + (q.z * tx - q.x * tz),
            v.z + q.w * tz + (q.x * ty - q.y * tx)};
}

// quat.py:276-317: row-major m[3*r + c]; no normalisation inside.
template <typename T>
__device__ __forceinline__ void q_to_matrix(const Quat<T> &q, T m[9]) {
    const T x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    const T xx = q.x * x2, yy = q.y * y2, zz = q.z * z2;
    const T xy = q.x * y2, xz = q.x * z2, yz = q.y * z2;
    const T wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    m[0] = T(1) - (yy + zz); m[1] = xy - wz;          m[2] = xz + wy;
    m[3] = xy + wz;          m[4] = T(1) - (xx + zz); m[5] = yz - wx;
    m[6] = xz - wy;          m[7] = yz + wx;          m[8] = T(1) - (xx + yy);
}

// quat.py:85-156: branch selection on m22 < 0, m00 > m11, m00 < -m11; then normalize (eps 1e-8).
template <typename T>
__device__ __forceinline__ Quat<T> q_from_matrix(const T m[9]) {
    const T m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7],
            m22 = m[8];
    Quat<T> q;
    if (m22 < T(0)) {
        if (m00 > m11) q = {m21 - m12, T(1) + m00 - m11 - m22, m10 + m01, m02 + m20};
        else           q = {m02 - m20, m10 + m01, T(1) - m00 + m11 - m22, m21 + m12};
    } else {
        if (m00 < -m11) q = {m10 - m01, m02 + m20, m21 + m12, T(1) - m00 - m11 + m22};
        else            q = {T(1) + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01};
    }
    return q_normalize(q, T(1e-8));
}

// Same branch selection, normalisation through q_normalize_fast (used inside the fused fk_quat kernel).
__device__ __forceinline__ Quat<float> q_from_matrix_fast(const float m[9]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7],
                m22 = m[8];
    const bool neg = m22 < 0.f;
    const bool alt = neg ? (m00 > m11) : (m00 < -m11);
    // the four candidates of quat.py:115-153 selected without divergent control flow
    const Quat<float> a = neg ? Quat<float>{m21 - m12, 1.f + m00 - m11 - m22, m10 + m01, m02 + m20}
                              : Quat<float>{m10 - m01, m02 + m20, m21 + m12, 1.f - m00 - m11 + m22};
    const Quat<float> b = neg ? Quat<float>{m02 - m20, m10 + m01, 1.f - m00 + m11 - m22, m21 + m12}
                              : Quat<float>{1.f + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01};
    return q_normalize_fast(alt ? a : b, 1e-8f);
}

// Rigid transform [R | p], R row-major.  12 registers.
template <typename T>
struct Xform {
    T r[9];
    T p[3];
};

// G = P * [L | off]   (ops/skeleton.py:55-58 restricted to the 3x4 block that is not constant)
template <typename T>
__device__ __forceinline__ void xf_compose(Xform<T> &g, const Xform<T> &par, const T l[9], T ox, T oy, T oz) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const T p0 = par.r[3 * a], p1 = par.r[3 * a + 1], p2 = par.r[3 * a + 2];
        g.r[3 * a + 0] = p0 * l[0] + p1 * l[3] + p2 * l[6];
        g.r[3 * a + 1] = p0 * l[1] + p1 * l[4] + p2 * l[7];
        g.r[3 * a + 2] = p0 * l[2] + p1 * l[5] + p2 * l[8];
        g.p[a] = p0 * ox + p1 * oy + p2 * oz + par.p[a];
    }
}

}  // namespace pmb
