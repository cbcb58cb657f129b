// Element-wise quaternion / dual-quaternion primitives (pymotion/rotations/quat.py,
// dual_quat.py).  One thread per element, grid-stride, 16-byte accesses for the
// quaternion operands.  Pure streaming kernels: the roofline is the byte count of
// their operands.
#pragma once
#include "common.cuh"

namespace pmb {

__device__ __forceinline__ Quat<float> ldq(const float4 *p, long long i) {
    const float4 a = __ldcs(p + i);
    return {a.x, a.y, a.z, a.w};
}
__device__ __forceinline__ void stq(float4 *p, long long i, const Quat<float> &q) {
    __stcs(p + i, make_float4(q.w, q.x, q.y, q.z));
}
__device__ __forceinline__ Vec3<float> ldv(const float *p, long long i) {
    return {__ldcs(p + 3 * i), __ldcs(p + 3 * i + 1), __ldcs(p + 3 * i + 2)};
}
__device__ __forceinline__ void stv(float *p, long long i, const Vec3<float> &v) {
    p[3 * i] = v.x, p[3 * i + 1] = v.y, p[3 * i + 2] = v.z;
}

#define PMB_GRID_STRIDE(i, n)                                                                     \
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < (n);    \
         i += static_cast<long long>(gridDim.x) * blockDim.x)

__global__ void quat_mul_kernel(const float4 *a, const float4 *b, float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) stq(o, i, q_mul(ldq(a, i), ldq(b, i)));
}
__global__ void quat_mul_vec_kernel(const float4 *q, const float *v, float *o, long long n) {
    PMB_GRID_STRIDE(i, n) stv(o, i, q_rotate(ldq(q, i), ldv(v, i)));
}
__global__ void quat_length_kernel(const float4 *q, float *o, long long n) {
    PMB_GRID_STRIDE(i, n) o[i] = q_length(ldq(q, i));
}
__global__ void quat_normalize_kernel(const float4 *q, float eps, float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) {
        // true divisions, like quat.py:423, so the result is the correctly rounded quotient
        const Quat<float> a = ldq(q, i);
        const float d = q_length(a) + eps;
        stq(o, i, Quat<float>{a.w / d, a.x / d, a.y / d, a.z / d});
    }
}
__global__ void quat_conjugate_kernel(const float4 *q, float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) stq(o, i, q_conj(ldq(q, i)));
}
// 36-byte records: a block stages its 256 matrices (9216 contiguous bytes, 16-byte aligned since 256 * 36 is a
// multiple of 16) in shared memory with a conflict-free odd stride of 9 words and writes them out as float4
// (measured: scalar 36-byte-strided stores ran at 1.3 TB/s, 22M quaternions).
__global__ void __launch_bounds__(256) quat_to_matrix_kernel(const float4 *q, float *o, long long n) {
    __shared__ __align__(16) float stage[256 * 9];
    const long long n_tiles = (n + 255) / 256;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * 256, i = base + threadIdx.x;
        if (i < n) q_to_matrix(ldq(q, i), stage + 9 * threadIdx.x);
        __syncthreads();
        const int cnt = static_cast<int>(min(256LL, n - base)) * 9;  // floats in this tile
        float *out = o + base * 9;
        for (int k = threadIdx.x; k < cnt / 4; k += 256) __stcs(reinterpret_cast<float4 *>(out) + k, reinterpret_cast<const float4 *>(stage)[k]);
        for (int k = (cnt & ~3) + threadIdx.x; k < cnt; k += 256) out[k] = stage[k];
        __syncthreads();
    }
}
__global__ void quat_from_matrix_kernel(const float *m, float4 *o, long long n) {
    PMB_GRID_STRIDE(i, n) {
        float a[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = __ldcs(m + 9 * i + k);
        stq(o, i, q_from_matrix(a));
    }
}
// dual_quat.py:12-36
template <bool A32>
__global__ void dq_from_rt_kernel(const float4 *r, const float *t, float4 *dq, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const Quat<float> qr = ldq(r, i);
        const Vec3<float> v = ldv(t, i);
        const Quat<float> d = q_mul(Quat<float>{0.f, v.x, v.y, v.z}, qr);
        st_dq<A32>(dq, i, make_float4(qr.w, qr.x, qr.y, qr.z), make_float4(0.5f * d.w, 0.5f * d.x, 0.5f * d.y, 0.5f * d.z));
    }
}
// dual_quat.py:39-59
template <bool A32>
__global__ void dq_from_t_kernel(const float *t, float4 *dq, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const Vec3<float> v = ldv(t, i);
        st_dq<A32>(dq, i, make_float4(1.f, 0.f, 0.f, 0.f), make_float4(0.f, v.x * 0.5f, v.y * 0.5f, v.z * 0.5f));
    }
}
// dual_quat.py:62-83
template <bool A32>
__global__ void dq_to_rt_kernel(const float4 *dq, float4 *r, float *t, long long n) {
    PMB_GRID_STRIDE(i, n) {
        const F8 x = ld_dq<A32>(dq, i);
        const Quat<float> qr{x.lo.x, x.lo.y, x.lo.z, x.lo.w}, qd{x.hi.x, x.hi.y, x.hi.z, x.hi.w};
        const Quat<float> m = q_mul(qd, q_conj(qr));
        stq(r, i, qr);
        stv(t, i, Vec3<float>{2.f * m.x, 2.f * m.y, 2.f * m.z});
    }
}

}  // namespace pmb
