// The numeric modules of the reference that sit either side of fk (SURVEY 8f rank 3 tail and rank 4):
//   pymotion/rotations/ortho6d.py      from_quat :14, from_matrix :31, to_quat :50, to_matrix :67
//   pymotion/ops/center_of_mass.py     center_of_mass :52 (human_center_of_mass :4 builds its weights on the host)
//   pymotion/ops/time.py               interpolate_positions :4
//   pymotion/ops/vector.py             normalize :4
// Element-wise / gather kernels; HBM-bound streaming, sized by their operand bytes.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"
#include "rotations_ext.cuh"

namespace pmb {

// ---- ortho6d: [n][3][2] = the first two columns of the rotation matrix ---------------------------------
__device__ __forceinline__ void o6_store(float *o, long long i, const float m[9]) {
    float2 *p = reinterpret_cast<float2 *>(o + 6 * i);  // 24-byte records: 8-byte aligned
    p[0] = make_float2(m[0], m[1]), p[1] = make_float2(m[3], m[4]), p[2] = make_float2(m[6], m[7]);
}
__global__ void ortho6d_from_matrix_kernel(const float *m, float *o, long long n) {
    PMB_GRID_STRIDE(i, n) {
        float a[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = __ldcs(m + 9 * i + k);
        o6_store(o, i, a);
    }
}
__global__ void ortho6d_from_quat_kernel(const float4 *q, float *o, long long n) {
    PMB_GRID_STRIDE(i, n) {
        float a[9];
        q_to_matrix(ldq(q, i), a);
        o6_store(o, i, a);
    }
}
// ortho6d.py:67-90: Gram-Schmidt on the two columns (plain norms, no eps), third column = cross product
__device__ __forceinline__ void o6_to_matrix(const float *o, long long i, float m[9]) {
    const float2 *p = reinterpret_cast<const float2 *>(o + 6 * i);
    const float2 r0 = p[0], r1 = p[1], r2 = p[2];
    Vec3<float> a{r0.x, r1.x, r2.x}, b{r0.y, r1.y, r2.y};
    const float na = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    const Vec3<float> c1{a.x / na, a.y / na, a.z / na};
    const float d = dot3_np(c1, b);
    Vec3<float> c2{b.x - d * c1.x, b.y - d * c1.y, b.z - d * c1.z};
    const float nb = sqrtf(c2.x * c2.x + c2.y * c2.y + c2.z * c2.z);
    c2 = {c2.x / nb, c2.y / nb, c2.z / nb};
    const Vec3<float> c3 = cross3(c1, c2);
    m[0] = c1.x, m[1] = c2.x, m[2] = c3.x;
    m[3] = c1.y, m[4] = c2.y, m[5] = c3.y;
    m[6] = c1.z, m[7] = c2.z, m[8] = c3.z;
}
// 36-byte output records: staged per block and written as float4, like quat_to_matrix_kernel
__global__ void __launch_bounds__(256) ortho6d_to_matrix_kernel(const float *o, float *m, long long n) {
    __shared__ __align__(16) float stage[256 * 9];
    const long long n_tiles = (n + 255) / 256;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * 256, i = base + threadIdx.x;
        if (i < n) o6_to_matrix(o, i, stage + 9 * threadIdx.x);
        __syncthreads();
        const int cnt = static_cast<int>(min(256LL, n - base)) * 9;
        float *out = m + base * 9;
        for (int k = threadIdx.x; k < cnt / 4; k += 256) __stcs(reinterpret_cast<float4 *>(out) + k, reinterpret_cast<const float4 *>(stage)[k]);
        for (int k = (cnt & ~3) + threadIdx.x; k < cnt; k += 256) out[k] = stage[k];
        __syncthreads();
    }
}
__global__ void ortho6d_to_quat_kernel(const float *o, float4 *q, long long n) {
    PMB_GRID_STRIDE(i, n) {
        float a[9];
        o6_to_matrix(o, i, a);
        stq(q, i, q_from_matrix(a));
    }
}

// ---- center_of_mass: out[f][c] = sum_j joints[f][j][c] * weights[f * wstride + j] ------------------------
// One thread per (frame, component); the sum runs in joint order with the product rounded first, like
// np.sum(joints * weights[..., None], axis=-2).  The three threads of a frame read neighbouring words and a
// warp's footprint (11 frames x 12 J bytes) stays in L1 across the joint loop, so DRAM sees each byte once.
__global__ void center_of_mass_kernel(const float *joints, const float *weights, long long wstride, float *out,
                                      long long n_frames, int n_joints) {
    PMB_GRID_STRIDE(i, 3 * n_frames) {
        const long long f = i / 3;
        const int c = static_cast<int>(i - 3 * f);
        const float *p = joints + f * 3 * n_joints + c;
        const float *w = weights + f * wstride;
        float acc = __fmul_rn(__ldg(p), __ldg(w));
        for (int j = 1; j < n_joints; ++j) acc = __fadd_rn(acc, __fmul_rn(__ldg(p + 3 * j), __ldg(w + j)));
        out[i] = acc;
    }
}

// ---- interpolate_positions (time.py:4-66) ---------------------------------------------------------------
// Times are float64 (frame stamps of long clips do not fit float32); one thread per sample finds its interval
// with np.searchsorted's "left" rule clamped to [0, T-2] and the weight; a second kernel blends the rows.
__global__ void interp_coeff_kernel(const double *sample, const double *orig, long long n_samples, long long n_orig,
                                    int *idx, float *w) {
    PMB_GRID_STRIDE(s, n_samples) {
        const double v = sample[s];
        long long lo = 0, hi = n_orig;  // first position with orig[pos] >= v
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (orig[mid] < v) lo = mid + 1;
            else hi = mid;
        }
        long long k = lo - 1;
        k = k < 0 ? 0 : (k > n_orig - 2 ? n_orig - 2 : k);
        idx[s] = static_cast<int>(k);
        w[s] = static_cast<float>((v - orig[k]) / (orig[k + 1] - orig[k]));
    }
}
// pos [outer][T][inner] -> out [outer][S][inner]
__global__ void interp_apply_kernel(const float *pos, const int *idx, const float *w, float *out, long long outer,
                                    long long n_orig, long long n_samples, long long inner) {
    const long long n = outer * n_samples * inner;
    PMB_GRID_STRIDE(i, n) {
        const long long m = i % inner, s = (i / inner) % n_samples, o = i / (inner * n_samples);
        const float ws = w[s];
        const float *row = pos + (o * n_orig + idx[s]) * inner + m;
        out[i] = __fadd_rn(__fmul_rn(1.f - ws, __ldg(row)), __fmul_rn(ws, __ldg(row + inner)));
    }
}

// ---- vector.normalize: v / (|v| + eps) over the last axis of length k --------------------------------------
// (a block-staged float4 variant for 3-vectors was measured SLOWER than this direct form: 2.66 against 3.83 TB/s)
__global__ void vec_normalize_kernel(const float *v, float eps, float *out, long long n, int k) {
    PMB_GRID_STRIDE(i, n) {
        const float *p = v + i * k;
        float n2 = 0.f;
        for (int c = 0; c < k; ++c) {
            const float x = __ldg(p + c);
            n2 += x * x;
        }
        const float d = sqrtf(n2) + eps;
        for (int c = 0; c < k; ++c) out[i * k + c] = __ldg(p + c) / d;
    }
}

// 3-vectors, four per thread: 48 contiguous bytes = three 16-byte loads in flight per thread instead of three 4-byte
// ones, same arithmetic as above (16-byte aligned arrays; the host sends the last n % 4 vectors through the generic
// kernel).
__global__ void vec3_normalize_x4_kernel(const float4 *__restrict__ v, float eps, float4 *__restrict__ out, long long n4) {
    PMB_GRID_STRIDE(i, n4) {
        const float4 a = __ldcs(v + 3 * i), b = __ldcs(v + 3 * i + 1), c = __ldcs(v + 3 * i + 2);
        // vectors: (a.x a.y a.z) (a.w b.x b.y) (b.z b.w c.x) (c.y c.z c.w)
        auto inv_len = [eps](float x, float y, float z) {
            float n2 = 0.f;
            n2 += x * x, n2 += y * y, n2 += z * z;
            return sqrtf(n2) + eps;
        };
        const float d0 = inv_len(a.x, a.y, a.z), d1 = inv_len(a.w, b.x, b.y), d2 = inv_len(b.z, b.w, c.x),
                    d3 = inv_len(c.y, c.z, c.w);
        __stcs(out + 3 * i, make_float4(a.x / d0, a.y / d0, a.z / d0, a.w / d1));
        __stcs(out + 3 * i + 1, make_float4(b.x / d1, b.y / d1, b.z / d2, b.w / d2));
        __stcs(out + 3 * i + 2, make_float4(c.x / d2, c.y / d3, c.z / d3, c.w / d3));
    }
}

}  // namespace pmb
