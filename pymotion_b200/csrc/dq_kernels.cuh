// Root-centred dual-quaternion encode / decode and the inverse-FK helper.
//   to_root_dual_quat     ops/skeleton.py:207-244  chain kernel, thread per frame
//   from_root_dual_quat   ops/skeleton.py:173-204  no chain: thread per (frame, joint)
//   from_global_rotations ops/skeleton.py:64-93    no chain: thread per (frame, joint)
#pragma once
#include "common.cuh"

namespace pmb {

// ---------------------------------------------------------------------------
// to_root_dual_quat.  Same thread-per-frame / warp-per-32-frames scheme as fk,
// but the chain state is (quaternion, translation) = 7 registers and a dual
// quaternion is exactly one 32-byte sector, so both staging and copy-out are
// 16-byte vector accesses: stage row stride 8C+4 words -> (2C+1) float4, odd,
// conflict-free for the per-thread STS.128.
// ---------------------------------------------------------------------------
template <int C>
struct DqTile {
    static constexpr int SQ = 2 * C + 1;                       // row stride in float4
    static constexpr int kStageBytesPerWarp = kWarp * SQ * 16;
    static constexpr int kSlotBytesPerWarp = 2 * kWarp * 16;   // (R, t) = 2 float4 per lane
    __host__ __device__ static constexpr int warp_bytes(int n_slots) {
        return kStageBytesPerWarp + n_slots * kSlotBytesPerWarp;
    }
};

template <int C, int WARPS>
__global__ void __launch_bounds__(WARPS *kWarp)
to_root_dq_kernel(const float4 *__restrict__ rot, const float *__restrict__ gpos, long long gstride,
                  const float *__restrict__ offsets, float4 *__restrict__ dq, long long n_frames, int n_joints,
                  int n_slots, const __grid_constant__ JointProgram prog) {
    using Tile = DqTile<C>;
    constexpr int SQ = Tile::SQ;
    static_assert((C & (C - 1)) == 0 && 2 * C <= 32, "C must be a power of two <= 16");
    constexpr int QW = 2 * C;          // float4 per full stage row
    constexpr int RPI = 32 / QW;       // rows copied per warp iteration

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tab = reinterpret_cast<float4 *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < n_joints; j += WARPS * kWarp)
        tab[j] = make_float4(offsets[3 * j], offsets[3 * j + 1], offsets[3 * j + 2], __uint_as_float(prog.code[j]));
    __syncthreads();

    const long long f0 = (static_cast<long long>(blockIdx.x) * WARPS + warp) * kWarp;
    if (f0 >= n_frames) return;
    const int nrows = static_cast<int>(min(static_cast<long long>(kWarp), n_frames - f0));
    const long long f = f0 + min(lane, nrows - 1);

    unsigned char *wbase = smem_raw + ((n_joints * 16 + 127) & ~127) + warp * Tile::warp_bytes(n_slots);
    float4 *stage = reinterpret_cast<float4 *>(wbase);
    float4 *slots = stage + kWarp * SQ;
    const float4 *qrow = rot + f * n_joints;

    Quat<float> cr{1.f, 0.f, 0.f, 0.f};
    Vec3<float> ct{0.f, 0.f, 0.f};

    for (int c0 = 0; c0 < n_joints; c0 += C) {
        const int cnt = min(C, n_joints - c0);
        float4 q[C];
#pragma unroll
        for (int jj = 0; jj < C; ++jj)
            if (jj < cnt) q[jj] = __ldg(qrow + c0 + jj);
#pragma unroll
        for (int jj = 0; jj < C; ++jj) {
            if (jj < cnt) {
                const int j = c0 + jj;
                const float4 e = tab[j];
                const uint32_t code = __float_as_uint(e.w);
                const Quat<float> r{q[jj].x, q[jj].y, q[jj].z, q[jj].w};
                if (jj == 0 && c0 == 0) {  // joint 0 carries the root's global rotation and position (:232)
                    const float *g = gpos + f * gstride;
                    cr = r;
                    ct = {__ldg(g), __ldg(g + 1), __ldg(g + 2)};
                } else if (prog_parent(code) == 0) {  // children of the root stay as they are (:236-237)
                    cr = r;
                    ct = {e.x, e.y, e.z};
                } else {
                    const uint32_t src = prog_src(code);
                    if (src != kSrcReg) {
                        const float4 a = slots[src * 2 * kWarp + lane], b = slots[(src * 2 + 1) * kWarp + lane];
                        cr = {a.x, a.y, a.z, a.w};
                        ct = {b.x, b.y, b.z};
                    }
                    const Vec3<float> v = q_rotate(cr, Vec3<float>{e.x, e.y, e.z});  // :238-240
                    ct = {v.x + ct.x, v.y + ct.y, v.z + ct.z};
                    cr = q_mul(cr, r);                                                  // :241
                }
                const uint32_t sv = prog_save(code);
                if (sv != kNoSave) {
                    slots[sv * 2 * kWarp + lane] = make_float4(cr.w, cr.x, cr.y, cr.z);
                    slots[(sv * 2 + 1) * kWarp + lane] = make_float4(ct.x, ct.y, ct.z, 0.f);
                }
                // dual_quat.py:28-35: q_d = 0.5 * ((0, t) (x) q_r)
                const Quat<float> d = q_mul(Quat<float>{0.f, ct.x, ct.y, ct.z}, cr);
                stage[lane * SQ + 2 * jj] = make_float4(cr.w, cr.x, cr.y, cr.z);
                stage[lane * SQ + 2 * jj + 1] = make_float4(0.5f * d.w, 0.5f * d.x, 0.5f * d.y, 0.5f * d.z);
            }
        }
        __syncwarp();
        const int col = lane & (QW - 1), sub = lane / QW;
        float4 *g = dq + ((f0 + sub) * n_joints + c0) * 2 + col;
        const float4 *s = stage + sub * SQ + col;
        const long long gstep = static_cast<long long>(RPI) * n_joints * 2;
        if (col < 2 * cnt) {
#pragma unroll 4
            for (int r = sub; r < nrows; r += RPI) {
                __stcs(g, *s);
                g += gstep, s += RPI * SQ;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Tiles of FB frames x J joints handled by one block, one thread per element.
// i / J for i < 2^16 through a multiply-high (magic = floor(2^32 / J) + 1).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int div_small(int i, uint32_t magic) { return static_cast<int>(__umulhi(static_cast<uint32_t>(i), magic)); }

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
from_root_dq_kernel(const float4 *__restrict__ dq, float *__restrict__ trans, float4 *__restrict__ rots,
                    long long n_frames, int n_joints, int frames_per_block, uint32_t magic,
                    const __grid_constant__ JointProgram prog) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tstage = reinterpret_cast<float *>(smem_raw);                                      // [FB*J*3]
    short *par = reinterpret_cast<short *>(smem_raw + ((frames_per_block * n_joints * 12 + 15) & ~15));  // [J]
    for (int j = threadIdx.x; j < n_joints; j += THREADS) par[j] = static_cast<short>(prog_parent(prog.code[j]));
    __syncthreads();

    const long long fbase = static_cast<long long>(blockIdx.x) * frames_per_block;
    const int nf = static_cast<int>(min(static_cast<long long>(frames_per_block), n_frames - fbase));
    const int n_el = nf * n_joints;
    const float4 *dtile = dq + fbase * n_joints * 2;
    float4 *rtile = rots + fbase * n_joints;

    for (int i = threadIdx.x; i < n_el; i += THREADS) {
        const int fl = div_small(i, magic);
        const int j = i - fl * n_joints;
        const float4 a = __ldg(dtile + 2 * i), b = __ldg(dtile + 2 * i + 1);  // L1-allocating: a neighbour re-reads it as its parent
        Quat<float> r{a.x, a.y, a.z, a.w};
        // dual_quat.py:82: t = vector part of 2 * (q_d (x) conj(q_r))
        Quat<float> m = q_mul(Quat<float>{b.x, b.y, b.z, b.w}, q_conj(r));
        Vec3<float> t{2.f * m.x, 2.f * m.y, 2.f * m.z};
        const int p = par[j];
        if (j > 0 && p != 0) {  // :196-203, parent taken in ROOT space (it has not been localised yet)
            const int ip = i - j + p;
            const float4 pa = __ldg(dtile + 2 * ip), pb = __ldg(dtile + 2 * ip + 1);
            const Quat<float> pr{pa.x, pa.y, pa.z, pa.w};
            const Quat<float> pm = q_mul(Quat<float>{pb.x, pb.y, pb.z, pb.w}, q_conj(pr));
            const Quat<float> inv = q_conj(pr);
            t = q_rotate(inv, Vec3<float>{t.x - 2.f * pm.x, t.y - 2.f * pm.y, t.z - 2.f * pm.z});
            r = q_mul(inv, r);
        }
        __stcs(rtile + i, make_float4(r.w, r.x, r.y, r.z));
        tstage[3 * i] = t.x, tstage[3 * i + 1] = t.y, tstage[3 * i + 2] = t.z;
    }
    __syncthreads();
    // translations: 12-byte records -> one flat, 16-byte aligned span per tile (fbase % 4 == 0)
    float *ttile = trans + fbase * n_joints * 3;
    const int n_w = 3 * n_el, n_v = n_w >> 2;
    for (int i = threadIdx.x; i < n_v; i += THREADS)
        __stcs(reinterpret_cast<float4 *>(ttile) + i, reinterpret_cast<const float4 *>(tstage)[i]);
    for (int i = 4 * n_v + threadIdx.x; i < n_w; i += THREADS) ttile[i] = tstage[i];
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
from_global_rotations_kernel(const float4 *__restrict__ gq, float4 *__restrict__ lq, long long n_frames,
                             int n_joints, int frames_per_block, uint32_t magic,
                             const __grid_constant__ JointProgram prog) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    short *par = reinterpret_cast<short *>(smem_raw);
    for (int j = threadIdx.x; j < n_joints; j += THREADS) par[j] = static_cast<short>(prog_parent(prog.code[j]));
    __syncthreads();
    const long long fbase = static_cast<long long>(blockIdx.x) * frames_per_block;
    const int nf = static_cast<int>(min(static_cast<long long>(frames_per_block), n_frames - fbase));
    const int n_el = nf * n_joints;
    const float4 *gt = gq + fbase * n_joints;
    float4 *lt = lq + fbase * n_joints;
    for (int i = threadIdx.x; i < n_el; i += THREADS) {
        const int fl = div_small(i, magic);
        const int j = i - fl * n_joints;
        const float4 a = __ldg(gt + i);
        Quat<float> r{a.x, a.y, a.z, a.w};
        if (j > 0) {
            const float4 pa = __ldg(gt + i - j + par[j]);
            r = q_mul(q_conj(Quat<float>{pa.x, pa.y, pa.z, pa.w}), r);
        }
        __stcs(lt + i, make_float4(r.w, r.x, r.y, r.z));
    }
}

}  // namespace pmb
