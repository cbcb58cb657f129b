// Root-centred dual-quaternion encode / decode and the inverse-FK helper.
//   to_root_dual_quat     ops/skeleton.py:207-244  chain kernel, thread per frame
//   from_root_dual_quat   ops/skeleton.py:173-204  no chain: thread per (frame, joint)
//   from_global_rotations ops/skeleton.py:64-93    no chain: thread per (frame, joint)
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace pmb {

// ---------------------------------------------------------------------------
// to_root_dual_quat.  Same machinery as fk (fk_kernel.cuh): thread per frame, persistent warp per tile of
// 32 frames, quaternions through double-buffered TMA boxes, joint program with registers / slots.  The chain
// state is (quaternion, translation) = 7 registers and a dual quaternion is exactly one 32-byte sector, so
// the stage holds `group` joints per row as 2*group float4 plus one float4 of padding (odd stride ->
// conflict-free STS.128) and is copied out with 16-byte accesses.  group = n_joints (whole rows) makes the
// tile's output one contiguous span; smaller groups are used when that does not fit in shared memory.
// ---------------------------------------------------------------------------
struct DqGeom {
    int stride4;      // stage row stride in float4
    int warp_bytes;   // stage + slots of one warp
    int block_bytes;
};
__host__ __device__ inline DqGeom dq_geom(int group, int warps, int n_joints, int n_slots) {
    DqGeom g;
    g.stride4 = 2 * group + 1;
    g.warp_bytes = kWarp * g.stride4 * 16 + n_slots * 2 * kWarp * 16;
    g.block_bytes = 1024 + warps * kBoxStages * kBoxBytes + ((n_joints * 16 + 127) & ~127) + warps * g.warp_bytes +
                    warps * kBoxStages * 8 + warps * kWarp * 4;
    return g;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS *kWarp)
to_root_dq_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
                  const float *__restrict__ offsets, float4 *__restrict__ dq, long long n_frames, int n_joints,
                  int n_slots, int group, uint32_t magic_full, uint32_t magic_tail,
                  const __grid_constant__ JointProgram prog) {
    constexpr int C = kChunk;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const DqGeom geo = dq_geom(group, WARPS, n_joints, n_slots);
    const int S4 = geo.stride4;

    float4 *in_stage = reinterpret_cast<float4 *>(smem_raw + warp * kBoxStages * kBoxBytes);
    float4 *tab = reinterpret_cast<float4 *>(smem_raw + WARPS * kBoxStages * kBoxBytes);
    unsigned char *after_tab = reinterpret_cast<unsigned char *>(tab) + ((n_joints * 16 + 127) & ~127);
    float4 *stage = reinterpret_cast<float4 *>(after_tab + warp * geo.warp_bytes);
    float4 *slots = stage + kWarp * S4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(after_tab + WARPS * geo.warp_bytes);
    const uint32_t bar0 = smem_u32(bars + warp * kBoxStages);
    const uint32_t in0 = smem_u32(in_stage);
    const uint32_t fence_word = smem_u32(reinterpret_cast<uint32_t *>(bars + WARPS * kBoxStages) + threadIdx.x);

    const long long n_tiles = (n_frames + kWarp - 1) / kWarp;
    const long long tile_stride = static_cast<long long>(gridDim.x) * WARPS;
    long long tile = static_cast<long long>(blockIdx.x) * WARPS + warp;
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kBoxStages; ++b) mbar_init(bar0 + 8 * b, 1);
        fence_barrier_init();
    }
    __syncwarp();
    long long la_tile = tile;
    int la_c0 = 0;
    auto issue_next = [&](int buf) {  // lane 0 only: the warp's chunks in processing order, across its tiles
        if (la_tile < n_tiles) {
            mbar_arrive_expect_tx(bar0 + 8 * buf, kBoxBytes);
            tma_load_2d(in0 + buf * kBoxBytes, &tm_rot, 4 * la_c0, static_cast<int>(la_tile * kWarp), bar0 + 8 * buf);
            la_c0 += C;
            if (la_c0 >= n_joints) la_c0 = 0, la_tile += tile_stride;
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kBoxStages; ++b) issue_next(b);
    }
    for (int j = threadIdx.x; j < n_joints; j += WARPS * kWarp)
        tab[j] = make_float4(offsets[3 * j], offsets[3 * j + 1], offsets[3 * j + 2], 0.f);
    __syncthreads();

    const int swz = lane & 7;
    uint32_t kchunk = 0;
    float gnext[3] = {0.f, 0.f, 0.f};
    if (tile < n_tiles) {
        const float *g = gpos + min(tile * kWarp + lane, n_frames - 1) * gstride;
        gnext[0] = __ldg(g), gnext[1] = __ldg(g + 1), gnext[2] = __ldg(g + 2);
    }

    for (; tile < n_tiles; tile += tile_stride) {
        const long long f0 = tile * kWarp;
        const int nrows = static_cast<int>(min(static_cast<long long>(kWarp), n_frames - f0));
        Quat<float> cr{1.f, 0.f, 0.f, 0.f};
        Vec3<float> ct{gnext[0], gnext[1], gnext[2]};
        int gj = 0;

        for (int c0 = 0; c0 < n_joints; c0 += C) {
            const int cnt = min(C, n_joints - c0);
            const bool last_chunk = c0 + C >= n_joints;
            const int buf = kchunk % kBoxStages;
            mbar_wait(bar0 + 8 * buf, (kchunk / kBoxStages) & 1);
            ++kchunk;
            const float4 *in_row = in_stage + buf * (kBoxBytes / 16) + lane * C;
            float4 q[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) q[jj] = in_row[jj ^ swz];
            {   // loads must have LANDED before the box is refilled through the async proxy (see fk_kernel.cuh)
                uint32_t acc = 0;
#pragma unroll
                for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x) | __float_as_uint(q[jj].w);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
            }
            __syncwarp();
            const long long next_tile = tile + tile_stride;
            if (lane == 0) issue_next(buf);
            if (last_chunk && next_tile < n_tiles) {
                const float *g = gpos + min(next_tile * kWarp + lane, n_frames - 1) * gstride;
                gnext[0] = __ldg(g), gnext[1] = __ldg(g + 1), gnext[2] = __ldg(g + 2);
            }

            float4 *st = stage + lane * S4 + 2 * gj;
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                if (jj < cnt) {
                    const int j = c0 + jj;
                    const uint32_t code = prog.code[j];  // constant bank, warp-uniform
                    const float4 e = tab[j];
                    const Quat<float> r{q[jj].x, q[jj].y, q[jj].z, q[jj].w};
                    if (jj == 0 && c0 == 0) {  // joint 0 carries the root's global rotation and position (:232)
                        cr = r;
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
                    st[2 * jj] = make_float4(cr.w, cr.x, cr.y, cr.z);
                    st[2 * jj + 1] = make_float4(0.5f * d.w, 0.5f * d.x, 0.5f * d.y, 0.5f * d.z);
                }
            }
            gj += cnt;

            if (gj == group || last_chunk) {
                __syncwarp();
                // rows of 2*gj float4 -> global rows of pitch 2*n_joints float4; flat unit i = row * (2*gj) + col
                const int w4 = 2 * gj;
                const uint32_t magic = (gj == group) ? magic_full : magic_tail;
                const int n4 = nrows * w4;
                float4 *g = dq + (f0 * n_joints + (c0 + cnt - gj)) * 2;
                const int pitch4 = 2 * n_joints;
#pragma unroll 4
                for (int i = lane; i < n4; i += kWarp) {
                    const int r = static_cast<int>(__umulhi(static_cast<uint32_t>(i), magic));
                    const int c = i - r * w4;
                    g[static_cast<long long>(r) * pitch4 + c] = stage[r * S4 + c];
                }
                __syncwarp();
                gj = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Tiles of FB frames x J joints handled by one block, one thread per element.
// i / J for i < 2^16 through a multiply-high (magic = floor(2^32 / J) + 1; J = 1 has no 32-bit magic and is
// passed as 0).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int div_small(int i, uint32_t magic) {
    return magic ? static_cast<int>(__umulhi(static_cast<uint32_t>(i), magic)) : i;
}

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
