// fk emitting global QUATERNIONS instead of rotation matrices: what every in-repo consumer
// of the reference's fk computes next (quat.from_matrix(rotmats), ops/skeleton.py:322-323, :140, :410-411).
//
// No matrix is ever formed: the chain state is (global quaternion, position) = 7 registers,
//     Q_j = Q_p (x) q^_j            p_j = p_p + Q_p (*) offset_j            q^ = q / (|q| + eps)
// and the sign of Q_j is made the one quat.from_matrix would return (the component it pivots on is >= 0,
// quat.py:111-155, with the branch conditions m22 < 0, m00 > m11, m00 < -m11 written in terms of the unit
// quaternion: x^2 + y^2 > 1/2, x^2 > y^2, z^2 > w^2).
//
// Same machinery as to_root_dual_quat (dq_kernels.cuh): thread per frame, persistent warp per tile of 32
// frames, quaternions through double-buffered TMA boxes, joint program with registers / slots, outputs staged
// per warp in shared memory (quaternions as float4 rows with an odd float4 stride, positions as float rows
// with an odd stride: conflict-free thread-per-frame stores) and copied out by the whole warp in groups of
// `group` joints.  Algorithmic HBM traffic: 44*J + 12 bytes per pose (16J + 12 in, 16J + 12J out).
#pragma once
#include "common.cuh"
#include "fk_kernel.cuh"  // copy_out_periodic
#include "tma.cuh"

namespace pmb {

// Full flush groups of the common sizes go out with the short-period lane map of fk_kernel.cuh (a lane's stage /
// global offsets repeat every few stores, so each store costs LDS + STG instead of a division and two multiplies:
// the generic loop was ~30 % of this kernel's instructions).
template <int G, bool POS>
__device__ __forceinline__ void fkq_flush_full(const float4 *qstage, const float *pstage, float4 *gq, float *gp, int n_joints,
                                               int lane) {
    copy_out_periodic<4 * G, 4 * (G | 1), 4>(reinterpret_cast<const float *>(qstage), reinterpret_cast<float *>(gq), 4 * n_joints, lane);
    if (POS) copy_out_periodic<3 * G, (3 * G) | 1, 1>(pstage, gp, 3 * n_joints, lane);
}

struct FkqGeom {
    int stride4;      // quaternion stage row stride in float4
    int stride_p;     // position stage row stride in floats
    int warp_bytes;   // stages + slots of one warp
    int block_bytes;
};
__host__ __device__ inline FkqGeom fkq_geom(int group, int warps, int n_joints, int n_slots) {
    FkqGeom g;
    g.stride4 = group | 1;
    g.stride_p = (3 * group) | 1;
    g.warp_bytes = kWarp * g.stride4 * 16 + ((kWarp * g.stride_p * 4 + 15) & ~15) + n_slots * 2 * kWarp * 16;
    g.block_bytes = 1024 + warps * kBoxStages * kBoxBytes + ((n_joints * 16 + 127) & ~127) + warps * g.warp_bytes +
                    warps * kBoxStages * 8 + warps * kWarp * 4;
    return g;
}

// The sign quat.from_matrix gives the quaternion of this rotation.
__device__ __forceinline__ Quat<float> q_from_matrix_sign(const Quat<float> &q) {
    const float xx = q.x * q.x, yy = q.y * q.y;
    const float pivot = (xx + yy > 0.5f) ? (xx > yy ? q.x : q.y) : (q.z * q.z > q.w * q.w ? q.z : q.w);
    return pivot < 0.f ? Quat<float>{-q.w, -q.x, -q.y, -q.z} : q;
}

// POS = false: rotations only (positions == nullptr; what mirror needs: no translation chain, no position stage)
template <int WARPS, bool POS>
__global__ void __launch_bounds__(WARPS *kWarp)
fk_quat_chain_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
                     const float *__restrict__ offsets, float *__restrict__ pos, float4 *__restrict__ grot,
                     long long n_frames, int n_joints, int n_slots, int group, uint32_t magic_q_full,
                     uint32_t magic_q_tail, uint32_t magic_p_full, uint32_t magic_p_tail,
                     const __grid_constant__ JointProgram prog) {
    constexpr int C = kChunk;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const FkqGeom geo = fkq_geom(group, WARPS, n_joints, n_slots);
    const int S4 = geo.stride4, SP = geo.stride_p;

    float4 *in_stage = reinterpret_cast<float4 *>(smem_raw + warp * kBoxStages * kBoxBytes);
    float4 *tab = reinterpret_cast<float4 *>(smem_raw + WARPS * kBoxStages * kBoxBytes);
    unsigned char *after_tab = reinterpret_cast<unsigned char *>(tab) + ((n_joints * 16 + 127) & ~127);
    float4 *qstage = reinterpret_cast<float4 *>(after_tab + warp * geo.warp_bytes);
    float *pstage = reinterpret_cast<float *>(qstage + kWarp * S4);
    float4 *slots = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(pstage) + ((kWarp * SP * 4 + 15) & ~15));
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
                for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
            }
            __syncwarp();
            const long long next_tile = tile + tile_stride;
            if (lane == 0) issue_next(buf);
            if (last_chunk && next_tile < n_tiles) {
                const float *g = gpos + min(next_tile * kWarp + lane, n_frames - 1) * gstride;
                gnext[0] = __ldg(g), gnext[1] = __ldg(g + 1), gnext[2] = __ldg(g + 2);
            }

            float4 *qs = qstage + lane * S4 + gj;
            float *ps = pstage + lane * SP + 3 * gj;
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                if (jj < cnt) {
                    const int j = c0 + jj;
                    const uint32_t code = prog.code[j];  // constant bank, warp-uniform
                    const float4 e = tab[j];
                    // the reference turns q / (|q| + eps) into a MATRIX: the zero quaternion becomes the identity
                    // (quat.py:293-315 with all products zero), so it has to here too
                    Quat<float> r = q_normalize_fast(Quat<float>{q[jj].x, q[jj].y, q[jj].z, q[jj].w}, 1e-8f);
                    if (r.w == 0.f && r.x == 0.f && r.y == 0.f && r.z == 0.f) r.w = 1.f;
                    if (jj == 0 && c0 == 0) {  // root: [R(q^_0) | global_pos] (skeleton.py:49), offsets[0] ignored
                        cr = r;
                    } else {
                        const uint32_t src = prog_src(code);
                        if (src != kSrcReg) {
                            const float4 a = slots[src * 2 * kWarp + lane], b = slots[(src * 2 + 1) * kWarp + lane];
                            cr = {a.x, a.y, a.z, a.w};
                            ct = {b.x, b.y, b.z};
                        }
                        if (POS) {
                            const Vec3<float> v = q_rotate(cr, Vec3<float>{e.x, e.y, e.z});
                            ct = {v.x + ct.x, v.y + ct.y, v.z + ct.z};
                        }
                        cr = q_mul(cr, r);
                    }
                    const uint32_t sv = prog_save(code);
                    if (sv != kNoSave) {
                        slots[sv * 2 * kWarp + lane] = make_float4(cr.w, cr.x, cr.y, cr.z);
                        slots[(sv * 2 + 1) * kWarp + lane] = make_float4(ct.x, ct.y, ct.z, 0.f);
                    }
                    const Quat<float> o = q_from_matrix_sign(cr);
                    qs[jj] = make_float4(o.w, o.x, o.y, o.z);
                    if (POS) ps[3 * jj] = ct.x, ps[3 * jj + 1] = ct.y, ps[3 * jj + 2] = ct.z;
                }
            }
            gj += cnt;

            if (gj == group || last_chunk) {
                __syncwarp();
                const int g0 = c0 + cnt - gj;  // first joint of the group
                float4 *gq = grot + f0 * n_joints + g0;
                float *gpp = POS ? pos + (f0 * n_joints + g0) * 3 : nullptr;
                const bool full = gj == group && nrows == kWarp;
                if (full && group == 16) {
                    fkq_flush_full<16, POS>(qstage, pstage, gq, gpp, n_joints, lane);
                } else if (full && group == 8) {
                    fkq_flush_full<8, POS>(qstage, pstage, gq, gpp, n_joints, lane);
                } else if (full && group == 24) {
                    fkq_flush_full<24, POS>(qstage, pstage, gq, gpp, n_joints, lane);
                } else if (full && group == 32) {
                    fkq_flush_full<32, POS>(qstage, pstage, gq, gpp, n_joints, lane);
                } else {
                    {   // quaternions: rows of gj float4 -> global rows of pitch n_joints float4
                        const uint32_t magic = (gj == group) ? magic_q_full : magic_q_tail;
                        const int n4 = nrows * gj;
#pragma unroll 4
                        for (int i = lane; i < n4; i += kWarp) {
                            // (a remainder group of ONE joint has no 32-bit magic: 2^32 / 1 + 1 wraps)
                            const int r = gj == 1 ? i : static_cast<int>(__umulhi(static_cast<uint32_t>(i), magic));
                            const int c = i - r * gj;
                            gq[static_cast<long long>(r) * n_joints + c] = qstage[r * S4 + c];
                        }
                    }
                    if (POS) {  // positions: rows of 3*gj floats -> global rows of pitch 3*n_joints floats
                        const uint32_t magic = (gj == group) ? magic_p_full : magic_p_tail;
                        const int w = 3 * gj, n1 = nrows * w;
                        const int pitch = 3 * n_joints;
#pragma unroll 4
                        for (int i = lane; i < n1; i += kWarp) {
                            const int r = static_cast<int>(__umulhi(static_cast<uint32_t>(i), magic));
                            const int c = i - r * w;
                            gpp[static_cast<long long>(r) * pitch + c] = pstage[r * SP + c];
                        }
                    }
                }
                __syncwarp();
                gj = 0;
            }
        }
    }
}

}  // namespace pmb
