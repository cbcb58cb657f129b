// Forward kinematics, one fused kernel (ops/skeleton.py:16-61 of the reference).
//
// Mapping: one THREAD per frame, one WARP per tile of 32 consecutive frames.
//   * the joint table (offsets + per-joint program word) is staged in shared
//     memory once per block;
//   * a thread streams its frame's quaternions from HBM as float4 (C joints =
//     C independent 16-byte loads in flight per thread), normalises, builds the
//     local 3x3 and composes it with its parent's global transform, which is
//     either still in registers (parent == previous joint: the chain case) or in
//     a per-warp shared-memory slot written when the parent was computed (branch
//     points of the tree; slots are allocated on the host, see joint_program.cuh);
//   * results of C joints are staged in a per-warp shared-memory tile with an
//     ODD row stride (conflict-free for the thread-per-frame writes), then
//     written to HBM as fully coalesced 128-byte runs by the whole warp.
// Warps never synchronise with each other after the table load: each walks its
// own load / compose / store phases so the SM overlaps them across warps.
//
// Algorithmic HBM traffic: 64*J + 12 bytes per pose (16J in, 48J + 12 ... out),
// no scratch in global memory.  See DESIGN.md for the roofline.
#pragma once
#include "common.cuh"

namespace pmb {

template <int C, int RW>
struct FkTile {
    static constexpr int SR = (RW * C) | 1;  // row stride (words) of the rotation stage, odd
    static constexpr int SP = (3 * C) | 1;   // row stride (words) of the position stage, odd
    static constexpr int kStageBytesPerWarp = kWarp * (SR + SP) * 4;
    static constexpr int kSlotBytesPerWarp = 3 * kWarp * 16;  // one slot = 12 floats x 32 lanes
    __host__ __device__ static constexpr int warp_bytes(int n_slots) {
        return kStageBytesPerWarp + n_slots * kSlotBytesPerWarp;
    }
};

// QUAT_OUT = false: rout is rotmats [F][J][9];  true: rout is global quaternions [F][J][4].
template <int C, int WARPS, bool PF_OFFSETS, bool QUAT_OUT>
__global__ void __launch_bounds__(WARPS *kWarp)
fk_chain_kernel(const float4 *__restrict__ rot, const float *__restrict__ gpos, long long gstride,
                const float *__restrict__ offsets, long long ostride, float *__restrict__ pos,
                float *__restrict__ rout, long long n_frames, int n_joints, int n_slots,
                const __grid_constant__ JointProgram prog) {
    constexpr int RW = QUAT_OUT ? 4 : 9;
    using Tile = FkTile<C, RW>;
    constexpr int SR = Tile::SR, SP = Tile::SP;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tab = reinterpret_cast<float4 *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < n_joints; j += WARPS * kWarp) {
        float4 e;
        if (PF_OFFSETS) {
            e.x = e.y = e.z = 0.f;
        } else {
            e.x = offsets[3 * j], e.y = offsets[3 * j + 1], e.z = offsets[3 * j + 2];
        }
        e.w = __uint_as_float(prog.code[j]);
        tab[j] = e;
    }
    __syncthreads();

    const long long f0 = (static_cast<long long>(blockIdx.x) * WARPS + warp) * kWarp;
    if (f0 >= n_frames) return;
    const int nrows = static_cast<int>(min(static_cast<long long>(kWarp), n_frames - f0));
    const long long f = f0 + min(lane, nrows - 1);  // tail lanes recompute the last frame, never stored

    unsigned char *wbase = smem_raw + ((n_joints * 16 + 127) & ~127) + warp * Tile::warp_bytes(n_slots);
    float *Rst = reinterpret_cast<float *>(wbase);
    float *Pst = Rst + kWarp * SR;
    float4 *slots = reinterpret_cast<float4 *>(Pst + kWarp * SP);

    const float4 *qrow = rot + f * n_joints;
    const float *orow = PF_OFFSETS ? offsets + f * ostride : nullptr;
    Xform<float> cur;

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
                float ox = e.x, oy = e.y, oz = e.z;
                if (PF_OFFSETS) {
                    ox = __ldg(orow + 3 * j), oy = __ldg(orow + 3 * j + 1), oz = __ldg(orow + 3 * j + 2);
                }
                float l[9];
                q_to_matrix(q_normalize(Quat<float>{q[jj].x, q[jj].y, q[jj].z, q[jj].w}, 1e-8f), l);
                if (jj == 0 && c0 == 0) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) cur.r[k] = l[k];
                    const float *g = gpos + f * gstride;
                    cur.p[0] = __ldg(g), cur.p[1] = __ldg(g + 1), cur.p[2] = __ldg(g + 2);
                } else {
                    const uint32_t src = prog_src(code);
                    if (src != kSrcReg) {  // warp-uniform: every lane runs the same program
                        const float4 *s = slots + src * 3 * kWarp + lane;
                        const float4 a = s[0], b = s[kWarp], c = s[2 * kWarp];
                        cur.r[0] = a.x, cur.r[1] = a.y, cur.r[2] = a.z, cur.r[3] = a.w;
                        cur.r[4] = b.x, cur.r[5] = b.y, cur.r[6] = b.z, cur.r[7] = b.w;
                        cur.r[8] = c.x, cur.p[0] = c.y, cur.p[1] = c.z, cur.p[2] = c.w;
                    }
                    xf_compose(cur, cur, l, ox, oy, oz);
                }
                const uint32_t sv = prog_save(code);
                if (sv != kNoSave) {
                    float4 *s = slots + sv * 3 * kWarp + lane;
                    s[0] = make_float4(cur.r[0], cur.r[1], cur.r[2], cur.r[3]);
                    s[kWarp] = make_float4(cur.r[4], cur.r[5], cur.r[6], cur.r[7]);
                    s[2 * kWarp] = make_float4(cur.r[8], cur.p[0], cur.p[1], cur.p[2]);
                }
                float *rs = Rst + lane * SR + RW * jj;
                if (QUAT_OUT) {
                    const Quat<float> gq = q_from_matrix(cur.r);
                    rs[0] = gq.w, rs[1] = gq.x, rs[2] = gq.y, rs[3] = gq.z;
                } else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) rs[k] = cur.r[k];
                }
                float *ps = Pst + lane * SP + 3 * jj;
                ps[0] = cur.p[0], ps[1] = cur.p[1], ps[2] = cur.p[2];
            }
        }
        __syncwarp();

        // coalesced copy-out: each warp store covers one contiguous run of a frame's row
        const int wr = RW * cnt, wp = 3 * cnt;
        float *rg = rout + (f0 * n_joints + c0) * RW + lane;
        float *pg = pos + (f0 * n_joints + c0) * 3 + lane;
        const long long rrow = static_cast<long long>(n_joints) * RW, prow = static_cast<long long>(n_joints) * 3;
        const float *rsm = Rst + lane, *psm = Pst + lane;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) {
#pragma unroll
            for (int u = 0; u < (RW * C + 31) / 32; ++u)
                if (lane + 32 * u < wr) __stcs(rg + 32 * u, rsm[32 * u]);
            if (lane < wp) __stcs(pg, psm[0]);
            rg += rrow, pg += prow, rsm += SR, psm += SP;
        }
        __syncwarp();
    }
}

}  // namespace pmb
