// Forward kinematics, row-team kernel (ops/skeleton.py:16-61 of the reference).
//
// The rows of a global transform are independent chains: with G = P * [L | off],
//     G[a][:] = P[a][:] * L          p[a] = P[a][:] . off + p_parent[a]          a = 0, 1, 2
// so row a of every joint only ever needs row a of its ancestors.  A TEAM of three compute warps owns a
// tile of 32 consecutive frames; warp a walks the tree for row a, lane = frame.  Compared with one thread
// per frame (fk_kernel.cuh) the dependent chain per thread is a third as long, the same shared memory
// carries three times as many warps (what limits 52- and 65-joint skeletons: the stage is 1536 J bytes per
// tile whatever the mapping), and the chain state is 4 registers.  No local matrix is formed: a row times
// R(q^) is the row rotated by the conjugate quaternion (two cross products), and the normalisation of q
// collapses into one scale 2 / (|q| + eps)^2 per joint, computed for a whole chunk ahead of the branchy
// tree walk.
//
//   in    a loader thread streams the tile's quaternions as TMA boxes of 8 joints x 32 frames (128-byte
//         swizzle) through an S-deep ring; full / empty mbarriers, the three row warps release a box as soon
//         as its quaternions are in registers;
//   walk  parents[] compiled on the host (joint_program.h); the parent row is either still in registers
//         (parent == previous joint) or read back from the stage -- the stage is the dense image of the
//         tile's output, so it already holds every ancestor: no slots;
//   out   the stage (32 x 36J and 32 x 12J bytes, both multiples of 128) is handed to the TMA engine by a
//         drainer thread as two contiguous line-aligned bulk stores while the row warps wait for
//         `stage_free`; the other teams on the SM cover the drain.
//
// Block = 1 team = 5 warps: rows 0..2, loader, drainer.  Algorithmic HBM traffic 64*J + 12 bytes per pose.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"

namespace pmb {

constexpr int kRowThreads = 160;

struct FkRowsGeom {
    int tab_bytes, stage_bytes, block_bytes;
};
__host__ __device__ inline FkRowsGeom fk_rows_geom(int stages, int n_joints, bool quat_out = false) {
    FkRowsGeom g;
    g.tab_bytes = (n_joints * 16 + 127) & ~127;
    g.stage_bytes = kWarp * (quat_out ? 7 : 12) * n_joints * 4;
    // 1 KB slack to align the boxes | boxes | joint table | stage | barriers (full[S], empty[S], stage_full, stage_free) | fence words
    g.block_bytes = 1024 + stages * kBoxBytes + g.tab_bytes + g.stage_bytes + (2 * stages + 2) * 8 + 96 * 4;
    g.block_bytes = (g.block_bytes + 15) & ~15;
    return g;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float a) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}

// 2 / (|q| + eps)^2 with three bare MUFU-class ops: the scale that turns the products of the RAW quaternion
// into those of q / (|q| + eps) (quat.py:411-423; eps joins the NORM).  sqrt.approx(0) = 0, so the zero
// quaternion needs no guard: its products are all zero and the local rotation is the identity, as in the
// reference.
__device__ __forceinline__ float rot_scale(const float4 &q, float eps) {
    const float n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float n, inv;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(n) : "f"(n2));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(n + eps));
    return (inv + inv) * inv;
}

// VEC = 2 (even joint count: every stage row is 8-byte aligned): a row's three numbers go out as one 64-bit
// and one 32-bit store -- the minimum number of shared-memory wavefronts for this layout; VEC = 1: 32-bit
// stores (odd row stride, conflict free as they are).
template <int S, int VEC>
__global__ void __launch_bounds__(kRowThreads, 5)
fk_rows_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
               const float *__restrict__ offsets, float *__restrict__ pos, float *__restrict__ rout,
               long long n_frames, int n_joints, const __grid_constant__ JointProgram prog) {
    constexpr int C = kChunk;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const FkRowsGeom geo = fk_rows_geom(S, n_joints);

    float4 *boxes = reinterpret_cast<float4 *>(smem_raw);
    float4 *tab = reinterpret_cast<float4 *>(smem_raw + S * kBoxBytes);
    float *Rst = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(tab) + geo.tab_bytes);
    float *Pst = Rst + kWarp * 9 * n_joints;
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(Rst) + geo.stage_bytes);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, stage_full = full0 + 16 * S,
                   stage_free = stage_full + 8;
    const uint32_t box0 = smem_u32(boxes);

    const long long n_tiles = (n_frames + kWarp - 1) / kWarp;
    const long long tile_stride = gridDim.x;
    const int rpitch = 9 * n_joints, ppitch = 3 * n_joints;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < S; ++b) mbar_init(full0 + 8 * b, 1), mbar_init(empty0 + 8 * b, 3);
        mbar_init(stage_full, 3);
        mbar_init(stage_free, 1);
        fence_barrier_init();
    }
    // offsets[0] is ignored by the reference (the root translation is global_pos, skeleton.py:49): with a zero
    // entry the root is an ordinary joint whose parent is the identity placed at global_pos
    for (int j = threadIdx.x; j < n_joints; j += kRowThreads)
        tab[j] = j == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(offsets[3 * j], offsets[3 * j + 1], offsets[3 * j + 2], 0.f);
    __syncthreads();  // the only block-wide barrier: the roles below never meet again

    if (warp == 3) {
        // ---- loader: the team's chunks in processing order, across its tiles -----------------------
        if (lane == 0) {
            uint32_t k = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += tile_stride) {
                for (int c0 = 0; c0 < n_joints; c0 += C, ++k) {
                    const uint32_t buf = k % S;
                    if (k >= S) mbar_wait(empty0 + 8 * buf, ((k / S) - 1) & 1);  // all three row warps have read it
                    mbar_arrive_expect_tx(full0 + 8 * buf, kBoxBytes);
                    tma_load_2d(box0 + buf * kBoxBytes, &tm_rot, 4 * c0, static_cast<int>(t * kWarp), full0 + 8 * buf);
                }
            }
        }
        return;
    }
    if (warp == 4) {
        // ---- drainer: stage -> HBM through the TMA engine ------------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += tile_stride, ++it) {
                const long long f0 = t * kWarp;
                const int nrows = static_cast<int>(min(static_cast<long long>(kWarp), n_frames - f0));
                mbar_wait(stage_full, it & 1);
                float *rg = rout + f0 * rpitch, *pg = pos + f0 * ppitch;
                const uint32_t rbytes = static_cast<uint32_t>(nrows * rpitch * 4), pbytes = static_cast<uint32_t>(nrows * ppitch * 4);
                // a full tile is two multiples of 128 bytes; a remainder tile can leave up to 3 words past the
                // last 16-byte unit, stored directly
                if (rbytes & ~15u) bulk_store(rg, smem_u32(Rst), rbytes & ~15u);
                if (pbytes & ~15u) bulk_store(pg, smem_u32(Pst), pbytes & ~15u);
                bulk_commit();
                for (uint32_t w = (rbytes & ~15u) / 4; w < rbytes / 4; ++w) rg[w] = Rst[w];
                for (uint32_t w = (pbytes & ~15u) / 4; w < pbytes / 4; ++w) pg[w] = Pst[w];
                bulk_wait_read0();  // the engine has read the stage: the row warps may overwrite it
                mbar_arrive(stage_free);
            }
            bulk_wait0();  // global writes of the last tile are complete at exit
        }
        return;
    }

    // ---- row warps ----------------------------------------------------------------------------------
    const int a = warp;  // the row of the transform this warp computes
    const uint32_t fence_word = stage_free + 8 + 4 * threadIdx.x;
    const int swz = lane & 7;
    float *Rrow = Rst + lane * rpitch + 3 * a;
    float *Prow = Pst + lane * ppitch + a;
    const int pa = a & 1;
    const float id0 = a == 0 ? 1.f : 0.f, id1 = a == 1 ? 1.f : 0.f, id2 = a == 2 ? 1.f : 0.f;

    long long tile = blockIdx.x;
    float gnext = 0.f;  // root position component of the NEXT tile, fetched a tile early
    if (tile < n_tiles) gnext = __ldg(gpos + min(tile * kWarp + lane, n_frames - 1) * gstride + a);
    uint32_t k = 0, it = 0;

    for (; tile < n_tiles; tile += tile_stride, ++it) {
        // row a of the "parent" of the root: the identity placed at global_pos
        float r0 = id0, r1 = id1, r2 = id2, pp = gnext;

        for (int c0 = 0; c0 < n_joints; c0 += C) {
            const int cnt = min(C, n_joints - c0);
            const uint32_t buf = k % S;
            mbar_wait(full0 + 8 * buf, (k / S) & 1);
            ++k;
            const float4 *in_row = boxes + buf * (kBoxBytes / 16) + lane * C;
            float4 q[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) q[jj] = in_row[jj ^ swz];
            {   // the loads must have LANDED before the box is released to the async proxy (see fk_kernel.cuh);
                // one component per 16-byte load is enough, the four arrive together
                uint32_t acc = 0;
#pragma unroll
                for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * buf);
            // branch-free part of the chunk: the eight normalisation scales, eight independent MUFU chains
            // (joints past the end of the skeleton are zero-filled by the TMA unit: finite, unused)
            float sc[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) sc[jj] = rot_scale(q[jj], 1e-8f);
            if (c0 == 0) {
                const long long next_tile = tile + tile_stride;
                if (next_tile < n_tiles) gnext = __ldg(gpos + min(next_tile * kWarp + lane, n_frames - 1) * gstride + a);
                if (it > 0) mbar_wait(stage_free, (it - 1) & 1);  // the previous tile has left the stage
            }

#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                if (jj < cnt) {
                    const int j = c0 + jj;
                    const uint32_t code = prog.code[j];  // constant bank, warp-uniform
                    const float4 e = tab[j];
                    if (prog_src(code) != kSrcReg) {  // parent is not the previous joint: its row is in the stage
                        const int p = static_cast<int>(prog_parent(code));
                        r0 = Rrow[9 * p], r1 = Rrow[9 * p + 1], r2 = Rrow[9 * p + 2];
                        pp = Prow[3 * p];
                    }
                    // row' = row * R(q^) = the row rotated by the conjugate of q^:
                    //   c = row x v,  row' = row + s (w c + c x v),  s = 2 / (|q| + eps)^2, q = (w, v) as loaded
                    const float w = q[jj].x, x = q[jj].y, y = q[jj].z, z = q[jj].w;
                    const float cx = r1 * z - r2 * y, cy = r2 * x - r0 * z, cz = r0 * y - r1 * x;
                    const float ex = w * cx + (cy * z - cz * y);
                    const float ey = w * cy + (cz * x - cx * z);
                    const float ez = w * cz + (cx * y - cy * x);
                    pp = r0 * e.x + r1 * e.y + r2 * e.z + pp;  // p[a] = parent row . offset + parent p[a]
                    r0 = sc[jj] * ex + r0, r1 = sc[jj] * ey + r1, r2 = sc[jj] * ez + r2;
                    float *rs = Rrow + 9 * j;
                    if (VEC == 2) {
                        // word 9j + 3a of an even-stride row: 8-byte aligned iff j + a is even (j and jj have the
                        // same parity: chunks start at multiples of 8)
                        // (PTX stores: the compiler otherwise merges the two arms back into three 32-bit stores)
                        const uint32_t ra = smem_u32(rs);
                        if (((jj & 1) ^ pa) == 0) {
                            sts64(ra, r0, r1);
                            sts32(ra + 8, r2);
                        } else {
                            sts32(ra, r0);
                            sts64(ra + 4, r1, r2);
                        }
                    } else {
                        rs[0] = r0, rs[1] = r1, rs[2] = r2;
                    }
                    Prow[3 * j] = pp;
                }
            }
        }
        fence_proxy_async_smem();  // this lane's stage writes -> visible to the async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(stage_full);
    }
}

}  // namespace pmb
