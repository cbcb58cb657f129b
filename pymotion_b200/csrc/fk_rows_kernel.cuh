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
// collapses into one scale 2 / (|q| + eps)^2 per joint.  The walk is branch free (predicated PTX for the parent
// fetch and the tail-chunk guard).
//
//   in    a loader thread streams the tile's quaternions as TMA boxes of 8 joints x 32 frames (128-byte
//         swizzle) through an S-deep ring; full / empty mbarriers, the three row warps release a box as soon
//         as its quaternions are in registers;
//   walk  parents[] compiled on the host (joint_program.h); the parent row is either still in registers
//         (parent == previous joint) or read back from the stage -- the stage is the dense image of the
//         tile's output, so it already holds every ancestor: no slots;
//   out   the stage (32 x 36J and 32 x 12J bytes, both multiples of 128) is handed to the TMA engine by a
//         drainer thread as two contiguous line-aligned bulk stores while the row warps wait for
//         `stage_free`; the other teams on the SM cover the drain.  The remainder tile of a batch (fewer than 32
//         frames: its span need not be a multiple of 16 bytes) is copied out by the lanes that wrote it.
//
// Block = 1 team = 5 warps: rows 0..2, loader, drainer.  Algorithmic HBM traffic 64*J + 12 bytes per pose.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"

// Build-time variants of the walk (measured on B200): chunk-wide hoisting of the normalisation
// scales, one-joint-early fetch of parent rows, and the algebraic form of the row rotation.
#ifndef PMB_ROWS_HOIST
#define PMB_ROWS_HOIST 0
#endif
#ifndef PMB_ROWS_PREFETCH
#define PMB_ROWS_PREFETCH 0
#endif
#ifndef PMB_ROWS_FORM
#define PMB_ROWS_FORM 0
#endif

namespace pmb {

constexpr int kRowThreads = 160;

struct FkRowsGeom {
    int tab_bytes, stage_bytes, block_bytes;
};
__host__ __device__ inline FkRowsGeom fk_rows_geom(int stages, int n_joints, bool quat_out = false) {
    FkRowsGeom g;
    g.tab_bytes = ((n_joints + kChunk) * 16 + 127) & ~127;  // padded: the tail chunk and the one-ahead prefetch read past J
    g.stage_bytes = kWarp * (quat_out ? 7 : 12) * n_joints * 4;
    // 1 KB slack to align the boxes | boxes | joint table | stage | barriers (full[S], empty[S], stage_full, stage_free) | fence words
    g.block_bytes = 1024 + stages * kBoxBytes + g.tab_bytes + g.stage_bytes + (2 * stages + 2) * 8 + 96 * 4;
    g.block_bytes = (g.block_bytes + 15) & ~15;
    return g;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// The tree walk below has NO branches: the parent fetch and the tail-chunk guard are predicated PTX, so
// the compiler is free to overlap the independent parts of consecutive joints (measured: with a uniform
// branch per joint every joint paid the constant-load -> compare -> branch latency, ~280 cycles per joint).
// Deliberately no "memory" clobber: the stage is touched only through these two helpers (volatile asm keeps
// their mutual order) and the joint table is read-only, so its loads may float.
__device__ __forceinline__ void load_parent_row_if(int parent /* taken iff >= 0 */, uint32_t raddr, uint32_t paddr, float &r0, float &r1,
                                                   float &r2, float &pp) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %4, 0;\n"
        "@p ld.shared.f32 %0, [%5];\n"
        "@p ld.shared.f32 %1, [%5+4];\n"
        "@p ld.shared.f32 %2, [%5+8];\n"
        "@p ld.shared.f32 %3, [%6];\n"
        "}"
        : "+f"(r0), "+f"(r1), "+f"(r2), "+f"(pp)
        : "r"(parent), "r"(raddr), "r"(paddr));
}
// ALIGNED: raddr is 8-byte aligned -> (r0, r1) as one 64-bit store; otherwise (r1, r2) is the aligned pair.
template <int VEC, bool ALIGNED>
__device__ __forceinline__ void store_row_if(int keep /* stored iff > 0 */, uint32_t raddr, uint32_t paddr, float r0, float r1, float r2,
                                             float pp) {
    if (VEC == 2 && ALIGNED) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.gt.s32 p, %0, 0;\n"
            "@p st.shared.v2.f32 [%1], {%3, %4};\n"
            "@p st.shared.f32 [%1+8], %5;\n"
            "@p st.shared.f32 [%2], %6;\n"
            "}" ::"r"(keep), "r"(raddr), "r"(paddr), "f"(r0), "f"(r1), "f"(r2), "f"(pp));
    } else if (VEC == 2) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.gt.s32 p, %0, 0;\n"
            "@p st.shared.f32 [%1], %3;\n"
            "@p st.shared.v2.f32 [%1+4], {%4, %5};\n"
            "@p st.shared.f32 [%2], %6;\n"
            "}" ::"r"(keep), "r"(raddr), "r"(paddr), "f"(r0), "f"(r1), "f"(r2), "f"(pp));
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.gt.s32 p, %0, 0;\n"
            "@p st.shared.f32 [%1], %3;\n"
            "@p st.shared.f32 [%1+4], %4;\n"
            "@p st.shared.f32 [%1+8], %5;\n"
            "@p st.shared.f32 [%2], %6;\n"
            "}" ::"r"(keep), "r"(raddr), "r"(paddr), "f"(r0), "f"(r1), "f"(r2), "f"(pp));
    }
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

// Everything a row warp needs from the kernel prologue.
struct RowCtx {
    const float4 *boxes, *tab;
    uint32_t rst, pst;            // shared addresses of the stage (rotation rows, positions)
    uint32_t full0, empty0, stage_full, stage_free, fence_word;
    const float *gpos;
    float *pos, *rout;
    long long gstride, n_frames, n_tiles, tile_stride, first_tile;
    int n_joints, lane;
};

// One row warp: row A of every joint of the tile, lane = frame.
// VEC = 2 (even joint count: every stage row is 8-byte aligned): a row's three numbers go out as one 64-bit
// and one 32-bit store -- the minimum number of shared-memory wavefronts for this layout; VEC = 1: 32-bit
// stores (odd row stride, conflict free as they are).
template <int S, int VEC, int A>
__device__ __forceinline__ void fk_row_walk(const RowCtx &cx) {
    constexpr int C = kChunk;
    const int lane = cx.lane, n_joints = cx.n_joints;
    const int swz = lane & 7;
    const uint32_t rrow = cx.rst + (lane * 9 * n_joints + 3 * A) * 4;  // this lane's row A of joint 0
    const uint32_t prow = cx.pst + (lane * 3 * n_joints + A) * 4;

    long long tile = cx.first_tile;
    float gnext = 0.f;  // root position component of the NEXT tile, fetched a tile early
    if (tile < cx.n_tiles) gnext = __ldg(cx.gpos + min(tile * kWarp + lane, cx.n_frames - 1) * cx.gstride + A);
    uint32_t k = 0, it = 0;

    for (; tile < cx.n_tiles; tile += cx.tile_stride, ++it) {
        // row A of the "parent" of the root: the identity placed at global_pos
        float r0 = A == 0 ? 1.f : 0.f, r1 = A == 1 ? 1.f : 0.f, r2 = A == 2 ? 1.f : 0.f, pp = gnext;

        for (int c0 = 0; c0 < n_joints; c0 += C) {
            const int cnt = n_joints - c0;  // >= 8 for every chunk but a partial last one
            const uint32_t buf = k % S;
            mbar_wait(cx.full0 + 8 * buf, (k / S) & 1);
            ++k;
            const float4 *in_row = cx.boxes + buf * (kBoxBytes / 16) + lane * C;
            float4 q[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) q[jj] = in_row[jj ^ swz];
            {   // the loads must have LANDED before the box is released to the async proxy (see fk_kernel.cuh);
                // one component per 16-byte load is enough, the four arrive together
                uint32_t acc = 0;
#pragma unroll
                for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(cx.fence_word), "r"(acc) : "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.empty0 + 8 * buf);
#if PMB_ROWS_HOIST
            // The eight normalisation scales of the chunk ahead of the walk (and of the block boundary below,
            // which keeps ptxas from sinking them back into it).
            float sc[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) sc[jj] = rot_scale(q[jj], 1e-8f);
#endif
            if (c0 == 0) {
                const long long next_tile = tile + cx.tile_stride;
                if (next_tile < cx.n_tiles)
                    gnext = __ldg(cx.gpos + min(next_tile * kWarp + lane, cx.n_frames - 1) * cx.gstride + A);
                if (it > 0) mbar_wait(cx.stage_free, (it - 1) & 1);  // the previous tile has left the stage
            }

            // Branch-free walk over the chunk.  Joints past the end of the skeleton (partial last chunk) are
            // zero quaternions (TMA fill) with padded table entries: identity steps whose stores are predicated off.
            float4 e = cx.tab[c0];
#if PMB_ROWS_PREFETCH
            // A parent row that has to come from the stage is fetched ONE JOINT EARLY into (t0, t1, t2, tp): such a
            // parent is at most joint j - 2, so its row is already in the stage when joint j - 1 starts.
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, tp = 0.f;
            {
                const int p = __float_as_int(e.w);
                load_parent_row_if(p, rrow + 36 * p, prow + 12 * p, t0, t1, t2, tp);
            }
#endif
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                const int j = c0 + jj;
                const float4 e_next = cx.tab[j + 1];  // one joint ahead (the table is padded by a chunk)
#if PMB_ROWS_PREFETCH
                if (__float_as_int(e.w) >= 0) r0 = t0, r1 = t1, r2 = t2, pp = tp;  // selects, not a branch
                if (jj + 1 < C) {
                    const int pn = __float_as_int(e_next.w);
                    load_parent_row_if(pn, rrow + 36 * pn, prow + 12 * pn, t0, t1, t2, tp);
                }
#else
                const int p = __float_as_int(e.w);    // parent whose row must come from the stage, or -1
                load_parent_row_if(p, rrow + 36 * p, prow + 12 * p, r0, r1, r2, pp);
#endif
#if PMB_ROWS_HOIST
                const float s = sc[jj];
#else
                const float s = rot_scale(q[jj], 1e-8f);
#endif
                const float w = q[jj].x, x = q[jj].y, y = q[jj].z, z = q[jj].w;
                pp = r0 * e.x + r1 * e.y + r2 * e.z + pp;  // p[A] = parent row . offset + parent p[A]
#if PMB_ROWS_FORM == 2
                // experiment: no arithmetic at all (data-movement ceiling of the kernel structure)
                r0 = x + s, r1 = y, r2 = z + w;
#elif PMB_ROWS_FORM == 0
                // row' = row * R(q^) = the row rotated by the conjugate of q^:
                //   c = row x v,  row' = row + s (w c + c x v),  s = 2 / (|q| + eps)^2, q = (w, v) as loaded
                const float cx_ = r1 * z - r2 * y, cy_ = r2 * x - r0 * z, cz_ = r0 * y - r1 * x;
                const float ex = w * cx_ + (cy_ * z - cz_ * y);
                const float ey = w * cy_ + (cz_ * x - cx_ * z);
                const float ez = w * cz_ + (cx_ * y - cy_ * x);
                r0 = s * ex + r0, r1 = s * ey + r1, r2 = s * ez + r2;
#else
                // the same rotation with (row x v) x v = v (row . v) - row |v|^2 expanded: the part that depends on
                // the row is 4 operations deep instead of 7,
                //   row' = a row + row x b + (row . v) d,   a = 1 - s |v|^2,  b = s w v,  d = s v
                const float a_ = 1.f - s * (x * x + y * y + z * z), sw = s * w;
                const float bx = sw * x, by = sw * y, bz = sw * z, dx = s * x, dy = s * y, dz = s * z;
                const float kd = r0 * x + r1 * y + r2 * z;
                const float n0 = a_ * r0 + (r1 * bz - r2 * by), n1 = a_ * r1 + (r2 * bx - r0 * bz), n2 = a_ * r2 + (r0 * by - r1 * bx);
                r0 = kd * dx + n0, r1 = kd * dy + n1, r2 = kd * dz + n2;
#endif
                // word 9j + 3A of an even-stride row is 8-byte aligned iff j + A is even (j and jj have the same
                // parity: chunks start at multiples of 8); constant after unrolling
                if (((jj + A) & 1) == 0) store_row_if<VEC, true>(cnt - jj, rrow + 36 * j, prow + 12 * j, r0, r1, r2, pp);
                else store_row_if<VEC, false>(cnt - jj, rrow + 36 * j, prow + 12 * j, r0, r1, r2, pp);
                e = e_next;
            }
        }
        const long long f0 = tile * kWarp;
        if (cx.n_frames - f0 >= kWarp) {
            fence_proxy_async_smem();  // this lane's stage writes -> visible to the async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(cx.stage_full);
        } else if (f0 + lane < cx.n_frames) {
            // Remainder tile (the last tile of the batch, fewer than 32 frames): its span need not be a multiple of
            // 16 bytes, so it does not go through the TMA engine.  Every lane copies out the rows IT wrote: no other
            // thread touches them, nothing to synchronise.
            float *rg = cx.rout + (f0 + lane) * 9 * n_joints + 3 * A;
            float *pg = cx.pos + (f0 + lane) * 3 * n_joints + A;
            for (int j = 0; j < n_joints; ++j) {
                float v0, v1, v2, vp;
                asm volatile("ld.shared.f32 %0, [%4];\n ld.shared.f32 %1, [%4+4];\n ld.shared.f32 %2, [%4+8];\n ld.shared.f32 %3, [%5];"
                             : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(vp)
                             : "r"(rrow + 36 * j), "r"(prow + 12 * j));
                rg[9 * j] = v0, rg[9 * j + 1] = v1, rg[9 * j + 2] = v2;
                pg[3 * j] = vp;
            }
        }
    }
}

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

    // tiles go round robin over the teams: a window of consecutive tiles slides through the arrays (measured: one
    // contiguous range of tiles per team is 4 % slower)
    const long long n_tiles = (n_frames + kWarp - 1) / kWarp;
    const long long tile_stride = gridDim.x;
    const long long tile_first = blockIdx.x;
    const int rpitch = 9 * n_joints, ppitch = 3 * n_joints;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < S; ++b) mbar_init(full0 + 8 * b, 1), mbar_init(empty0 + 8 * b, 3);
        mbar_init(stage_full, 3);
        mbar_init(stage_free, 1);
        fence_barrier_init();
    }
    // Joint table: offset (x, y, z) | parent index if the parent's row has to be fetched from the stage, -1 if it
    // is the previous joint (still in registers).  offsets[0] is ignored by the reference (the root translation
    // is global_pos, skeleton.py:49): with a zero entry the root is an ordinary joint whose parent is the
    // identity placed at global_pos.  Padding entries are identity steps.
    for (int j = threadIdx.x; j < n_joints + kChunk; j += kRowThreads) {
        float4 e = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (j > 0 && j < n_joints) {
            e.x = offsets[3 * j], e.y = offsets[3 * j + 1], e.z = offsets[3 * j + 2];
            const uint32_t code = prog.code[j];
            if (prog_src(code) != kSrcReg) e.w = __int_as_float(static_cast<int>(prog_parent(code)));
        }
        tab[j] = e;
    }
    __syncthreads();  // the only block-wide barrier: the roles below never meet again

    if (warp == 3) {
        // ---- loader: the team's chunks in processing order, across its tiles -----------------------
        if (lane == 0) {
            uint32_t k = 0;
            for (long long t = tile_first; t < n_tiles; t += tile_stride) {
                for (int c0 = 0; c0 < n_joints; c0 += C, ++k) {
                    const uint32_t buf = k % S;
                    if (k >= S) mbar_wait_long(empty0 + 8 * buf, ((k / S) - 1) & 1);  // all three row warps have read it
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
            for (long long t = tile_first; t < n_tiles; t += tile_stride, ++it) {
                const long long f0 = t * kWarp;
                if (n_frames - f0 < kWarp) break;  // the remainder tile is written by the row warps themselves
                mbar_wait_long(stage_full, it & 1);
                float *rg = rout + f0 * rpitch, *pg = pos + f0 * ppitch;
                const uint32_t rbytes = static_cast<uint32_t>(kWarp * rpitch * 4), pbytes = static_cast<uint32_t>(kWarp * ppitch * 4);
                // two contiguous, 128-byte aligned spans
                bulk_store(rg, smem_u32(Rst), rbytes);
                bulk_store(pg, smem_u32(Pst), pbytes);
                bulk_commit();
                bulk_wait_read0();  // the engine has read the stage: the row warps may overwrite it
                mbar_arrive(stage_free);
            }
            bulk_wait0();  // global writes of the last tile are complete at exit
        }
        return;
    }

    // ---- row warps: warp a computes row a of every transform ------------------------------------------
    RowCtx cx;
    cx.boxes = boxes, cx.tab = tab;
    cx.rst = smem_u32(Rst), cx.pst = smem_u32(Pst);
    cx.full0 = full0, cx.empty0 = empty0, cx.stage_full = stage_full, cx.stage_free = stage_free;
    cx.fence_word = stage_free + 8 + 4 * threadIdx.x;
    cx.pos = pos, cx.rout = rout;
    cx.gpos = gpos, cx.gstride = gstride, cx.n_frames = n_frames, cx.n_tiles = n_tiles, cx.tile_stride = tile_stride;
    cx.first_tile = tile_first, cx.n_joints = n_joints, cx.lane = lane;
    if (warp == 0) fk_row_walk<S, VEC, 0>(cx);
    else if (warp == 1) fk_row_walk<S, VEC, 1>(cx);
    else fk_row_walk<S, VEC, 2>(cx);
}

}  // namespace pmb
