// Forward kinematics, track kernel (ops/skeleton.py:16-61 of the reference).
//
// Same mapping as the lane kernel -- a warp owns a tile of FR consecutive frames, lane = 3 f + a walks row a of
// frame f, G[a][:] = P[a][:] * R(q^), p[a] = P[a][:] . off + p_parent[a]; quaternions arrive as TMA boxes of
// 8 joints x FR frames -- but inside a box the tree is no longer walked in index order, one joint after the other.
// ncu on the lane kernel (profiles/r1_fk_52_ncu_*): 8 warps per SM, every one of them a single dependent chain,
// issue slots 48 % used, 22 % of the samples on fixed-latency dependencies, 22 % on shared-memory / MUFU results;
// shared memory (the dense output stage) rules out more warps.  So the parallelism has to come from inside the warp:
//
//   tracks   the host compiles parents[] into steps of U independent joints (track_schedule.h: Hu's
//            highest-level-first list schedule, window = one box of 8 joints, the part of the input that is resident).
//            A lane runs the U items of a step interleaved -- U independent register chains -- so latencies overlap
//            without more warps: 52 joints take 29 steps of two, 65 take 34.  An item whose parent was the same
//            track's previous item keeps it in registers; every other parent row is read back from the stage, which
//            already holds every joint processed so far (the same lane wrote it).
//   input    ring of NB boxes per warp, refilled by lane 0 as soon as the box's last step has read it; an item
//            reads its quaternion straight from the box (one swizzled LDS.128 per item, the three lanes of a frame
//            share it).  (Measured and retired, experiments/retired/fk_tracks_ldg_ring_kernel.cuh: a whole-skeleton
//            schedule fed by per-item 16-byte global loads through a register ring -- half of every 32-byte sector
//            wasted, the ring too shallow for the latency: 2.95 ms at 4M x 52 against 2.50 for the lane kernel.)
//   output   the dense image of the tile's output (FR x 36 J and FR x 12 J bytes) goes to HBM as two contiguous TMA
//            bulk stores.  The spans of a tile need not be 16-byte multiples (FR = 10 with an odd joint count): the
//            stage is placed at the same 16-byte phase as the global span, the aligned middle goes through the TMA
//            engine and the up-to-three head / tail words are stored by single lanes.  The remainder tile of a batch
//            is just a shorter span.
//
// Algorithmic HBM traffic 64 J + 12 bytes per pose.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "fk_rows_kernel.cuh"   // rot_scale
#include "tma.cuh"
#include "track_schedule.h"

namespace pmb {

struct FkTracksGeom {
    int tab_bytes, box_bytes, rst_bytes, warp_bytes, block_bytes;
};
// per block: schedule table (16 bytes per item) | chunk table | per warp: NB boxes | R stage (+16 bytes of phase slack) |
// P stage (+16) | NB mbarriers (64 bytes reserved) + fence words
__host__ __device__ inline FkTracksGeom fk_tracks_geom(int fr, int warps, int n_joints, int n_items, int n_boxes) {
    FkTracksGeom g;
    const int n_chunks = (n_joints + kChunk - 1) / kChunk;
    g.tab_bytes = ((n_items * 16 + 127) & ~127) + (((n_chunks + 2) * 4 + 127) & ~127);
    g.box_bytes = fr * 128;  // FR frames x 8 joints x 16 bytes
    g.rst_bytes = (fr * 36 * n_joints + 16 + 15) & ~15;
    g.warp_bytes = (n_boxes * g.box_bytes + g.rst_bytes + ((fr * 12 * n_joints + 16 + 15) & ~15) + 64 + 128 + 127) & ~127;
    g.block_bytes = 128 + g.tab_bytes + warps * g.warp_bytes;
    return g;
}

// Table word of an item: bits 0-9 joint | 10-19 parent | 30 parent kept in registers | 31 no-op.  The predicates compare
// it (shifted so that the flag is the sign bit) with a per-lane threshold: 0 for the lanes that own a (frame, row),
// INT_MAX for the idle lanes of the warp (which therefore never touch the stage) -- no extra instruction to mask them.
__device__ __forceinline__ void track_load_parent_if(uint32_t flag /* taken iff (int)flag >= thr */, int thr, uint32_t raddr, uint32_t paddr,
                                                     float &r0, float &r1, float &r2, float &pp) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %4, %7;\n"
        "@p ld.shared.f32 %0, [%5];\n"
        "@p ld.shared.f32 %1, [%5+4];\n"
        "@p ld.shared.f32 %2, [%5+8];\n"
        "@p ld.shared.f32 %3, [%6];\n"
        "}"
        : "+f"(r0), "+f"(r1), "+f"(r2), "+f"(pp)
        : "r"(flag), "r"(raddr), "r"(paddr), "r"(thr));
}
__device__ __forceinline__ void track_store_if(uint32_t flag /* stored iff (int)flag >= thr */, int thr, uint32_t raddr, uint32_t paddr, float r0,
                                               float r1, float r2, float pp) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %0, %7;\n"
        "@p st.shared.f32 [%1], %3;\n"
        "@p st.shared.f32 [%1+4], %4;\n"
        "@p st.shared.f32 [%1+8], %5;\n"
        "@p st.shared.f32 [%2], %6;\n"
        "}" ::"r"(flag), "r"(raddr), "r"(paddr), "f"(r0), "f"(r1), "f"(r2), "f"(pp), "r"(thr));
}

// U: tracks per lane (independent chains interleaved in the instruction stream).  UL: tracks as LANE GROUPS (1 or 2):
// with UL = 2 lanes 0 .. 15 walk one joint and lanes 16 .. 31 another one of the same 5 frames -- half the stage per warp,
// so twice as many warps (dependent chains in flight) for the same shared memory at the same instructions per
// frame and joint.  The schedule has U * UL tracks; track (t, u) is column t * U + u of a step.  NB: TMA boxes in flight.
template <int U, int NB, int UL = 1>
__global__ void __launch_bounds__(512, 1)
fk_tracks_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
                 const float *__restrict__ offsets, float *__restrict__ pos, float *__restrict__ rout,
                 long long n_frames, int n_joints, int n_steps, int fr, const float *__restrict__ rot_l2_prefetch,
                 const __grid_constant__ TrackProgram prog) {
    static_assert(UL == 1 || UL == 2, "one or two lane groups");
    static_assert(NB >= 2 && NB <= 8, "ring depth");
    constexpr int C = kChunk;
    extern __shared__ __align__(128) unsigned char smem_trk[];
    unsigned char *smem_raw = smem_trk + ((128u - (smem_u32(smem_trk) & 127u)) & 127u);
    const int warps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_items = n_steps * U * UL;
    const int n_chunks = (n_joints + C - 1) / C;
    const FkTracksGeom geo = fk_tracks_geom(fr, warps, n_joints, n_items, NB);
    const int BOX = geo.box_bytes;

    // Schedule table, 16 bytes per item: offset (x, y, z) | word (above).  One shared-memory wavefront per item; the
    // byte offsets of the joint / parent rows in the two stages and of the quaternion in its box row are multiplies of
    // the word's fields (issue slots are there, shared-memory bandwidth is what the walk is short of).
    // offsets[0] is ignored by the reference (the root translation is global_pos, skeleton.py:49).
    uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
    int *cfirst = reinterpret_cast<int *>(smem_raw + ((n_items * 16 + 127) & ~127));
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
        const uint32_t c = prog.code[i];
        const uint32_t j = track_joint(c), p = track_parent(c);
        uint4 e = make_uint4(0u, 0u, 0u, 0xC0000000u);
        if (!(c & kTrackNoop)) {
            if (j > 0) e.x = __float_as_uint(offsets[3 * j]), e.y = __float_as_uint(offsets[3 * j + 1]), e.z = __float_as_uint(offsets[3 * j + 2]);
            e.w = j | (p << 10) | ((c & kTrackCarry) ? 0x40000000u : 0u);
        }
        tab[i] = e;
    }
    for (int i = threadIdx.x; i <= n_chunks; i += blockDim.x) cfirst[i] = prog.chunk_first[i];

    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    const uint32_t box0 = smem_u32(mine);
    const uint32_t rst0 = box0 + NB * BOX;
    const uint32_t pst0 = rst0 + geo.rst_bytes;
    const uint32_t bar0 = pst0 + ((fr * 12 * n_joints + 16 + 15) & ~15);  // NB mbarriers (16-byte aligned)
    const uint32_t fence_word = bar0 + 64 + 4 * lane;
    const uint32_t tab0 = smem_u32(tab);
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) mbar_init(bar0 + 8 * b, 1);
        fence_barrier_init();
    }
    __syncthreads();  // the tables; from here on the warps never meet again

    const long long n_tiles = (n_frames + fr - 1) / fr;
    const long long tile_stride = static_cast<long long>(gridDim.x) * warps;
    long long tile = static_cast<long long>(blockIdx.x) * warps + warp;
    if (tile >= n_tiles) return;
    const int rpitch = 36 * n_joints, ppitch = 12 * n_joints;  // bytes per frame row of the two stages

    // lane -> (lane group, frame, row); lanes past 3 FR of their group shadow its first lane and never touch the stage
    const int grp = UL == 2 ? lane >> 4 : 0, gl = UL == 2 ? (lane & 15) : lane;
    const bool active = gl < 3 * fr;
    const int f = active ? gl / 3 : 0, a = active ? gl - 3 * f : 0;
    const int thr = active ? 0 : 0x7FFFFFFF;  // predicate threshold: the idle lanes never pass
    const float id0 = a == 0 ? 1.f : 0.f, id1 = a == 1 ? 1.f : 0.f, id2 = a == 2 ? 1.f : 0.f;

    // TMA producer (lane 0): the warp's boxes in processing order, across its tiles
    long long la_tile = tile;
    int la_c0 = 0;
    auto issue_next = [&](int buf) {
        if (la_tile < n_tiles) {
            mbar_arrive_expect_tx(bar0 + 8 * buf, BOX);
            tma_load_2d(box0 + buf * BOX, &tm_rot, 4 * la_c0, static_cast<int>(la_tile * fr), bar0 + 8 * buf);
            la_c0 += C;
            if (la_c0 >= n_joints) la_c0 = 0, la_tile += tile_stride;
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) issue_next(b);
    }

    float gnext = __ldg(gpos + min(tile * fr + f, n_frames - 1) * gstride + a);
    uint32_t k = 0;
    bool draining = false;  // lane 0: a bulk store of the stage may still be in flight

    for (; tile < n_tiles; tile += tile_stride) {
        const long long f0 = tile * fr;
        const int nrows = static_cast<int>(min(static_cast<long long>(fr), n_frames - f0));
        // the stage sits at the 16-byte phase of the tile's global spans
        const uint32_t rphase = static_cast<uint32_t>((f0 * rpitch) & 15), pphase = static_cast<uint32_t>((f0 * ppitch) & 15);
        const uint32_t rrow = rst0 + rphase + f * rpitch + 12 * a;
        const uint32_t prow = pst0 + pphase + f * ppitch + 4 * a;

        // track 0 starts from the "parent" of the root: row a of the identity placed at global_pos
        float r0[U], r1[U], r2[U], pp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r0[u] = id0, r1[u] = id1, r2[u] = id2, pp[u] = gnext;
        {
            const long long next_tile = tile + tile_stride;
            if (next_tile < n_tiles) gnext = __ldg(gpos + min(next_tile * fr + f, n_frames - 1) * gstride + a);
        }
        if (lane == 0) {
            // the warp's next tile: its quaternions (FR x 16 J contiguous bytes) into L2 as one burst, a tile ahead of the boxes
            const long long next_tile = tile + tile_stride;
            if (rot_l2_prefetch && next_tile < n_tiles - 1)
                bulk_prefetch_l2(rot_l2_prefetch + next_tile * fr * 4 * n_joints, static_cast<uint32_t>(fr * 16 * n_joints));
            if (draining) bulk_wait_read0();  // the previous tile has left the stage
        }
        __syncwarp();

        int step = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const uint32_t buf = k % NB;
            mbar_wait(bar0 + 8 * buf, (k / NB) & 1);
            ++k;
            // the box row of this lane's frame is 128-byte aligned, so 16-byte chunk (jj ^ swz) of it is at
            // (row | swz << 4) ^ (jj << 4): one XOR per read
            const uint32_t row = box0 + buf * BOX + f * 128;
            const uint32_t row_swz = row | (((row >> 7) & 7u) << 4);
            const int step_end = cfirst[c + 1];
            uint32_t acc = 0;
            for (; step < step_end; ++step) {
                const uint32_t t0 = tab0 + (step * UL + grp) * (U * 16);
                float4 e[U], qq[U];
                uint32_t wj[U], wp[U];
#pragma unroll
                for (int u = 0; u < U; ++u) e[u] = lds128_ro(t0 + u * 16);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t w = __float_as_uint(e[u].w);
                    qq[u] = lds128(row_swz ^ ((w & 7u) << 4));
                    wj[u] = w & 0x3FFu, wp[u] = (w >> 10) & 0x3FFu;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    track_load_parent_if(__float_as_uint(e[u].w) << 1, thr, rrow + 36 * wp[u], prow + 12 * wp[u], r0[u], r1[u], r2[u], pp[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc |= __float_as_uint(qq[u].x);
                    const float sc = rot_scale(qq[u], 1e-8f);
                    const float w = qq[u].x, x = qq[u].y, y = qq[u].z, z = qq[u].w;
                    pp[u] = r0[u] * e[u].x + r1[u] * e[u].y + r2[u] * e[u].z + pp[u];
                    const float cx_ = r1[u] * z - r2[u] * y, cy_ = r2[u] * x - r0[u] * z, cz_ = r0[u] * y - r1[u] * x;
                    const float ex = w * cx_ + (cy_ * z - cz_ * y);
                    const float ey = w * cy_ + (cz_ * x - cx_ * z);
                    const float ez = w * cz_ + (cx_ * y - cy_ * x);
                    r0[u] = sc * ex + r0[u], r1[u] = sc * ey + r1[u], r2[u] = sc * ez + r2[u];
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    track_store_if(__float_as_uint(e[u].w), thr, rrow + 36 * wj[u], prow + 12 * wj[u], r0[u], r1[u], r2[u], pp[u]);
                if (UL > 1) __syncwarp();  // a parent row may have been stored by the other lane group
            }
            // the box's loads must have LANDED before it is refilled through the async proxy (see fk_kernel.cuh): a
            // store that depends on every loaded quaternion precedes the refill
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
            __syncwarp();
            if (lane == 0) issue_next(buf);  // refill with the box NB ahead (this tile's or the next tile's)
        }

        // ---- the tile's output: two contiguous spans, aligned middle through the TMA engine --------------------
        fence_proxy_async_smem();  // this lane's stage writes -> visible to the async proxy
        __syncwarp();
        {
            const long long ra = f0 * rpitch, rb = ra + static_cast<long long>(nrows) * rpitch;  // byte span of rotmats
            const long long pa = f0 * ppitch, pb = pa + static_cast<long long>(nrows) * ppitch;  // byte span of positions
            const long long ra16 = (ra + 15) & ~15LL, rb16 = rb & ~15LL, pa16 = (pa + 15) & ~15LL, pb16 = pb & ~15LL;
            unsigned char *rg = reinterpret_cast<unsigned char *>(rout), *pg = reinterpret_cast<unsigned char *>(pos);
            if (lane == 0) {
                if (rb16 > ra16) bulk_store(rg + ra16, rst0 + rphase + static_cast<uint32_t>(ra16 - ra), static_cast<uint32_t>(rb16 - ra16));
                if (pb16 > pa16) bulk_store(pg + pa16, pst0 + pphase + static_cast<uint32_t>(pa16 - pa), static_cast<uint32_t>(pb16 - pa16));
                bulk_commit();
                draining = true;
            }
            // head / tail words (at most 3 + 3 per span when the middle exists; a span shorter than its alignment
            // gap is all "head").  Lanes 1 .. 31 copy them straight from the stage.
            if (((ra | rb | pa | pb) & 15) != 0) {
                const long long rh_end = rb16 > ra16 ? ra16 : rb, ph_end = pb16 > pa16 ? pa16 : pb;
                const int rh = static_cast<int>(rh_end - ra) >> 2, rt = rb16 > ra16 ? static_cast<int>(rb - rb16) >> 2 : 0;
                const int ph = static_cast<int>(ph_end - pa) >> 2, pt = pb16 > pa16 ? static_cast<int>(pb - pb16) >> 2 : 0;
                const int i = lane - 1;
                if (i >= 0) {
                    float v;
                    if (i < rh) {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(rst0 + rphase + 4 * i));
                        reinterpret_cast<float *>(rg + ra)[i] = v;
                    } else if (i - 8 >= 0 && i - 8 < rt) {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(rst0 + rphase + static_cast<uint32_t>(rb16 - ra) + 4 * (i - 8)));
                        reinterpret_cast<float *>(rg + rb16)[i - 8] = v;
                    } else if (i - 16 >= 0 && i - 16 < ph) {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pst0 + pphase + 4 * (i - 16)));
                        reinterpret_cast<float *>(pg + pa)[i - 16] = v;
                    } else if (i - 24 >= 0 && i - 24 < pt) {
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pst0 + pphase + static_cast<uint32_t>(pb16 - pa) + 4 * (i - 24)));
                        reinterpret_cast<float *>(pg + pb16)[i - 24] = v;
                    }
                }
                __syncwarp();  // the stage is rewritten only after these reads
            }
        }
    }
    if (lane == 0 && draining) bulk_wait0();  // global writes of the last tile are complete at exit
}

}  // namespace pmb
