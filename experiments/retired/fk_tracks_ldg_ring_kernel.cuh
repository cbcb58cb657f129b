// Forward kinematics, track kernel (ops/skeleton.py:16-61 of the reference).
//
// Same mapping as the lane kernel -- a warp owns a tile of FR consecutive frames, lane = 3 f + a walks row a of
// frame f, G[a][:] = P[a][:] * R(q^), p[a] = P[a][:] . off + p_parent[a] -- but the tree is no longer walked in
// index order, one joint after the other.  ncu on the lane kernel (profiles/r1_fk_52_ncu_*): 8 warps per SM,
// every one of them a single dependent chain, issue slots 48 % used, 22 % of the samples on fixed-latency
// dependencies, 22 % on shared-memory / MUFU results, 16 % waiting for the next TMA box; shared memory (the
// dense output stage) rules out more warps.  So the parallelism has to come from inside the warp:
//
//   tracks   the host compiles parents[] into T steps of U independent joints (track_schedule.h, Hu's
//            highest-level-first list schedule: minimum T for U tracks).  A lane runs the U items of a step
//            interleaved -- U independent register chains -- so latencies overlap without more warps.  An item
//            whose parent was the same track's previous item keeps it in registers; every other parent row is
//            read back from the stage, which already holds every joint processed so far (same lane wrote it).
//   input    the U quaternions of a step are fetched with plain 16-byte loads, D steps ahead, into a register
//            ring (the three lanes of a frame share the address).  No shared memory for the input at all: what
//            the TMA boxes and their mbarriers used now stages output, and registers are plentiful at 6 .. 12
//            warps per SM.  The next tile's quaternions (FR x 16 J contiguous bytes) are pulled into L2 by one
//            cp.async.bulk.prefetch a tile ahead, so DRAM sees one burst per tile and the loads hit L2.
//   output   the dense image of the tile's output (FR x 36 J and FR x 12 J bytes) goes to HBM as two contiguous
//            TMA bulk stores.  The spans of a tile need not be 16-byte multiples any more (FR = 10 with an odd
//            joint count): the stage is placed at the same 16-byte phase as the global span, the aligned middle
//            goes through the TMA engine and the up-to-three head / tail words are stored by single lanes.
//            The remainder tile of a batch is just a shorter span.
//
// Algorithmic HBM traffic 64 J + 12 bytes per pose.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "fk_rows_kernel.cuh"  // rot_scale
#include "tma.cuh"
#include "track_schedule.h"

namespace pmb {

struct FkTracksGeom {
    int tab_bytes, warp_bytes, block_bytes;
};
// per block: schedule table (32 bytes per item) | per warp: R stage (+16 bytes of phase slack) | P stage (+16)
__host__ __device__ inline FkTracksGeom fk_tracks_geom(int fr, int warps, int n_joints, int n_items) {
    FkTracksGeom g;
    g.tab_bytes = (n_items * 32 + 127) & ~127;
    g.warp_bytes = ((fr * 36 * n_joints + 16 + 15) & ~15) + ((fr * 12 * n_joints + 16 + 15) & ~15);
    g.block_bytes = 128 + g.tab_bytes + warps * g.warp_bytes;
    return g;
}

// The predicates compare the table word with a per-lane threshold: 0 for the lanes that own a (frame, row), INT_MAX for
// the idle lanes of the warp (which therefore never touch the stage) -- no extra instruction to mask them.
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
// Ring refill: volatile, so that it is issued where it is written (right after the step that freed the ring entry) and not
// sunk towards its use D steps later -- the whole point of the ring is the distance.
__device__ __forceinline__ float4 track_ldg_q(const void *p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// U: tracks per lane (independent chains interleaved).  D: steps of quaternion prefetch (register ring depth).
template <int U, int D>
__global__ void __launch_bounds__(256, 1)
fk_tracks_kernel(const float4 *__restrict__ rot, const float *__restrict__ gpos, long long gstride,
                 const float *__restrict__ offsets, float *__restrict__ pos, float *__restrict__ rout,
                 long long n_frames, int n_joints, int n_steps, int fr, int l2_prefetch,
                 const __grid_constant__ TrackProgram prog) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int warps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_items = n_steps * U;
    const FkTracksGeom geo = fk_tracks_geom(fr, warps, n_joints, n_items);

    // Schedule table, 32 bytes per item:  A = offset (x, y, z) | 16 j      B = 36 j | 12 j | 36 p | 12 p
    // (byte offsets of the item's quaternion in its frame's input row and of the joint / parent rows in the two
    // stages).  B.x < 0: no-op item, nothing is stored.  B.z < 0: the parent is in the track's registers.
    // offsets[0] is ignored by the reference (the root translation is global_pos, skeleton.py:49).
    uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
        const uint32_t c = prog.code[i];
        const uint32_t j = track_joint(c), p = track_parent(c);
        uint4 A = make_uint4(0u, 0u, 0u, 0u), B = make_uint4(0x80000000u, 0u, 0x80000000u, 0u);
        if (!(c & kTrackNoop)) {
            if (j > 0) A.x = __float_as_uint(offsets[3 * j]), A.y = __float_as_uint(offsets[3 * j + 1]), A.z = __float_as_uint(offsets[3 * j + 2]);
            A.w = 16u * j;
            B.x = 36u * j, B.y = 12u * j;
            if (!(c & kTrackCarry)) B.z = 36u * p, B.w = 12u * p;
        }
        tab[2 * i] = A, tab[2 * i + 1] = B;
    }
    __syncthreads();  // the table; from here on the warps never meet again

    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    const uint32_t rst0 = smem_u32(mine);
    const uint32_t pst0 = rst0 + ((fr * 36 * n_joints + 16 + 15) & ~15);
    const uint32_t tab0 = smem_u32(tab);

    const long long n_tiles = (n_frames + fr - 1) / fr;
    const long long tile_stride = static_cast<long long>(gridDim.x) * warps;
    long long tile = static_cast<long long>(blockIdx.x) * warps + warp;
    if (tile >= n_tiles) return;
    const int rpitch = 36 * n_joints, ppitch = 12 * n_joints;  // bytes per frame row of the two stages

    // lane -> (frame, row); lanes past 3 FR shadow lane 0 and never touch the stage
    const bool active = lane < 3 * fr;
    const int f = active ? lane / 3 : 0, a = active ? lane - 3 * f : 0;
    const int thr = active ? 0 : 0x7FFFFFFF;  // predicate threshold: the idle lanes never pass
    const float id0 = a == 0 ? 1.f : 0.f, id1 = a == 1 ? 1.f : 0.f, id2 = a == 2 ? 1.f : 0.f;
    const int n_slots = ((n_steps + D - 1) / D) * D;  // steps padded to whole ring turns: the ring phase is the same for every tile

    auto qrow_of = [&](long long t) { return rot + min(t * fr + f, n_frames - 1) * n_joints; };
    auto q16_of = [&](int step, int u) { return tab[2 * (step * U + u)].w; };

    // quaternion ring: q[d][u] holds the input of the slot that is d slots ahead (mod D)
    float4 q[D][U];
    const float4 *qrow = qrow_of(tile);
    {
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int step = min(d, n_steps - 1);
                q[d][u] = track_ldg_q(reinterpret_cast<const unsigned char *>(qrow) + q16_of(step, u));
            }
    }
    float gnext = __ldg(gpos + min(tile * fr + f, n_frames - 1) * gstride + a);
    bool draining = false;  // lane 0: a bulk store of the stage may still be in flight

    for (; tile < n_tiles; tile += tile_stride) {
        const long long f0 = tile * fr;
        const int nrows = static_cast<int>(min(static_cast<long long>(fr), n_frames - f0));
        const long long next_tile = tile + tile_stride;
        const bool has_next = next_tile < n_tiles;
        const float4 *qrow_next = qrow_of(has_next ? next_tile : tile);
        // the stage sits at the 16-byte phase of the tile's global spans
        const uint32_t rphase = static_cast<uint32_t>((f0 * rpitch) & 15), pphase = static_cast<uint32_t>((f0 * ppitch) & 15);
        const uint32_t rrow = rst0 + rphase + f * rpitch + 12 * a;
        const uint32_t prow = pst0 + pphase + f * ppitch + 4 * a;

        // track 0 starts from the "parent" of the root: row a of the identity placed at global_pos
        float r0[U], r1[U], r2[U], pp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r0[u] = id0, r1[u] = id1, r2[u] = id2, pp[u] = gnext;
        if (has_next) {
            gnext = __ldg(gpos + min(next_tile * fr + f, n_frames - 1) * gstride + a);
            // the warp's next tile (l2_prefetch = 1) or the one after it (2): its quaternions into L2 as one burst,
            // well ahead of the ring
            const long long far_tile = l2_prefetch == 2 ? next_tile + tile_stride : next_tile;
            if (l2_prefetch && lane == 0 && far_tile < n_tiles - 1)
                bulk_prefetch_l2(rot + far_tile * fr * n_joints, static_cast<uint32_t>(fr * 16 * n_joints));
        }
        if (lane == 0 && draining) bulk_wait_read0();  // the previous tile has left the stage
        __syncwarp();

        for (int s0 = 0; s0 < n_slots; s0 += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int s = s0 + d;
                if (s < n_steps) {
                    uint4 A[U], B[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) A[u] = tab[2 * (s * U + u)], B[u] = tab[2 * (s * U + u) + 1];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        track_load_parent_if(B[u].z, thr, rrow + B[u].z, prow + B[u].w, r0[u], r1[u], r2[u], pp[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float4 qq = q[d][u];
                        const float sc = rot_scale(qq, 1e-8f);
                        const float w = qq.x, x = qq.y, y = qq.z, z = qq.w;
                        pp[u] = r0[u] * __uint_as_float(A[u].x) + r1[u] * __uint_as_float(A[u].y) + r2[u] * __uint_as_float(A[u].z) + pp[u];
                        const float cx_ = r1[u] * z - r2[u] * y, cy_ = r2[u] * x - r0[u] * z, cz_ = r0[u] * y - r1[u] * x;
                        const float ex = w * cx_ + (cy_ * z - cz_ * y);
                        const float ey = w * cy_ + (cz_ * x - cx_ * z);
                        const float ez = w * cz_ + (cx_ * y - cy_ * x);
                        r0[u] = sc * ex + r0[u], r1[u] = sc * ey + r1[u], r2[u] = sc * ez + r2[u];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        track_store_if(B[u].x, thr, rrow + B[u].x, prow + B[u].y, r0[u], r1[u], r2[u], pp[u]);
                }
                // refill ring entry d with the input of the slot D ahead: a later step of this tile, or -- past the
                // end -- the matching step of the warp's next tile
                {
                    const int sn = s + D;
                    const bool same = sn < n_slots;
                    const int step = min(same ? sn : sn - n_slots, n_steps - 1);
                    const unsigned char *base = reinterpret_cast<const unsigned char *>(same ? qrow : qrow_next);
#pragma unroll
                    for (int u = 0; u < U; ++u) q[d][u] = track_ldg_q(base + q16_of(step, u));
                    __syncwarp();  // keeps ptxas from sinking the loads towards their use (it schedules for an occupancy we do not have)
                }
            }
        }
        qrow = qrow_next;

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
