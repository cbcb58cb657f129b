// RETIRED (round 2): fk lane = (frame, row) kernel of round 1, superseded by fk_tracks_kernel.cuh (31 .. 59 joints) and
// fk_mtracks_kernel.cuh (60 joints and more, small skeletons the row teams do not take): 2M x 24 0.599 ms against 0.550,
// 2M x 40 0.907 against 0.877, 4M x 52 2.504 against 2.30, 4M x 65 3.442 against 2.69.  Not compiled into the library.
// Forward kinematics, lane = (frame, row) kernel (ops/skeleton.py:16-61 of the reference).
//
// What bounds fk on a B200 is how many frames an SM keeps in flight, and that is set by shared memory: the
// output of a frame (48 J bytes) has to be staged so that HBM sees long contiguous writes (fk_kernel.cuh).
// The row-team kernel (fk_rows_kernel.cuh) spends three warps on a tile of 32 frames, so a 65-joint skeleton
// gets two tiles = six walking warps per SM and each of them is latency-bound (measured: removing ALL the
// arithmetic from the walk changed the time by 12 %).  This kernel keeps the same staged bytes per frame but
// cuts the tile to FR = 8 (or 10) frames and puts the three rows of a frame on three LANES of one warp:
//
//     lane = 3 f + a      f = frame of the tile (0 .. FR-1),  a = row of the 3x4 transform (0 .. 2)
//
// A warp owns a whole tile (24 or 30 active lanes) and is autonomous -- own TMA boxes, own mbarriers, own
// dense stage, own bulk stores -- so the same shared memory now carries 8 .. 20 walking warps per SM instead
// of 6 .. 12, every one of them with the short 4-register row chain of the row-team kernel:
//     G[a][:] = P[a][:] * R(q^)   (the row rotated by the conjugate quaternion),   p[a] = P[a][:] . off + p_parent[a]
// The three lanes of a frame read the same quaternion (one shared-memory broadcast), so the box traffic is a
// third of the row-team kernel's.  A tile's output is still one contiguous span (FR * 36 J and FR * 12 J bytes,
// 18 .. 25 KB for the large skeletons), handed to the TMA engine by lane 0; the wait for the engine to have
// read the stage is deferred to just before the next tile's first store.
//
// FR: the spans must be multiples of 16 bytes for cp.async.bulk: FR = 8 always works (288 J, 96 J), FR = 10
// (360 J, 120 J; 30 of 32 lanes busy) needs an even joint count.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "fk_rows_kernel.cuh"  // rot_scale, load_parent_row_if, store_row_if, mbar_arrive
#include "tma.cuh"

#ifndef PMB_LANES_HOIST
#define PMB_LANES_HOIST 0
#endif
#ifndef PMB_LANES_XOR
#define PMB_LANES_XOR 1  // box / table reads through 32-bit shared addresses (one XOR per swizzled read)
#endif

namespace pmb {

struct FkLanesGeom {
    int box_bytes, tab_bytes, warp_bytes, block_bytes;
};
__host__ __device__ inline FkLanesGeom fk_lanes_geom(int fr, int warps, int n_joints, int n_boxes = 2) {
    FkLanesGeom g;
    g.box_bytes = fr * 128;  // FR frames x 8 joints x 16 bytes
    g.tab_bytes = ((n_joints + kChunk) * 16 + 127) & ~127;     // padded: the tail chunk and the one-ahead prefetch read past J
    // per warp: NB boxes | rotation stage | position stage | NB mbarriers (32 bytes reserved) | 32 fence words
    g.warp_bytes = ((n_boxes * g.box_bytes + fr * 48 * n_joints + 32 + 128) + 127) & ~127;
    g.block_bytes = 128 + g.tab_bytes + warps * g.warp_bytes;
    return g;
}

// NB: TMA boxes in flight per warp (ring depth, 2 .. 4).
// (Measured and retired, see experiments/retired/README.md: whole-tile contiguous input, L2 prefetch of the next
// tile, evict-first stores.)
template <int FR, int WARPS, int NB>
__global__ void __launch_bounds__(WARPS *kWarp)
fk_lanes_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
                const float *__restrict__ offsets, float *__restrict__ pos, float *__restrict__ rout,
                long long n_frames, int n_joints, const __grid_constant__ JointProgram prog) {
    constexpr int C = kChunk;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const FkLanesGeom geo = fk_lanes_geom(FR, WARPS, n_joints, NB);
    const int BOX = geo.box_bytes;  // bytes of one input buffer (a box, or a whole tile)

    float4 *tab = reinterpret_cast<float4 *>(smem_raw);
    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    const float4 *boxes = reinterpret_cast<const float4 *>(mine);
    float *Rst = reinterpret_cast<float *>(mine + NB * BOX);
    float *Pst = Rst + FR * 9 * n_joints;
    const uint32_t box0 = smem_u32(mine);
    const uint32_t tab0 = smem_u32(tab);
    const uint32_t bar0 = smem_u32(Pst + FR * 3 * n_joints);   // NB mbarriers (8-byte aligned: all sizes above are multiples of 16)
    const uint32_t fence_word = bar0 + 32 + 4 * lane;

    // Joint table: offset (x, y, z) | parent index if the parent's row has to be fetched from the stage, -1 if it
    // is the previous joint (still in registers).  offsets[0] is ignored by the reference (the root translation
    // is global_pos, skeleton.py:49): with a zero entry the root is an ordinary joint whose parent is the
    // identity placed at global_pos.  Padding entries are identity steps.
    for (int j = threadIdx.x; j < n_joints + kChunk; j += WARPS * kWarp) {
        float4 e = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (j > 0 && j < n_joints) {
            e.x = offsets[3 * j], e.y = offsets[3 * j + 1], e.z = offsets[3 * j + 2];
            const uint32_t code = prog.code[j];
            if (prog_src(code) != kSrcReg) e.w = __int_as_float(static_cast<int>(prog_parent(code)));
        }
        tab[j] = e;
    }
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) mbar_init(bar0 + 8 * b, 1);
        fence_barrier_init();
    }
    __syncthreads();  // the table; from here on the warps never meet again

    const long long n_tiles = (n_frames + FR - 1) / FR;
    const long long tile_stride = static_cast<long long>(gridDim.x) * WARPS;
    long long tile = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const int rpitch = 9 * n_joints, ppitch = 3 * n_joints;

    // lane -> (frame, row); lanes past 3 FR walk a copy of lane 0's chain and never store
    const bool active = lane < 3 * FR;
    const int f = active ? lane / 3 : 0, a = active ? lane - 3 * f : 0;
    const uint32_t rrow = smem_u32(Rst) + (f * rpitch + 3 * a) * 4;
    const uint32_t prow = smem_u32(Pst) + (f * ppitch + a) * 4;
    // 128-byte hardware swizzle: 16-byte chunk c of box row r sits at chunk c ^ ((address >> 7) & 7); boxes are
    // only 128-byte aligned here, so the row's phase includes the box base
    const float4 *in_row0 = boxes + f * C;
    const int swz_base = (box0 >> 7) + f;  // + box index * (BOX / 128)
    const float id0 = a == 0 ? 1.f : 0.f, id1 = a == 1 ? 1.f : 0.f, id2 = a == 2 ? 1.f : 0.f;

    // TMA producer (lane 0): the warp's chunks in processing order, across its tiles
    long long la_tile = tile;
    int la_c0 = 0;
    auto issue_next = [&](int buf) {
        if (la_tile < n_tiles) {
            mbar_arrive_expect_tx(bar0 + 8 * buf, BOX);
            tma_load_2d(box0 + buf * BOX, &tm_rot, 4 * la_c0, static_cast<int>(la_tile * FR), bar0 + 8 * buf);
            la_c0 += C;
            if (la_c0 >= n_joints) la_c0 = 0, la_tile += tile_stride;
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) issue_next(b);
    }

    float gnext = 0.f;  // root position component of the NEXT tile, fetched a tile early
    if (tile < n_tiles) gnext = __ldg(gpos + min(tile * FR + f, n_frames - 1) * gstride + a);
    uint32_t k = 0;
    bool draining = false;  // lane 0: a bulk store of the stage may still be in flight

    for (; tile < n_tiles; tile += tile_stride) {
        const long long f0 = tile * FR;
        const int nrows = static_cast<int>(min(static_cast<long long>(FR), n_frames - f0));
        // row a of the "parent" of the root: the identity placed at global_pos
        float r0 = id0, r1 = id1, r2 = id2, pp = gnext;

        for (int c0 = 0; c0 < n_joints; c0 += C) {
            const int cnt = active ? n_joints - c0 : 0;  // joints left (>= 8 except in a partial last chunk); 0 = never store
            float4 q[C];
            {
                const uint32_t buf = k % NB;
                mbar_wait(bar0 + 8 * buf, (k / NB) & 1);
                ++k;
#if PMB_LANES_XOR
                // the box row of this lane's frame is 128-byte aligned, so chunk (jj ^ swz) of it is at
                // (row | swz << 4) ^ (jj << 4): one XOR per read
                const uint32_t row = box0 + buf * BOX + f * 128;
                const uint32_t row_swz = row | (((row >> 7) & 7u) << 4);
#pragma unroll
                for (int jj = 0; jj < C; ++jj) q[jj] = lds128(row_swz ^ (jj << 4));
#else
                const float4 *in_row = in_row0 + buf * (BOX / 16);
                const int swz = (swz_base + buf * (BOX / 128)) & 7;
#pragma unroll
                for (int jj = 0; jj < C; ++jj) q[jj] = in_row[jj ^ swz];
#endif
                {   // the loads must have LANDED before the box is refilled through the async proxy (see fk_kernel.cuh)
                    uint32_t acc = 0;
#pragma unroll
                    for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
                }
                __syncwarp();
                if (lane == 0) issue_next(buf);  // refill with the chunk NB ahead (this tile's or the next tile's)
            }
#if PMB_LANES_HOIST
            // the eight normalisation scales of the chunk, ahead of the walk and of the block boundary below (which
            // keeps ptxas from sinking the MUFU chains back into the walk, where their latency showed as ~15 % of
            // the warp's stall samples)
            float sc[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) sc[jj] = rot_scale(q[jj], 1e-8f);
#endif
            if (c0 == 0) {
                const long long next_tile = tile + tile_stride;
                if (next_tile < n_tiles) gnext = __ldg(gpos + min(next_tile * FR + f, n_frames - 1) * gstride + a);
                if (lane == 0 && draining) bulk_wait_read0();  // the previous tile has left the stage
                __syncwarp();
            }

            // Branch-free walk over the chunk (see fk_rows_kernel.cuh).  Joints past the end of the skeleton are
            // zero quaternions (TMA fill) with padded table entries: identity steps whose stores are predicated off.
#if PMB_LANES_XOR
            const uint32_t tab_c0 = tab0 + c0 * 16;
            float4 e = lds128_ro(tab_c0);
#else
            float4 e = tab[c0];
#endif
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                const int j = c0 + jj;
#if PMB_LANES_XOR
                const float4 e_next = lds128_ro(tab_c0 + 16 * (jj + 1));
#else
                const float4 e_next = tab[j + 1];
#endif
                // parent whose row must come from the stage, or -1; the idle lanes (which shadow lane 0's addresses)
                // never read the stage
                const int p = active ? __float_as_int(e.w) : -1;
                load_parent_row_if(p, rrow + 36 * p, prow + 12 * p, r0, r1, r2, pp);
#if PMB_LANES_HOIST
                const float s = sc[jj];
#else
                const float s = rot_scale(q[jj], 1e-8f);
#endif
                const float w = q[jj].x, x = q[jj].y, y = q[jj].z, z = q[jj].w;
                pp = r0 * e.x + r1 * e.y + r2 * e.z + pp;
                const float cx_ = r1 * z - r2 * y, cy_ = r2 * x - r0 * z, cz_ = r0 * y - r1 * x;
                const float ex = w * cx_ + (cy_ * z - cz_ * y);
                const float ey = w * cy_ + (cz_ * x - cx_ * z);
                const float ez = w * cz_ + (cx_ * y - cy_ * x);
                r0 = s * ex + r0, r1 = s * ey + r1, r2 = s * ez + r2;
                store_row_if<1, false>(cnt - jj, rrow + 36 * j, prow + 12 * j, r0, r1, r2, pp);
                e = e_next;
            }
        }

        // the tile's output: two contiguous spans
        float *rg = rout + f0 * rpitch, *pg = pos + f0 * ppitch;
        if (nrows == FR) {
            fence_proxy_async_smem();  // this lane's stage writes -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                bulk_store(rg, smem_u32(Rst), static_cast<uint32_t>(FR * rpitch * 4));
                bulk_store(pg, smem_u32(Pst), static_cast<uint32_t>(FR * ppitch * 4));
                bulk_commit();
                draining = true;
            }
        } else {  // remainder tile (the last one): plain copies
            __syncwarp();
            for (int i = lane; i < nrows * rpitch; i += kWarp) rg[i] = Rst[i];
            for (int i = lane; i < nrows * ppitch; i += kWarp) pg[i] = Pst[i];
            __syncwarp();
        }
    }
    if (lane == 0 && draining) bulk_wait0();  // global writes of the last tile are complete at exit
}

}  // namespace pmb
