// Forward kinematics, matrix track kernel (ops/skeleton.py:16-61 of the reference): the quaternion track kernel's mapping
// (qtracks_kernel.cuh) with the 3x4 transform as the chain state.
//
// The row kernels (fk_rows / fk_tracks) give a lane ONE ROW of a frame's transform: three lanes per (frame,
// joint), 30 of 32 lanes busy, ~57 instructions per step for 10 (frame, joint) items -- ncu at 4M x 65: 2.24e9 warp
// instructions, issue slots 66 % busy, and shared memory rules out more warps.  Here a lane owns a WHOLE (frame, joint):
//   lanes    4 tracks x 8 frames.  The host's whole-skeleton four-track level schedule (track_schedule.h) puts four
//            independent joints of the tile's 8 frames into every step (65 joints: 18 steps, 52: 16, 22: 8): ~100 instructions
//            per step for 32 items -- less than half the instructions per pose of the row kernels.
//   walk     G = P [R(q^) | off]: the local matrix of step s + 1 (normalisation q / (|q| + 1e-8) and the nine products of
//            quat.py:293-315) is formed while step s computes; the step-to-step chain is parent read -> 27 + 9 FMA -> store.
//            A joint whose parent was the same track's previous item keeps it in registers, every other parent is read back
//            from the stage, which already holds every joint processed so far.  to_matrix(0) = I: the zero quaternion needs no
//            special case.
//   stage    the dense image of the tile's output -- 8 x 36 J bytes of matrices, 8 x 12 J bytes of positions -- so both spans
//            of a full tile are multiples of 16 bytes for EVERY joint count and leave as two TMA bulk stores.  The remainder
//            tile of a batch is copied out by the lanes.
//   input    the tile's quaternions (8 rows of 16 J contiguous bytes) as bulk copies into a double buffer, a tile ahead; rows
//            padded to an odd number of 16-byte units (conflict-free 16-byte reads of a quarter warp).
//   tiles    claimed from a per-block counter (see qtracks_kernel.cuh).
// Algorithmic HBM traffic 64 J + 12 bytes per pose.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"
#include "track_schedule.h"

namespace pmb {

constexpr int kMtFrames = 8, kMtTracks = 4;

struct MtGeom {
    int in_pitch, in_bytes, r_bytes, p_bytes, tab_bytes, warp_bytes, block_bytes;
};
__host__ __device__ inline MtGeom mt_geom(int warps, int n_joints, int n_items) {
    MtGeom g;
    g.in_pitch = 16 * (n_joints | 1);
    g.in_bytes = kMtFrames * g.in_pitch;
    g.r_bytes = kMtFrames * 36 * n_joints;  // 288 J: a multiple of 16
    g.p_bytes = kMtFrames * 12 * n_joints;  //  96 J: a multiple of 16
    g.tab_bytes = (((n_items + kMtTracks) * 16 + 127) & ~127) + 128;  // + one step of no-ops (look-ahead) + the tile counter
    g.warp_bytes = (2 * g.in_bytes + g.r_bytes + g.p_bytes + 16 + 128 + 127) & ~127;  // + 2 mbarriers + fence words
    g.block_bytes = 128 + g.tab_bytes + warps * g.warp_bytes;
    return g;
}

// A transform is 9 + 3 words of the dense stage at word offsets 9 (J f + j) and 3 (J f + j): only 4-byte aligned, one 32-bit
// access per word.  The 32 lanes of an access -- 8 frames x 4 joints -- collide in the banks whenever two of the step's joints
// are fewer than 8 indices apart modulo 32 (ncu at 4M x 65: 3.2 wavefronts per STS, 2.1 per LDS).  Measured alternative: 64-bit
// pairs chosen per lane by the parity of J f + j (10 predicated instructions of 16 lanes each instead of 9 of 32) halves the
// wavefronts but adds 16 instructions per step: slower (4M x 65: 3.07 - 3.38 ms against 2.75 - 2.94).
//
// parent transform from the stage, taken iff (int)flag >= 0 (the carry flag sits in the sign bit)
__device__ __forceinline__ void mt_load_parent_if(uint32_t flag, uint32_t raddr, uint32_t paddr, float (&r)[9], float (&p)[3]) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ge.s32 q, %12, 0;\n"
        "@q ld.shared.f32 %0, [%13];\n"
        "@q ld.shared.f32 %1, [%13+4];\n"
        "@q ld.shared.f32 %2, [%13+8];\n"
        "@q ld.shared.f32 %3, [%13+12];\n"
        "@q ld.shared.f32 %4, [%13+16];\n"
        "@q ld.shared.f32 %5, [%13+20];\n"
        "@q ld.shared.f32 %6, [%13+24];\n"
        "@q ld.shared.f32 %7, [%13+28];\n"
        "@q ld.shared.f32 %8, [%13+32];\n"
        "@q ld.shared.f32 %9, [%14];\n"
        "@q ld.shared.f32 %10, [%14+4];\n"
        "@q ld.shared.f32 %11, [%14+8];\n"
        "}"
        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]), "+f"(p[0]),
          "+f"(p[1]), "+f"(p[2])
        : "r"(flag), "r"(raddr), "r"(paddr));
}
// stored iff (int)flag >= 0 (the no-op flag sits in the sign bit)
__device__ __forceinline__ void mt_store_if(uint32_t flag, uint32_t raddr, uint32_t paddr, const float (&r)[9], const float (&p)[3]) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ge.s32 q, %0, 0;\n"
        "@q st.shared.f32 [%1], %3;\n"
        "@q st.shared.f32 [%1+4], %4;\n"
        "@q st.shared.f32 [%1+8], %5;\n"
        "@q st.shared.f32 [%1+12], %6;\n"
        "@q st.shared.f32 [%1+16], %7;\n"
        "@q st.shared.f32 [%1+20], %8;\n"
        "@q st.shared.f32 [%1+24], %9;\n"
        "@q st.shared.f32 [%1+28], %10;\n"
        "@q st.shared.f32 [%1+32], %11;\n"
        "@q st.shared.f32 [%2], %12;\n"
        "@q st.shared.f32 [%2+4], %13;\n"
        "@q st.shared.f32 [%2+8], %14;\n"
        "}" ::"r"(flag), "r"(raddr), "r"(paddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
        "f"(r[8]), "f"(p[0]), "f"(p[1]), "f"(p[2])
        : "memory");
}

__global__ void __launch_bounds__(512, 1)
fk_mtracks_kernel(const float4 *__restrict__ rot, const float *__restrict__ gpos, long long gstride, const float *__restrict__ offsets,
                  float *__restrict__ out_p, float *__restrict__ out_r, long long n_frames, int n_joints, int n_steps,
                  const __grid_constant__ TrackProgram prog) {
    constexpr int FQ = kMtFrames, NT = kMtTracks;
    extern __shared__ __align__(128) unsigned char smem_mt[];
    unsigned char *smem_raw = smem_mt + ((128u - (smem_u32(smem_mt) & 127u)) & 127u);
    const int warps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_items = n_steps * NT;
    const MtGeom geo = mt_geom(warps, n_joints, n_items);

    // Item table, 16 bytes per item: offset (x, y, z) | word: bits 0-9 joint, 10-19 parent, 30 = parent in the track's
    // registers, 31 (+ 30) = no-op.  offsets[0] is ignored by the reference (the root translation is global_pos, skeleton.py:49).
    uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
    uint32_t *tile_counter = reinterpret_cast<uint32_t *>(smem_raw + geo.tab_bytes - 128);
    if (threadIdx.x == 0) *tile_counter = 0u;
    for (int i = threadIdx.x; i < n_items + NT; i += blockDim.x) {
        const uint32_t c = i < n_items ? prog.code[i] : kTrackNoop;
        const uint32_t j = track_joint(c), p = track_parent(c);
        uint4 e = make_uint4(0u, 0u, 0u, 0xC0000000u);
        if (!(c & kTrackNoop)) {
            if (j > 0) e.x = __float_as_uint(offsets[3 * j]), e.y = __float_as_uint(offsets[3 * j + 1]), e.z = __float_as_uint(offsets[3 * j + 2]);
            e.w = j | (p << 10) | ((c & kTrackCarry) ? 0x40000000u : 0u);
        }
        tab[i] = e;
    }
    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    const uint32_t in0 = smem_u32(mine);
    const uint32_t rst = in0 + 2 * geo.in_bytes;
    const uint32_t pst = rst + geo.r_bytes;
    const uint32_t bar0 = pst + geo.p_bytes;  // two mbarriers
    const uint32_t fence_word = bar0 + 16 + 4 * lane;
    const uint32_t tab0 = smem_u32(tab);
    if (lane == 0) {
        mbar_init(bar0, 1), mbar_init(bar0 + 8, 1);
        fence_barrier_init();
    }
    const int trk = lane >> 3, f = lane & 7;
    __syncthreads();  // table, counter, barriers; from here on the warps never meet again

    const long long n_tiles = (n_frames + FQ - 1) / FQ;
    auto claim = [&]() -> long long {
        uint32_t n = 0;
        if (lane == 0) n = atomicAdd(tile_counter, 1u);
        n = __shfl_sync(0xffffffffu, n, 0);
        return static_cast<long long>(n) * gridDim.x + blockIdx.x;
    };
    long long tile = claim(), tile_next = claim();
    if (tile >= n_tiles) return;
    const int row_bytes = 16 * n_joints;
    const int rpitch = 36 * n_joints, ppitch = 12 * n_joints;
    const uint32_t in_row0 = in0 + f * geo.in_pitch;
    const uint32_t r_row = rst + f * rpitch, p_row = pst + f * ppitch;

    // lanes 0 .. 7 load the frame rows of a tile; lane 0 announces the bytes
    auto issue_tile = [&](long long t, int buf) {
        if (t < n_tiles) {
            const int rows = static_cast<int>(min(static_cast<long long>(FQ), n_frames - t * FQ));
            if (lane == 0) mbar_arrive_expect_tx(bar0 + 8 * buf, static_cast<uint32_t>(rows * row_bytes));
            if (lane < rows) bulk_load_1d(in0 + buf * geo.in_bytes + lane * geo.in_pitch, rot + (t * FQ + lane) * n_joints,
                                          static_cast<uint32_t>(row_bytes), bar0 + 8 * buf);
        }
    };
    issue_tile(tile, 0);
    issue_tile(tile_next, 1);

    float gn0, gn1, gn2;  // root position of this lane's frame in the NEXT tile, fetched a tile early
    {
        const float *g = gpos + min(tile * FQ + f, n_frames - 1) * gstride;
        gn0 = __ldg(g), gn1 = __ldg(g + 1), gn2 = __ldg(g + 2);
    }
    uint32_t k = 0;
    bool draining = false;  // lane 0: a bulk store of the stage may still be in flight

    // local matrix of an item: R(q / (|q| + 1e-8)), quat.py:411-423 and :293-315
    auto local_matrix = [&](const float4 &qv, float (&m)[9]) {
        const Quat<float> q = q_normalize_fast(Quat<float>{qv.x, qv.y, qv.z, qv.w}, 1e-8f);
        q_to_matrix(q, m);
    };

    for (; tile < n_tiles; ++k) {
        const long long f0 = tile * FQ;
        const int nrows = static_cast<int>(min(static_cast<long long>(FQ), n_frames - f0));
        const int buf = k & 1;
        const uint32_t in_row = in_row0 + buf * geo.in_bytes;
        // track 0 starts from the "parent" of the root: the identity placed at global_pos
        float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        float T[3] = {gn0, gn1, gn2};
        if (tile_next < n_tiles) {
            const float *g = gpos + min(tile_next * FQ + f, n_frames - 1) * gstride;
            gn0 = __ldg(g), gn1 = __ldg(g + 1), gn2 = __ldg(g + 2);
        }
        mbar_wait(bar0 + 8 * buf, (k >> 1) & 1);
        if (lane == 0 && draining) bulk_wait_read0();  // the previous tile has left the stage
        __syncwarp();

        uint32_t acc = 0;
        uint32_t tab_addr = tab0 + trk * 16;
        float4 e = lds128_ro(tab_addr);
        float L[9];
        {
            const float4 qv = lds128(in_row + 16 * (__float_as_uint(e.w) & 0x3FFu));
            acc |= __float_as_uint(qv.x);
            local_matrix(qv, L);
        }
        for (int step = 0; step < n_steps; ++step) {
            const uint32_t w = __float_as_uint(e.w);
            const uint32_t j = w & 0x3FFu, p = (w >> 10) & 0x3FFu;
            // parent: the track's registers, or the stage (a joint stored in an earlier step)
            mt_load_parent_if(w << 1, r_row + 36 * p, p_row + 12 * p, R, T);
            // look-ahead, issued after the parent reads (see qtracks_kernel.cuh): entry, quaternion and local matrix of step s + 1
            tab_addr += NT * 16;
            const float4 e_next = lds128_ro(tab_addr);
            const float4 qv_next = lds128(in_row + 16 * (__float_as_uint(e_next.w) & 0x3FFu));
            acc |= __float_as_uint(qv_next.x);
            float L_next[9];
            local_matrix(qv_next, L_next);

            float G[9];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float p0 = R[3 * a], p1 = R[3 * a + 1], p2 = R[3 * a + 2];
                G[3 * a + 0] = p0 * L[0] + p1 * L[3] + p2 * L[6];
                G[3 * a + 1] = p0 * L[1] + p1 * L[4] + p2 * L[7];
                G[3 * a + 2] = p0 * L[2] + p1 * L[5] + p2 * L[8];
                T[a] = p0 * e.x + p1 * e.y + p2 * e.z + T[a];
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = G[i];
            mt_store_if(w, r_row + 36 * j, p_row + 12 * j, R, T);
            __syncwarp();  // a parent may have been stored by another track
            e = e_next;
#pragma unroll
            for (int i = 0; i < 9; ++i) L[i] = L_next[i];
        }
        // the tile's quaternions have been READ (a store that depends on all of them precedes the refill through the async
        // proxy, see fk_kernel.cuh); fetch the tile after the next one into this buffer
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
        __syncwarp();
        tile = tile_next;
        tile_next = claim();
        issue_tile(tile_next, buf);

        // ---- output: the two dense spans of the tile ----------------------------------------------------------------
        if (nrows == FQ) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store(reinterpret_cast<unsigned char *>(out_r) + f0 * rpitch, rst, static_cast<uint32_t>(FQ * rpitch));
                bulk_store(reinterpret_cast<unsigned char *>(out_p) + f0 * ppitch, pst, static_cast<uint32_t>(FQ * ppitch));
                bulk_commit();
                draining = true;
            }
        } else {  // remainder tile (the last one of the batch): its spans need not be multiples of 16 bytes
            __syncwarp();
            const int nr = nrows * 9 * n_joints, np = nrows * 3 * n_joints;
            float *gr = out_r + f0 * 9 * n_joints, *gp = out_p + f0 * 3 * n_joints;
            for (int i = lane; i < nr; i += 32) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(rst + 4 * i));
                gr[i] = v;
            }
            for (int i = lane; i < np; i += 32) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pst + 4 * i));
                gp[i] = v;
            }
            __syncwarp();
        }
    }
    if (lane == 0 && draining) bulk_wait0();  // global writes of the last tile are complete at exit
}

}  // namespace pmb
