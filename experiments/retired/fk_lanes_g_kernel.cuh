// Forward kinematics, lane = (frame, row) kernel with a GROUPED stage (ops/skeleton.py:16-61 of the reference).
//
// fk_lanes_kernel.cuh stages the whole output of a tile (FR x 48 J bytes) so that the TMA engine can write it as
// one contiguous span; for large skeletons that stage is what caps the frames an SM keeps in flight (64 .. 80 at
// 52 .. 65 joints against the ~120 the 22-joint headline runs best with), and throughput follows frames in flight
// (DESIGN.md section 4.1).  This variant stages only G = 32 (or 16) joints at a time:
//
//   * the stage is FR x 48 G bytes whatever the joint count, padded to bank-conflict-free row strides (no dense
//     image is needed any more, so ANY joint count is conflict free, including the multiples of 8 the dense
//     kernels cannot take);
//   * after every G joints the warp copies the group out itself: per frame one contiguous piece of 36 G bytes
//     (1152 bytes for G = 32) of the rotation rows and 12 G bytes of the positions, with the short-period lane
//     map of fk_kernel.cuh (a lane's stage / global offsets repeat every row, so a store costs LDS + STG);
//   * ancestors outside the current group are no longer in the stage, so rows of branch joints live in per-warp
//     slots (float4 = row + position component per lane, allocated by the host's linear scan, joint_program.h):
//     one predicated 16-byte load / store instead of the four 4-byte parent loads of the dense kernels.
//
// Same walk otherwise: lane = 3 f + a, 4-register row chain, rows rotated by the conjugate quaternion, branch
// free, TMA boxes of 8 joints x FR frames through an NB-deep ring.  No TMA store, no drain wait.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "fk_kernel.cuh"       // copy_out_periodic
#include "fk_rows_kernel.cuh"  // rot_scale
#include "tma.cuh"

namespace pmb {

// conflict-free row strides (words) for lane = (frame, row): rotation rows 9 G words + pad with stride = 9 (mod 32),
// positions 3 G words + pad with stride = 3 (mod 32): the 30 active lanes then hit 30 different banks
__host__ __device__ constexpr int fk_lg_sr(int g) { return ((9 * g + 22) / 32) * 32 + 9; }
__host__ __device__ constexpr int fk_lg_sp(int g) { return ((3 * g + 28) / 32) * 32 + 3; }

struct FkLanesGGeom {
    int box_bytes, tab_bytes, warp_bytes, block_bytes;
};
__host__ __device__ inline FkLanesGGeom fk_lanes_g_geom(int fr, int warps, int n_joints, int n_slots, int n_boxes, int g) {
    FkLanesGGeom geo;
    geo.box_bytes = fr * 128;
    geo.tab_bytes = ((n_joints + kChunk) * 16 + 127) & ~127;
    // per warp: NB boxes | rotation stage | position stage | slots | NB mbarriers (32 bytes reserved) | 32 fence words
    geo.warp_bytes = ((n_boxes * geo.box_bytes + fr * (fk_lg_sr(g) + fk_lg_sp(g)) * 4 + 15) & ~15) + n_slots * kWarp * 16 + 32 + 128;
    geo.warp_bytes = (geo.warp_bytes + 127) & ~127;
    geo.block_bytes = 128 + geo.tab_bytes + warps * geo.warp_bytes;
    return geo;
}

__device__ __forceinline__ void load_slot_if(int take /* iff >= 0 */, uint32_t addr, float &r0, float &r1, float &r2, float &pp) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %4, 0;\n"
        "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%5];\n"
        "}"
        : "+f"(r0), "+f"(r1), "+f"(r2), "+f"(pp)
        : "r"(take), "r"(addr));
}
__device__ __forceinline__ void store_slot_if(int keep /* iff >= 0 */, uint32_t addr, float r0, float r1, float r2, float pp) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %0, 0;\n"
        "@p st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n"
        "}" ::"r"(keep), "r"(addr), "f"(r0), "f"(r1), "f"(r2), "f"(pp));
}
__device__ __forceinline__ void store_row4_if(int keep /* iff > 0 */, uint32_t raddr, uint32_t paddr, float r0, float r1, float r2,
                                              float pp) {
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

template <int FR, int WARPS, int NB, int G>
__global__ void __launch_bounds__(WARPS *kWarp)
fk_lanes_g_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
                  const float *__restrict__ offsets, float *__restrict__ pos, float *__restrict__ rout,
                  long long n_frames, int n_joints, int n_slots, const __grid_constant__ JointProgram prog) {
    constexpr int C = kChunk;
    constexpr int BOX = FR * 128;
    constexpr int SR = fk_lg_sr(G), SP = fk_lg_sp(G);
    static_assert(G % C == 0, "a flush group is a whole number of TMA chunks");
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const FkLanesGGeom geo = fk_lanes_g_geom(FR, WARPS, n_joints, n_slots, NB, G);

    float4 *tab = reinterpret_cast<float4 *>(smem_raw);
    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    const float4 *boxes = reinterpret_cast<const float4 *>(mine);
    float *Rst = reinterpret_cast<float *>(mine + NB * BOX);
    float *Pst = Rst + FR * SR;
    unsigned char *after_stage = mine + ((NB * BOX + FR * (SR + SP) * 4 + 15) & ~15);
    const uint32_t slots0 = smem_u32(after_stage) + lane * 16;                   // slot s of this lane: + s * 512
    const uint32_t bar0 = smem_u32(after_stage + n_slots * kWarp * 16);          // NB mbarriers
    const uint32_t fence_word = bar0 + 32 + 4 * lane;
    const uint32_t box0 = smem_u32(mine);

    // Joint table: offset (x, y, z) | slot program: bits 0-7 the slot the parent row comes from (0xFF: the previous
    // joint, still in registers), bits 8-15 the slot this joint's row is saved to (0xFF: none).  offsets[0] is
    // ignored by the reference (skeleton.py:49): zero entry, the root's parent is the identity placed at global_pos.
    for (int j = threadIdx.x; j < n_joints + kChunk; j += WARPS * kWarp) {
        float4 e = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xFFFFu));
        if (j < n_joints) {
            if (j > 0) e.x = offsets[3 * j], e.y = offsets[3 * j + 1], e.z = offsets[3 * j + 2];
            e.w = __uint_as_float(prog.code[j] & 0xFFFFu);
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

    const bool active = lane < 3 * FR;
    const int f = active ? lane / 3 : 0, a = active ? lane - 3 * f : 0;
    const uint32_t rrow = smem_u32(Rst) + (f * SR + 3 * a) * 4;  // this lane's row of the group's first joint
    const uint32_t prow = smem_u32(Pst) + (f * SP + a) * 4;
    const float4 *in_row0 = boxes + f * C;
    const int swz_base = (box0 >> 7) + f;
    const float id0 = a == 0 ? 1.f : 0.f, id1 = a == 1 ? 1.f : 0.f, id2 = a == 2 ? 1.f : 0.f;

    long long la_tile = tile;
    int la_c0 = 0;
    auto issue_next = [&](int buf) {  // lane 0: the warp's chunks in processing order, across its tiles
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

    float gnext = 0.f;
    if (tile < n_tiles) gnext = __ldg(gpos + min(tile * FR + f, n_frames - 1) * gstride + a);
    uint32_t k = 0;

    for (; tile < n_tiles; tile += tile_stride) {
        const long long f0 = tile * FR;
        const int nrows = static_cast<int>(min(static_cast<long long>(FR), n_frames - f0));
        float r0 = id0, r1 = id1, r2 = id2, pp = gnext;
        int gj = 0;  // joints already staged in the current flush group

        for (int c0 = 0; c0 < n_joints; c0 += C) {
            const int cnt_all = n_joints - c0;             // joints left (>= 8 except in a partial last chunk)
            const int cnt = active ? cnt_all : 0;           // 0 = never store
            const bool last_chunk = cnt_all <= C;
            const uint32_t buf = k % NB;
            mbar_wait(bar0 + 8 * buf, (k / NB) & 1);
            ++k;
            const float4 *in_row = in_row0 + buf * (BOX / 16);
            const int swz = (swz_base + buf * (BOX / 128)) & 7;
            float4 q[C];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) q[jj] = in_row[jj ^ swz];
            {   // the loads must have LANDED before the box is refilled through the async proxy (see fk_kernel.cuh)
                uint32_t acc = 0;
#pragma unroll
                for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
            }
            __syncwarp();
            if (lane == 0) issue_next(buf);
            if (c0 == 0) {
                const long long next_tile = tile + tile_stride;
                if (next_tile < n_tiles) gnext = __ldg(gpos + min(next_tile * FR + f, n_frames - 1) * gstride + a);
            }

            // Branch-free walk over the chunk (see fk_rows_kernel.cuh); joints past the end of the skeleton are zero
            // quaternions with padded table entries: identity steps that neither load, save nor store.
            const uint32_t rs = rrow + 36 * gj, ps = prow + 12 * gj;
            float4 e = tab[c0];
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                const float4 e_next = tab[c0 + jj + 1];
                const uint32_t code = __float_as_uint(e.w);
                const int src = static_cast<int>(code & 0xFFu), sav = static_cast<int>((code >> 8) & 0xFFu);
                // 0xFF -> -1 (not taken); the idle lanes never touch the slots
                load_slot_if(active && src != 0xFF ? src : -1, slots0 + src * (kWarp * 16), r0, r1, r2, pp);
                const float s = rot_scale(q[jj], 1e-8f);
                const float w = q[jj].x, x = q[jj].y, y = q[jj].z, z = q[jj].w;
                pp = r0 * e.x + r1 * e.y + r2 * e.z + pp;
                const float cx_ = r1 * z - r2 * y, cy_ = r2 * x - r0 * z, cz_ = r0 * y - r1 * x;
                const float ex = w * cx_ + (cy_ * z - cz_ * y);
                const float ey = w * cy_ + (cz_ * x - cx_ * z);
                const float ez = w * cz_ + (cx_ * y - cy_ * x);
                r0 = s * ex + r0, r1 = s * ey + r1, r2 = s * ez + r2;
                store_slot_if(active && sav != 0xFF ? sav : -1, slots0 + sav * (kWarp * 16), r0, r1, r2, pp);
                store_row4_if(cnt - jj, rs + 36 * jj, ps + 12 * jj, r0, r1, r2, pp);
                e = e_next;
            }
            gj += min(C, cnt_all);

            if (gj == G || last_chunk) {
                __syncwarp();
                const int g0 = c0 + min(C, cnt_all) - gj;  // first joint of the group
                float *rg = rout + (f0 * n_joints + g0) * 9;
                float *pg = pos + (f0 * n_joints + g0) * 3;
                if (gj == G && nrows == FR) {
                    copy_out_periodic<9 * G, SR, 1, FR>(Rst, rg, rpitch, lane);
                    copy_out_periodic<3 * G, SP, 1, FR>(Pst, pg, ppitch, lane);
                } else {  // remainder group / remainder tile: one row at a time, lanes across the row
                    const int wr = 9 * gj, wp = 3 * gj;
                    for (int r = 0; r < nrows; ++r) {
                        for (int c = lane; c < wr; c += kWarp) rg[static_cast<long long>(r) * rpitch + c] = Rst[r * SR + c];
                        for (int c = lane; c < wp; c += kWarp) pg[static_cast<long long>(r) * ppitch + c] = Pst[r * SP + c];
                    }
                }
                __syncwarp();
                gj = 0;
            }
        }
    }
}

}  // namespace pmb
