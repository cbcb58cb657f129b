// RETIRED (round 2): from_root_positions as a tile kernel -- measured alternative, not part of the library.
//
// Measured on B200 (profiles/r2_sweep_frp.jsonl), tile kernel against the shipping thread-per-frame kernel with paired
// 32-byte stores:   1M x 22  0.253 ms vs 0.182,   4M x 52  3.12 vs 2.11,   4M x 65  5.00 vs 3.06.
// ncu at 4M x 65 (profiles/r2_frp_tile_4m_x_65_ncu_summary.txt): DRAM traffic = algorithmic (3.13 + 4.11 GB), no global
// store instructions -- but 5 warps per SM (shared memory: 1.2 KB per frame in flight), every frame one dependent chain of
// ~218 instructions per alignment: issue slots 37 % used, 'wait' (fixed-latency dependency) the top stall.  More warps need
// less shared memory per frame or parallelism inside a frame (independent subtrees), and the latter breaks the in-order
// 8-joint output stage.  What actually bounded the thread-per-frame kernel were its half-sector stores, which pairing fixed.
// Depends on ik_kernels.cuh (ik_from_to, ik_from_to_axis, ...) and tma.cuh of the library.
#pragma once
#include "../../pymotion_b200/csrc/ik_kernels.cuh"

namespace pmb {

// Tile kernel: a warp owns tiles of 32 consecutive frames (thread = frame, as above) but never touches global memory
// with a strided access.
//   input    the tile's positions -- 32 x 12 J contiguous bytes -- arrive as ONE bulk copy into shared memory (compulsory
//            DRAM traffic only; every point is then an LDS);
//   output   rotations go through a double-buffered stage of 8 joints per frame (rows 144 bytes apart: conflict-free
//            16-byte stores); after every 8 joints each lane hands its own 128-byte row piece to the TMA engine
//            (cp.async.bulk shared -> global): whole sectors, no store instructions, no cross-lane traffic;
//   tiles    claimed from a per-block counter (the blocks own interleaved tiles), so warps never wait for each other;
//   walk     shared memory leaves room for 5 .. 12 such warps per SM, so a warp has to run without stalls on its own: the
//            host flattens the tree into ITEMS (one alignment each: the first child of a joint, a further child, or a
//            leaf), one 16-byte table entry per item holding the child's rest direction and every index the item needs.
//            The entry of item i + 2, the two points and the parent slot of item i + 1 are fetched while item i computes:
//            nothing but arithmetic is left on the dependent chain.  (With the per-joint tables walked as in the kernel
//            above -- program word -> child range -> child index -> rest direction / points, four dependent look-ups
//            per joint -- the same tile structure took ~900 cycles per joint: 4.5 ms at 4M x 65.)
// Shared memory per warp: 384 J (positions) + 9 KB (stage) + 512 per live branch slot.
constexpr int kFrpGroup = 8;                          // joints per flush
constexpr int kFrpStagePitch = kFrpGroup * 16 + 16;   // bytes per frame row of the stage (odd multiple of 16)
constexpr int kFrpMaxItems = 2 * PMB_MAX_JOINTS;
constexpr uint32_t kFrpLeaf = 0u, kFrpFirst = 1u, kFrpFurther = 2u, kFrpNoSlot = 31u;
// Item word: bits 0-8 joint | 9-17 child (the joint itself for a leaf) | 18-19 kind | 20 last item of its joint |
// 21-25 slot the parent's global rotation comes from (31: the previous joint, in registers) | 26-30 slot this joint's
// global rotation is saved to (31: none).
struct FrpItems {
    uint32_t word[kFrpMaxItems];
};
__host__ __device__ __forceinline__ uint32_t frp_item(uint32_t j, uint32_t c, uint32_t kind, bool last, uint32_t src, uint32_t save) {
    return j | (c << 9) | (kind << 18) | (last ? 1u << 20 : 0u) | (src << 21) | (save << 26);
}
struct FrpTileGeom {
    int tab_bytes, in_bytes, stage_bytes, slot_bytes, warp_bytes, block_bytes;
};
__host__ __device__ inline FrpTileGeom frp_tile_geom(int warps, int n_joints, int n_slots, int n_items) {
    FrpTileGeom g;
    g.tab_bytes = (((n_items + 2) * 16 + 127) & ~127) + 128;  // items (+ two of padding for the look-ahead) + the tile counter
    g.in_bytes = (32 * 12 * n_joints + 127) & ~127;
    g.stage_bytes = 2 * 32 * kFrpStagePitch;
    g.slot_bytes = n_slots * 32 * 16;
    g.warp_bytes = g.in_bytes + g.stage_bytes + g.slot_bytes + 128;  // + the mbarrier
    g.block_bytes = 128 + g.tab_bytes + warps * g.warp_bytes;
    return g;
}

template <bool FAST>
__global__ void __launch_bounds__(512, 1)
from_root_positions_tile_kernel(const float *__restrict__ pos, const float *__restrict__ offsets, float4 *__restrict__ rots,
                                long long n_frames, int n_joints, int n_slots, int n_items,
                                const __grid_constant__ FrpItems items) {
    extern __shared__ __align__(128) unsigned char smem_frp[];
    unsigned char *smem_raw = smem_frp + ((128u - (smem_u32(smem_frp) & 127u)) & 127u);
    const int warps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const FrpTileGeom geo = frp_tile_geom(warps, n_joints, n_slots, n_items);
    // item table: rest direction of the item's child -- from_to / from_to_axis normalise their first argument (quat.py:541,
    // :616), the same value for every frame -- and the item word
    uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
    for (int i = threadIdx.x; i < n_items + 2; i += blockDim.x) {
        uint4 e = make_uint4(0u, 0u, 0u, frp_item(0, 0, kFrpLeaf, false, kFrpNoSlot, kFrpNoSlot));
        if (i < n_items) {
            const uint32_t w = items.word[i];
            const int c = (w >> 9) & 0x1FFu;
            const Vec3<float> d = v_normalize(Vec3<float>{offsets[3 * c], offsets[3 * c + 1], offsets[3 * c + 2]}, 1e-8f);
            e = make_uint4(__float_as_uint(d.x), __float_as_uint(d.y), __float_as_uint(d.z), w);
        }
        tab[i] = e;
    }
    uint32_t *tile_counter = reinterpret_cast<uint32_t *>(smem_raw + geo.tab_bytes - 128);
    if (threadIdx.x == 0) *tile_counter = 0u;
    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    float *in = reinterpret_cast<float *>(mine);
    const uint32_t in_addr = smem_u32(mine);
    const uint32_t stage0 = in_addr + geo.in_bytes;
    float4 *slots = reinterpret_cast<float4 *>(mine + geo.in_bytes + geo.stage_bytes);
    const uint32_t bar = stage0 + geo.stage_bytes + geo.slot_bytes;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();  // table, counter, barriers; from here on the warps never meet again

    const long long n_tiles = (n_frames + 31) / 32;
    auto claim = [&]() -> long long {
        uint32_t n = 0;
        if (lane == 0) n = atomicAdd(tile_counter, 1u);
        n = __shfl_sync(0xffffffffu, n, 0);
        return static_cast<long long>(n) * gridDim.x + blockIdx.x;
    };
    const int row_floats = 3 * n_joints;
    const float *my_row = in + lane * row_floats;
    const uint32_t my_stage = stage0 + lane * kFrpStagePitch;
    uint32_t phase = 0;
    uint32_t half = 0;      // which half of the stage the current group of joints goes to: alternates flush by flush
    bool stored = false;    // this lane has handed something to the engine
    auto point = [&](uint32_t j) { return Vec3<float>{my_row[3 * j], my_row[3 * j + 1], my_row[3 * j + 2]}; };

    for (long long tile = claim(); tile < n_tiles; tile = claim()) {
        const long long f0 = tile * 32;
        const int nrows = static_cast<int>(min(32LL, n_frames - f0));
        const uint32_t bytes = static_cast<uint32_t>(nrows) * 12u * n_joints;
        const float *src = pos + f0 * row_floats;
        if ((bytes & 15u) == 0) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, bytes);
                bulk_load_1d(in_addr, src, bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {  // the remainder tile of a batch whose byte count is not a multiple of 16: plain copies
            for (int i = lane; i < nrows * row_floats; i += 32) in[i] = __ldg(src + i);
            __syncwarp();
        }
        if (lane < nrows) {
            float4 *R = rots + (f0 + lane) * n_joints;
            uint4 e0 = tab[0], e1 = tab[1];
            Vec3<float> a0 = point(e0.w & 0x1FFu), b0 = point((e0.w >> 9) & 0x1FFu);
            float4 g0 = make_float4(1.f, 0.f, 0.f, 0.f);  // the root's parent is the world
            Quat<float> cur{1.f, 0.f, 0.f, 0.f}, G{1.f, 0.f, 0.f, 0.f}, rot{1.f, 0.f, 0.f, 0.f};
            Vec3<float> pj{0.f, 0.f, 0.f}, to_c{0.f, 0.f, 0.f};
            for (int i = 0; i < n_items; ++i) {
                // look-ahead: entry of item i + 2; points and parent slot of item i + 1
                const uint4 e2 = tab[i + 2];
                const Vec3<float> a1 = point(e1.w & 0x1FFu), b1 = point((e1.w >> 9) & 0x1FFu);
                const uint32_t src1 = (e1.w >> 21) & 31u;
                float4 g1 = make_float4(1.f, 0.f, 0.f, 0.f);
                if (src1 != kFrpNoSlot) g1 = slots[src1 * 32 + lane];

                const uint32_t w = e0.w;
                const uint32_t kind = (w >> 18) & 3u;
                const Vec3<float> dir{__uint_as_float(e0.x), __uint_as_float(e0.y), __uint_as_float(e0.z)};
                if (kind != kFrpFurther) {  // first item of its joint
                    G = ((w >> 21) & 31u) == kFrpNoSlot ? cur : Quat<float>{g0.x, g0.y, g0.z, g0.w};
                    rot = {1.f, 0.f, 0.f, 0.f};
                }
                if (kind == kFrpFirst) {
                    pj = a0;
                    to_c = {b0.x - a0.x, b0.y - a0.y, b0.z - a0.z};
                    rot = ik_from_to<FAST>(dir, q_rotate(q_conj(G), to_c));
                } else if (kind == kFrpFurther) {
                    const Quat<float> inv = q_conj(q_mul(G, ik_q_normalize<FAST>(rot)));  // fk normalises local rotations (quat.py:411)
                    const Vec3<float> pred = q_rotate(inv, Vec3<float>{b0.x - pj.x, b0.y - pj.y, b0.z - pj.z});
                    const Vec3<float> axis = q_rotate(inv, ik_v_normalize<FAST>(to_c));
                    rot = q_mul(rot, ik_from_to_axis<FAST>(dir, pred, axis));
                }
                if (w & (1u << 20)) {  // last item of the joint: its rotation is final
                    const int j = static_cast<int>(w & 0x1FFu);
                    const int jj = j & (kFrpGroup - 1);
                    const uint32_t base = my_stage + half * (32 * kFrpStagePitch);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + 16 * jj), "f"(rot.w), "f"(rot.x), "f"(rot.y), "f"(rot.z) : "memory");
                    if (jj == kFrpGroup - 1 || j == n_joints - 1) {
                        fence_proxy_async_smem();
                        bulk_store(R + (j - jj), base, static_cast<uint32_t>(jj + 1) * 16u);
                        bulk_commit();
                        // the OTHER half -- handed over one flush (8 joints of arithmetic) ago -- must have been read
                        // before the next group is written there
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        half ^= 1u, stored = true;
                    }
                    cur = q_mul(G, ik_q_normalize<FAST>(rot));
                    const uint32_t sv = (w >> 26) & 31u;
                    if (sv != kFrpNoSlot) slots[sv * 32 + lane] = make_float4(cur.w, cur.x, cur.y, cur.z);
                }
                e0 = e1, e1 = e2, a0 = a1, b0 = b1, g0 = g1;
            }
        }
        __syncwarp();  // every lane has read its row: the buffer may be refilled
    }
    if (stored) bulk_wait0();  // global writes complete at exit
}

}  // namespace pmb
