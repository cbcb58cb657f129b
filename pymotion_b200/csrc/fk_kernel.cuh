// Forward kinematics, one fused kernel (ops/skeleton.py:16-61 of the reference).
//
// Mapping: one THREAD per frame, one WARP per tile of 32 consecutive frames; warps are
// persistent (tile t, t + stride, ...) and autonomous (own staging buffers, own mbarrier)
// so load / compose / store phases of different warps overlap on the SM.
//
//   in   the tile's quaternions arrive 8 joints x 32 frames at a time through TMA
//        (cp.async.bulk.tensor.2d, 128-byte hardware swizzle so the thread-per-frame
//        16-byte reads are bank-conflict free); the copy of the next chunk -- or of the
//        warp's next tile -- is in flight while the current chunk is composed;
//   walk the joint table (offsets + per-joint program word) sits in shared memory; a
//        thread composes joint after joint in registers; the parent transform is either
//        still in registers (parent == previous joint, the chain case) or in a per-warp
//        shared-memory slot written when the parent was computed (branch points of the
//        tree; slots are allocated on the host, joint_program.h);
//   out  results are staged per warp in shared memory and written to HBM by the whole
//        warp.  HBM wants every 128-byte line -- better, every DRAM page -- written in
//        one go (measured: flushing 8 joints at a time leaves DRAM ~90 % busy at ~45 % of
//        its byte rate, because a 288-byte piece per frame row splits most lines between
//        two flushes).  So the flush group is as large as shared memory allows:
//          G == 0  whole rows: the stage is the dense image of the tile's output
//                  (32 x 36J and 32 x 12J bytes, both multiples of 128), copied out as one
//                  contiguous, line-aligned span with 16-byte accesses;
//          G > 0   G joints per flush (32, 16 or 8), rows padded to a conflict-free stride,
//                  copied out with the short-period lane map of copy_out_periodic.
//        Thread-per-frame writes into the stage are 64-bit when the joint count is even
//        (rows are then 8-byte aligned) and conflict free in both layouts.
//
// Algorithmic HBM traffic: 64*J + 12 bytes per pose (16J + 12 in, 12J + 36J out), no scratch
// in global memory.  See DESIGN.md for the roofline and the measurements behind the choices.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "tma.cuh"

#ifndef PMB_ST_POLICY
#define PMB_ST_POLICY 0
#endif
#ifndef PMB_FK_UNIFORM_CODE
#define PMB_FK_UNIFORM_CODE 1
#endif
#ifndef PMB_FK_BULK_STORE
#define PMB_FK_BULK_STORE 1
#endif
#ifndef PMB_FK_FAST_NORM
#define PMB_FK_FAST_NORM 1
#endif
#if PMB_FK_FAST_NORM
#define PMB_FK_NORMALIZE q_normalize_fast
#else
#define PMB_FK_NORMALIZE q_normalize
#endif

namespace pmb {

// Output store policy (measured): 0 = default (write-back, normal L2 priority),
// 1 = .cs streaming / evict-first.
template <typename V>
__device__ __forceinline__ void out_store(V *p, const V &v) {
#if PMB_ST_POLICY == 1
    __stcs(p, v);
#else
    *p = v;
#endif
}

__host__ __device__ constexpr int ce_gcd(int a, int b) { return b == 0 ? a : ce_gcd(b, a % b); }
__host__ __device__ constexpr int ce_lcm(int a, int b) { return a / ce_gcd(a, b) * b; }


// Shared-memory geometry, shared by host (sizing) and device.  group = 0: dense rows of n_joints joints.
struct FkGeom {
    int sr, sp;         // row strides of the rotation / position stage, in words
    int warp_bytes;     // stage + slots of one warp
    int block_bytes;    // whole dynamic allocation of a block
};
__host__ __device__ constexpr int fk_pad(int w, int vec) { return vec == 1 ? (w | 1) : (((w / 2) & 1) ? w : w + 2); }
__host__ __device__ inline FkGeom fk_geom(int group, int vec, int rw, int warps, int n_joints, int n_slots) {
    FkGeom g;
    g.sr = group ? fk_pad(rw * group, vec) : rw * n_joints;
    g.sp = group ? fk_pad(3 * group, vec) : 3 * n_joints;
    g.warp_bytes = ((kWarp * (g.sr + g.sp) * 4 + 15) & ~15) + n_slots * 3 * kWarp * 16;
    // 1 KB slack to align the TMA boxes to 1024 | boxes | joint table | per-warp stage + slots | mbarriers | fence words
    g.block_bytes = 1024 + warps * kBoxStages * kBoxBytes + ((n_joints * 16 + 127) & ~127) + warps * g.warp_bytes +
                    warps * kBoxStages * 8 + warps * kWarp * 4;
    return g;
}

// Warp-cooperative copy of a FULL padded stage (ROWS rows x W words, row stride S) to global rows of
// pitch `pitch` words, in units of VEC words.  Store k of the warp serves flat unit 32k + lane ->
// (row, column); that map repeats every P stores / RPP rows, so a lane computes its P offsets once
// and each store costs LDS + STG.
template <int W, int S, int VEC, int ROWS = 32>
__device__ __forceinline__ void copy_out_periodic(const float *__restrict__ stage, float *__restrict__ gtile, int pitch,
                                                  int lane) {
    constexpr int WV = W / VEC;
    constexpr int L = ce_lcm(32, WV);
    constexpr int P = L / 32;     // stores per period
    constexpr int RPP = L / WV;   // rows per period
    static_assert(ROWS % RPP == 0 && P <= 9, "group geometry must give a short period");
    using V = typename std::conditional<VEC == 4, float4, typename std::conditional<VEC == 2, float2, float>::type>::type;
    int soff[P], goff[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int i = 32 * k + lane;
        const int dr = i / WV, col = i - dr * WV;
        soff[k] = dr * S + col * VEC;
        goff[k] = dr * pitch + col * VEC;
    }
#pragma unroll 2
    for (int g = 0; g < ROWS / RPP; ++g) {
        V v[P];
#pragma unroll
        for (int k = 0; k < P; ++k) v[k] = *reinterpret_cast<const V *>(stage + g * RPP * S + soff[k]);
        float *gp = gtile + static_cast<long long>(g * RPP) * pitch;
#pragma unroll
        for (int k = 0; k < P; ++k) out_store(reinterpret_cast<V *>(gp + goff[k]), v[k]);
    }
}

// Remainder group (w < W words per row) and remainder tile (nrows < 32) of the padded layout: one row
// at a time, lanes across the row.
template <int WMAX, int S, int VEC>
__device__ __forceinline__ void copy_out_rows(const float *__restrict__ stage, float *__restrict__ gtile, int pitch,
                                              int nrows, int w, int lane) {
    using V = typename std::conditional<VEC == 2, float2, float>::type;
    constexpr int MAXU = (WMAX / VEC + 31) / 32;
    const int wv = w / VEC;
    const V *sp = reinterpret_cast<const V *>(stage) + lane;
    V *gp = reinterpret_cast<V *>(gtile) + lane;
#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
        V v[MAXU];
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
            if (lane + 32 * u < wv) v[u] = sp[32 * u];
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
            if (lane + 32 * u < wv) out_store(gp + 32 * u, v[u]);
        sp += S / VEC;
        gp += pitch / VEC;
    }
}

// Dense layout: the stage IS the tile's output image -> one contiguous, 16-byte aligned span of n words.
__device__ __forceinline__ void copy_out_flat(const float *__restrict__ stage, float *__restrict__ gtile, int n, int lane) {
    const float4 *s4 = reinterpret_cast<const float4 *>(stage);
    float4 *g4 = reinterpret_cast<float4 *>(gtile);
    const int n4 = n >> 2;
    int i = lane;
    for (; i + 96 < n4; i += 128) {  // 4 x 512 bytes in flight per warp
        const float4 a = s4[i], b = s4[i + 32], c = s4[i + 64], d = s4[i + 96];
        out_store(g4 + i, a), out_store(g4 + i + 32, b), out_store(g4 + i + 64, c), out_store(g4 + i + 96, d);
    }
    for (; i < n4; i += 32) out_store(g4 + i, s4[i]);
    for (int k = 4 * n4 + lane; k < n; k += 32) gtile[k] = stage[k];  // only a remainder tile can leave 1..3 words
}

// G: joints per flush group (0 = whole rows, dense).  QO = false: rout is rotmats [F][J][9]; true: global
// quaternions [F][J][4].
template <int G, int WARPS, int VEC, bool PF_OFFSETS, bool QO>
__global__ void __launch_bounds__(WARPS *kWarp)
fk_chain_kernel(const __grid_constant__ CUtensorMap tm_rot, const float *__restrict__ gpos, long long gstride,
                const float *__restrict__ offsets, long long ostride, float *__restrict__ pos,
                float *__restrict__ rout, long long n_frames, int n_joints, int n_slots,
                const __grid_constant__ JointProgram prog) {
    constexpr int RW = QO ? 4 : 9;
    constexpr int C = kChunk;
    constexpr bool DENSE = (G == 0);
    static_assert(G % C == 0, "a flush group is a whole number of TMA chunks");

    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // align by OFFSETTING the array (pointer arithmetic keeps the shared address space; an integer round trip
    // would demote every access below to generic LD/ST)
    unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const FkGeom geo = fk_geom(G, VEC, RW, WARPS, n_joints, n_slots);
    const int SR = DENSE ? geo.sr : fk_pad(RW * G, VEC), SP = DENSE ? geo.sp : fk_pad(3 * G, VEC);

    float4 *in_stage = reinterpret_cast<float4 *>(smem_raw + warp * kBoxStages * kBoxBytes);  // kBoxStages boxes
    float4 *tab = reinterpret_cast<float4 *>(smem_raw + WARPS * kBoxStages * kBoxBytes);
    unsigned char *after_tab = reinterpret_cast<unsigned char *>(tab) + ((n_joints * 16 + 127) & ~127);
    float *Rst = reinterpret_cast<float *>(after_tab + warp * geo.warp_bytes);
    float *Pst = Rst + kWarp * SR;
    float4 *slots = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(Rst) + ((kWarp * (SR + SP) * 4 + 15) & ~15));
    uint64_t *bars = reinterpret_cast<uint64_t *>(after_tab + WARPS * geo.warp_bytes);
    const uint32_t bar0 = smem_u32(bars + warp * kBoxStages);
    const uint32_t in0 = smem_u32(in_stage);
    const uint32_t fence_word = smem_u32(reinterpret_cast<uint32_t *>(bars + WARPS * kBoxStages) + threadIdx.x);  // see the chunk loop

    // Persistent warps: all tiles cost the same, so a static round robin balances.
    const long long n_tiles = (n_frames + kWarp - 1) / kWarp;
    const long long tile_stride = static_cast<long long>(gridDim.x) * WARPS;
    long long tile = static_cast<long long>(blockIdx.x) * WARPS + warp;
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kBoxStages; ++b) mbar_init(bar0 + 8 * b, 1);
        fence_barrier_init();
    }
    __syncwarp();
    // Look-ahead cursor of the TMA producer (lane 0): the warp's chunks in processing order, across its tiles.
    long long la_tile = tile;
    int la_c0 = 0;
    auto issue_next = [&](int buf) {  // lane 0 only
        if (la_tile < n_tiles) {
            mbar_arrive_expect_tx(bar0 + 8 * buf, kBoxBytes);
            tma_load_2d(in0 + buf * kBoxBytes, &tm_rot, 4 * la_c0, static_cast<int>(la_tile * kWarp), bar0 + 8 * buf);
            la_c0 += C;
            if (la_c0 >= n_joints) la_c0 = 0, la_tile += tile_stride;
        }
    };
    if (lane == 0) {  // the first kBoxStages chunks: in flight while the block loads its joint table
#pragma unroll
        for (int b = 0; b < kBoxStages; ++b) issue_next(b);
    }
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

    // TMA 128-byte swizzle: 16-byte chunk jj of row r lands at chunk jj ^ (r & 7)
    const int swz = lane & 7;
    const int rpitch = n_joints * RW, ppitch = n_joints * 3;
    uint32_t kchunk = 0;  // chunks consumed so far: buffer = kchunk % kBoxStages, parity = (kchunk / kBoxStages) & 1
    // root position of the warp's NEXT tile, fetched one chunk early like its quaternions
    float gnext[3] = {0.f, 0.f, 0.f};
    if (tile < n_tiles) {
        const float *g = gpos + min(tile * kWarp + lane, n_frames - 1) * gstride;
        gnext[0] = __ldg(g), gnext[1] = __ldg(g + 1), gnext[2] = __ldg(g + 2);
    }

    for (; tile < n_tiles; tile += tile_stride) {
        const long long f0 = tile * kWarp;
        const int nrows = static_cast<int>(min(static_cast<long long>(kWarp), n_frames - f0));
        const long long f = f0 + min(lane, nrows - 1);  // tail lanes recompute the last frame, never stored
        const float *orow = PF_OFFSETS ? offsets + f * ostride : nullptr;
        Xform<float> cur;
        cur.p[0] = gnext[0], cur.p[1] = gnext[1], cur.p[2] = gnext[2];
        int gj = 0;  // joints already staged in the current flush group

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
            // The refill below goes through the async proxy and is not ordered after shared-memory loads that
            // have merely been ISSUED (measured: ~1e-5 of the tiles read the next chunk's data).  A store whose
            // operand depends on every loaded register cannot issue before all of them have landed, and the TMA
            // instruction issues after it.
            {
                uint32_t acc = 0;
#pragma unroll
                for (int jj = 0; jj < C; ++jj) acc |= __float_as_uint(q[jj].x) | __float_as_uint(q[jj].w);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
            }
            __syncwarp();  // every lane has its quaternions in registers: the buffer can be refilled
            const long long next_tile = tile + tile_stride;
            if (lane == 0) issue_next(buf);  // refill with the chunk kBoxStages ahead (this tile's or the next tile's)
            if (last_chunk && next_tile < n_tiles) {
                const float *g = gpos + min(next_tile * kWarp + lane, n_frames - 1) * gstride;
                gnext[0] = __ldg(g), gnext[1] = __ldg(g + 1), gnext[2] = __ldg(g + 2);
            }

            float *rs0 = Rst + lane * SR + RW * gj;
            float *ps0 = Pst + lane * SP + 3 * gj;
            float carry_r = 0.f, carry_p = 0.f;  // VEC == 2: odd word waiting for its 8-byte partner
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                if (jj < cnt) {
                    const int j = c0 + jj;
                    // program word from the kernel-parameter constant bank: j is warp-uniform, so the branches on
                    // it are uniform branches (no divergence bookkeeping); offsets come from the shared table
                    const uint32_t code = PMB_FK_UNIFORM_CODE ? prog.code[j] : __float_as_uint(tab[j].w);
                    const float4 e = tab[j];
                    float ox = e.x, oy = e.y, oz = e.z;
                    if (PF_OFFSETS) {
                        ox = __ldg(orow + 3 * j), oy = __ldg(orow + 3 * j + 1), oz = __ldg(orow + 3 * j + 2);
                    }
                    float l[9];
                    q_to_matrix(PMB_FK_NORMALIZE(Quat<float>{q[jj].x, q[jj].y, q[jj].z, q[jj].w}, 1e-8f), l);
                    if (jj == 0 && c0 == 0) {
#pragma unroll
                        for (int k = 0; k < 9; ++k) cur.r[k] = l[k];  // root: [R | global_pos] (skeleton.py:49)
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
                    float o[RW];
                    if (QO) {
                        const Quat<float> gq = q_from_matrix_fast(cur.r);
                        o[0] = gq.w, o[1] = gq.x, o[2] = gq.y, o[3] = gq.z;
                    } else {
#pragma unroll
                        for (int k = 0; k < RW; ++k) o[k] = cur.r[k];
                    }
                    float *rs = rs0 + RW * jj;
                    float *ps = ps0 + 3 * jj;
                    if (VEC == 1) {
#pragma unroll
                        for (int k = 0; k < RW; ++k) rs[k] = o[k];
                        ps[0] = cur.p[0], ps[1] = cur.p[1], ps[2] = cur.p[2];
                    } else {
                        // 8-byte stores; a chunk starts 8-byte aligned (row strides and 8-joint chunks are even)
                        // and joints alternate parity when RW is odd
                        if ((RW * jj) % 2 == 0) {
#pragma unroll
                            for (int k = 0; k + 1 < RW; k += 2) *reinterpret_cast<float2 *>(rs + k) = make_float2(o[k], o[k + 1]);
                            if (RW % 2) carry_r = o[RW - 1];
                        } else {
                            *reinterpret_cast<float2 *>(rs - 1) = make_float2(carry_r, o[0]);
#pragma unroll
                            for (int k = 1; k + 1 < RW; k += 2) *reinterpret_cast<float2 *>(rs + k) = make_float2(o[k], o[k + 1]);
                        }
                        if ((3 * jj) % 2 == 0) {
                            *reinterpret_cast<float2 *>(ps) = make_float2(cur.p[0], cur.p[1]);
                            carry_p = cur.p[2];
                        } else {
                            *reinterpret_cast<float2 *>(ps - 1) = make_float2(carry_p, cur.p[0]);
                            *reinterpret_cast<float2 *>(ps + 1) = make_float2(cur.p[1], cur.p[2]);
                        }
                    }
                }
            }
            gj += cnt;

            if (DENSE ? last_chunk : (gj == G || last_chunk)) {
                __syncwarp();
                if (DENSE) {
                    if (PMB_FK_BULK_STORE && nrows == kWarp) {
                        // whole tile = two contiguous, 128-byte aligned spans: hand them to the TMA engine and go on
                        fence_proxy_async_smem();  // this lane's stage writes -> visible to the async proxy
                        __syncwarp();
                        if (lane == 0) {
                            bulk_store(rout + f0 * rpitch, smem_u32(Rst), static_cast<uint32_t>(kWarp * rpitch * 4));
                            bulk_store(pos + f0 * ppitch, smem_u32(Pst), static_cast<uint32_t>(kWarp * ppitch * 4));
                            bulk_commit();
                            bulk_wait_read0();  // the stage may be overwritten once the engine has read it
                        }
                    } else {
                        copy_out_flat(Rst, rout + f0 * rpitch, nrows * rpitch, lane);
                        copy_out_flat(Pst, pos + f0 * ppitch, nrows * ppitch, lane);
                    }
                } else {
                    const int g0 = c0 + cnt - gj;  // first joint of the group
                    float *rg = rout + (f0 * n_joints + g0) * RW;
                    float *pg = pos + (f0 * n_joints + g0) * 3;
                    constexpr int GG = DENSE ? C : G;  // (dense never gets here; keeps the templates well-formed)
                    if (gj == GG && nrows == kWarp) {
                        copy_out_periodic<RW * GG, fk_pad(RW * GG, VEC), VEC>(Rst, rg, rpitch, lane);
                        copy_out_periodic<3 * GG, fk_pad(3 * GG, VEC), VEC>(Pst, pg, ppitch, lane);
                    } else {  // VEC == 2 only ever sees even counts (even joint count, even chunks)
                        copy_out_rows<RW * GG, fk_pad(RW * GG, VEC), VEC>(Rst, rg, rpitch, nrows, RW * gj, lane);
                        copy_out_rows<3 * GG, fk_pad(3 * GG, VEC), VEC>(Pst, pg, ppitch, nrows, 3 * gj, lane);
                    }
                }
                __syncwarp();
                gj = 0;
            }
        }
    }
    if (DENSE && PMB_FK_BULK_STORE && lane == 0) bulk_wait0();  // global writes of the last tile are complete at exit
}

}  // namespace pmb
