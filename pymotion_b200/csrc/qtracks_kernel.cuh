// Quaternion-chain walks with the tree spread over the LANES of a warp: to_root_dual_quat (ops/skeleton.py:207-244) and
// fk emitting global quaternions (fk followed by quat.from_matrix, ops/skeleton.py:322-323, :134-140).
//
// The thread-per-frame chain kernels (dq_kernels.cuh, fk_quat_kernel.cuh) keep 32 frames per warp and walk the joints one
// after the other; their shared memory (flush stage + one slot per live branch + boxes, all times 32 frames) leaves
// 4 .. 8 warps per SM at 52 / 65 joints, and ncu shows ~145 instructions per joint and warp for ~70 of arithmetic (slot
// traffic, the copy-out of the stage through registers): 4.8 TB/s (to_root_dual_quat) and 3.6 TB/s (fk_quat) at 4M x 65.
// Here a warp owns a tile of 8 frames and its lanes are (track, frame): the host's level schedule (track_schedule.h, four
// tracks, whole skeleton) puts four independent joints of the same 8 frames into every step.
//   * chain state of a quaternion walk is small (rotation + translation), so a lane per (frame, joint) wastes nothing --
//     unlike the 3x4 transform of fk, which the row kernels split over three lanes;
//   * the stage is the dense image of the tile's OUTPUT rows (padded to an odd multiple of 16 bytes per frame: every
//     16-byte access of a quarter warp hits eight different bank groups) and doubles as the parent store: a joint whose
//     parent was not the same track's previous item reads it back from there -- no slots, no extra traffic;
//   * input: the tile's quaternions (8 rows of 16 J contiguous bytes) arrive as bulk copies into a double buffer, a tile
//     ahead; output: one bulk store per frame row (rows of quaternions / dual quaternions are 16-byte multiples for every
//     joint count), positions as one dense span with head / tail words like the fk track kernel.
// to_root_dual_quat detaches the children of the root (skeleton.py:236-237: "already in root space"): their parent is a
// constant identity record kept one joint past the end of every stage row.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"
#include "track_schedule.h"

namespace pmb {

constexpr int kQtDq = 0, kQtFkQuat = 1, kQtFkQuatRot = 2;  // dual quaternions | global quaternions + positions | quaternions only
constexpr int kQtMirror = 3;  // mirrored LOCAL rotations: the quaternion walk followed, per tile, by the second half of mirror
                              // (ops/skeleton.py:324-331, :410-416): re-index the global quaternions by the joints mapping,
                              // flip two components, back to local space -- without the global quaternions ever leaving
                              // the SM (the two-kernel path moves 64 J bytes per pose for 32 J of input and output)
// joint j of the mirrored result reads the global quaternions of joints (word & 0xFFFF) and, for j > 0, (word >> 16):
// joints_mapping[j] and joints_mapping[parents[j]]; (fx, fy, fz) are the signs of the vector part under the mirror axis
struct QtMirrorTable {
    uint32_t word[PMB_MAX_JOINTS];
    float fx, fy, fz;
};
struct QtNoMirror {};
template <int MODE> struct QtMirrorArg { using type = QtNoMirror; };
template <> struct QtMirrorArg<kQtMirror> { using type = QtMirrorTable; };
// Lane shapes: NT tracks x FQ frames <= 32 lanes.  4 x 8 is the general one; 3 x 10 fills the steps of skeletons whose level
// schedule is no shorter with four tracks than with three (the 22-joint body: 8 steps either way, 22 of 24 slots used instead of
// 22 of 32) -- the walk at 22 joints is ISSUE bound (ncu: 78 % of the issue slots, 'not selected' the top stall), so the
// instructions per frame, steps / FQ, are what counts there.
constexpr int kQtFrames = 8, kQtTracks = 4;

struct QtGeom {
    int in_pitch, in_bytes, q_pitch, q_bytes, p_bytes, o_bytes, tab_bytes, warp_bytes, block_bytes;
};
__host__ __device__ inline QtGeom qt_geom(int mode, int warps, int n_joints, int n_items, int n_tracks = kQtTracks, int tile_frames = kQtFrames) {
    QtGeom g;
    g.in_pitch = 16 * (n_joints | 1);                                        // odd number of 16-byte units per frame row
    g.in_bytes = tile_frames * g.in_pitch;
    g.q_pitch = mode == kQtDq ? 32 * (n_joints + 1) + 16 : 16 * ((n_joints + 1) | 1);  // + the identity record (dq) / padding
    g.q_bytes = tile_frames * g.q_pitch;
    g.p_bytes = mode == kQtFkQuat ? ((tile_frames * 12 * n_joints + 16 + 15) & ~15) : 0;  // dense, + 16 bytes of phase slack
    g.o_bytes = mode == kQtMirror ? g.in_bytes : 0;                          // mirrored local rotations of the tile, rows like the input's
    g.tab_bytes = (((n_items + n_tracks) * 16 + 127) & ~127) + 128;  // + one step of no-ops (prefetch overrun) + the tile counter
    if (mode == kQtMirror) g.tab_bytes += (n_joints * 4 + 127) & ~127;       // + the mirror table
    g.warp_bytes = (2 * g.in_bytes + g.q_bytes + g.p_bytes + g.o_bytes + 16 + 128 + 127) & ~127;  // + 2 mbarriers + fence words
    g.block_bytes = 128 + g.tab_bytes + warps * g.warp_bytes;
    return g;
}

__device__ __forceinline__ void qt_lds128_if(uint32_t flag /* taken iff (int)flag >= 0 */, uint32_t addr, float4 &v) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %4, 0;\n"
        "@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%5];\n"
        "}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "r"(flag), "r"(addr));
}
__device__ __forceinline__ void qt_lds3_if(uint32_t flag, uint32_t addr, float &a, float &b, float &c) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %3, 0;\n"
        "@p ld.shared.f32 %0, [%4];\n"
        "@p ld.shared.f32 %1, [%4+4];\n"
        "@p ld.shared.f32 %2, [%4+8];\n"
        "}"
        : "+f"(a), "+f"(b), "+f"(c)
        : "r"(flag), "r"(addr));
}
__device__ __forceinline__ void qt_sts128_if(uint32_t flag, uint32_t addr, float a, float b, float c, float d) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %0, 0;\n"
        "@p st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n"
        "}" ::"r"(flag), "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void qt_sts3_if(uint32_t flag, uint32_t addr, float a, float b, float c) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ge.s32 p, %0, 0;\n"
        "@p st.shared.f32 [%1], %2;\n"
        "@p st.shared.f32 [%1+4], %3;\n"
        "@p st.shared.f32 [%1+8], %4;\n"
        "}" ::"r"(flag), "r"(addr), "f"(a), "f"(b), "f"(c));
}

// The sign quat.from_matrix gives the quaternion of this rotation (quat.py:111-155 in terms of the unit quaternion).
__device__ __forceinline__ Quat<float> qt_from_matrix_sign(const Quat<float> &q) {
    const float xx = q.x * q.x, yy = q.y * q.y;
    const float pivot = (xx + yy > 0.5f) ? (xx > yy ? q.x : q.y) : (q.z * q.z > q.w * q.w ? q.z : q.w);
    return pivot < 0.f ? Quat<float>{-q.w, -q.x, -q.y, -q.z} : q;
}

template <int MODE, bool PIPE, int NT = kQtTracks, int FQ = kQtFrames>
__global__ void __launch_bounds__(512, 1)
qtracks_kernel(const float4 *__restrict__ rot, const float *__restrict__ gpos, long long gstride, const float *__restrict__ offsets,
               float4 *__restrict__ out_q, float *__restrict__ out_p, long long n_frames, int n_joints, int n_steps,
               int dynamic_claims, const __grid_constant__ TrackProgram prog,
               const __grid_constant__ typename QtMirrorArg<MODE>::type mir) {
    static_assert(NT * FQ <= 32 && NT >= 1 && FQ >= 1, "lanes = tracks x frames");
    constexpr bool POS = MODE == kQtFkQuat;
    constexpr bool MIRROR = MODE == kQtMirror;
    extern __shared__ __align__(128) unsigned char smem_qt[];
    unsigned char *smem_raw = smem_qt + ((128u - (smem_u32(smem_qt) & 127u)) & 127u);
    const int warps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_items = n_steps * NT;
    const QtGeom geo = qt_geom(MODE, warps, n_joints, n_items, NT, FQ);

    // Item table, 16 bytes per item: offset (x, y, z) | word: bits 0-9 joint, 10-19 parent, 30 = parent in the track's
    // registers, 31 + 30 = no-op.  to_root_dual_quat: the children of the root get the identity record (joint index J).
    uint4 *tab = reinterpret_cast<uint4 *>(smem_raw);
    uint32_t *tile_counter = reinterpret_cast<uint32_t *>(smem_raw + geo.tab_bytes - 128);
    if (threadIdx.x == 0) *tile_counter = 0u;
    uint32_t *mtab = reinterpret_cast<uint32_t *>(smem_raw + (((n_items + NT) * 16 + 127) & ~127));  // kQtMirror only
    if constexpr (MIRROR)
        for (int j = threadIdx.x; j < n_joints; j += blockDim.x) mtab[j] = mir.word[j];
    for (int i = threadIdx.x; i < n_items + NT; i += blockDim.x) {
        const uint32_t c = i < n_items ? prog.code[i] : kTrackNoop;
        const uint32_t j = track_joint(c);
        uint32_t p = track_parent(c);
        uint4 e = make_uint4(0u, 0u, 0u, 0xC0000000u);
        if (!(c & kTrackNoop)) {
            if (j > 0) e.x = __float_as_uint(offsets[3 * j]), e.y = __float_as_uint(offsets[3 * j + 1]), e.z = __float_as_uint(offsets[3 * j + 2]);
            uint32_t carry = (c & kTrackCarry) ? 0x40000000u : 0u;
            if (MODE == kQtDq && j > 0 && p == 0) p = static_cast<uint32_t>(n_joints), carry = 0u;
            e.w = j | (p << 10) | carry;
        }
        tab[i] = e;
    }
    unsigned char *mine = smem_raw + geo.tab_bytes + warp * geo.warp_bytes;
    const uint32_t in0 = smem_u32(mine);
    const uint32_t qst = in0 + 2 * geo.in_bytes;
    const uint32_t pst = qst + geo.q_bytes;
    const uint32_t ost = pst + geo.p_bytes;   // kQtMirror: the tile's output rows
    const uint32_t bar0 = ost + geo.o_bytes;  // two mbarriers
    const uint32_t fence_word = bar0 + 16 + 4 * lane;
    const uint32_t tab0 = smem_u32(tab);
    if (lane == 0) {
        mbar_init(bar0, 1), mbar_init(bar0 + 8, 1);
        fence_barrier_init();
    }
    // lane -> (track, frame); lanes past NT * FQ shadow lane 0 (same addresses: a broadcast) and never store
    const bool idle = lane >= NT * FQ;
    const int trk = idle ? 0 : lane / FQ, f = idle ? 0 : lane - (lane / FQ) * FQ;
    const uint32_t idle_mask = idle ? 0x80000000u : 0u;
    if (MODE == kQtDq && lane < FQ) {  // the identity record of this frame's stage row: rotation (1, 0, 0, 0), dual part 0
        float4 *rec = reinterpret_cast<float4 *>(mine + 2 * geo.in_bytes + lane * geo.q_pitch + 32 * n_joints);
        rec[0] = make_float4(1.f, 0.f, 0.f, 0.f), rec[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();  // table, barriers, identity records; from here on the warps never meet again

    // Tiles: block b owns tiles b, b + grid, b + 2 grid, ...; its warps CLAIM them from a shared-memory counter.  (With a
    // static round robin the warps of an SM whose count is not a multiple of four -- 6 or 7 at 65 joints, 9 at 52 -- finish
    // at the pace of the sub-partition that holds one warp more: measured 1.67 ms with 9 warps against 1.52 ms with 8 at
    // 4M x 52.)  A warp holds two claims: the tile it walks and the one whose quaternions are on their way.
    const long long n_tiles = (n_frames + FQ - 1) / FQ;
    uint32_t static_n = warp;
    auto claim = [&]() -> long long {
        uint32_t n = 0;
        if (dynamic_claims) {
            if (lane == 0) n = atomicAdd(tile_counter, 1u);
            n = __shfl_sync(0xffffffffu, n, 0);
            return static_cast<long long>(n) * gridDim.x + blockIdx.x;
        }
        n = static_n, static_n += warps;
        return (static_cast<long long>(n / warps) * gridDim.x + blockIdx.x) * warps + (n % warps);
    };
    long long tile = claim(), tile_next = claim();
    if (tile >= n_tiles) return;
    const int row_bytes = 16 * n_joints;
    const uint32_t in_row0 = in0 + f * geo.in_pitch, q_row = qst + f * geo.q_pitch;
    const int ppitch = 12 * n_joints;

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
    bool draining = false;

    for (; tile < n_tiles; ++k) {
        const long long f0 = tile * FQ;
        const int nrows = static_cast<int>(min(static_cast<long long>(FQ), n_frames - f0));
        const int buf = k & 1;
        const uint32_t in_row = in_row0 + buf * geo.in_bytes;
        const uint32_t pphase = POS ? static_cast<uint32_t>((f0 * ppitch) & 15) : 0u;
        const uint32_t p_row = pst + pphase + f * ppitch;
        // track 0 starts from the "parent" of the root: the identity placed at global_pos
        Quat<float> R{1.f, 0.f, 0.f, 0.f};
        float4 D = make_float4(0.f, 0.f, 0.f, 0.f);
        float t0 = gn0, t1 = gn1, t2 = gn2;
        if (tile_next < n_tiles) {
            const float *g = gpos + min(tile_next * FQ + f, n_frames - 1) * gstride;
            gn0 = __ldg(g), gn1 = __ldg(g + 1), gn2 = __ldg(g + 2);
        }
        mbar_wait(bar0 + 8 * buf, (k >> 1) & 1);
        if (!MIRROR && draining) bulk_wait_read0();  // (lanes that stored) the previous tile has left the stage
        __syncwarp();

        // The walk, software pipelined: the table entry and the local quaternion of step s + 1 (for fk_quat already
        // normalised) are fetched while step s computes, so what is left on the step-to-step critical path is parent
        // read -> products -> store -> __syncwarp.  The table ends with one step of no-ops for the last prefetch.
        uint32_t acc = 0;
        uint32_t tab_addr = tab0 + trk * 16;
        float4 e = lds128_ro(tab_addr);
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PIPE) {
            qv = lds128(in_row + 16 * (__float_as_uint(e.w) & 0x3FFu));
            acc |= __float_as_uint(qv.x);
            if (MODE != kQtDq) {
                Quat<float> r = q_normalize_fast(Quat<float>{qv.x, qv.y, qv.z, qv.w}, 1e-8f);
                qv = make_float4(r.w, r.x, r.y, r.z);
            }
        }
        for (int step = 0; step < n_steps; ++step) {
            tab_addr += NT * 16;
            float4 e_next = e, qv_next = qv;
            if (!PIPE) {
                qv = lds128(in_row + 16 * (__float_as_uint(e.w) & 0x3FFu));
                acc |= __float_as_uint(qv.x);
                if (MODE != kQtDq) {
                    Quat<float> r = q_normalize_fast(Quat<float>{qv.x, qv.y, qv.z, qv.w}, 1e-8f);
                    qv = make_float4(r.w, r.x, r.y, r.z);
                }
            }
            // Issued AFTER the parent reads of this step (shared-memory instructions keep their order): the address of the
            // quaternion depends on the table entry, and an in-order warp would otherwise sit on that dependency with the
            // parent reads -- the critical path -- queued behind it.
            auto prefetch = [&]() {
                if (PIPE) {
                    e_next = lds128_ro(tab_addr);
                    qv_next = lds128(in_row + 16 * (__float_as_uint(e_next.w) & 0x3FFu));
                    acc |= __float_as_uint(qv_next.x);
                    if (MODE != kQtDq) {
                        // the reference turns q / (|q| + eps) into a MATRIX: the zero quaternion becomes the identity (below)
                        Quat<float> r = q_normalize_fast(Quat<float>{qv_next.x, qv_next.y, qv_next.z, qv_next.w}, 1e-8f);
                        qv_next = make_float4(r.w, r.x, r.y, r.z);
                    }
                }
            };
            const uint32_t w = __float_as_uint(e.w);
            const uint32_t j = w & 0x3FFu, p = (w >> 10) & 0x3FFu;
            if (MODE == kQtDq) {
                // parent (rotation, dual part): registers, or the stage (a joint stored earlier, or the identity record)
                float4 a = make_float4(R.w, R.x, R.y, R.z), b = D;
                qt_lds128_if(w << 1, q_row + 32 * p, a);
                qt_lds128_if(w << 1, q_row + 32 * p + 16, b);
                prefetch();
                {
                    // translation back from the dual part: t = 2 (d (x) conj(r)) / |r|^2 (dual_quat.py:62-83; rotations are
                    // not normalised on this path).  Branch free: the tracks of a warp disagree about where the parent is.
                    const float n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
                    const float s = __fdividef(2.f, n2);
                    const float u0 = s * (-b.x * a.y + b.y * a.x - b.z * a.w + b.w * a.z);
                    const float u1 = s * (-b.x * a.z + b.y * a.w + b.z * a.x - b.w * a.y);
                    const float u2 = s * (-b.x * a.w - b.y * a.z + b.z * a.y + b.w * a.x);
                    const bool from_stage = static_cast<int>(w << 1) >= 0;
                    t0 = from_stage ? u0 : t0, t1 = from_stage ? u1 : t1, t2 = from_stage ? u2 : t2;
                }
                const Quat<float> Rp{a.x, a.y, a.z, a.w};
                const Vec3<float> v = q_rotate(Rp, Vec3<float>{e.x, e.y, e.z});   // skeleton.py:238-240
                t0 += v.x, t1 += v.y, t2 += v.z;
                R = q_mul(Rp, Quat<float>{qv.x, qv.y, qv.z, qv.w});                // :241
                const Quat<float> d = q_mul(Quat<float>{0.f, t0, t1, t2}, R);      // dual_quat.py:28-35
                D = make_float4(0.5f * d.w, 0.5f * d.x, 0.5f * d.y, 0.5f * d.z);
                qt_sts128_if(w | idle_mask, q_row + 32 * j, R.w, R.x, R.y, R.z);
                qt_sts128_if(w | idle_mask, q_row + 32 * j + 16, D.x, D.y, D.z, D.w);
            } else {
                float4 a = make_float4(R.w, R.x, R.y, R.z);
                qt_lds128_if(w << 1, q_row + 16 * p, a);
                if (POS) qt_lds3_if(w << 1, p_row + 12 * p, t0, t1, t2);
                prefetch();
                Quat<float> r{qv.x, qv.y, qv.z, qv.w};
                if (r.w == 0.f && r.x == 0.f && r.y == 0.f && r.z == 0.f) r.w = 1.f;
                const Quat<float> Qp{a.x, a.y, a.z, a.w};
                if (POS) {
                    const Vec3<float> v = q_rotate(Qp, Vec3<float>{e.x, e.y, e.z});
                    t0 += v.x, t1 += v.y, t2 += v.z;
                }
                R = qt_from_matrix_sign(q_mul(Qp, r));
                qt_sts128_if(w | idle_mask, q_row + 16 * j, R.w, R.x, R.y, R.z);
                if (POS) qt_sts3_if(w | idle_mask, p_row + 12 * j, t0, t1, t2);
            }
            __syncwarp();  // a parent may have been stored by another track
            if (PIPE) e = e_next, qv = qv_next;
            else e = lds128_ro(tab_addr);
        }
        // the tile's quaternions have been READ (a store that depends on all of them precedes the refill through the async
        // proxy, see fk_kernel.cuh); fetch the tile after the next one into this buffer
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(fence_word), "r"(acc) : "memory");
        __syncwarp();
        tile = tile_next;
        tile_next = claim();
        issue_tile(tile_next, buf);

        if constexpr (MIRROR) {
            // ---- second half of mirror on the tile's global quaternions (all in the stage): track t takes joints
            // t, t + 4, ...; results go to their own rows (the stage is still being read by the other tracks)
            if (draining) bulk_wait_read0();  // (lanes that stored) the previous tile has left the output rows
            __syncwarp();
            const uint32_t o_row = ost + f * geo.in_pitch;
            for (int j = idle ? n_joints : trk; j < n_joints; j += NT) {
                const uint32_t m = mtab[j];
                const float4 a = lds128(q_row + 16 * (m & 0xFFFFu));
                Quat<float> r{a.x, mir.fx * a.y, mir.fy * a.z, mir.fz * a.w};
                if (j > 0) {
                    const float4 b = lds128(q_row + 16 * (m >> 16));
                    r = q_mul(Quat<float>{b.x, -mir.fx * b.y, -mir.fy * b.z, -mir.fz * b.w}, r);  // conj(flip(parent)) (x) flip(joint)
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(o_row + 16 * j), "f"(r.w), "f"(r.x), "f"(r.y), "f"(r.z) : "memory");
            }
        }
        // ---- output: one bulk store per frame row; positions as one dense span ---------------------------------
        fence_proxy_async_smem();
        __syncwarp();
        if (lane < nrows) {
            const int out_row = MODE == kQtDq ? 2 * row_bytes : row_bytes;
            const uint32_t src_row = MIRROR ? ost + lane * geo.in_pitch : qst + lane * geo.q_pitch;
            bulk_store(reinterpret_cast<unsigned char *>(out_q) + (f0 + lane) * out_row, src_row, static_cast<uint32_t>(out_row));
        }
        if (POS) {
            const long long pa = f0 * ppitch, pb = pa + static_cast<long long>(nrows) * ppitch;
            const long long pa16 = (pa + 15) & ~15LL, pb16 = pb & ~15LL;
            unsigned char *pg = reinterpret_cast<unsigned char *>(out_p);
            if (lane == 0 && pb16 > pa16)
                bulk_store(pg + pa16, pst + pphase + static_cast<uint32_t>(pa16 - pa), static_cast<uint32_t>(pb16 - pa16));
            if (((pa | pb) & 15) != 0) {  // head / tail words around the 16-byte aligned middle
                const long long h_end = pb16 > pa16 ? pa16 : pb;
                const int ph = static_cast<int>(h_end - pa) >> 2, pt = pb16 > pa16 ? static_cast<int>(pb - pb16) >> 2 : 0;
                float v;
                if (lane >= 8 && lane - 8 < ph) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pst + pphase + 4 * (lane - 8)));
                    reinterpret_cast<float *>(pg + pa)[lane - 8] = v;
                } else if (lane >= 24 && lane - 24 < pt) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pst + pphase + static_cast<uint32_t>(pb16 - pa) + 4 * (lane - 24)));
                    reinterpret_cast<float *>(pg + pb16)[lane - 24] = v;
                }
            }
        }
        if (lane < FQ) {
            bulk_commit();
            draining = true;
        }
        __syncwarp();  // the stage is rewritten only after the head / tail reads
    }
    if (draining) bulk_wait0();  // global writes of the last tile are complete at exit
}

}  // namespace pmb
