// libpymotion_b200.so, fk translation unit: pmb_fk_f32 / pmb_fk_quat_f32 and the launch policy that picks one of
// the fk kernels (track, row-team, lane, thread-per-frame) from the joint count and the topology.
// Host side only validates, looks up the per-topology program, picks a launch configuration and launches.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "fk_kernel.cuh"
#include "fk_mtracks_kernel.cuh"
#include "fk_quat_kernel.cuh"
#include "fk_rows_kernel.cuh"
#include "fk_tracks_kernel.cuh"
#include "host_common.h"
#include "qtracks_host.h"

using namespace pmbh;

namespace {

struct FkArgs {
    const float *rot, *gpos, *offsets;
    long long gstride, ostride;
    float *pos, *rout;
    long long n_frames;
    int n_joints, n_slots;
    const pmb::JointProgram *prog;
    const int64_t *parents_host;
    cudaStream_t stream;
};

// ---- thread-per-frame kernel (fk_kernel.cuh) ---------------------------------------------------
template <int G, int WARPS, int VEC, bool PF, bool QO>
int launch_fk_cfg(const FkArgs &a, const DeviceProps &dp) {
    auto kernel = pmb::fk_chain_kernel<G, WARPS, VEC, PF, QO>;
    const int smem = pmb::fk_geom(G, VEC, QO ? 4 : 9, WARPS, a.n_joints, a.n_slots).block_bytes;
    int per_sm = 0, rc = kernel_fit(kernel, dp, WARPS * 32, smem, per_sm);
    if (rc) return rc;
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk kernel does not fit on an SM (%d bytes of shared memory)", smem);
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk))) return rc;
    // persistent grid: as many blocks as are resident at once; warps walk the tiles round robin
    const long long tiles = (a.n_frames + 31) / 32;
    per_sm = std::max(1, std::min(per_sm, knob(K_FK_BLOCKS_PER_SM, per_sm)));
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_chain_kernel<G=%d,WARPS=%d,VEC=%d,PF=%d,QO=%d> grid=%lld smem=%d", G, WARPS, VEC, int(PF), int(QO), blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.ostride, a.pos, a.rout,
                                                                        a.n_frames, a.n_joints, a.n_slots, *a.prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- row-team kernel (fk_rows_kernel.cuh) ------------------------------------------------------
template <int S, int VEC>
int launch_fk_rows_cfg(const FkArgs &a, const DeviceProps &dp, int team_cap) {
    auto kernel = pmb::fk_rows_kernel<S, VEC>;
    const int smem = pmb::fk_rows_geom(S, a.n_joints).block_bytes;
    int per_sm = 0, rc = kernel_fit(kernel, dp, pmb::kRowThreads, smem, per_sm);
    if (rc) return rc;
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk row kernel does not fit on an SM (%d bytes of shared memory)", smem);
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk))) return rc;
    const long long tiles = (a.n_frames + 31) / 32;
    if (team_cap > 0) per_sm = std::min(per_sm, team_cap);
    per_sm = std::max(1, std::min(per_sm, knob(K_FK_BLOCKS_PER_SM, per_sm)));
    const long long blocks = std::min<long long>(tiles, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_rows_kernel<S=%d,VEC=%d> grid=%lld (%d teams/SM) smem=%d", S, VEC, blocks, per_sm, smem);
    kernel<<<static_cast<unsigned>(blocks), pmb::kRowThreads, smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.pos, a.rout,
                                                                                a.n_frames, a.n_joints, *a.prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// Teams (blocks) of the row kernel that fit on an SM with S box stages: every block also costs 1 KB of
// system shared memory.
inline int fk_rows_teams(int stages, const FkArgs &a, const DeviceProps &dp) {
    const int bytes = pmb::fk_rows_geom(stages, a.n_joints).block_bytes;
    if (bytes > dp.smem_optin) return 0;
    return std::min(12, (dp.smem_optin + 1024) / (bytes + 1024));
}

// ---- track kernel (fk_tracks_kernel.cuh) -------------------------------------------------------
struct FkTracksShape {
    int fr = 0, warps = 0, blocks = 0, smem = 0;  // frames per tile, warps per block, blocks per SM
};
// The block shape that puts the most warps (= frames in flight) on an SM: every block carries its own copy of the
// schedule table and costs 1 KB of system shared memory.
inline FkTracksShape fk_tracks_shape(int fr, int n_joints, int n_items, int n_boxes, int warps_cap, const DeviceProps &dp) {
    FkTracksShape best;
    for (int blocks = 1; blocks <= 4; ++blocks)
        for (int warps = 16; warps >= 1; --warps) {
            const int smem = pmb::fk_tracks_geom(fr, warps, n_joints, n_items, n_boxes).block_bytes;
            if (smem > dp.smem_optin || blocks * (smem + 1024) > dp.smem_sm) continue;
            if (blocks * warps > warps_cap) continue;
            if (blocks * warps > best.blocks * best.warps) best = {fr, warps, blocks, smem};
            break;  // fewer warps per block only lowers the product for this block count
        }
    return best;
}

template <int U, int NB, int UL>
int launch_fk_tracks_cfg(const FkArgs &a, const DeviceProps &dp, const FkTracksShape &sh, const pmb::TrackProgram &tp, int n_steps) {
    auto kernel = pmb::fk_tracks_kernel<U, NB, UL>;
    int per_sm = 0, rc = kernel_fit(kernel, dp, sh.warps * 32, sh.smem, per_sm);
    if (rc) return rc;
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk track kernel does not fit on an SM (%d bytes of shared memory)", sh.smem);
    per_sm = std::min(per_sm, sh.blocks);
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk, sh.fr))) return rc;
    const long long tiles = (a.n_frames + sh.fr - 1) / sh.fr;
    const long long blocks = std::min<long long>((tiles + sh.warps - 1) / sh.warps, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_tracks_kernel<U=%d,NB=%d,UL=%d> FR=%d steps=%d grid=%lld x %d warps (%d warps/SM) smem=%d", U, NB, UL, sh.fr,
                 n_steps, blocks, sh.warps, per_sm * sh.warps, sh.smem);
    kernel<<<static_cast<unsigned>(blocks), sh.warps * 32, sh.smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.pos, a.rout, a.n_frames,
                                                                              a.n_joints, n_steps, sh.fr,
                                                                              knob(K_FK_L2_PREFETCH, 0) ? a.rot : nullptr, tp);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// Worst-case number of lanes of one lane group whose stage stores fall into the same shared-memory bank: the frames of a
// tile are 9J words apart (only 9 J mod 32 matters; 1 = conflict free).
inline int fk_tracks_bank_degree(int fr, int n_joints) {
    int count[32] = {0}, worst = 0;
    for (int f = 0; f < fr; ++f) worst = std::max(worst, ++count[(f * 9 * n_joints) & 31]);
    return worst;
}

// Default policy (measured on B200, profiles/r2_sweep_tracks_*.jsonl): skeletons too large for four row teams per SM take
// the track kernel with two lane groups (tiles of 5 frames, one track per lane) --
//     2M x 40: 0.880 ms against 0.907 for the lane kernel of round 1;  4M x 52: 2.331 against 2.504;  4M x 65: 3.090 against 3.442
// -- unless the frames of a group collide in 3 or more banks (J = 32, 48, 64, ...), which stay with the padded stage of
// the thread-per-frame kernel.  Four boxes in flight from 48 joints up (52: 2.360 -> 2.331, 65: 3.220 -> 3.090 against
// three), three below (40: 0.880 against 0.892).
// PMB_FK_TRACKS = 0 / 1 forces; PMB_FK_U = 1 | 2 tracks per lane, PMB_FK_UL = 1 | 2 lane groups, PMB_FK_NB = 2 .. 6 boxes in
// flight, PMB_FK_FR frames per tile, PMB_FK_WARPS_PER_SM caps the resident warps.
bool try_fk_tracks(const FkArgs &a, const DeviceProps &dp, int &rc, bool rows_preferred) {
    const int force = knob(K_FK_TRACKS, -1);
    if (force == 0) return false;
    if (force != 1) {
        if (knob_set(K_FK_GROUP) || knob_set(K_FK_ROWS) || knob_set(K_FK_WARPS)) return false;  // another kernel is forced
        if (rows_preferred || a.n_joints <= 30) return false;
        if (fk_tracks_bank_degree(5, a.n_joints) >= 3) return false;
    }
    const int U = knob(K_FK_U, force == 1 ? 2 : 1), UL = knob(K_FK_UL, force == 1 ? 1 : 2);
    const int NB = knob(K_FK_NB, force == 1 ? 3 : (a.n_joints >= 48 ? 4 : 3));
    if (UL != 1 && UL != 2) { rc = fail(PMB_ERR_SHAPE, "PMB_FK_UL must be 1 or 2"); return true; }
    const pmb::TrackProgram *tp = nullptr;
    int n_steps = 0;
    if ((rc = track_program(a.parents_host, a.n_joints, U * UL, pmb::kChunk, tp, n_steps))) return true;
    if (n_steps == 0) { rc = fail(PMB_ERR_SHAPE, "track schedule with %d tracks does not fit", U * UL); return true; }
    const int fr = knob(K_FK_FR, UL == 2 ? 5 : 10);
    if (fr < 1 || 3 * fr * UL > 32) { rc = fail(PMB_ERR_SHAPE, "PMB_FK_FR=%d does not fit %d lane group(s)", fr, UL); return true; }
    const FkTracksShape sh = fk_tracks_shape(fr, a.n_joints, n_steps * U * UL, NB, knob(K_FK_WARPS_PER_SM, 16), dp);
    if (sh.fr == 0) {
        if (force != 1) return false;  // too many joints for a tile of 5 frames: the older kernels decide
        rc = fail(PMB_ERR_SHAPE, "fk track kernel: %d joints do not fit in shared memory", a.n_joints);
        return true;
    }
#define PMB_TRACKS_CASE(u, nb, ul) \
    if (U == u && NB == nb && UL == ul) { rc = launch_fk_tracks_cfg<u, nb, ul>(a, dp, sh, *tp, n_steps); return true; }
    PMB_TRACKS_CASE(1, 2, 1) PMB_TRACKS_CASE(1, 3, 1) PMB_TRACKS_CASE(2, 2, 1) PMB_TRACKS_CASE(2, 3, 1) PMB_TRACKS_CASE(2, 4, 1)
    PMB_TRACKS_CASE(1, 2, 2) PMB_TRACKS_CASE(1, 3, 2) PMB_TRACKS_CASE(1, 4, 2) PMB_TRACKS_CASE(1, 5, 2) PMB_TRACKS_CASE(1, 6, 2)
    PMB_TRACKS_CASE(2, 4, 2)
#undef PMB_TRACKS_CASE
    rc = fail(PMB_ERR_SHAPE, "PMB_FK_U=%d / PMB_FK_NB=%d / PMB_FK_UL=%d select no available variant", U, NB, UL);
    return true;
}

// ---- matrix track kernel (fk_mtracks_kernel.cuh) ------------------------------------------------
// Worst-case number of the 8 frames of a tile whose stage rows fall into the same shared-memory bank (rows 9 J words apart).
inline int fk_mtracks_bank_degree(int n_joints) {
    int count[32] = {0}, worst = 0;
    for (int f = 0; f < pmb::kMtFrames; ++f) worst = std::max(worst, ++count[(f * 9 * n_joints) & 31]);
    return worst;
}
// PMB_FK_MTRACKS = 0 / 1 forces; PMB_FK_WARPS_PER_SM caps the resident warps.
bool try_fk_mtracks(const FkArgs &a, const DeviceProps &dp, int &rc) {
    const int force = knob(K_FK_MTRACKS, -1);
    if (force == 0) return false;
    if (force != 1 && (knob_set(K_FK_GROUP) || knob_set(K_FK_ROWS) || knob_set(K_FK_WARPS) || knob_set(K_FK_TRACKS)))
        return false;  // another kernel is being forced
    const pmb::TrackProgram *tp = nullptr;
    int n_steps = 0;
    if ((rc = track_program(a.parents_host, a.n_joints, pmb::kMtTracks, 0, tp, n_steps))) return true;
    const int n_items = n_steps * pmb::kMtTracks;
    // the block shape that puts the most warps on an SM (every block carries its own table and costs 1 KB of system shared memory)
    int best_warps = 0, best_blocks = 0, best_smem = 0;
    const int cap = knob(K_FK_WARPS_PER_SM, 32);
    if (n_steps > 0)
        for (int blocks = 1; blocks <= 4; ++blocks)
            for (int warps = 16; warps >= 1; --warps) {
                const int smem = pmb::mt_geom(warps, a.n_joints, n_items).block_bytes;
                if (smem > dp.smem_optin || blocks * (smem + 1024) > dp.smem_sm || blocks * warps > cap) continue;
                if (blocks * warps > best_blocks * best_warps) best_warps = warps, best_blocks = blocks, best_smem = smem;
                break;
            }
    if (best_warps == 0) {
        if (force != 1) return false;
        rc = fail(PMB_ERR_SHAPE, "PMB_FK_MTRACKS=1: %d joints do not fit the matrix track kernel", a.n_joints);
        return true;
    }
    if (force != 1) {
        // a schedule less than half full (chains), fewer than four warps per SM (large skeletons), or stage rows that collide
        // 3-way and worse in the banks (J = 16, 48, 64, ...) belong to the other kernels
        if (2 * a.n_joints < n_items || best_warps * best_blocks < 4 || fk_mtracks_bank_degree(a.n_joints) >= 3) return false;
        // Between 31 and 59 joints the two track kernels are within +-5 % of each other, the order depending on the GPU's power state
        // (profiles/r2_placement_probe.jsonl: not on where the arrays sit) (one process running the workloads in sequence: matrix tracks 0.82 / 2.18 ms at
        // 2M x 40 / 4M x 52 against 0.94 / 2.48 for the row tracks; each workload in a process of its own, as in bench.py: 0.90 /
        // 2.37 against 0.88 / 2.30).  The row tracks keep that range; from 60 joints up the matrix tracks are ahead everywhere
        // (4M x 65: 2.69 - 2.94 ms against 2.97 - 3.23), and below 31 joints they replace the lane kernel (2M x 24: 0.550 against 0.599).
        if (a.n_joints > 30 && a.n_joints < 60) return false;
    }
    auto kernel = pmb::fk_mtracks_kernel;
    int per_sm = 0;
    if ((rc = kernel_fit(kernel, dp, best_warps * 32, best_smem, per_sm))) return true;
    if (per_sm < 1) { rc = fail(PMB_ERR_CUDA, "fk matrix track kernel does not fit on an SM (%d bytes of shared memory)", best_smem); return true; }
    per_sm = std::min(per_sm, best_blocks);
    const long long tiles = (a.n_frames + pmb::kMtFrames - 1) / pmb::kMtFrames;
    const long long blocks = std::min<long long>((tiles + best_warps - 1) / best_warps, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_mtracks_kernel steps=%d grid=%lld x %d warps (%d warps/SM) smem=%d", n_steps, blocks, best_warps, per_sm * best_warps, best_smem);
    kernel<<<static_cast<unsigned>(blocks), best_warps * 32, best_smem, a.stream>>>(reinterpret_cast<const float4 *>(a.rot), a.gpos, a.gstride,
                                                                                   a.offsets, a.pos, a.rout, a.n_frames, a.n_joints, n_steps, *tp);
    rc = PMB_OK;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "fk_mtracks_kernel launch");
    return true;
}

// Which fk kernel (shared offsets, matrices out) -- measured on B200, DESIGN.md section 4:
//   row-team kernel      skeletons small enough for >= 4 teams per SM (J <= 30) with J not a multiple of 4
//                        (1M x 22: 0.2304 ms against 0.2565 ms matrix tracks, 0.2415 ms thread-per-frame, 0.2326 ms lanes);
//   matrix track kernel  60 joints and more, and small skeletons the row teams do not take (J a multiple of 4), provided the level
//                        schedule is at least half full, >= 4 warps fit an SM and the stage rows collide at most 2-way in the
//                        banks (try_fk_mtracks above: 4M x 65 2.69 ms against 2.99 row tracks, 2M x 24 0.550 against 0.599 lanes);
//   track kernel         31 .. 59 joints and what the matrix track kernel does not take (try_fk_tracks above);
//   thread-per-frame     what is left: joint counts whose dense stage rows collide 3-way and worse in the banks (J = 16, 32, 48,
//                        64, ...: its padded stage has no conflicts, 5.6 TB/s at J = 32), sparse level schedules (chains),
//                        per-frame offsets, or a forced variant.
// (The lane = (frame, row) kernel of round 1 is retired: experiments/retired/fk_lanes_kernel.cuh.)
// Bank conflicts of the row-team kernel (lane = frame, 32 lanes 9J words apart): J odd: none; J = 2 (mod 4):
// none with its 64-bit stores; J = 0 (mod 4): 4-way (J = 52: 4.5 TB/s) up to 32-way (J = 32: 0.84 TB/s).
bool fk_rows_preferred(const FkArgs &a, const DeviceProps &dp) {
    return fk_rows_teams(2, a, dp) >= 4 && a.n_joints % 4 != 0;
}

bool try_fk_rows(const FkArgs &a, const DeviceProps &dp, int &rc) {
    const int force = knob(K_FK_ROWS, -1);
    if (force == 0) return false;
    if (force != 1 && (knob_set(K_FK_GROUP) || knob_set(K_FK_WARPS))) return false;  // a chain-kernel variant is being forced
    int stages = knob(K_FK_STAGES, -1);
    int team_cap = 0;  // 0: as many as fit
    if (stages < 0) {
        const int t2 = fk_rows_teams(2, a, dp);
        if (t2 >= 4) {
            stages = 2, team_cap = 4;
        } else {
            stages = 2;
            for (int s = 3; s <= 4; ++s)
                if (fk_rows_teams(s, a, dp) == t2) stages = s;  // the deepest ring that does not cost a team
        }
    }
    const int teams = (stages >= 2 && stages <= 4) ? fk_rows_teams(stages, a, dp) : 0;
    if (teams < 1) {
        if (force == 1) { rc = fail(PMB_ERR_SHAPE, "PMB_FK_ROWS=1: the row kernel does not fit (stages %d)", stages); return true; }
        return false;
    }
    if (force != 1 && !fk_rows_preferred(a, dp)) return false;
    if (a.n_joints % 2 == 0)
        rc = stages == 2 ? launch_fk_rows_cfg<2, 2>(a, dp, team_cap) : stages == 3 ? launch_fk_rows_cfg<3, 2>(a, dp, team_cap) : launch_fk_rows_cfg<4, 2>(a, dp, team_cap);
    else
        rc = stages == 2 ? launch_fk_rows_cfg<2, 1>(a, dp, team_cap) : stages == 3 ? launch_fk_rows_cfg<3, 1>(a, dp, team_cap) : launch_fk_rows_cfg<4, 1>(a, dp, team_cap);
    return true;
}

// ---- fk_quat, quaternion-chain kernel (fk_quat_kernel.cuh) ------------------------------------
int launch_fk_quat_chain(const FkArgs &a, const DeviceProps &dp) {
    constexpr int WARPS = 4;
    // like to_root_dual_quat: the largest flush group that still lets TWO 4-warp blocks share an SM
    const int budget = (dp.smem_optin - 2048) / 2;
    int group = a.n_joints;
    if (pmb::fkq_geom(group, WARPS, a.n_joints, a.n_slots).block_bytes > budget) {
        group = ((a.n_joints + 7) / 8) * 8;
        while (group > 8 && pmb::fkq_geom(group, WARPS, a.n_joints, a.n_slots).block_bytes > budget) group -= 8;
    }
    // ... unless that leaves flushes of 8 joints (deep orderings with many live slots): measured at 4M x 65,
    // one block per SM flushing 24 joints at a time beats two blocks flushing 8 (3.16 ms vs 3.87 ms)
    if (group < 16 && a.n_joints > 16) {
        int g1 = 24;
        while (g1 > 8 && pmb::fkq_geom(g1, WARPS, a.n_joints, a.n_slots).block_bytes > dp.smem_optin) g1 -= 8;
        if (g1 > group) group = g1;
    }
    if (knob_set(K_FKQ_GROUP)) {
        const int v = knob(K_FKQ_GROUP, 0);
        if (v >= a.n_joints) group = a.n_joints;  // whole rows
        else if (v >= 8 && v % 8 == 0) group = v;
    }
    const int smem = pmb::fkq_geom(group, WARPS, a.n_joints, a.n_slots).block_bytes;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", a.n_slots);
    auto kernel = a.pos ? pmb::fk_quat_chain_kernel<WARPS, true> : pmb::fk_quat_chain_kernel<WARPS, false>;
    int per_sm = 0, rc = kernel_fit(kernel, dp, WARPS * 32, smem, per_sm);
    if (rc) return rc;
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk_quat kernel does not fit on an SM");
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk))) return rc;
    per_sm = std::max(1, std::min(per_sm, knob(K_FKQ_BLOCKS_PER_SM, per_sm)));
    const long long tiles = (a.n_frames + 31) / 32;
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    auto magic_of = [](int d) { return static_cast<uint32_t>((1ULL << 32) / static_cast<uint32_t>(d)) + 1u; };
    const int tail = a.n_joints % group ? a.n_joints % group : group;
    note_variant("fk_quat_chain_kernel<WARPS=%d,POS=%d> group=%d grid=%lld smem=%d", WARPS, a.pos ? 1 : 0, group, blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, a.stream>>>(
        tm, a.gpos, a.gstride, a.offsets, a.pos, reinterpret_cast<float4 *>(a.rout), a.n_frames, a.n_joints, a.n_slots, group,
        magic_of(group), magic_of(tail), magic_of(3 * group), magic_of(3 * tail), *a.prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

inline bool fk_fits(int group, int vec, int rw, int warps, const FkArgs &a, const DeviceProps &dp) {
    return pmb::fk_geom(group, vec, rw, warps, a.n_joints, a.n_slots).block_bytes <= dp.smem_optin;
}

// Output flush group (see fk_kernel.cuh): whole rows if a block of >= 4 warps fits, else the largest group
// that does.  FULL = the main entry point gets every variant; the rarer ones (per-frame offsets, quaternion
// output) get the two extremes only, to keep the build small.
template <int VEC, bool PF, bool QO, bool FULL>
int launch_fk_group(const FkArgs &a, const DeviceProps &dp) {
    constexpr int RW = QO ? 4 : 9;
    const int force_group = knob(K_FK_GROUP, -1);
    const int force_warps = knob(K_FK_WARPS, -1);
    auto want = [&](int group, int warps) {
        if (force_group >= 0 && force_group != group) return false;
        if (force_warps >= 0 && force_warps != warps) return false;
        return fk_fits(group, VEC, RW, warps, a, dp);
    };
    // Measured on B200: whole-row staging with 4 warps per SM is the fastest layout whenever it
    // fits (1M x 22: 0.242 ms; 5 warps 0.262, 3 warps 0.298); otherwise the largest flush group that keeps
    // 4 warps per block wins (4M x 52: G = 16 4.77 TB/s vs dense with 2 warps 3.2 TB/s).
    // (whole rows are bank-conflict free only for odd J or J = 2 mod 4; unless forced, other joint counts take a
    // padded flush group)
    const bool dense_ok = QO || force_group == 0 || a.n_joints % 4 != 0;
    if (dense_ok && want(0, 4)) return launch_fk_cfg<0, 4, VEC, PF, QO>(a, dp);
    if constexpr (FULL) {
        if (force_warps == 5 && want(0, 5)) return launch_fk_cfg<0, 5, VEC, PF, QO>(a, dp);
        if (force_warps == 2 && want(0, 2)) return launch_fk_cfg<0, 2, VEC, PF, QO>(a, dp);
        if (want(32, 4)) return launch_fk_cfg<32, 4, VEC, PF, QO>(a, dp);
        if (want(16, 4)) return launch_fk_cfg<16, 4, VEC, PF, QO>(a, dp);
    }
    if (want(8, 4)) return launch_fk_cfg<8, 4, VEC, PF, QO>(a, dp);
    if (want(8, 1)) return launch_fk_cfg<8, 1, VEC, PF, QO>(a, dp);
    if (force_group >= 0 || force_warps >= 0) return fail(PMB_ERR_SHAPE, "PMB_FK_GROUP / PMB_FK_WARPS select no available variant");
    return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", a.n_slots);
}

template <bool PF, bool QO>
int launch_fk(const FkArgs &a, const DeviceProps &dp) {
    // 64-bit staging / stores need 8-byte aligned rows: even joint count.
    if ((a.n_joints % 2) == 0) return launch_fk_group<2, PF, QO, !PF && !QO>(a, dp);
    return launch_fk_group<1, PF, QO, !PF && !QO>(a, dp);
}

int fk_common(const float *rot, const float *gpos, int64_t gstride, const float *offsets, int64_t ostride,
              const int64_t *parents_host, int64_t n_frames, int32_t n_joints, float *pos, float *rout, bool quat_out,
              void *stream) {
    const bool rotations_only = quat_out && ostride == 0 && !pos;  // fk_quat without positions (mirror)
    if (!rot || !gpos || !offsets || (!pos && !rotations_only) || !rout) return fail(PMB_ERR_NULL, "fk: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (n_frames > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames must be below 2^31 per call");
    if (gstride != 0 && gstride != 3) return fail(PMB_ERR_SHAPE, "gpos_frame_stride must be 0 or 3");
    if (ostride != 0 && ostride != 3LL * n_joints) return fail(PMB_ERR_SHAPE, "offsets_frame_stride must be 0 or 3*n_joints");
    if (!aligned16(rot)) return fail(PMB_ERR_ALIGN, "rot must be 16-byte aligned");
    if (!aligned16(rout) || (pos && !aligned16(pos))) return fail(PMB_ERR_ALIGN, "output arrays must be 16-byte aligned");
    const pmb::JointProgram *prog = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    FkArgs a{rot, gpos, offsets, gstride, ostride, pos, rout, n_frames, n_joints, n_slots, prog, parents_host,
             static_cast<cudaStream_t>(stream)};
    if (ostride == 0 && !quat_out) {
        int rrc = PMB_OK;
        const bool rows_first = fk_rows_preferred(a, dp) && knob(K_FK_ROWS, -1) != 0;
        if ((!rows_first || knob(K_FK_MTRACKS, -1) == 1) && try_fk_mtracks(a, dp, rrc)) return rrc;
        if (try_fk_tracks(a, dp, rrc, rows_first)) return rrc;
        if (try_fk_rows(a, dp, rrc)) return rrc;
    }
    // fk_quat: the quaternion track kernel (qtracks_kernel.cuh) unless it does not apply; PMB_FKQ_TRACKS = 0 / 1 forces,
    // and forcing a variant of the chain kernels (PMB_FKQ_GROUP / _BLOCKS_PER_SM / _MATRIX) selects those.
    const int qtracks = knob(K_FKQ_TRACKS, (knob_set(K_FKQ_GROUP) || knob_set(K_FKQ_BLOCKS_PER_SM) || knob_set(K_FKQ_MATRIX)) ? 0 : -1);
    if (ostride == 0 && quat_out && qtracks != 0) {
        int trc = PMB_OK;
        const bool f = qtracks == 1;
        const bool took = rotations_only
                              ? launch_qtracks<pmb::kQtFkQuatRot>(rot, gpos, gstride, offsets, parents_host, n_frames, n_joints, rout, nullptr, a.stream, dp, f, trc)
                              : launch_qtracks<pmb::kQtFkQuat>(rot, gpos, gstride, offsets, parents_host, n_frames, n_joints, rout, pos, a.stream, dp, f, trc);
        if (took) return trc;
    }
    if (ostride == 0 && quat_out && (rotations_only || knob(K_FKQ_MATRIX, 0) == 0)) return launch_fk_quat_chain(a, dp);
    if (ostride == 0) return quat_out ? launch_fk<false, true>(a, dp) : launch_fk<false, false>(a, dp);
    return quat_out ? launch_fk<true, true>(a, dp) : launch_fk<true, false>(a, dp);
}

}  // namespace

extern "C" {

int pmb_fk_f32(const float *rot, const float *global_pos, int64_t gpos_frame_stride, const float *offsets,
               int64_t offsets_frame_stride, const int64_t *parents_host, int64_t n_frames, int32_t n_joints, float *positions,
               float *rotmats, void *stream) {
    return fk_common(rot, global_pos, gpos_frame_stride, offsets, offsets_frame_stride, parents_host, n_frames, n_joints, positions,
                     rotmats, false, stream);
}

int pmb_fk_quat_f32(const float *rot, const float *global_pos, int64_t gpos_frame_stride, const float *offsets,
                    int64_t offsets_frame_stride, const int64_t *parents_host, int64_t n_frames, int32_t n_joints, float *positions,
                    float *global_rots, void *stream) {
    return fk_common(rot, global_pos, gpos_frame_stride, offsets, offsets_frame_stride, parents_host, n_frames, n_joints, positions,
                     global_rots, true, stream);
}

// fk (rotations only) -> sign of quat.from_matrix -> mirror flip / re-index -> local rotations (ops/skeleton.py:322-331,
// :410-416), as ONE kernel where the quaternion track kernel applies; else the two-kernel path through the caller's scratch.
namespace {
bool mirror_fused_applies(const int64_t *parents_host, int32_t n_joints, const DeviceProps &dp) {
    if (knob(K_MIRROR_FUSED, 1) == 0) return false;
    const pmb::TrackProgram *tp = nullptr;
    int n_steps = 0;
    bool three = false;
    if (qt_pick_shape(pmb::kQtMirror, parents_host, n_joints, tp, n_steps, three) || n_steps == 0) return false;
    const int nt = three ? 3 : 4, fq = three ? 10 : 8;
    const QtShape sh = qt_shape(pmb::kQtMirror, n_joints, n_steps * nt, knob(K_QT_WARPS_PER_SM, 32), dp, nt, fq);
    return sh.warps * sh.blocks >= 4 && 2 * n_joints >= n_steps * nt;
}
}  // namespace

int pmb_mirror_local_needs_scratch(const int64_t *parents_host, int32_t n_joints) {
    if (!parents_host) return fail(PMB_ERR_NULL, "parents_host is NULL");
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS) return fail(PMB_ERR_SHAPE, "n_joints = %d outside [1, %d]", n_joints, PMB_MAX_JOINTS);
    DeviceProps dp;
    int rc = device_props(dp);
    if (rc) return rc;
    return mirror_fused_applies(parents_host, n_joints, dp) ? 0 : 1;
}

int pmb_mirror_local_f32(const float *local_quats, const int64_t *parents_host, const int64_t *joints_mapping_host, int32_t mirror_axis,
                         int64_t n_frames, int32_t n_joints, float *global_quats_scratch, float *mirrored_local_quats, void *stream) {
    if (!local_quats || !mirrored_local_quats) return fail(PMB_ERR_NULL, "mirror_local: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (n_frames > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames must be below 2^31 per call");
    if (mirror_axis < 0 || mirror_axis > 2) return fail(PMB_ERR_SHAPE, "mirror_axis must be 0 (X), 1 (Y) or 2 (Z)");
    if (!aligned16(local_quats) || !aligned16(mirrored_local_quats)) return fail(PMB_ERR_ALIGN, "quaternion arrays must be 16-byte aligned");
    const pmb::JointProgram *prog = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, false, prog, n_slots);  // validates parents[]
    if (rc) return rc;
    static thread_local pmb::QtMirrorTable mt;
    for (int j = 0; j < n_joints; ++j) {
        const int64_t mj = joints_mapping_host ? joints_mapping_host[j] : j;
        const int64_t pj = j > 0 ? parents_host[j] : 0;
        const int64_t mp = joints_mapping_host ? joints_mapping_host[pj] : pj;
        if (mj < 0 || mj >= n_joints || mp < 0 || mp >= n_joints)
            return fail(PMB_ERR_SHAPE, "joints_mapping[%d] outside [0, %d)", j, n_joints);
        mt.word[j] = static_cast<uint32_t>(mj) | (static_cast<uint32_t>(mp) << 16);
    }
    // the two vector components that change sign (skeleton.py:307-315): X -> (y, z), Y -> (x, z), Z -> (x, y)
    mt.fx = mirror_axis == 0 ? 1.f : -1.f, mt.fy = mirror_axis == 1 ? 1.f : -1.f, mt.fz = mirror_axis == 2 ? 1.f : -1.f;
    if (n_frames == 0) return PMB_OK;
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    if (mirror_fused_applies(parents_host, n_joints, dp)) {
        int trc = PMB_OK;
        // root at the origin, offsets unused: the walk is rotations only (the pointers only have to be readable)
        if (launch_qtracks<pmb::kQtMirror>(local_quats, local_quats, 0, local_quats, parents_host, n_frames, n_joints, mirrored_local_quats,
                                           nullptr, static_cast<cudaStream_t>(stream), dp, true, trc, mt))
            return trc;
    }
    if (!global_quats_scratch)
        return fail(PMB_ERR_NULL, "mirror_local: this skeleton takes the two-kernel path and needs global_quats_scratch "
                                  "(pmb_mirror_local_needs_scratch)");
    if ((rc = pmb_fk_quat_f32(local_quats, local_quats, 0, local_quats, 0, parents_host, n_frames, n_joints, nullptr, global_quats_scratch, stream)))
        return rc;
    return pmb_mirror_to_local_f32(global_quats_scratch, parents_host, joints_mapping_host, mirror_axis, n_frames, n_joints,
                                   mirrored_local_quats, stream);
}

}  // extern "C"
