// extern "C" entry points of libpymotion_b200.so (see include/pymotion_b200.h).
// Host side only validates, builds the joint program, picks a launch
// configuration and launches; nothing here computes on the CPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "dq_kernels.cuh"
#include "elementwise.cuh"
#include "fk_kernel.cuh"
#include "fk_quat_kernel.cuh"
#include "fk_rows_kernel.cuh"
#include "fk_lanes_kernel.cuh"
#include "fk_lanes_g_kernel.cuh"
#include "ik_kernels.cuh"
#include "joint_program.h"
#include "misc_ops.cuh"
#include "rotations_ext.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local char g_variant[128] = "";  // which kernel variant the last launch on this thread picked

void note_variant(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_variant, sizeof(g_variant), fmt, ap);
    va_end(ap);
}

int fail(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

int cuda_fail(cudaError_t e, const char *what) {
    return fail(PMB_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define PMB_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

struct DeviceProps {
    int sm_count = 0;
    int smem_optin = 0;
    int cc_major = 0, cc_minor = 0;
    bool ok = false;
};

int device_props(DeviceProps &out) {
    // per-device cache; the table is tiny and devices are few
    static std::mutex mu;
    static DeviceProps cache[64];
    int dev = 0;
    PMB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) return fail(PMB_ERR_CUDA, "device ordinal %d out of range", dev);
    DeviceProps &p = cache[dev];
    if (!p.ok) {
        PMB_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        p.ok = true;
    }
    out = p;
    return PMB_OK;
}

int check_program(const int64_t *parents_host, int32_t n_joints, bool detach, pmb::JointProgram &prog, int &n_slots) {
    if (!parents_host) return fail(PMB_ERR_NULL, "parents_host is NULL");
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS)
        return fail(PMB_ERR_SHAPE, "n_joints = %d outside [1, %d]", n_joints, PMB_MAX_JOINTS);
    const pmb::ProgramInfo info = pmb::build_joint_program(parents_host, n_joints, detach, prog.code);
    if (info.status == PMB_ERR_TOPOLOGY)
        return fail(PMB_ERR_TOPOLOGY,
                    "parents[%d] = %lld is not in [0, %d): joints must come after their parent (BVH order)",
                    info.bad_joint, static_cast<long long>(parents_host[info.bad_joint]), info.bad_joint);
    if (info.status != PMB_OK) return fail(info.status, "cannot build the joint program");
    n_slots = info.n_slots;
    return PMB_OK;
}

template <typename K>
int set_smem(K kernel, int bytes) {
    if (bytes > 48 * 1024) PMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return PMB_OK;
}

// ---- TMA descriptor for the quaternion input -------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_fn(EncodeTiledFn &out) {
    static std::mutex mu;
    static EncodeTiledFn cached = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!cached) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        PMB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(PMB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        cached = reinterpret_cast<EncodeTiledFn>(fn);
    }
    out = cached;
    return PMB_OK;
}

// rot viewed as a 2-D float tensor [n_frames][4 * n_joints]; box = 32 frames x C joints, hardware swizzle
// matched to the box row (64 B for C = 4, 128 B for C = 8) so thread-per-frame 16-byte reads are conflict free.
int make_rot_map(CUtensorMap &tm, const float *rot, int64_t n_frames, int32_t n_joints, int chunk, int box_frames = 32) {
    EncodeTiledFn enc;
    int rc = encode_fn(enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(4) * n_joints, static_cast<cuuint64_t>(n_frames)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(16) * n_joints};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(4 * chunk), static_cast<cuuint32_t>(box_frames)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = chunk == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    // L2 promotion: the granule L2 fetches from DRAM for a box row (experiment knob PMB_TMA_L2PROMO = 0 none,
    // 1 = 64 B, 2 = 128 B, 3 = 256 B)
    // Measured (profiles/r1_sweep_l2promo.jsonl): 256 B granules are worth +0.6 % at 22 joints, +4 % at 52, +5 % at 65
    // (a frame's quaternion row spans 1.4 .. 4 granules and the next chunk of the same frames finds them in L2).
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char *env = getenv("PMB_TMA_L2PROMO")) {
        const int v = atoi(env);
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(rot), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PMB_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return PMB_OK;
}

// ---- fk launch ----------------------------------------------------------------
struct FkArgs {
    const float *rot, *gpos, *offsets;
    long long gstride, ostride;
    float *pos, *rout;
    long long n_frames;
    int n_joints, n_slots;
    const pmb::JointProgram *prog;
    cudaStream_t stream;
};

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v ? atoi(v) : fallback;
}

template <int G, int WARPS, int VEC, bool PF, bool QO>
int launch_fk_cfg(const FkArgs &a, const DeviceProps &dp) {
    auto kernel = pmb::fk_chain_kernel<G, WARPS, VEC, PF, QO>;
    const int smem = pmb::fk_geom(G, VEC, QO ? 4 : 9, WARPS, a.n_joints, a.n_slots).block_bytes;
    int rc = set_smem(kernel, smem);
    if (rc) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk))) return rc;
    // persistent grid: as many blocks as are resident at once; warps walk the tiles round robin
    const long long tiles = (a.n_frames + 31) / 32;
    int per_sm = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk kernel does not fit on an SM (%d bytes of shared memory)", smem);
    per_sm = std::max(1, std::min(per_sm, env_int("PMB_FK_BLOCKS_PER_SM", per_sm)));
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_chain_kernel<G=%d,WARPS=%d,VEC=%d,PF=%d,QO=%d> grid=%lld smem=%d", G, WARPS, VEC, int(PF), int(QO), blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.ostride,
                                                                        a.pos, a.rout, a.n_frames, a.n_joints,
                                                                        a.n_slots, *a.prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- fk, row-team kernel (fk_rows_kernel.cuh) ------------------------------------
template <int S, int VEC>
int launch_fk_rows_cfg(const FkArgs &a, const DeviceProps &dp, int team_cap) {
    auto kernel = pmb::fk_rows_kernel<S, VEC>;
    const int smem = pmb::fk_rows_geom(S, a.n_joints).block_bytes;
    int rc = set_smem(kernel, smem);
    if (rc) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk))) return rc;
    const long long tiles = (a.n_frames + 31) / 32;
    int per_sm = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, pmb::kRowThreads, smem));
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk row kernel does not fit on an SM (%d bytes of shared memory)", smem);
    if (team_cap > 0) per_sm = std::min(per_sm, team_cap);
    per_sm = std::max(1, std::min(per_sm, env_int("PMB_FK_BLOCKS_PER_SM", per_sm)));
    const long long blocks = std::min<long long>(tiles, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_rows_kernel<S=%d,VEC=%d> grid=%lld (%d teams/SM) smem=%d", S, VEC, blocks, per_sm, smem);
    kernel<<<static_cast<unsigned>(blocks), pmb::kRowThreads, smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.pos,
                                                                                a.rout, a.n_frames, a.n_joints,
                                                                                env_int("PMB_ST_HINT", 0) | (env_int("PMB_FK_TILE_ORDER", 0) << 1),
                                                                                env_int("PMB_L2_PREFETCH", 0) ? a.rot : nullptr, *a.prog);
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

// ---- fk, lane = (frame, row) kernel (fk_lanes_kernel.cuh) ---------------------------------
template <int FR, int WARPS, int NB, bool TILE_IN = false>
int launch_fk_lanes_nb(const FkArgs &a, const DeviceProps &dp, int block_cap) {
    auto kernel = pmb::fk_lanes_kernel<FR, WARPS, NB, TILE_IN>;
    const int smem = pmb::fk_lanes_geom(FR, WARPS, a.n_joints, NB, TILE_IN).block_bytes;
    if (smem > dp.smem_optin) return fail(PMB_ERR_SHAPE, "fk lane kernel: %d joints do not fit in shared memory", a.n_joints);
    int rc = set_smem(kernel, smem);
    if (rc) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk, FR))) return rc;
    const long long tiles = (a.n_frames + FR - 1) / FR;
    int per_sm = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk lane kernel does not fit on an SM (%d bytes of shared memory)", smem);
    per_sm = std::max(1, std::min(per_sm, block_cap));
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_lanes_kernel<FR=%d,WARPS=%d,NB=%d,TILE_IN=%d> grid=%lld (%d warps/SM) smem=%d", FR, WARPS, NB, int(TILE_IN), blocks,
                 per_sm * WARPS, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.pos, a.rout,
                                                                        a.n_frames, a.n_joints, env_int("PMB_ST_HINT", 0),
                                                                        (TILE_IN || env_int("PMB_L2_PREFETCH", 0)) ? a.rot : nullptr, *a.prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

template <int FR, int WARPS>
int launch_fk_lanes_cfg(const FkArgs &a, const DeviceProps &dp, int block_cap, int n_boxes) {
    const int nb = env_int("PMB_FK_NB", n_boxes);  // TMA boxes in flight per warp
    if (env_int("PMB_FK_TILE_IN", 0) == 1 && a.n_joints % 4 != 0) {  // whole-tile contiguous input (experiment)
        if (nb == 3) return launch_fk_lanes_nb<FR, WARPS, 3, true>(a, dp, block_cap);
        return launch_fk_lanes_nb<FR, WARPS, 2, true>(a, dp, block_cap);
    }
    if (nb == 3) return launch_fk_lanes_nb<FR, WARPS, 3>(a, dp, block_cap);
    if (nb == 4) return launch_fk_lanes_nb<FR, WARPS, 4>(a, dp, block_cap);
    return launch_fk_lanes_nb<FR, WARPS, 2>(a, dp, block_cap);
}

// ---- fk, lane kernel with a grouped stage (fk_lanes_g_kernel.cuh) ---------------------------
template <int NB, int G>
int launch_fk_lanes_g_cfg(const FkArgs &a, const DeviceProps &dp) {
    constexpr int FR = 10, WARPS = 4;
    auto kernel = pmb::fk_lanes_g_kernel<FR, WARPS, NB, G>;
    const int smem = pmb::fk_lanes_g_geom(FR, WARPS, a.n_joints, a.n_slots, NB, G).block_bytes;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", a.n_slots);
    int rc = set_smem(kernel, smem);
    if (rc) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk, FR))) return rc;
    const long long tiles = (a.n_frames + FR - 1) / FR;
    int per_sm = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk grouped lane kernel does not fit on an SM (%d bytes of shared memory)", smem);
    per_sm = std::max(1, std::min(per_sm, env_int("PMB_FK_BLOCKS_PER_SM", 3)));  // 12 walking warps per SM
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    note_variant("fk_lanes_g_kernel<FR=%d,WARPS=%d,NB=%d,G=%d> grid=%lld (%d warps/SM) smem=%d", FR, WARPS, NB, G, blocks,
                 per_sm * WARPS, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, a.stream>>>(tm, a.gpos, a.gstride, a.offsets, a.pos, a.rout,
                                                                        a.n_frames, a.n_joints, a.n_slots, *a.prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// PMB_FK_LG = 16 | 32 forces the grouped lane kernel with that flush group (experiments / tests); PMB_FK_NB = 2 | 3.
bool try_fk_lanes_g(const FkArgs &a, const DeviceProps &dp, int &rc) {
    const int g = env_int("PMB_FK_LG", 0);
    if (g != 16 && g != 32) return false;
    const int nb = env_int("PMB_FK_NB", 2);
    if (g == 32) rc = nb == 3 ? launch_fk_lanes_g_cfg<3, 32>(a, dp) : launch_fk_lanes_g_cfg<2, 32>(a, dp);
    else rc = nb == 3 ? launch_fk_lanes_g_cfg<3, 16>(a, dp) : launch_fk_lanes_g_cfg<2, 16>(a, dp);
    return true;
}

// Worst-case number of lanes of one stage store that fall into the same shared-memory bank: lane = (frame, row)
// puts the frames of a tile 9J words apart, so only (9J mod 32) matters.  1 = conflict free.
inline int fk_lanes_bank_degree(int fr, int n_joints) {
    int count[32] = {0}, worst = 0;
    for (int f = 0; f < fr; ++f) worst = std::max(worst, ++count[(f * 9 * n_joints) & 31]);
    return worst;
}

struct FkLanesPlan {
    int fr = 0, warps = 0, frames_in_flight = 0, n_boxes = 2, warps_sm = 0;
};
// The tile size / block shape that keeps the most frames in flight per SM (what the throughput of the large
// skeletons follows): FR = 10 needs an even joint count, blocks of 1, 2 or 4 warps.
inline FkLanesPlan fk_lanes_plan(const FkArgs &a, const DeviceProps &dp) {
    FkLanesPlan best;
    for (int fr : {10, 8}) {
        if (fr == 10 && a.n_joints % 2) continue;
        for (int warps : {4, 2, 1}) {
            const int bytes = pmb::fk_lanes_geom(fr, warps, a.n_joints).block_bytes;
            if (bytes > dp.smem_optin) continue;
            const int blocks = std::min(32, (dp.smem_optin + 1024) / (bytes + 1024));
            const int warps_sm = std::min(blocks * warps, 12);  // measured at 22 joints: beyond 12 walking warps per SM it gets slower
            // ... discounted by the bank conflicts of that tile size (measured: J = 40, 100 frames 3-way 5.33 TB/s
            // against 96 frames 2-way 5.61; J = 52, 80 frames 2-way 5.00 against 64 frames conflict-free 4.56)
            const int degree = fk_lanes_bank_degree(fr, a.n_joints);
            const int weight = degree <= 1 ? 100 : degree == 2 ? 90 : degree == 3 ? 80 : 50;
            const int fif = warps_sm * fr * weight;
            if (fif > best.frames_in_flight) best = {fr, warps, fif, 2, warps_sm};
        }
    }
    // A third TMA box per warp when it is free (same number of blocks per SM) and the SM is short of warps to hide
    // the load latency with: measured +3.5 % at 65 joints (8 warps per SM), nothing at 40 joints (12 warps).
    if (best.fr && best.warps_sm <= 8) {
        auto blocks_of = [&](int nb) {
            const int bytes = pmb::fk_lanes_geom(best.fr, best.warps, a.n_joints, nb).block_bytes;
            return bytes > dp.smem_optin ? 0 : std::min(32, (dp.smem_optin + 1024) / (bytes + 1024));
        };
        if (blocks_of(3) == blocks_of(2)) best.n_boxes = 3;
    }
    return best;
}

// PMB_FK_LANES = 0 / 1 forces; PMB_FK_FR = 8 | 10, PMB_FK_WARPS = 1 | 2 | 4 pick the shape.
bool try_fk_lanes(const FkArgs &a, const DeviceProps &dp, int &rc, bool rows_preferred) {
    const int force = env_int("PMB_FK_LANES", -1);
    if (force == 0) return false;
    if (force != 1 && (getenv("PMB_FK_GROUP") || getenv("PMB_FK_ROWS"))) return false;  // another kernel is being forced
    FkLanesPlan plan = fk_lanes_plan(a, dp);
    if (plan.fr == 0) {
        if (force == 1) { rc = fail(PMB_ERR_SHAPE, "PMB_FK_LANES=1: the lane kernel does not fit"); return true; }
        return false;
    }
    int fr = env_int("PMB_FK_FR", plan.fr);
    if (fr != 10 || a.n_joints % 2) fr = 8;  // the spans of 10 frames are 16-byte multiples only for an even joint count
    const int warps = env_int("PMB_FK_WARPS", plan.warps);
    if (force != 1) {
        if (rows_preferred) return false;
        // a dense stage whose frames collide in 4 or more banks (J = 16, 32, 48, 64, ...: measured 2.0 TB/s at
        // J = 32) belongs to the thread-per-frame kernel with its padded stage
        if (fk_lanes_bank_degree(fr, a.n_joints) >= 4) return false;
    }
    const int cap = env_int("PMB_FK_BLOCKS_PER_SM", std::max(1, 12 / std::max(1, warps)));
    // the plan's ring depth only holds for the plan's own shape
    const int nb = (fr == plan.fr && warps == plan.warps) ? plan.n_boxes : 2;
    if (fr == 10) rc = warps == 1 ? launch_fk_lanes_cfg<10, 1>(a, dp, cap, nb) : warps == 2 ? launch_fk_lanes_cfg<10, 2>(a, dp, cap, nb) : launch_fk_lanes_cfg<10, 4>(a, dp, cap, nb);
    else rc = warps == 1 ? launch_fk_lanes_cfg<8, 1>(a, dp, cap, nb) : warps == 2 ? launch_fk_lanes_cfg<8, 2>(a, dp, cap, nb) : launch_fk_lanes_cfg<8, 4>(a, dp, cap, nb);
    return true;
}

// Which fk kernel (shared offsets, matrices out) -- measured on B200, DESIGN.md section 4:
//   row-team kernel      skeletons small enough for >= 4 teams per SM (J <= 30) with J not a multiple of 4
//                        (1M x 22: 0.2316 ms against 0.2415 ms thread-per-frame, 0.2326 ms lanes);
//   lane kernel          everything larger (2M x 40: 5.6 TB/s against 5.15; 4M x 52: 5.0 against 4.76; 4M x 65:
//                        4.4, as the row kernel, against 3.9), unless
//   thread-per-frame     the dense stage of the lane kernel would put >= 4 frames in one bank (J = 16, 32, 48,
//                        64, ...: 2.0 TB/s at J = 32 against 5.6), per-frame offsets, or a forced variant.
// Bank conflicts of the row-team kernel (lane = frame, 32 lanes 9J words apart): J odd: none; J = 2 (mod 4):
// none with its 64-bit stores; J = 0 (mod 4): 4-way (J = 52: 4.5 TB/s) up to 32-way (J = 32: 0.84 TB/s).
bool fk_rows_preferred(const FkArgs &a, const DeviceProps &dp) {
    return fk_rows_teams(2, a, dp) >= 4 && a.n_joints % 4 != 0;
}

bool try_fk_rows(const FkArgs &a, const DeviceProps &dp, int &rc) {
    const int force = env_int("PMB_FK_ROWS", -1);
    if (force == 0) return false;
    if (force != 1 && (getenv("PMB_FK_GROUP") || getenv("PMB_FK_WARPS"))) return false;  // a chain-kernel variant is being forced
    int stages = env_int("PMB_FK_STAGES", -1);
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

// ---- fk_quat, quaternion-chain kernel (fk_quat_kernel.cuh) ------------------------------
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
    if (const char *env = getenv("PMB_FKQ_GROUP")) {
        const int v = atoi(env);
        if (v >= a.n_joints) group = a.n_joints;  // whole rows
        else if (v >= 8 && v % 8 == 0) group = v;
    }
    const int smem = pmb::fkq_geom(group, WARPS, a.n_joints, a.n_slots).block_bytes;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", a.n_slots);
    auto kernel = a.pos ? pmb::fk_quat_chain_kernel<WARPS, true> : pmb::fk_quat_chain_kernel<WARPS, false>;
    int rc = set_smem(kernel, smem);
    if (rc) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, a.rot, a.n_frames, a.n_joints, pmb::kChunk))) return rc;
    int per_sm = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "fk_quat kernel does not fit on an SM");
    per_sm = std::max(1, std::min(per_sm, env_int("PMB_FKQ_BLOCKS_PER_SM", per_sm)));
    const long long tiles = (a.n_frames + 31) / 32;
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    auto magic_of = [](int d) { return static_cast<uint32_t>((1ULL << 32) / static_cast<uint32_t>(d)) + 1u; };
    const int tail = a.n_joints % group ? a.n_joints % group : group;
    note_variant("fk_quat_chain_kernel<WARPS=%d,POS=%d> group=%d grid=%lld smem=%d", WARPS, a.pos ? 1 : 0, group, blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, a.stream>>>(
        tm, a.gpos, a.gstride, a.offsets, a.pos, reinterpret_cast<float4 *>(a.rout), a.n_frames, a.n_joints, a.n_slots,
        group, magic_of(group), magic_of(tail), magic_of(3 * group), magic_of(3 * tail), *a.prog);
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
    const int force_group = env_int("PMB_FK_GROUP", -1);
    const int force_warps = env_int("PMB_FK_WARPS", -1);
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
              const int64_t *parents_host, int64_t n_frames, int32_t n_joints, float *pos, float *rout,
              bool quat_out, void *stream) {
    const bool rotations_only = quat_out && ostride == 0 && !pos;  // fk_quat without positions (mirror)
    if (!rot || !gpos || !offsets || (!pos && !rotations_only) || !rout) return fail(PMB_ERR_NULL, "fk: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (n_frames > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames must be below 2^31 per call");
    if (gstride != 0 && gstride != 3) return fail(PMB_ERR_SHAPE, "gpos_frame_stride must be 0 or 3");
    if (ostride != 0 && ostride != 3LL * n_joints)
        return fail(PMB_ERR_SHAPE, "offsets_frame_stride must be 0 or 3*n_joints");
    if (!aligned16(rot)) return fail(PMB_ERR_ALIGN, "rot must be 16-byte aligned");
    if (!aligned16(rout) || (pos && !aligned16(pos))) return fail(PMB_ERR_ALIGN, "output arrays must be 16-byte aligned");
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    FkArgs a{rot, gpos, offsets, gstride, ostride, pos, rout, n_frames, n_joints, n_slots, &prog,
             static_cast<cudaStream_t>(stream)};
    if (ostride == 0 && !quat_out) {
        int rrc = PMB_OK;
        if (try_fk_lanes_g(a, dp, rrc)) return rrc;
        const bool rows_first = fk_rows_preferred(a, dp) && env_int("PMB_FK_ROWS", -1) != 0;
        if (try_fk_lanes(a, dp, rrc, rows_first || env_int("PMB_FK_ROWS", -1) == 1)) return rrc;
        if (try_fk_rows(a, dp, rrc)) return rrc;
    }
    if (ostride == 0 && quat_out && (rotations_only || env_int("PMB_FKQ_MATRIX", 0) == 0)) return launch_fk_quat_chain(a, dp);
    if (ostride == 0) return quat_out ? launch_fk<false, true>(a, dp) : launch_fk<false, false>(a, dp);
    return quat_out ? launch_fk<true, true>(a, dp) : launch_fk<true, false>(a, dp);
}

inline int ew_grid(int64_t n, int threads, const DeviceProps &dp) {
    const int64_t want = (n + threads - 1) / threads;
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(dp.sm_count) * 16)));
}

// divisor of div_small (dq_kernels.cuh): floor(2^32 / d) + 1, and 0 for d = 1 (which has no 32-bit magic)
inline uint32_t magic_small(int d) { return d <= 1 ? 0u : static_cast<uint32_t>((1ULL << 32) / static_cast<uint32_t>(d)) + 1u; }

int tile_frames(int n_joints, int cap_elems) {
    int fb = (cap_elems / n_joints) & ~3;
    return std::max(4, std::min(64, fb));
}

// ---- host-buffer pipeline workspace -------------------------------------------
struct HostPipe {
    std::mutex mu;
    int device = -1;
    cudaStream_t stream[2] = {nullptr, nullptr};
    float *rot[2] = {nullptr, nullptr}, *gpos[2] = {nullptr, nullptr}, *pos[2] = {nullptr, nullptr},
          *rotm[2] = {nullptr, nullptr};
    float *offsets = nullptr;
    size_t cap_frames = 0, cap_joints = 0;
    void release() {
        for (int s = 0; s < 2; ++s) {
            if (rot[s]) cudaFree(rot[s]);
            if (gpos[s]) cudaFree(gpos[s]);
            if (pos[s]) cudaFree(pos[s]);
            if (rotm[s]) cudaFree(rotm[s]);
            rot[s] = gpos[s] = pos[s] = rotm[s] = nullptr;
            if (stream[s]) cudaStreamDestroy(stream[s]);
            stream[s] = nullptr;
        }
        if (offsets) cudaFree(offsets);
        offsets = nullptr;
        cap_frames = cap_joints = 0;
        device = -1;
    }
};
HostPipe g_pipe;

}  // namespace

extern "C" {

int pmb_version(void) { return PMB_VERSION; }
const char *pmb_last_error(void) { return g_err; }
const char *pmb_last_variant(void) { return g_variant; }

const char *pmb_status_string(int status) {
    switch (status) {
        case PMB_OK: return "ok";
        case PMB_ERR_NULL: return "null pointer";
        case PMB_ERR_SHAPE: return "bad shape";
        case PMB_ERR_ALIGN: return "misaligned pointer";
        case PMB_ERR_TOPOLOGY: return "bad parents table";
        case PMB_ERR_CUDA: return "CUDA error";
        case PMB_ERR_ROOT_OFFSET: return "offsets[0] != 0";
        default: return "unknown status";
    }
}

int pmb_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len) {
    DeviceProps dp;
    int rc = device_props(dp);
    if (rc) return rc;
    if (sm_count) *sm_count = dp.sm_count;
    if (cc_major) *cc_major = dp.cc_major;
    if (cc_minor) *cc_minor = dp.cc_minor;
    if (name && name_len > 0) {
        int dev = 0;
        cudaDeviceProp prop;
        PMB_CUDA(cudaGetDevice(&dev));
        PMB_CUDA(cudaGetDeviceProperties(&prop, dev));
        snprintf(name, static_cast<size_t>(name_len), "%s", prop.name);
    }
    return PMB_OK;
}

int pmb_build_joint_program(const int64_t *parents_host, int32_t n_joints, uint32_t *codes_out) {
    if (!codes_out) return fail(PMB_ERR_NULL, "codes_out is NULL");
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    memcpy(codes_out, prog.code, sizeof(uint32_t) * static_cast<size_t>(n_joints));
    return n_slots;
}

int pmb_fk_f32(const float *rot, const float *global_pos, int64_t gpos_frame_stride, const float *offsets,
               int64_t offsets_frame_stride, const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
               float *positions, float *rotmats, void *stream) {
    return fk_common(rot, global_pos, gpos_frame_stride, offsets, offsets_frame_stride, parents_host, n_frames,
                     n_joints, positions, rotmats, false, stream);
}

int pmb_fk_quat_f32(const float *rot, const float *global_pos, int64_t gpos_frame_stride, const float *offsets,
                    int64_t offsets_frame_stride, const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                    float *positions, float *global_rots, void *stream) {
    return fk_common(rot, global_pos, gpos_frame_stride, offsets, offsets_frame_stride, parents_host, n_frames,
                     n_joints, positions, global_rots, true, stream);
}

int pmb_to_root_dual_quat_f32(const float *rotations, const float *global_pos, int64_t gpos_frame_stride,
                              const int64_t *parents_host, const float *offsets, const float *offsets_host0,
                              int64_t n_frames, int32_t n_joints, float *dq, void *stream) {
    if (!rotations || !global_pos || !offsets || !dq) return fail(PMB_ERR_NULL, "to_root_dual_quat: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (gpos_frame_stride != 0 && gpos_frame_stride != 3) return fail(PMB_ERR_SHAPE, "gpos_frame_stride must be 0 or 3");
    if (!aligned16(rotations) || !aligned16(dq)) return fail(PMB_ERR_ALIGN, "rotations and dq must be 16-byte aligned");
    if (offsets_host0 && (offsets_host0[0] != 0.f || offsets_host0[1] != 0.f || offsets_host0[2] != 0.f))
        return fail(PMB_ERR_ROOT_OFFSET, "offsets[0] must be zero (ops/skeleton.py:227)");
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, true, prog, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    if (n_frames > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames must be below 2^31 per call");
    // Joints per flush.  Dual quaternions are whole 32-byte sectors, so partial flushes cost DRAM little and
    // occupancy matters more than for fk (measured, 1M x 22: whole rows / 4 warps per SM 0.239 ms, 8 joints /
    // 12 warps 0.214 ms, 16 joints / 8 warps 0.186 ms): take the largest group that still lets TWO 4-warp
    // blocks share an SM.
    constexpr int WARPS = 4;
    const int budget = (dp.smem_optin - 2048) / 2;
    int group = n_joints;
    if (pmb::dq_geom(group, WARPS, n_joints, n_slots).block_bytes > budget) {
        group = ((n_joints + 7) / 8) * 8;
        while (group > 8 && pmb::dq_geom(group, WARPS, n_joints, n_slots).block_bytes > budget) group -= 8;
    }
    if (const char *env = getenv("PMB_DQ_GROUP")) {
        const int v = atoi(env);
        if (v >= 8 && v % 8 == 0 && v < n_joints) group = v;
    }
    const int smem = pmb::dq_geom(group, WARPS, n_joints, n_slots).block_bytes;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", n_slots);
    auto kernel = pmb::to_root_dq_kernel<WARPS>;
    if ((rc = set_smem(kernel, smem))) return rc;
    CUtensorMap tm;
    if ((rc = make_rot_map(tm, rotations, n_frames, n_joints, pmb::kChunk))) return rc;
    int per_sm = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, smem));
    if (per_sm < 1) return fail(PMB_ERR_CUDA, "to_root_dual_quat kernel does not fit on an SM");
    per_sm = std::max(1, std::min(per_sm, env_int("PMB_DQ_BLOCKS_PER_SM", per_sm)));
    const long long tiles = (n_frames + 31) / 32;
    const long long blocks = std::min<long long>((tiles + WARPS - 1) / WARPS, static_cast<long long>(per_sm) * dp.sm_count);
    auto magic_of = [](int d) { return static_cast<uint32_t>((1ULL << 32) / static_cast<uint32_t>(d)) + 1u; };
    const int tail = n_joints % group ? n_joints % group : group;
    note_variant("to_root_dq_kernel<WARPS=%d> group=%d grid=%lld smem=%d", WARPS, group, blocks, smem);
    kernel<<<static_cast<unsigned>(blocks), WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        tm, global_pos, gpos_frame_stride, offsets, reinterpret_cast<float4 *>(dq), n_frames, n_joints, n_slots, group,
        magic_of(2 * group), magic_of(2 * tail), prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_from_root_dual_quat_f32(const float *dq, const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                                float *translations, float *rotations, void *stream) {
    if (!dq || !translations || !rotations) return fail(PMB_ERR_NULL, "from_root_dual_quat: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (!aligned16(dq) || !aligned16(rotations) || !aligned16(translations))
        return fail(PMB_ERR_ALIGN, "dq, rotations and translations must be 16-byte aligned");
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, true, prog, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    constexpr int THREADS = 256;
    // Elements (frame, joint) per block tile.  The translations stage costs 12 bytes of shared memory per element and
    // every thread has one 32-byte element in flight, so the tile size sets the bytes in flight per SM: 2304 elements
    // (27 KB) lets the 8 blocks of 256 threads the SM can hold all be resident.  Measured against the 4096 of the
    // first half (profiles/r1_sweep_frdq_tile.jsonl): 4M x 52 2.249 -> 1.888 ms, 4M x 65 2.756 -> 2.319 ms, 1M x 22 unchanged.
    const int fb = tile_frames(n_joints, env_int("PMB_FRDQ_ELEMS", 2304));
    const int smem = ((fb * n_joints * 12 + 15) & ~15) + ((n_joints * 2 + 15) & ~15);
    auto kernel = pmb::from_root_dq_kernel<THREADS>;
    if ((rc = set_smem(kernel, smem))) return rc;
    const long long blocks = (n_frames + fb - 1) / fb;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    const uint32_t magic = magic_small(n_joints);
    kernel<<<static_cast<unsigned>(blocks), THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(dq), translations, reinterpret_cast<float4 *>(rotations), n_frames, n_joints,
        fb, magic, prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_from_global_rotations_f32(const float *global_quats, const int64_t *parents_host, int64_t n_frames,
                                  int32_t n_joints, float *local_quats, void *stream) {
    if (!global_quats || !local_quats) return fail(PMB_ERR_NULL, "from_global_rotations: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (!aligned16(global_quats) || !aligned16(local_quats))
        return fail(PMB_ERR_ALIGN, "quaternion arrays must be 16-byte aligned");
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    constexpr int THREADS = 256;
    const int fb = tile_frames(n_joints, 4096);
    const int smem = (n_joints * 2 + 15) & ~15;
    const long long blocks = (n_frames + fb - 1) / fb;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    const uint32_t magic = magic_small(n_joints);
    pmb::from_global_rotations_kernel<THREADS><<<static_cast<unsigned>(blocks), THREADS, smem,
                                                 static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(global_quats), reinterpret_cast<float4 *>(local_quats), n_frames, n_joints, fb,
        magic, prog);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- host-buffer pipeline -------------------------------------------------------
void pmb_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_pipe.mu);
    g_pipe.release();
}

int pmb_fk_f32_host(const float *rot_host, const float *global_pos_host, const float *offsets_host,
                    const int64_t *parents_host, int64_t n_frames, int32_t n_joints, float *positions_host,
                    float *rotmats_host, int64_t chunk_frames) {
    if (!rot_host || !global_pos_host || !offsets_host || !positions_host || !rotmats_host)
        return fail(PMB_ERR_NULL, "fk_host: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames < 0");
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS) return fail(PMB_ERR_SHAPE, "n_joints = %d out of range", n_joints);
    if (chunk_frames <= 0) chunk_frames = 1 << 16;
    chunk_frames = (chunk_frames + 31) & ~31LL;
    std::lock_guard<std::mutex> lock(g_pipe.mu);
    HostPipe &w = g_pipe;
    int dev = 0;
    PMB_CUDA(cudaGetDevice(&dev));
    const size_t J = static_cast<size_t>(n_joints);
    if (w.device != dev || w.cap_frames < static_cast<size_t>(chunk_frames) || w.cap_joints < J) {
        w.release();
        const size_t F = static_cast<size_t>(chunk_frames);
        for (int s = 0; s < 2; ++s) {
            PMB_CUDA(cudaStreamCreateWithFlags(&w.stream[s], cudaStreamNonBlocking));
            PMB_CUDA(cudaMalloc(&w.rot[s], F * J * 16));
            PMB_CUDA(cudaMalloc(&w.gpos[s], F * 12));
            PMB_CUDA(cudaMalloc(&w.pos[s], F * J * 12));
            PMB_CUDA(cudaMalloc(&w.rotm[s], F * J * 36));
        }
        PMB_CUDA(cudaMalloc(&w.offsets, static_cast<size_t>(PMB_MAX_JOINTS) * 12));
        w.device = dev, w.cap_frames = F, w.cap_joints = J;
    }
    PMB_CUDA(cudaMemcpyAsync(w.offsets, offsets_host, J * 12, cudaMemcpyHostToDevice, w.stream[0]));
    PMB_CUDA(cudaStreamSynchronize(w.stream[0]));
    int slot = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk_frames, slot ^= 1) {
        const size_t n = static_cast<size_t>(std::min<int64_t>(chunk_frames, n_frames - f0));
        cudaStream_t st = w.stream[slot];  // stream order makes the slot's buffers safe to reuse
        PMB_CUDA(cudaMemcpyAsync(w.rot[slot], rot_host + f0 * J * 4, n * J * 16, cudaMemcpyHostToDevice, st));
        PMB_CUDA(cudaMemcpyAsync(w.gpos[slot], global_pos_host + f0 * 3, n * 12, cudaMemcpyHostToDevice, st));
        int rc = pmb_fk_f32(w.rot[slot], w.gpos[slot], 3, w.offsets, 0, parents_host, static_cast<int64_t>(n), n_joints,
                            w.pos[slot], w.rotm[slot], st);
        if (rc) return rc;
        PMB_CUDA(cudaMemcpyAsync(positions_host + f0 * J * 3, w.pos[slot], n * J * 12, cudaMemcpyDeviceToHost, st));
        PMB_CUDA(cudaMemcpyAsync(rotmats_host + f0 * J * 9, w.rotm[slot], n * J * 36, cudaMemcpyDeviceToHost, st));
    }
    PMB_CUDA(cudaStreamSynchronize(w.stream[0]));
    PMB_CUDA(cudaStreamSynchronize(w.stream[1]));
    return PMB_OK;
}

// ---- element-wise ---------------------------------------------------------------
#define PMB_EW_PROLOGUE(n, ...)                                                                      \
    const void *ptrs_[] = {__VA_ARGS__};                                                             \
    for (const void *p_ : ptrs_)                                                                     \
        if (!p_) return fail(PMB_ERR_NULL, "%s: NULL array pointer", __func__);                      \
    if ((n) < 0) return fail(PMB_ERR_SHAPE, "%s: n < 0", __func__);                                  \
    if ((n) == 0) return PMB_OK;                                                                     \
    DeviceProps dp_;                                                                                 \
    {                                                                                                \
        int rc_ = device_props(dp_);                                                                 \
        if (rc_) return rc_;                                                                         \
    }                                                                                                \
    const int grid_ = ew_grid((n), 256, dp_);                                                        \
    cudaStream_t st_ = static_cast<cudaStream_t>(stream)

#define PMB_NEED16(p) \
    if (!aligned16(p)) return fail(PMB_ERR_ALIGN, "%s: " #p " must be 16-byte aligned", __func__)

int pmb_quat_mul_f32(const float *q0, const float *q1, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q0, q1, out);
    PMB_NEED16(q0); PMB_NEED16(q1); PMB_NEED16(out);
    pmb::quat_mul_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q0, (const float4 *)q1, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_mul_vec_f32(const float *q, const float *v, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, v, out);
    PMB_NEED16(q);
    pmb::quat_mul_vec_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, v, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_length_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q);
    pmb::quat_length_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_normalize_f32(const float *q, float eps, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q); PMB_NEED16(out);
    pmb::quat_normalize_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, eps, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_conjugate_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q); PMB_NEED16(out);
    pmb::quat_conjugate_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_matrix_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q);
    PMB_NEED16(out);
    pmb::quat_to_matrix_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_matrix_f32(const float *m, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, m, out);
    PMB_NEED16(out);
    pmb::quat_from_matrix_kernel<<<grid_, 256, 0, st_>>>(m, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_from_rotation_translation_f32(const float *rotations, const float *translations, float *dq, int64_t n,
                                         void *stream) {
    PMB_EW_PROLOGUE(n, rotations, translations, dq);
    PMB_NEED16(rotations); PMB_NEED16(dq);
    if (aligned32(dq)) pmb::dq_from_rt_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)rotations, translations, (float4 *)dq, n);
    else pmb::dq_from_rt_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)rotations, translations, (float4 *)dq, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_from_translation_f32(const float *translations, float *dq, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, translations, dq);
    PMB_NEED16(dq);
    if (aligned32(dq)) pmb::dq_from_t_kernel<true><<<grid_, 256, 0, st_>>>(translations, (float4 *)dq, n);
    else pmb::dq_from_t_kernel<false><<<grid_, 256, 0, st_>>>(translations, (float4 *)dq, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_to_rotation_translation_f32(const float *dq, float *rotations, float *translations, int64_t n,
                                       void *stream) {
    PMB_EW_PROLOGUE(n, dq, rotations, translations);
    PMB_NEED16(dq); PMB_NEED16(rotations);
    if (aligned32(dq)) pmb::dq_to_rt_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)rotations, translations, n);
    else pmb::dq_to_rt_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)rotations, translations, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- the rest of the quaternion / dual-quaternion surface (rotations_ext.cuh) ------------------
int pmb_quat_from_angle_axis_f32(const float *angle, const float *axis, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, angle, axis, out);
    PMB_NEED16(out);
    pmb::quat_from_angle_axis_kernel<<<grid_, 256, 0, st_>>>(angle, axis, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_scaled_angle_axis_f32(const float *scaled_axis, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, scaled_axis, out);
    PMB_NEED16(out);
    pmb::quat_from_scaled_angle_axis_kernel<<<grid_, 256, 0, st_>>>(scaled_axis, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_euler_f32(const float *euler, const uint8_t *order_codes, int64_t order_stride, float *out, int64_t n,
                            void *stream) {
    PMB_EW_PROLOGUE(n, euler, order_codes, out);
    PMB_NEED16(out);
    if (order_stride != 0 && order_stride != 1) return fail(PMB_ERR_SHAPE, "%s: order_stride must be 0 or 1", __func__);
    pmb::quat_from_euler_kernel<<<grid_, 256, 0, st_>>>(euler, order_codes, order_stride, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_euler_f32(const float *q, const uint8_t *order_codes, int64_t order_stride, float *out, int64_t n,
                          void *stream) {
    PMB_EW_PROLOGUE(n, q, order_codes, out);
    PMB_NEED16(q);
    if (order_stride != 0 && order_stride != 1) return fail(PMB_ERR_SHAPE, "%s: order_stride must be 0 or 1", __func__);
    pmb::quat_to_euler_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, order_codes, order_stride, out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_angle_axis_f32(const float *q, float *angle, float *axis, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, angle, axis);
    PMB_NEED16(q);
    pmb::quat_to_angle_axis_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, angle, axis, 0, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_to_scaled_angle_axis_f32(const float *q, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, out);
    PMB_NEED16(q);
    pmb::quat_to_angle_axis_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, nullptr, out, 1, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_slerp_f32(const float *q0, const float *q1, const float *t, int64_t t_stride, int32_t shortest, float *out,
                       int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q0, q1, t, out);
    PMB_NEED16(q0); PMB_NEED16(q1); PMB_NEED16(out);
    if (t_stride != 0 && t_stride != 1) return fail(PMB_ERR_SHAPE, "%s: t_stride must be 0 or 1", __func__);
    pmb::quat_slerp_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q0, (const float4 *)q1, t, t_stride, shortest,
                                                   (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_to_f32(const float *v1, const float *v2, int32_t normalize_input, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, v1, v2, out);
    PMB_NEED16(out);
    pmb::quat_from_to_kernel<<<grid_, 256, 0, st_>>>(v1, v2, nullptr, normalize_input, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_quat_from_to_axis_f32(const float *v1, const float *v2, const float *rot_axis, int32_t normalize_input, float *out,
                              int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, v1, v2, rot_axis, out);
    PMB_NEED16(out);
    pmb::quat_from_to_kernel<<<grid_, 256, 0, st_>>>(v1, v2, rot_axis, normalize_input, (float4 *)out, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int64_t pmb_unroll_workspace_bytes(int64_t n_steps, int64_t n_cols) {
    if (n_steps < 0 || n_cols < 0) return 0;
    const int64_t chunks = (n_steps + pmb::kUnrollChunk - 1) / pmb::kUnrollChunk;
    return n_steps * n_cols + chunks * n_cols + 16;
}
int pmb_unroll_f32(const float *x, int32_t width, int64_t n_steps, int64_t n_cols, float *out, void *workspace,
                   int64_t workspace_bytes, void *stream) {
    if (width != 4 && width != 8) return fail(PMB_ERR_SHAPE, "%s: width must be 4 (quaternions) or 8 (dual quaternions)", __func__);
    if (n_steps < 0 || n_cols < 0) return fail(PMB_ERR_SHAPE, "%s: negative size", __func__);
    if (n_steps == 0 || n_cols == 0) return PMB_OK;
    if (!x || !out || !workspace) return fail(PMB_ERR_NULL, "%s: NULL array pointer", __func__);
    PMB_NEED16(x); PMB_NEED16(out);
    if (workspace_bytes < pmb_unroll_workspace_bytes(n_steps, n_cols))
        return fail(PMB_ERR_SHAPE, "%s: workspace too small (%lld < %lld bytes)", __func__,
                    static_cast<long long>(workspace_bytes), static_cast<long long>(pmb_unroll_workspace_bytes(n_steps, n_cols)));
    DeviceProps dp;
    int rc = device_props(dp);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t chunks = (n_steps + pmb::kUnrollChunk - 1) / pmb::kUnrollChunk;
    if (n_cols > 0x7FFFFFFFLL || (chunks * n_cols + 127) / 128 > 0x7FFFFFFFLL)
        return fail(PMB_ERR_SHAPE, "%s: array too large for one launch", __func__);
    uint8_t *local = static_cast<uint8_t *>(workspace), *agg = local + n_steps * n_cols;
    const int w4 = width / 4;
    const long long local_threads = chunks * n_cols;
    pmb::unroll_local_kernel<<<static_cast<unsigned>((local_threads + 127) / 128), 128, 0, st>>>(
        reinterpret_cast<const float4 *>(x), w4, n_steps, n_cols, chunks, local, agg);
    PMB_CUDA(cudaGetLastError());
    pmb::unroll_chunks_kernel<<<static_cast<unsigned>(n_cols), 256, 0, st>>>(agg, chunks, n_cols);
    PMB_CUDA(cudaGetLastError());
    if (n_cols <= 512 && env_int("PMB_UNROLL_CHUNK_APPLY", 1))
        pmb::unroll_apply_chunk_kernel<<<static_cast<unsigned>(chunks), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(x), w4, n_steps, static_cast<int>(n_cols), magic_small(static_cast<int>(n_cols)),
            local, agg, reinterpret_cast<float4 *>(out));
    else
        pmb::unroll_apply_kernel<<<ew_grid(n_steps * n_cols, 256, dp), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(x), w4, n_steps, n_cols, local, agg, reinterpret_cast<float4 *>(out));
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_dq_is_unit_f32(const float *dq, float atol, int64_t n, int32_t *flags3, void *stream) {
    if (!flags3) return fail(PMB_ERR_NULL, "%s: flags3 is NULL", __func__);
    PMB_CUDA(cudaMemsetAsync(flags3, 0, 3 * sizeof(int32_t), static_cast<cudaStream_t>(stream)));
    PMB_EW_PROLOGUE(n, dq);
    PMB_NEED16(dq);
    pmb::dq_is_unit_kernel<<<grid_, 256, 0, st_>>>((const float4 *)dq, atol, flags3, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_dq_normalize_f32(const float *dq, float *out, int64_t n, int32_t *flags3, void *stream) {
    if (!flags3) return fail(PMB_ERR_NULL, "%s: flags3 is NULL", __func__);
    PMB_CUDA(cudaMemsetAsync(flags3, 0, 3 * sizeof(int32_t), static_cast<cudaStream_t>(stream)));
    PMB_EW_PROLOGUE(n, dq, out);
    PMB_NEED16(dq); PMB_NEED16(out);
    // pass 1 reads the array and reduces the reference's whole-array is_unit verdict into flags3; pass 2 reads it
    // again and writes every result once (96 bytes of traffic per element instead of the 128 of scale-then-fix)
    if (aligned32(dq) && aligned32(out)) {
        pmb::dq_normalize_flags_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)dq, flags3, n);
        PMB_CUDA(cudaGetLastError());
        pmb::dq_normalize_write_kernel<true><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)out, flags3, n);
    } else {
        pmb::dq_normalize_flags_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)dq, flags3, n);
        PMB_CUDA(cudaGetLastError());
        pmb::dq_normalize_write_kernel<false><<<grid_, 256, 0, st_>>>((const float4 *)dq, (float4 *)out, flags3, n);
    }
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- fk consumers: from_root_positions, mirror (ik_kernels.cuh) ------------------------------
int pmb_from_root_positions_f32(const float *positions, const int64_t *parents_host, const float *offsets,
                                int64_t n_frames, int32_t n_joints, float *rotations, void *stream) {
    if (!positions || !offsets || !rotations) return fail(PMB_ERR_NULL, "from_root_positions: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    PMB_NEED16(rotations);
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    // children in index order, like the reference's `children` lists (skeleton.py:122-126)
    pmb::ChildTable kids;
    {
        int count[PMB_MAX_JOINTS + 1] = {0};
        for (int i = 1; i < n_joints; ++i) ++count[parents_host[i]];
        kids.start[0] = 0;
        for (int j = 0; j < n_joints; ++j) kids.start[j + 1] = static_cast<uint16_t>(kids.start[j] + count[j]);
        int fill[PMB_MAX_JOINTS] = {0};
        for (int i = 1; i < n_joints; ++i) {
            const int p = static_cast<int>(parents_host[i]);
            kids.child[kids.start[p] + fill[p]++] = static_cast<uint16_t>(i);
        }
    }
    DeviceProps dp;
    if ((rc = device_props(dp))) return rc;
    constexpr int THREADS = 128;
    const int smem = n_joints * 16 + n_slots * THREADS * 16;
    if (smem > dp.smem_optin)
        return fail(PMB_ERR_TOPOLOGY, "joint order needs %d live branch slots; does not fit in shared memory", n_slots);
    auto kernel = pmb::from_root_positions_kernel<THREADS>;
    if ((rc = set_smem(kernel, smem))) return rc;
    // Thread = frame reads its positions row 12 bytes at a time, so the op lives on L1 hits, and L1 is what the
    // shared-memory carve-out leaves of the SM's 256 KB.  Left to the driver, the carve-out is sized for the nine
    // blocks the registers allow: with many live branch slots (2 KB each per block) that takes nearly all of it
    // (measured at 4M x 65, 10 slots: 46 % L1 hits, 7.5 x the input re-read from L2, 10.4 ms).  Ask for what six
    // blocks need instead; the occupancy follows the carve-out.  Measured (profiles/r1_sweep_frp_carveout.jsonl):
    // 4M x 65 10.42 -> 4.79 ms (4.95 for 3 .. 4 blocks, 8.2 for 2), 4M x 52 3.84 -> 3.66 ms, 1M x 22 0.366 -> 0.360 ms.
    {
        const int target = std::max(1, std::min(9, env_int("PMB_FRP_BLOCKS_PER_SM", 6)));
        const int want = target * (smem + 1024);
        const int pct = std::max(1, std::min(100, (want * 100 + dp.smem_optin - 1) / dp.smem_optin));
        PMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        note_variant("from_root_positions_kernel<%d> smem=%d carveout=%d%% (for %d blocks/SM)", THREADS, smem, pct, target);
    }
    const long long blocks = (n_frames + THREADS - 1) / THREADS;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    kernel<<<static_cast<unsigned>(blocks), THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        positions, offsets, reinterpret_cast<float4 *>(rotations), n_frames, n_joints, n_slots, prog, kids);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_mirror_to_local_f32(const float *global_quats, const int64_t *parents_host, const int64_t *joints_mapping_host,
                            int32_t mirror_axis, int64_t n_frames, int32_t n_joints, float *local_quats, void *stream) {
    if (!global_quats || !local_quats) return fail(PMB_ERR_NULL, "mirror_to_local: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames = %lld < 0", static_cast<long long>(n_frames));
    if (mirror_axis < 0 || mirror_axis > 2) return fail(PMB_ERR_SHAPE, "mirror_axis must be 0 (X), 1 (Y) or 2 (Z)");
    PMB_NEED16(global_quats); PMB_NEED16(local_quats);
    pmb::JointProgram prog;
    int n_slots = 0;
    int rc = check_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    pmb::JointMap jm;
    for (int j = 0; j < n_joints; ++j) {
        const int64_t m = joints_mapping_host ? joints_mapping_host[j] : j;
        if (m < 0 || m >= n_joints) return fail(PMB_ERR_SHAPE, "joints_mapping[%d] = %lld outside [0, %d)", j, static_cast<long long>(m), n_joints);
        jm.map[j] = static_cast<uint16_t>(m);
    }
    if (n_frames == 0) return PMB_OK;
    constexpr int THREADS = 256;
    const int fb = tile_frames(n_joints, 4096);
    const long long blocks = (n_frames + fb - 1) / fb;
    if (blocks > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "n_frames too large for one launch");
    const uint32_t magic = magic_small(n_joints);
    // the two vector components that change sign (skeleton.py:307-315): X -> (y, z), Y -> (x, z), Z -> (x, y)
    const float fx = mirror_axis == 0 ? 1.f : -1.f, fy = mirror_axis == 1 ? 1.f : -1.f, fz = mirror_axis == 2 ? 1.f : -1.f;
    pmb::mirror_to_local_kernel<THREADS><<<static_cast<unsigned>(blocks), THREADS, (4 * n_joints + 15) & ~15, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(global_quats), reinterpret_cast<float4 *>(local_quats), n_frames, n_joints, fb,
        magic, fx, fy, fz, prog, jm);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_vec_mirror_f32(const float *v, int32_t axis, float *out, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, v, out);
    if (axis < 0 || axis > 2) return fail(PMB_ERR_SHAPE, "%s: axis must be 0, 1 or 2", __func__);
    pmb::vec_mirror_kernel<<<ew_grid(3 * n, 256, dp_), 256, 0, st_>>>(v, out, axis, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

int pmb_root_center_f32(const float *positions, float *out, int64_t n_frames, int32_t n_joints, void *stream) {
    if (n_joints < 1) return fail(PMB_ERR_SHAPE, "%s: n_joints < 1", __func__);
    PMB_EW_PROLOGUE(n_frames, positions, out);
    pmb::root_center_kernel<<<ew_grid(n_frames * n_joints * 3, 256, dp_), 256, 0, st_>>>(positions, out, n_frames, n_joints);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

// ---- ortho6d, center_of_mass, interpolate_positions, vector.normalize (misc_ops.cuh) ---------
int pmb_ortho6d_from_matrix_f32(const float *rotmats, float *ortho6d, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, rotmats, ortho6d);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    pmb::ortho6d_from_matrix_kernel<<<grid_, 256, 0, st_>>>(rotmats, ortho6d, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_ortho6d_from_quat_f32(const float *q, float *ortho6d, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, q, ortho6d);
    PMB_NEED16(q);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    pmb::ortho6d_from_quat_kernel<<<grid_, 256, 0, st_>>>((const float4 *)q, ortho6d, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_ortho6d_to_matrix_f32(const float *ortho6d, float *rotmats, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, ortho6d, rotmats);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    PMB_NEED16(rotmats);
    pmb::ortho6d_to_matrix_kernel<<<grid_, 256, 0, st_>>>(ortho6d, rotmats, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_ortho6d_to_quat_f32(const float *ortho6d, float *q, int64_t n, void *stream) {
    PMB_EW_PROLOGUE(n, ortho6d, q);
    PMB_NEED16(q);
    if (reinterpret_cast<uintptr_t>(ortho6d) & 7u) return fail(PMB_ERR_ALIGN, "%s: ortho6d must be 8-byte aligned", __func__);
    pmb::ortho6d_to_quat_kernel<<<grid_, 256, 0, st_>>>(ortho6d, (float4 *)q, n);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_center_of_mass_f32(const float *joints, const float *weights, int64_t weights_frame_stride, int64_t n_frames,
                           int32_t n_joints, float *out, void *stream) {
    if (n_joints < 1) return fail(PMB_ERR_SHAPE, "%s: n_joints < 1", __func__);
    if (weights_frame_stride != 0 && weights_frame_stride != n_joints)
        return fail(PMB_ERR_SHAPE, "%s: weights_frame_stride must be 0 or n_joints", __func__);
    PMB_EW_PROLOGUE(n_frames, joints, weights, out);
    pmb::center_of_mass_kernel<<<ew_grid(3 * n_frames, 256, dp_), 256, 0, st_>>>(joints, weights, weights_frame_stride, out,
                                                                              n_frames, n_joints);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_interpolate_positions_f32(const double *sample_times, const double *original_times, const float *positions,
                                  int64_t outer, int64_t n_original, int64_t n_samples, int64_t inner, float *out,
                                  int32_t *idx_workspace, float *weight_workspace, void *stream) {
    if (n_original < 2) return fail(PMB_ERR_SHAPE, "%s: at least two original times are needed", __func__);
    if (outer < 0 || inner < 0 || n_samples < 0) return fail(PMB_ERR_SHAPE, "%s: negative size", __func__);
    if (n_original > 0x7FFFFFFFLL) return fail(PMB_ERR_SHAPE, "%s: n_original must be below 2^31", __func__);
    if (outer == 0 || inner == 0 || n_samples == 0) return PMB_OK;
    PMB_EW_PROLOGUE(n_samples, sample_times, original_times, positions, out, idx_workspace, weight_workspace);
    pmb::interp_coeff_kernel<<<grid_, 256, 0, st_>>>(sample_times, original_times, n_samples, n_original, idx_workspace,
                                                     weight_workspace);
    PMB_CUDA(cudaGetLastError());
    pmb::interp_apply_kernel<<<ew_grid(outer * n_samples * inner, 256, dp_), 256, 0, st_>>>(
        positions, idx_workspace, weight_workspace, out, outer, n_original, n_samples, inner);
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}
int pmb_vec_normalize_f32(const float *v, float eps, float *out, int64_t n, int32_t k, void *stream) {
    if (k < 1) return fail(PMB_ERR_SHAPE, "%s: k < 1", __func__);
    PMB_EW_PROLOGUE(n, v, out);
    if (k == 3 && aligned16(v) && aligned16(out) && n >= 4 && env_int("PMB_VEC3_X4", 1)) {
        const int64_t n4 = n / 4;
        pmb::vec3_normalize_x4_kernel<<<ew_grid(n4, 256, dp_), 256, 0, st_>>>((const float4 *)v, eps, (float4 *)out, n4);
        PMB_CUDA(cudaGetLastError());
        if (n % 4) pmb::vec_normalize_kernel<<<1, 32, 0, st_>>>(v + 12 * n4, eps, out + 12 * n4, n % 4, 3);
    } else {
        pmb::vec_normalize_kernel<<<grid_, 256, 0, st_>>>(v, eps, out, n, k);
    }
    PMB_CUDA(cudaGetLastError());
    return PMB_OK;
}

}  // extern "C"
