// Host-side plumbing shared by the translation units of libpymotion_b200.so (api_*.cu): status / error
// reporting, the experiment-knob snapshot, per-device properties, the launch caches (shared-memory attribute +
// occupancy per kernel, joint programs per topology, TMA descriptors) and small launch helpers.
// Definitions live in api_core.cu.  Nothing here computes on the CPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/pymotion_b200.h"
#include "common.cuh"
#include "track_schedule.h"

namespace pmbh {

// ---- status ---------------------------------------------------------------------------
int fail(int status, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void note_variant(const char *fmt, ...);  // which kernel variant the last launch on this thread picked

#define PMB_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return pmbh::cuda_fail(e_, #call); \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

// ---- experiment knobs -------------------------------------------------------------------
// Kernel-variant knobs exist for the sweeps and the forced-variant tests only.  They are read from the
// environment ONCE, and only when PMB_EXPERIMENT=1 is set: in production no PMB_* variable can change which
// kernel runs and a launch never calls getenv.  pmb_reload_knobs() re-reads them (tests / sweeps switch variants
// inside one process).
#define PMB_KNOB_LIST(X)                                                                                          \
    X(FK_ROWS) X(FK_TRACKS) X(FK_MTRACKS) X(FK_STAGES) X(FK_FR) X(FK_WARPS) X(FK_GROUP) X(FK_BLOCKS_PER_SM) X(FK_NB) \
    X(FK_U) X(FK_UL) X(FK_WARPS_PER_SM) X(FK_L2_PREFETCH) X(TMA_L2PROMO) X(FKQ_GROUP) X(FKQ_BLOCKS_PER_SM)         \
    X(FKQ_MATRIX) X(DQ_GROUP) X(DQ_BLOCKS_PER_SM) X(FRDQ_ELEMS) X(FRP_BLOCKS_PER_SM) X(UNROLL_CHUNK_APPLY)        \
    X(VEC3_X4) X(HOST_CHUNK_MB) X(HOST_THREADS) X(DQ_TRACKS) X(FKQ_TRACKS) X(QT_WARPS_PER_SM) X(QT_PIPE) X(QT_DYNAMIC) X(QT_SHAPE) X(FRP_FAST) X(FRP_WIDE) X(MIRROR_FUSED)
enum Knob {
#define X(name) K_##name,
    PMB_KNOB_LIST(X)
#undef X
    K_COUNT
};
struct KnobTable {
    bool present[K_COUNT] = {};
    int value[K_COUNT] = {};
};
const KnobTable &knobs();
void reload_knobs();
inline int knob(Knob k, int fallback) {
    const KnobTable &t = knobs();
    return t.present[k] ? t.value[k] : fallback;
}
inline bool knob_set(Knob k) { return knobs().present[k]; }

// ---- device -----------------------------------------------------------------------------
struct DeviceProps {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;   // largest dynamic shared memory of one block
    int smem_sm = 0;      // shared memory of one SM
    int cc_major = 0, cc_minor = 0;
    bool ok = false;
};
int device_props(DeviceProps &out);  // of the current device (cached)

// Resident blocks per SM of `kernel` launched with (threads, smem); the first use per (kernel, device) raises the
// kernel's dynamic shared-memory limit.  Cached: a launch costs one hash lookup, not two driver calls.
int kernel_fit_impl(const void *kernel, int device, int threads, int smem, int &per_sm);
template <typename K>
int kernel_fit(K kernel, const DeviceProps &dp, int threads, int smem, int &per_sm) {
    return kernel_fit_impl(reinterpret_cast<const void *>(kernel), dp.device, threads, smem, per_sm);
}

// ---- per-topology programs (cached by the bytes of parents[]) -----------------------------------
// The returned pointers stay valid until the calling thread's next lookup of the same kind.
int joint_program(const int64_t *parents_host, int32_t n_joints, bool detach_root_children, const pmb::JointProgram *&prog,
                  int &n_slots);
// window: joints are scheduled in windows of that many consecutive joints (0 = the whole skeleton at once)
int track_program(const int64_t *parents_host, int32_t n_joints, int n_tracks, int window, const pmb::TrackProgram *&prog,
                  int &n_steps);

// ---- TMA descriptor of the quaternion input: rot viewed as [n_frames][4 * n_joints] floats, box = box_frames x
// `chunk` joints, hardware swizzle matched to the box row.  The last descriptor of a thread is cached.
int make_rot_map(CUtensorMap &tm, const float *rot, int64_t n_frames, int32_t n_joints, int chunk, int box_frames = 32);

// ---- small launch helpers ---------------------------------------------------------------------
inline int ew_grid(int64_t n, int threads, const DeviceProps &dp) {
    const int64_t want = (n + threads - 1) / threads;
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(dp.sm_count) * 16)));
}
// divisor of div_small (dq_kernels.cuh): floor(2^32 / d) + 1, and 0 for d = 1 (which has no 32-bit magic)
inline uint32_t magic_small(int d) { return d <= 1 ? 0u : static_cast<uint32_t>((1ULL << 32) / static_cast<uint32_t>(d)) + 1u; }
inline int tile_frames(int n_joints, int cap_elems) {
    const int fb = (cap_elems / n_joints) & ~3;
    return std::max(4, std::min(64, fb));
}

}  // namespace pmbh

// ---- prologue of the element-wise entry points: NULL / size checks, device properties, grid, stream -------------
#define PMB_EW_PROLOGUE(n, ...)                                                                      \
    const void *ptrs_[] = {__VA_ARGS__};                                                             \
    for (const void *p_ : ptrs_)                                                                     \
        if (!p_) return pmbh::fail(PMB_ERR_NULL, "%s: NULL array pointer", __func__);                      \
    if ((n) < 0) return pmbh::fail(PMB_ERR_SHAPE, "%s: n < 0", __func__);                                  \
    if ((n) == 0) return PMB_OK;                                                                     \
    pmbh::DeviceProps dp_;                                                                                 \
    {                                                                                                \
        int rc_ = pmbh::device_props(dp_);                                                                 \
        if (rc_) return rc_;                                                                         \
    }                                                                                                \
    const int grid_ = pmbh::ew_grid((n), 256, dp_);                                                        \
    cudaStream_t st_ = static_cast<cudaStream_t>(stream)

#define PMB_NEED16(p) \
    if (!pmbh::aligned16(p)) return pmbh::fail(PMB_ERR_ALIGN, "%s: " #p " must be 16-byte aligned", __func__)
