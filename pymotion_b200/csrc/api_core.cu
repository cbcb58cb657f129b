// libpymotion_b200.so, core translation unit: error reporting, the knob snapshot, device properties, the launch
// caches declared in host_common.h and the diagnostic entry points of include/pymotion_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "host_common.h"
#include "joint_program.h"

namespace pmbh {

namespace {
thread_local char g_err[512] = "";
thread_local char g_variant[160] = "";
}  // namespace

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

// ---- knobs ------------------------------------------------------------------------------
namespace {
KnobTable g_knobs;
std::once_flag g_knobs_once;
void load_knobs() {
    static const char *names[K_COUNT] = {
#define X(name) "PMB_" #name,
        PMB_KNOB_LIST(X)
#undef X
    };
    KnobTable t;
    const char *experiment = getenv("PMB_EXPERIMENT");
    if (experiment && atoi(experiment) == 1)
        for (int k = 0; k < K_COUNT; ++k)
            if (const char *v = getenv(names[k])) t.present[k] = true, t.value[k] = atoi(v);
    g_knobs = t;
}
}  // namespace
const KnobTable &knobs() {
    std::call_once(g_knobs_once, load_knobs);
    return g_knobs;
}
void reload_knobs() {
    knobs();
    load_knobs();
}

// ---- device -----------------------------------------------------------------------------
int device_props(DeviceProps &out) {
    static std::mutex mu;
    static DeviceProps cache[64];
    int dev = 0;
    PMB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(PMB_ERR_CUDA, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    DeviceProps &p = cache[dev];
    if (!p.ok) {
        p.device = dev;
        PMB_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        PMB_CUDA(cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        p.ok = true;
    }
    out = p;
    return PMB_OK;
}

// ---- (kernel, device, threads, smem) -> resident blocks per SM -------------------------------
namespace {
struct FitKey {
    const void *kernel;
    int device, threads, smem;
    bool operator==(const FitKey &o) const { return kernel == o.kernel && device == o.device && threads == o.threads && smem == o.smem; }
};
struct FitHash {
    size_t operator()(const FitKey &k) const {
        size_t h = reinterpret_cast<size_t>(k.kernel);
        h = h * 1000003u ^ static_cast<size_t>(k.device);
        h = h * 1000003u ^ static_cast<size_t>(k.threads);
        return h * 1000003u ^ static_cast<size_t>(k.smem);
    }
};
std::mutex g_fit_mu;
std::unordered_map<FitKey, int, FitHash> g_fit;
std::unordered_map<FitKey, int, FitHash> g_smem_limit;  // (kernel, device, 0, 0) -> dynamic shared memory limit set so far
}  // namespace

int kernel_fit_impl(const void *kernel, int device, int threads, int smem, int &per_sm) {
    const FitKey key{kernel, device, threads, smem};
    std::lock_guard<std::mutex> lock(g_fit_mu);
    auto it = g_fit.find(key);
    if (it != g_fit.end()) {
        per_sm = it->second;
        return PMB_OK;
    }
    if (smem > 48 * 1024) {
        int &limit = g_smem_limit[FitKey{kernel, device, 0, 0}];
        if (smem > limit) {
            PMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            limit = smem;
        }
    }
    int n = 0;
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem));
    g_fit.emplace(key, n);
    per_sm = n;
    return PMB_OK;
}

// ---- per-topology programs --------------------------------------------------------------------
namespace {
template <typename P>
struct ProgEntry {
    int n_joints = 0, variant = -1, extra = 0;  // variant: detach flag / track count; extra: n_slots / n_steps
    uint64_t stamp = 0;
    std::vector<int64_t> parents;
    P prog;
};
template <typename P>
struct ProgCache {
    static constexpr int N = 8;
    ProgEntry<P> e[N];
    uint64_t clock = 0;
    ProgEntry<P> *find(const int64_t *parents, int n_joints, int variant) {
        for (auto &x : e)
            if (x.n_joints == n_joints && x.variant == variant &&
                memcmp(x.parents.data(), parents, sizeof(int64_t) * static_cast<size_t>(n_joints)) == 0) {
                x.stamp = ++clock;
                return &x;
            }
        return nullptr;
    }
    ProgEntry<P> *victim() {
        ProgEntry<P> *v = &e[0];
        for (auto &x : e)
            if (x.stamp < v->stamp) v = &x;
        v->stamp = ++clock;
        return v;
    }
};
}  // namespace

int joint_program(const int64_t *parents_host, int32_t n_joints, bool detach, const pmb::JointProgram *&prog, int &n_slots) {
    if (!parents_host) return fail(PMB_ERR_NULL, "parents_host is NULL");
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS)
        return fail(PMB_ERR_SHAPE, "n_joints = %d outside [1, %d]", n_joints, PMB_MAX_JOINTS);
    thread_local ProgCache<pmb::JointProgram> cache;
    if (auto *hit = cache.find(parents_host, n_joints, detach ? 1 : 0)) {
        prog = &hit->prog, n_slots = hit->extra;
        return PMB_OK;
    }
    pmb::JointProgram built;
    const pmb::ProgramInfo info = pmb::build_joint_program(parents_host, n_joints, detach, built.code);
    if (info.status == PMB_ERR_TOPOLOGY)
        return fail(PMB_ERR_TOPOLOGY, "parents[%d] = %lld is not in [0, %d): joints must come after their parent (BVH order)",
                    info.bad_joint, static_cast<long long>(parents_host[info.bad_joint]), info.bad_joint);
    if (info.status != PMB_OK) return fail(info.status, "cannot build the joint program");
    auto *slot = cache.victim();
    slot->n_joints = n_joints, slot->variant = detach ? 1 : 0, slot->extra = info.n_slots;
    slot->parents.assign(parents_host, parents_host + n_joints);
    slot->prog = built;
    prog = &slot->prog, n_slots = info.n_slots;
    return PMB_OK;
}

int track_program(const int64_t *parents_host, int32_t n_joints, int n_tracks, int window, const pmb::TrackProgram *&prog,
                  int &n_steps) {
    if (!parents_host) return fail(PMB_ERR_NULL, "parents_host is NULL");
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS)
        return fail(PMB_ERR_SHAPE, "n_joints = %d outside [1, %d]", n_joints, PMB_MAX_JOINTS);
    thread_local ProgCache<pmb::TrackProgram> cache;
    const int variant = n_tracks | (window << 8);
    if (auto *hit = cache.find(parents_host, n_joints, variant)) {
        prog = &hit->prog, n_steps = hit->extra;
        return PMB_OK;
    }
    auto *slot = cache.victim();
    slot->n_joints = 0, slot->variant = -1;  // invalid while it is being rebuilt
    int bad = -1;
    const int steps = pmb::build_track_schedule(parents_host, n_joints, n_tracks, slot->prog.code, &bad, window, slot->prog.chunk_first);
    if (steps < 0)
        return fail(PMB_ERR_TOPOLOGY, "parents[%d] = %lld is not in [0, %d): joints must come after their parent (BVH order)", bad,
                    bad >= 0 ? static_cast<long long>(parents_host[bad]) : -1LL, bad);
    slot->parents.assign(parents_host, parents_host + n_joints);
    slot->n_joints = n_joints, slot->variant = variant, slot->extra = steps;  // steps == 0: does not fit kTrackCap (cached too)
    prog = &slot->prog, n_steps = steps;
    return PMB_OK;
}

// ---- TMA descriptor ---------------------------------------------------------------------------
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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
struct MapKey {
    const float *rot = nullptr;
    int64_t n_frames = -1;
    int n_joints = 0, chunk = 0, box_frames = 0, promo = 0;
    bool operator==(const MapKey &o) const {
        return rot == o.rot && n_frames == o.n_frames && n_joints == o.n_joints && chunk == o.chunk && box_frames == o.box_frames &&
               promo == o.promo;
    }
};
}  // namespace

int make_rot_map(CUtensorMap &tm, const float *rot, int64_t n_frames, int32_t n_joints, int chunk, int box_frames) {
    // L2 promotion: the granule L2 fetches from DRAM for a box row.  Measured (profiles/r1_sweep_l2promo.jsonl): 256-byte
    // granules are worth +0.6 % at 22 joints, +4 % at 52, +5 % at 65 (a frame's quaternion row spans 1.4 .. 4 granules and
    // the next chunk of the same frames finds them in L2).  Knob PMB_TMA_L2PROMO = 0 none, 1 = 64 B, 2 = 128 B, 3 = 256 B.
    const int pv = knob(K_TMA_L2PROMO, 3);
    thread_local MapKey last_key;
    thread_local CUtensorMap last_map;
    const MapKey key{rot, n_frames, n_joints, chunk, box_frames, pv};
    if (key == last_key) {
        tm = last_map;
        return PMB_OK;
    }
    EncodeTiledFn enc;
    int rc = encode_fn(enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(4) * n_joints, static_cast<cuuint64_t>(n_frames)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(16) * n_joints};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(4 * chunk), static_cast<cuuint32_t>(box_frames)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = chunk == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUtensorMapL2promotion promo = pv == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                         : pv == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : pv == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                   : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(rot), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PMB_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    last_key = key, last_map = tm;
    return PMB_OK;
}

}  // namespace pmbh

using namespace pmbh;

extern "C" {

int pmb_version(void) { return PMB_VERSION; }
const char *pmb_last_error(void) { return g_err; }
const char *pmb_last_variant(void) { return g_variant; }
void pmb_reload_knobs(void) { reload_knobs(); }

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
        cudaDeviceProp prop;
        PMB_CUDA(cudaGetDeviceProperties(&prop, dp.device));
        snprintf(name, static_cast<size_t>(name_len), "%s", prop.name);
    }
    return PMB_OK;
}

int pmb_build_joint_program(const int64_t *parents_host, int32_t n_joints, uint32_t *codes_out) {
    if (!codes_out) return fail(PMB_ERR_NULL, "codes_out is NULL");
    const pmb::JointProgram *prog = nullptr;
    int n_slots = 0;
    int rc = joint_program(parents_host, n_joints, false, prog, n_slots);
    if (rc) return rc;
    memcpy(codes_out, prog->code, sizeof(uint32_t) * static_cast<size_t>(n_joints));
    return n_slots;
}

int pmb_build_track_schedule(const int64_t *parents_host, int32_t n_joints, int32_t n_tracks, int32_t window, uint32_t *codes_out,
                             int32_t codes_capacity, int32_t *window_first_out) {
    if (!codes_out) return fail(PMB_ERR_NULL, "codes_out is NULL");
    if (n_tracks < 1 || n_tracks > 8) return fail(PMB_ERR_SHAPE, "n_tracks = %d outside [1, 8]", n_tracks);
    if (window != 0 && window != 8) return fail(PMB_ERR_SHAPE, "window = %d: 0 (whole skeleton) or 8 (one box of quaternions)", window);
    const pmb::TrackProgram *prog = nullptr;
    int n_steps = 0;
    int rc = track_program(parents_host, n_joints, n_tracks, window, prog, n_steps);
    if (rc) return rc;
    if (n_steps == 0) return fail(PMB_ERR_SHAPE, "schedule of %d tracks does not fit %d items", n_tracks, pmb::kTrackCap);
    if (n_steps * n_tracks > codes_capacity) return fail(PMB_ERR_SHAPE, "codes_out holds %d items, the schedule has %d", codes_capacity, n_steps * n_tracks);
    memcpy(codes_out, prog->code, sizeof(uint32_t) * static_cast<size_t>(n_steps) * n_tracks);
    if (window_first_out) {
        const int n_windows = window ? (n_joints + window - 1) / window : 1;
        for (int w = 0; w <= n_windows; ++w) window_first_out[w] = prog->chunk_first[w];
    }
    return n_steps;
}

}  // extern "C"
