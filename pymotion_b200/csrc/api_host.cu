// libpymotion_b200.so, host-buffer pipeline: fk / fk_quat on arrays that live in HOST memory (the reference's own
// calling convention: NumPy arrays in, NumPy arrays out -- ops/skeleton.py:16).
//
// The frame axis is cut into chunks; chunk k is copied in, computed and copied out on stream k mod 2, so the
// H2D copy of one chunk, the kernel of another and the D2H copy of a third overlap (PCIe is full duplex and the
// B200 has separate copy engines per direction).  Page-locked caller buffers are DMA'd directly.  Pageable buffers
// (plain NumPy arrays) go through a page-locked staging ring owned by the workspace: a few host threads copy
// user memory <-> ring while the GPU works on the neighbouring chunk, instead of letting cudaMemcpy bounce every
// pageable transfer through the driver's single staging buffer synchronously.
//
// Workspaces are per device and per concurrent caller (a small pool, no lock held while a call runs); chunks are
// sized in bytes, not frames.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "host_common.h"

using namespace pmbh;

namespace {

// ---- a few helper threads for user memory <-> staging ring copies -------------------------------------
class CopyPool {
public:
    explicit CopyPool(int n_threads) {
        for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    // copies [src, src + bytes) to dst in slices; the caller works too and returns when every slice is done
    void copy(void *dst, const void *src, size_t bytes) {
        const size_t slice = 1u << 20;
        if (bytes <= 2 * slice || workers_.empty()) {
            memcpy(dst, src, bytes);
            return;
        }
        std::unique_lock<std::mutex> lock(mu_);
        dst_ = static_cast<char *>(dst), src_ = static_cast<const char *>(src), bytes_ = bytes, slice_ = slice;
        next_.store(0);
        n_slices_ = (bytes + slice - 1) / slice;
        pending_ = n_slices_;
        ++generation_;
        lock.unlock();
        cv_.notify_all();
        work();
        lock.lock();
        done_cv_.wait(lock, [this] { return pending_ == 0; });
    }

private:
    void work() {
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= n_slices_) return;
            const size_t off = i * slice_, n = std::min(slice_, bytes_ - off);
            memcpy(dst_ + off, src_ + off, n);
            std::lock_guard<std::mutex> lock(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
    uint64_t generation_ = 0;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, slice_ = 0, n_slices_ = 0, pending_ = 0;
    std::atomic<size_t> next_{0};
};

struct HostPipe {
    int device = -1;
    bool busy = false;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr}, offsets_ready = nullptr;
    char *dev_in[2] = {nullptr, nullptr}, *dev_out[2] = {nullptr, nullptr};
    char *pin_in[2] = {nullptr, nullptr}, *pin_out[2] = {nullptr, nullptr};
    size_t dev_in_cap = 0, dev_out_cap = 0, pin_in_cap = 0, pin_out_cap = 0;
    float *offsets = nullptr;
    std::unique_ptr<CopyPool> pool;

    int grow_dev(size_t in_bytes, size_t out_bytes) {
        if (in_bytes > dev_in_cap) {
            for (int s = 0; s < 2; ++s) {
                if (dev_in[s]) cudaFree(dev_in[s]);
                dev_in[s] = nullptr;
                PMB_CUDA(cudaMalloc(&dev_in[s], in_bytes));
            }
            dev_in_cap = in_bytes;
        }
        if (out_bytes > dev_out_cap) {
            for (int s = 0; s < 2; ++s) {
                if (dev_out[s]) cudaFree(dev_out[s]);
                dev_out[s] = nullptr;
                PMB_CUDA(cudaMalloc(&dev_out[s], out_bytes));
            }
            dev_out_cap = out_bytes;
        }
        return PMB_OK;
    }
    int grow_pin(char *(&buf)[2], size_t &cap, size_t bytes) {
        if (bytes > cap) {
            for (int s = 0; s < 2; ++s) {
                if (buf[s]) cudaFreeHost(buf[s]);
                buf[s] = nullptr;
                PMB_CUDA(cudaHostAlloc(&buf[s], bytes, cudaHostAllocDefault));
            }
            cap = bytes;
        }
        return PMB_OK;
    }
    int init(int dev) {
        device = dev;
        for (int s = 0; s < 2; ++s) {
            PMB_CUDA(cudaStreamCreateWithFlags(&stream[s], cudaStreamNonBlocking));
            PMB_CUDA(cudaEventCreateWithFlags(&h2d_done[s], cudaEventDisableTiming));
            PMB_CUDA(cudaEventCreateWithFlags(&d2h_done[s], cudaEventDisableTiming));
        }
        PMB_CUDA(cudaEventCreateWithFlags(&offsets_ready, cudaEventDisableTiming));
        PMB_CUDA(cudaMalloc(&offsets, static_cast<size_t>(PMB_MAX_JOINTS) * 12));
        return PMB_OK;
    }
    void release() {
        for (int s = 0; s < 2; ++s) {
            if (dev_in[s]) cudaFree(dev_in[s]);
            if (dev_out[s]) cudaFree(dev_out[s]);
            if (pin_in[s]) cudaFreeHost(pin_in[s]);
            if (pin_out[s]) cudaFreeHost(pin_out[s]);
            if (stream[s]) cudaStreamDestroy(stream[s]);
            if (h2d_done[s]) cudaEventDestroy(h2d_done[s]);
            if (d2h_done[s]) cudaEventDestroy(d2h_done[s]);
        }
        if (offsets_ready) cudaEventDestroy(offsets_ready);
        if (offsets) cudaFree(offsets);
        pool.reset();
    }
};

std::mutex g_pipes_mu;
std::vector<std::unique_ptr<HostPipe>> g_pipes;

HostPipe *acquire_pipe(int device, int &rc) {
    std::lock_guard<std::mutex> lock(g_pipes_mu);
    for (auto &p : g_pipes)
        if (!p->busy && p->device == device) {
            p->busy = true;
            return p.get();
        }
    std::unique_ptr<HostPipe> p(new HostPipe);
    if ((rc = p->init(device))) {
        p->release();
        return nullptr;
    }
    p->busy = true;
    g_pipes.push_back(std::move(p));
    return g_pipes.back().get();
}
struct PipeLease {  // gives the workspace back, and never lets a call return with copies still in flight
    HostPipe *p;
    ~PipeLease() {
        if (!p) return;
        cudaStreamSynchronize(p->stream[0]);
        cudaStreamSynchronize(p->stream[1]);
        std::lock_guard<std::mutex> lock(g_pipes_mu);
        p->busy = false;
    }
};

bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// One pipelined pass of a per-frame op over HOST arrays: up to two per-frame inputs (in_a, in_b bytes per frame) and up to two
// per-frame outputs (out_a, out_b bytes per frame); `launch(d_in_a, d_in_b, d_out_a, d_out_b, d_offsets, n, stream)` enqueues the
// op for n frames of device-resident staging.  offsets_host (n_joints x 3 floats, or NULL) is uploaded once.
struct HostOp {
    size_t in_a, in_b, out_a, out_b;
};
template <class Launch>
int host_pipeline(const HostOp &op, const char *in_a_host, const char *in_b_host, const float *offsets_host, int32_t n_joints,
                  int64_t n_frames, char *out_a_host, char *out_b_host, int64_t chunk_frames, Launch launch) {
    const size_t J = static_cast<size_t>(n_joints);
    const size_t in_frame = op.in_a + op.in_b, out_frame = op.out_a + op.out_b;  // bytes per frame, in / out
    if (chunk_frames <= 0) {
        // chunks are sized in bytes: ~48 MB of traffic (in + out) each, so that the pipeline fills after a few per cent
        // of a large batch whatever the joint count, and the device staging stays ~100 MB
        const size_t target = static_cast<size_t>(knob(K_HOST_CHUNK_MB, 48)) << 20;
        chunk_frames = static_cast<int64_t>(std::max<size_t>(1024, target / (in_frame + out_frame)));
    }
    chunk_frames = std::min<int64_t>((chunk_frames + 31) & ~31LL, (n_frames + 31) & ~31LL);
    const size_t F = static_cast<size_t>(chunk_frames);

    int dev = 0, rc = PMB_OK;
    PMB_CUDA(cudaGetDevice(&dev));
    PipeLease lease{acquire_pipe(dev, rc)};
    if (!lease.p) return rc;
    HostPipe &w = *lease.p;
    // device staging of one chunk: [in_a | in_b] and [out_a | out_b] (sub-buffers 256-byte aligned)
    const size_t ia_bytes = (F * op.in_a + 255) & ~size_t(255), oa_bytes = (F * op.out_a + 255) & ~size_t(255);
    if ((rc = w.grow_dev(ia_bytes + F * op.in_b + 256, oa_bytes + F * op.out_b + 256))) return rc;
    const bool in_pinned = is_pinned(in_a_host) && (!op.in_b || is_pinned(in_b_host));
    const bool out_pinned = is_pinned(out_a_host) && (!op.out_b || is_pinned(out_b_host));
    if (!in_pinned && (rc = w.grow_pin(w.pin_in, w.pin_in_cap, ia_bytes + F * op.in_b + 256))) return rc;
    if (!out_pinned && (rc = w.grow_pin(w.pin_out, w.pin_out_cap, oa_bytes + F * op.out_b + 256))) return rc;
    if ((!in_pinned || !out_pinned) && !w.pool) {
        const int hw = static_cast<int>(std::thread::hardware_concurrency());
        const int n = std::max(1, std::min(knob(K_HOST_THREADS, 4), hw > 0 ? hw : 1));
        w.pool.reset(new CopyPool(n - 1));
    }

    if (offsets_host) {
        PMB_CUDA(cudaMemcpyAsync(w.offsets, offsets_host, J * 12, cudaMemcpyHostToDevice, w.stream[0]));
        PMB_CUDA(cudaEventRecord(w.offsets_ready, w.stream[0]));
        PMB_CUDA(cudaStreamWaitEvent(w.stream[1], w.offsets_ready, 0));
        PMB_CUDA(cudaStreamSynchronize(w.stream[0]));  // offsets_host may be a temporary of the caller
    }

    struct Pending {  // a chunk whose results sit in the output ring and still have to reach the caller's memory
        bool live = false;
        size_t f0 = 0, n = 0;
    } pending[2];
    auto copy_out = [&](int slot) -> int {
        Pending &pd = pending[slot];
        if (!pd.live) return PMB_OK;
        PMB_CUDA(cudaEventSynchronize(w.d2h_done[slot]));
        w.pool->copy(out_a_host + pd.f0 * op.out_a, w.pin_out[slot], pd.n * op.out_a);
        if (op.out_b) w.pool->copy(out_b_host + pd.f0 * op.out_b, w.pin_out[slot] + oa_bytes, pd.n * op.out_b);
        pd.live = false;
        return PMB_OK;
    };

    int slot = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk_frames, slot ^= 1) {
        const size_t n = static_cast<size_t>(std::min<int64_t>(chunk_frames, n_frames - f0));
        cudaStream_t st = w.stream[slot];  // stream order makes the slot's device buffers safe to reuse
        char *d_ia = w.dev_in[slot], *d_ib = w.dev_in[slot] + ia_bytes;
        char *d_oa = w.dev_out[slot], *d_ob = w.dev_out[slot] + oa_bytes;
        if (in_pinned) {
            PMB_CUDA(cudaMemcpyAsync(d_ia, in_a_host + f0 * op.in_a, n * op.in_a, cudaMemcpyHostToDevice, st));
            if (op.in_b) PMB_CUDA(cudaMemcpyAsync(d_ib, in_b_host + f0 * op.in_b, n * op.in_b, cudaMemcpyHostToDevice, st));
        } else {
            PMB_CUDA(cudaEventSynchronize(w.h2d_done[slot]));  // the ring slot's previous chunk has left for the device
            w.pool->copy(w.pin_in[slot], in_a_host + f0 * op.in_a, n * op.in_a);
            if (op.in_b) memcpy(w.pin_in[slot] + ia_bytes, in_b_host + f0 * op.in_b, n * op.in_b);
            PMB_CUDA(cudaMemcpyAsync(d_ia, w.pin_in[slot], n * op.in_a, cudaMemcpyHostToDevice, st));
            if (op.in_b) PMB_CUDA(cudaMemcpyAsync(d_ib, w.pin_in[slot] + ia_bytes, n * op.in_b, cudaMemcpyHostToDevice, st));
            PMB_CUDA(cudaEventRecord(w.h2d_done[slot], st));
        }
        if ((rc = launch(d_ia, d_ib, d_oa, d_ob, w.offsets, static_cast<int64_t>(n), st))) return rc;  // (the lease drains both streams)
        if (out_pinned) {
            PMB_CUDA(cudaMemcpyAsync(out_a_host + f0 * op.out_a, d_oa, n * op.out_a, cudaMemcpyDeviceToHost, st));
            if (op.out_b) PMB_CUDA(cudaMemcpyAsync(out_b_host + f0 * op.out_b, d_ob, n * op.out_b, cudaMemcpyDeviceToHost, st));
        } else {
            if ((rc = copy_out(slot))) return rc;  // the ring slot must be empty before the GPU refills it (normally done below)
            PMB_CUDA(cudaMemcpyAsync(w.pin_out[slot], d_oa, n * op.out_a, cudaMemcpyDeviceToHost, st));
            if (op.out_b) PMB_CUDA(cudaMemcpyAsync(w.pin_out[slot] + oa_bytes, d_ob, n * op.out_b, cudaMemcpyDeviceToHost, st));
            PMB_CUDA(cudaEventRecord(w.d2h_done[slot], st));
            pending[slot] = {true, static_cast<size_t>(f0), n};
            if ((rc = copy_out(slot ^ 1))) return rc;  // the previous chunk's results, while the GPU works on this one
        }
    }
    PMB_CUDA(cudaStreamSynchronize(w.stream[0]));
    PMB_CUDA(cudaStreamSynchronize(w.stream[1]));
    if (!out_pinned)
        for (int s = 0; s < 2; ++s)
            if ((rc = copy_out(s))) return rc;
    return PMB_OK;
}

int validate_host_call(const char *what, const int64_t *parents_host, int64_t n_frames, int32_t n_joints, bool detach) {
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "%s: n_frames < 0", what);
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS) return fail(PMB_ERR_SHAPE, "%s: n_joints = %d out of range", what, n_joints);
    const pmb::JointProgram *prog = nullptr;  // validate the topology before any allocation or copy
    int n_slots = 0;
    return joint_program(parents_host, n_joints, detach, prog, n_slots);
}

// kind 0: fk (rotation matrices out, 36 J bytes per frame); kind 1: fk_quat (global quaternions, 16 J)
int fk_host_common(int kind, const float *rot_host, const float *gpos_host, const float *offsets_host, const int64_t *parents_host,
                   int64_t n_frames, int32_t n_joints, float *pos_host, float *rout_host, int64_t chunk_frames) {
    if (!rot_host || !gpos_host || !offsets_host || !pos_host || !rout_host) return fail(PMB_ERR_NULL, "fk_host: NULL array pointer");
    int rc = validate_host_call("fk_host", parents_host, n_frames, n_joints, false);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    const size_t J = static_cast<size_t>(n_joints);
    const HostOp op{J * 16, 12, J * 12, J * (kind == 0 ? 36u : 16u)};
    return host_pipeline(op, reinterpret_cast<const char *>(rot_host), reinterpret_cast<const char *>(gpos_host), offsets_host, n_joints,
                         n_frames, reinterpret_cast<char *>(pos_host), reinterpret_cast<char *>(rout_host), chunk_frames,
                         [&](char *d_rot, char *d_gpos, char *d_pos, char *d_rout, float *d_off, int64_t n, cudaStream_t st) {
                             auto fn = kind == 0 ? pmb_fk_f32 : pmb_fk_quat_f32;
                             return fn(reinterpret_cast<float *>(d_rot), reinterpret_cast<float *>(d_gpos), 3, d_off, 0, parents_host, n, n_joints,
                                       reinterpret_cast<float *>(d_pos), reinterpret_cast<float *>(d_rout), st);
                         });
}

}  // namespace

extern "C" {

void pmb_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_pipes_mu);
    for (auto it = g_pipes.begin(); it != g_pipes.end();) {
        if (!(*it)->busy) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice((*it)->device);
            (*it)->release();
            cudaSetDevice(cur);
            it = g_pipes.erase(it);
        } else {
            ++it;
        }
    }
}

int pmb_fk_f32_host(const float *rot_host, const float *global_pos_host, const float *offsets_host, const int64_t *parents_host,
                    int64_t n_frames, int32_t n_joints, float *positions_host, float *rotmats_host, int64_t chunk_frames) {
    return fk_host_common(0, rot_host, global_pos_host, offsets_host, parents_host, n_frames, n_joints, positions_host, rotmats_host,
                          chunk_frames);
}

int pmb_fk_quat_f32_host(const float *rot_host, const float *global_pos_host, const float *offsets_host, const int64_t *parents_host,
                         int64_t n_frames, int32_t n_joints, float *positions_host, float *global_rots_host, int64_t chunk_frames) {
    return fk_host_common(1, rot_host, global_pos_host, offsets_host, parents_host, n_frames, n_joints, positions_host,
                          global_rots_host, chunk_frames);
}

// to_root_dual_quat / from_root_dual_quat on HOST arrays (ops/skeleton.py:207, :173 with the reference's own calling convention,
// NumPy in / NumPy out): the same chunked, two-stream pipeline.
int pmb_to_root_dual_quat_f32_host(const float *rotations_host, const float *global_pos_host, const int64_t *parents_host,
                                   const float *offsets_host, int64_t n_frames, int32_t n_joints, float *dq_host, int64_t chunk_frames) {
    if (!rotations_host || !global_pos_host || !offsets_host || !dq_host) return fail(PMB_ERR_NULL, "to_root_dual_quat_host: NULL array pointer");
    if (offsets_host[0] != 0.f || offsets_host[1] != 0.f || offsets_host[2] != 0.f)
        return fail(PMB_ERR_ROOT_OFFSET, "offsets[0] must be zero (ops/skeleton.py:227)");
    int rc = validate_host_call("to_root_dual_quat_host", parents_host, n_frames, n_joints, true);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    const size_t J = static_cast<size_t>(n_joints);
    const HostOp op{J * 16, 12, J * 32, 0};
    return host_pipeline(op, reinterpret_cast<const char *>(rotations_host), reinterpret_cast<const char *>(global_pos_host), offsets_host,
                         n_joints, n_frames, reinterpret_cast<char *>(dq_host), nullptr, chunk_frames,
                         [&](char *d_rot, char *d_gpos, char *d_dq, char *, float *d_off, int64_t n, cudaStream_t st) {
                             return pmb_to_root_dual_quat_f32(reinterpret_cast<float *>(d_rot), reinterpret_cast<float *>(d_gpos), 3, parents_host,
                                                              d_off, offsets_host, n, n_joints, reinterpret_cast<float *>(d_dq), st);
                         });
}

int pmb_from_root_dual_quat_f32_host(const float *dq_host, const int64_t *parents_host, int64_t n_frames, int32_t n_joints,
                                     float *translations_host, float *rotations_host, int64_t chunk_frames) {
    if (!dq_host || !translations_host || !rotations_host) return fail(PMB_ERR_NULL, "from_root_dual_quat_host: NULL array pointer");
    int rc = validate_host_call("from_root_dual_quat_host", parents_host, n_frames, n_joints, true);
    if (rc) return rc;
    if (n_frames == 0) return PMB_OK;
    const size_t J = static_cast<size_t>(n_joints);
    const HostOp op{J * 32, 0, J * 12, J * 16};
    return host_pipeline(op, reinterpret_cast<const char *>(dq_host), nullptr, nullptr, n_joints, n_frames,
                         reinterpret_cast<char *>(translations_host), reinterpret_cast<char *>(rotations_host), chunk_frames,
                         [&](char *d_dq, char *, char *d_t, char *d_r, float *, int64_t n, cudaStream_t st) {
                             return pmb_from_root_dual_quat_f32(reinterpret_cast<float *>(d_dq), parents_host, n, n_joints,
                                                                reinterpret_cast<float *>(d_t), reinterpret_cast<float *>(d_r), st);
                         });
}

}  // extern "C"
