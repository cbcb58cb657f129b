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

// kind 0: fk (rotation matrices out, 36 J bytes per frame); kind 1: fk_quat (global quaternions, 16 J)
int fk_host_common(int kind, const float *rot_host, const float *gpos_host, const float *offsets_host, const int64_t *parents_host,
                   int64_t n_frames, int32_t n_joints, float *pos_host, float *rout_host, int64_t chunk_frames) {
    if (!rot_host || !gpos_host || !offsets_host || !pos_host || !rout_host) return fail(PMB_ERR_NULL, "fk_host: NULL array pointer");
    if (n_frames < 0) return fail(PMB_ERR_SHAPE, "n_frames < 0");
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS) return fail(PMB_ERR_SHAPE, "n_joints = %d out of range", n_joints);
    {   // validate the topology before any allocation or copy
        const pmb::JointProgram *prog = nullptr;
        int n_slots = 0;
        int rc = joint_program(parents_host, n_joints, false, prog, n_slots);
        if (rc) return rc;
    }
    if (n_frames == 0) return PMB_OK;
    const size_t J = static_cast<size_t>(n_joints);
    const size_t rot_w = kind == 0 ? 9 : 4;                           // floats per joint of the rotation output
    const size_t in_frame = J * 16 + 12, out_frame = J * 12 + J * rot_w * 4;  // bytes per frame, in / out
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
    // device staging of one chunk: [rot | gpos] in, [positions | rotations] out (sub-buffers 256-byte aligned)
    const size_t rot_bytes = (F * J * 16 + 255) & ~size_t(255), pos_bytes = (F * J * 12 + 255) & ~size_t(255);
    if ((rc = w.grow_dev(rot_bytes + F * 12, pos_bytes + F * J * rot_w * 4))) return rc;
    const bool in_pinned = is_pinned(rot_host) && is_pinned(gpos_host);
    const bool out_pinned = is_pinned(pos_host) && is_pinned(rout_host);
    if (!in_pinned && (rc = w.grow_pin(w.pin_in, w.pin_in_cap, rot_bytes + F * 12))) return rc;
    if (!out_pinned && (rc = w.grow_pin(w.pin_out, w.pin_out_cap, pos_bytes + F * J * rot_w * 4))) return rc;
    if ((!in_pinned || !out_pinned) && !w.pool) {
        const int hw = static_cast<int>(std::thread::hardware_concurrency());
        const int n = std::max(1, std::min(knob(K_HOST_THREADS, 4), hw > 0 ? hw : 1));
        w.pool.reset(new CopyPool(n - 1));
    }

    PMB_CUDA(cudaMemcpyAsync(w.offsets, offsets_host, J * 12, cudaMemcpyHostToDevice, w.stream[0]));
    PMB_CUDA(cudaEventRecord(w.offsets_ready, w.stream[0]));
    PMB_CUDA(cudaStreamWaitEvent(w.stream[1], w.offsets_ready, 0));
    PMB_CUDA(cudaStreamSynchronize(w.stream[0]));  // offsets_host may be a temporary of the caller

    struct Pending {  // a chunk whose results sit in the output ring and still have to reach the caller's memory
        bool live = false;
        size_t f0 = 0, n = 0;
    } pending[2];
    auto copy_out = [&](int slot) -> int {
        Pending &pd = pending[slot];
        if (!pd.live) return PMB_OK;
        PMB_CUDA(cudaEventSynchronize(w.d2h_done[slot]));
        w.pool->copy(pos_host + pd.f0 * J * 3, w.pin_out[slot], pd.n * J * 12);
        w.pool->copy(rout_host + pd.f0 * J * rot_w, w.pin_out[slot] + pos_bytes, pd.n * J * rot_w * 4);
        pd.live = false;
        return PMB_OK;
    };

    int slot = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk_frames, slot ^= 1) {
        const size_t n = static_cast<size_t>(std::min<int64_t>(chunk_frames, n_frames - f0));
        cudaStream_t st = w.stream[slot];  // stream order makes the slot's device buffers safe to reuse
        float *d_rot = reinterpret_cast<float *>(w.dev_in[slot]), *d_gpos = reinterpret_cast<float *>(w.dev_in[slot] + rot_bytes);
        float *d_pos = reinterpret_cast<float *>(w.dev_out[slot]), *d_rout = reinterpret_cast<float *>(w.dev_out[slot] + pos_bytes);
        if (in_pinned) {
            PMB_CUDA(cudaMemcpyAsync(d_rot, rot_host + f0 * J * 4, n * J * 16, cudaMemcpyHostToDevice, st));
            PMB_CUDA(cudaMemcpyAsync(d_gpos, gpos_host + f0 * 3, n * 12, cudaMemcpyHostToDevice, st));
        } else {
            PMB_CUDA(cudaEventSynchronize(w.h2d_done[slot]));  // the ring slot's previous chunk has left for the device
            w.pool->copy(w.pin_in[slot], rot_host + f0 * J * 4, n * J * 16);
            memcpy(w.pin_in[slot] + rot_bytes, gpos_host + f0 * 3, n * 12);
            PMB_CUDA(cudaMemcpyAsync(d_rot, w.pin_in[slot], n * J * 16, cudaMemcpyHostToDevice, st));
            PMB_CUDA(cudaMemcpyAsync(d_gpos, w.pin_in[slot] + rot_bytes, n * 12, cudaMemcpyHostToDevice, st));
            PMB_CUDA(cudaEventRecord(w.h2d_done[slot], st));
        }
        rc = kind == 0 ? pmb_fk_f32(d_rot, d_gpos, 3, w.offsets, 0, parents_host, static_cast<int64_t>(n), n_joints, d_pos, d_rout, st)
                       : pmb_fk_quat_f32(d_rot, d_gpos, 3, w.offsets, 0, parents_host, static_cast<int64_t>(n), n_joints, d_pos, d_rout, st);
        if (rc) return rc;  // (the lease drains both streams)
        if (out_pinned) {
            PMB_CUDA(cudaMemcpyAsync(pos_host + f0 * J * 3, d_pos, n * J * 12, cudaMemcpyDeviceToHost, st));
            PMB_CUDA(cudaMemcpyAsync(rout_host + f0 * J * rot_w, d_rout, n * J * rot_w * 4, cudaMemcpyDeviceToHost, st));
        } else {
            if ((rc = copy_out(slot))) return rc;  // the ring slot must be empty before the GPU refills it (normally done below)
            PMB_CUDA(cudaMemcpyAsync(w.pin_out[slot], d_pos, n * J * 12, cudaMemcpyDeviceToHost, st));
            PMB_CUDA(cudaMemcpyAsync(w.pin_out[slot] + pos_bytes, d_rout, n * J * rot_w * 4, cudaMemcpyDeviceToHost, st));
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

}  // extern "C"
