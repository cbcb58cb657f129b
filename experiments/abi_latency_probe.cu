// Host cost of one pmb_fk_f32 call measured from C (no ctypes in the way): back-to-back calls on a 1000 x 22 batch.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I include -o abi_latency_probe abi_latency_probe.cu -L pymotion_b200 -lpymotion_b200
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../include/pymotion_b200.h"

int main() {
    const int J = 22;
    const long long F = 1000;
    const int64_t par[J] = {0, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 11, 18, 19, 20};
    float *rot, *gp, *off, *pos, *rotm, *dq;
    cudaMalloc(&rot, F * J * 16); cudaMalloc(&gp, F * 12); cudaMalloc(&off, J * 12); cudaMalloc(&pos, F * J * 12); cudaMalloc(&rotm, F * J * 36);
    cudaMalloc(&dq, F * J * 32);
    std::vector<float> h(F * J * 4, 0.5f);
    cudaMemcpy(rot, h.data(), F * J * 16, cudaMemcpyHostToDevice);
    cudaMemset(gp, 0, F * 12); cudaMemset(off, 0, J * 12);
    const float off0[3] = {0, 0, 0};
    auto bench = [&](const char *name, auto call) {
        for (int i = 0; i < 2000; ++i) if (call()) { printf("error: %s\n", pmb_last_error()); return; }
        cudaDeviceSynchronize();
        // (a) 20000 calls back to back: the launch queue fills, so this is the larger of the host cost and the GPU's time per
        // tiny kernel; (b) bursts of 400 calls into an empty queue: the host cost alone
        const int it = 20000;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < it; ++i) call();
        auto t1 = std::chrono::steady_clock::now();
        cudaDeviceSynchronize();
        double burst = 0;
        const int bursts = 20, per = 400;
        for (int b = 0; b < bursts; ++b) {
            auto s0 = std::chrono::steady_clock::now();
            for (int i = 0; i < per; ++i) call();
            auto s1 = std::chrono::steady_clock::now();
            cudaDeviceSynchronize();
            burst += std::chrono::duration<double, std::micro>(s1 - s0).count();
        }
        printf("{\"call\": \"%s\", \"frames\": %lld, \"us_per_call_queue_full\": %.3f, \"us_per_call_host_only\": %.3f, \"variant\": \"%s\"}\n", name, F,
               std::chrono::duration<double, std::micro>(t1 - t0).count() / it, burst / (bursts * per), pmb_last_variant());
    };
    bench("pmb_fk_f32", [&] { return pmb_fk_f32(rot, gp, 3, off, 0, par, F, J, pos, rotm, nullptr); });
    bench("pmb_to_root_dual_quat_f32", [&] { return pmb_to_root_dual_quat_f32(rot, gp, 3, par, off, off0, F, J, dq, nullptr); });
    bench("pmb_from_root_dual_quat_f32", [&] { return pmb_from_root_dual_quat_f32(dq, par, F, J, pos, rot, nullptr); });
    bench("pmb_quat_normalize_f32", [&] { return pmb_quat_normalize_f32(rot, 1e-8f, rot, F * J, nullptr); });
    return 0;
}
