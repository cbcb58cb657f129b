// Host cost of back-to-back kernel launches as a function of the parameter block size (the fk kernels carry their joint
// program -- 2 KB -- and a 128-byte tensor map as __grid_constant__ parameters).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o launch_param_probe launch_param_probe.cu && ./launch_param_probe
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

template <int N> struct Blob { unsigned w[N]; };
template <int N> __global__ void k(const __grid_constant__ Blob<N> b, unsigned *out) { if (out && b.w[N - 1] == 12345u) out[0] = b.w[0]; }

template <int N> double run(int iters) {
    Blob<N> b{};
    for (int i = 0; i < 1000; ++i) k<N><<<1, 32>>>(b, nullptr);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) k<N><<<1, 32>>>(b, nullptr);
    auto t1 = std::chrono::steady_clock::now();
    cudaDeviceSynchronize();
    return std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
}
int main() {
    const int it = 20000;
    printf("{\"param_bytes\": 16, \"us_per_launch\": %.3f}\n", run<4>(it));
    printf("{\"param_bytes\": 256, \"us_per_launch\": %.3f}\n", run<64>(it));
    printf("{\"param_bytes\": 1024, \"us_per_launch\": %.3f}\n", run<256>(it));
    printf("{\"param_bytes\": 2304, \"us_per_launch\": %.3f}\n", run<576>(it));
    printf("{\"param_bytes\": 4224, \"us_per_launch\": %.3f}\n", run<1056>(it));
    return 0;
}
