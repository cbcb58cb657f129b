// Probe: issue rate of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.
// If FFMA2 issues at the rate of FFMA, packing two independent chains halves the FP issue slots of the fk walk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/ffma2_probe experiments/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float *out, int iters, long long *cycles) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 1.0001f, c = 1e-4f;
    unsigned long long A0, A1, A2, A3, M, C;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(A1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(A2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(A3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %1};" : "=l"(M) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(c));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent scalar chains
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a0 = fmaf(a0, m, c), a1 = fmaf(a1, m, c), a2 = fmaf(a2, m, c), a3 = fmaf(a3, m, c);
                a4 = fmaf(a4, m, c), a5 = fmaf(a5, m, c), a6 = fmaf(a6, m, c), a7 = fmaf(a7, m, c);
            }
        } else {  // 4 independent packed chains = the same 8 FMAs per round
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A0) : "l"(M), "l"(C));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A1) : "l"(M), "l"(C));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A2) : "l"(M), "l"(C));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A3) : "l"(M), "l"(C));
            }
        }
    }
    long long t1 = clock64();
    if (MODE == 1) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(A0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(A1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(A2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(A3));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float *out;
    long long *cyc, h[4];
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 64);
    const int iters = 4096;
    for (int warps : {4, 8, 16}) {
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) probe<0><<<1, warps * 32>>>(out, iters, cyc);
            else probe<1><<<1, warps * 32>>>(out, iters, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
            const double fma_per_warp = 64.0 * iters;  // scalar FMAs per thread
            const double instr = mode == 0 ? fma_per_warp : fma_per_warp / 2;
            printf("{\"probe\": \"%s\", \"warps\": %d, \"cycles\": %lld, \"warp_instr_per_clk_per_sm\": %.3f, \"fma_lanes_per_clk_per_sm\": %.1f}\n",
                   mode ? "fma.rn.f32x2" : "fma.rn.f32", warps, h[0], instr * warps / h[0], fma_per_warp * warps * 32 / h[0]);
        }
    }
    printf("# %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
