// Probe: are TMA tensor stores of small boxes at 4-byte-granular inner coordinates (a) legal,
// (b) fast enough to carry fk's output?  Output viewed as a 2-D tensor with FOUR frames per row so the
// row pitch (4 * 36 * J bytes) is a multiple of 16 for every J.  Each warp owns 32 frames and, per chunk
// of C joints, issues 4 stores of an [8 rows x 9C floats] box (one per frame residue mod 4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tm), "r"(smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// value written for flat output index i (exact in fp32)
__host__ __device__ inline float val_of(long long i) { return (float)(i % 16777213LL); }

template <int C, int WARPS, int WORDS /*9 or 3*/>
__global__ void __launch_bounds__(WARPS * 32) store_probe(const __grid_constant__ CUtensorMap tm, long long F, int J, int Jw) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int BOXW = WORDS * C;                 // floats per box row
    constexpr int STAGE = 4 * 8 * BOXW * 4;         // bytes per warp (4 residues x 8 rows)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *stage = reinterpret_cast<float *>(smem + warp * ((STAGE + 127) & ~127));
    const long long f0 = ((long long)blockIdx.x * WARPS + warp) * 32;
    if (f0 >= F) return;
    const int r = lane >> 3, m = lane & 7;
    const long long f = f0 + 4 * m + r;
    float *mine = stage + (r * 8 + m) * BOXW;
    for (int c0 = 0; c0 + C <= Jw; c0 += C) {
        // fill: thread writes its frame's C*WORDS floats with 16-byte stores
        const long long base = (f * J + c0) * WORDS;
#pragma unroll
        for (int k = 0; k < BOXW; k += 4)
            *reinterpret_cast<float4 *>(mine + k) = make_float4(val_of(base + k), val_of(base + k + 1), val_of(base + k + 2), val_of(base + k + 3));
        fence_async_smem();
        __syncwarp();
        if (lane < 4) {
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(stage + lane * 8 * BOXW);
            tma_store_2d(&tm, s, lane * WORDS * J + WORDS * c0, (int)(f0 >> 2));
            tma_commit();
            tma_wait_read0();
        }
        __syncwarp();
    }
    if (lane < 4) tma_wait0();
}

// plain coalesced writer of the same bytes, as the speed reference
__global__ void plain_store(float4 *out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        __stcs(out + i, make_float4(1.f, 2.f, 3.f, 4.f));
}

template <int C, int WORDS>
void run(EncodeFn enc, long long F, int J, const char *label, bool verify) {
    const int Jw = (J / C) * C;  // joints written by full chunks; the remainder is left untouched
    const long long n = F * J * WORDS;
    float *out;
    CK(cudaMalloc(&out, n * 4));
    CK(cudaMemset(out, 0xFF, n * 4));
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)(4LL * WORDS * J), (cuuint64_t)(F / 4)};
    cuuint64_t strides[1] = {(cuuint64_t)(4LL * WORDS * J * 4)};
    cuuint32_t box[2] = {(cuuint32_t)(WORDS * C), 8};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("%s: cuTensorMapEncodeTiled failed rc=%d\n", label, (int)rc); cudaFree(out); return; }
    constexpr int WARPS = 4;
    const int stage = ((4 * 8 * WORDS * C * 4 + 127) & ~127);
    const int smem = WARPS * stage + 1024;
    auto k = store_probe<C, WARPS, WORDS>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long tiles = F / 32;
    const int blocks = (int)((tiles + WARPS - 1) / WARPS);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) k<<<blocks, WARPS * 32, smem>>>(tm, F, J, Jw);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k<<<blocks, WARPS * 32, smem>>>(tm, F, J, Jw);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("%-28s F=%lld J=%d (written %d) C=%d words=%d : %.4f ms  %.1f GB/s written\n", label, F, J, Jw, C, WORDS, ms, (double)F * Jw * WORDS * 4 / ms / 1e6);
    if (verify) {
        std::vector<float> h((size_t)n);
        CK(cudaMemcpy(h.data(), out, n * 4, cudaMemcpyDeviceToHost));
        long long bad = 0, first = -1;
        for (long long i = 0; i < n; ++i) {
            const int joint = (int)((i / WORDS) % J);
            const bool ok = joint < Jw ? (h[i] == val_of(i)) : (h[i] != h[i]);  // untouched = NaN pattern of the memset
            if (!ok) { if (first < 0) first = i; ++bad; }
        }
        printf("   verify: %lld mismatches of %lld (first at %lld)\n", bad, n, first);
    }
    // reference: plain coalesced store of the same bytes
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) plain_store<<<148 * 8, 256>>>((float4 *)out, n / 4);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("   plain coalesced float4 store of the same bytes: %.4f ms  %.1f GB/s\n", ms, n * 4 / ms / 1e6);
    CK(cudaFree(out));
}

int main() {
    EncodeFn enc = get_encode();
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    // small verified cases (J = 22: row pitch 792 B per frame, only 8-byte aligned per frame; 65: 4-byte)
    run<4, 9>(enc, 4096, 20, "rotmats C=4 small", true);
    run<4, 3>(enc, 4096, 20, "positions C=4 small", true);
    run<8, 9>(enc, 4096, 24, "rotmats C=8 J=24 small", true);
    run<4, 9>(enc, 1000000 / 32 * 32, 20, "rotmats C=4 1M x 20", false);
    run<4, 9>(enc, 1000000 / 32 * 32, 24, "rotmats C=4 1M x 24", false);
    run<8, 9>(enc, 1000000 / 32 * 32, 24, "rotmats C=8 1M x 24", false);
    run<4, 9>(enc, 1000000 / 32 * 32, 52, "rotmats C=4 1M x 52", false);
    run<4, 3>(enc, 1000000 / 32 * 32, 24, "positions C=4 1M x 24", false);
    run<8, 3>(enc, 1000000 / 32 * 32, 24, "positions C=8 1M x 24", false);
    return 0;
}
