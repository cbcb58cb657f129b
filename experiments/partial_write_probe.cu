// Probe: does a stream of 12-byte-per-frame partial-sector global writes (the pattern a lane = (frame, row) fk warp
// would produce if it wrote positions straight from registers) make L2 fetch the sectors from DRAM first?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/partial_write_probe experiments/partial_write_probe.cu
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ./experiments/partial_write_probe
// mode 0: lane (f, a) writes pos[f][j][a] for j = 0 .. J-1, one 4-byte store per joint (10 frames per warp)
// mode 1: the same bytes written as full coalesced float4 rows (baseline: what the bulk store achieves)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void partial(float *pos, long long n_frames, int J, int delay) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = lane / 3, a = lane - 3 * f;
    const long long tiles = n_frames / 10;
    for (long long t = blockIdx.x * (long long)(blockDim.x >> 5) + warp; t < tiles; t += (long long)gridDim.x * (blockDim.x >> 5)) {
        if (lane >= 30) continue;
        float *p = pos + ((t * 10 + f) * J) * 3 + a;
        float v = (float)t;
        for (int j = 0; j < J; ++j) {
            for (int d = 0; d < delay; ++d) v = v * 1.0001f + 0.5f;  // stand-in for the walk between two stores
            p[3 * j] = v;
        }
    }
}
__global__ void full(float4 *pos, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        pos[i] = make_float4((float)i, 1.f, 2.f, 3.f);
}
int main() {
    const long long F = 4000000;
    for (int J : {52, 65}) {
        float *pos;
        cudaMalloc(&pos, F * J * 12);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        for (int mode = 0; mode < 3; ++mode) {
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) partial<<<148 * 4, 256>>>(pos, F, J, 0);
                else if (mode == 1) partial<<<148 * 4, 256>>>(pos, F, J, 16);
                else full<<<148 * 8, 256>>>((float4 *)pos, F * J * 3 / 4);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("J=%d mode=%d  %.3f ms  %.0f GB/s written\n", J, mode, ms, F * J * 12 / ms / 1e6);
        }
        cudaFree(pos);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
