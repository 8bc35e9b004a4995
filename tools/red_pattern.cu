// red_pattern.cu — microbenchmark: cost of warp-wide red.global.add.f64 as a function of the address pattern.
// Decides how the assembly kernels should lay out their scatter: are 32 lanes hitting 8 lines x 4
// sectors cheaper than 32 lanes hitting 32 lines?  Footprint 64 MB (L2 resident) and 2 GB (HBM).
// Patterns (per warp instruction, base line chosen pseudo-randomly per warp and iteration):
//   0: 32 consecutive doubles (2 lines, 8 sectors)
//   1: 32 different lines (32 sectors)
//   2: 8 lines x 4 doubles in ONE sector of each line (8 sectors)
//   3: 8 lines x 4 doubles in 4 different sectors of each line (32 sectors)
//   4: 4 lines x 8 doubles spread over the 4 sectors of each line (16 sectors)
//   5: 2 lines, 32 doubles permuted randomly inside them (8 sectors)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/red_pattern.cu -o tools/bin/red_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int PATTERN>
__global__ void red_kernel(double *buf, uint32_t nlines_mask, int iters) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = hash32(warp * 9781u + it * 6271u);
        uint64_t idx;  // in doubles; a line = 16 doubles, a sector = 4 doubles
        if (PATTERN == 0) idx = (uint64_t)(r & nlines_mask & ~1u) * 16 + lane;
        else if (PATTERN == 1) idx = (uint64_t)(hash32(r + lane * 7919u) & nlines_mask) * 16 + (lane & 15);
        else if (PATTERN == 2) idx = (uint64_t)(hash32(r + (lane >> 2) * 7919u) & nlines_mask) * 16 + (lane & 3);
        else if (PATTERN == 3) idx = (uint64_t)(hash32(r + (lane >> 2) * 7919u) & nlines_mask) * 16 + (lane & 3) * 4 + ((lane >> 2) & 3);
        else if (PATTERN == 4) idx = (uint64_t)(hash32(r + (lane >> 3) * 7919u) & nlines_mask) * 16 + (lane & 7) * 2;
        else idx = (uint64_t)(r & nlines_mask & ~1u) * 16 + ((lane * 13 + 5) & 31);
        atomicAdd(buf + idx, 1.0);
    }
}

template <int P>
double run(double *buf, uint32_t mask, int sms) {
    const int iters = 256, threads = 256, grid = sms * 8;
    red_kernel<P><<<grid, threads>>>(buf, mask, 8);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    red_kernel<P><<<grid, threads>>>(buf, mask, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return (double)grid * threads * iters / (ms * 1e-3) / 1e9;  // G lane-ops / s
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"sms\": %d", sms);
    for (int big = 0; big < 2; big++) {
        const size_t bytes = big ? (size_t)2 << 30 : (size_t)64 << 20;
        double *buf;
        cudaMalloc(&buf, bytes);
        cudaMemset(buf, 0, bytes);
        const uint32_t mask = (uint32_t)(bytes / 128 - 1);
        const char *tag = big ? "hbm2g" : "l2_64m";
        printf(", \"%s\": [%.1f, %.1f, %.1f, %.1f, %.1f, %.1f]", tag, run<0>(buf, mask, sms), run<1>(buf, mask, sms), run<2>(buf, mask, sms),
               run<3>(buf, mask, sms), run<4>(buf, mask, sms), run<5>(buf, mask, sms));
        cudaFree(buf);
    }
    printf(", \"unit\": \"G lane-atomics/s\", \"patterns\": [\"32 consecutive\", \"32 lines\", \"8 lines x 1 sector\", \"8 lines x 4 sectors\", \"4 lines x 4 sectors\", \"2 lines permuted\"]}\n");
    return 0;
}
