// Cost of a warp shuffle next to a shared-memory load (one B200): cycles per warp-wide instruction and SM, 8 warps per SM issuing
// back to back.  Question: would exchanging the nine stage-1 sums of the sum-factorisation kernel through SHFL (18 per (e,f))
// be cheaper than through shared memory (1 STS.64 + 9 LDS.64)?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/shfl_cost.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, long long *cycles, int iters) {
    __shared__ double s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = i * 0.5;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc = lane * 0.25;
    const int grp = (lane / 9) * 9;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 9; u++) {
            if (MODE == 0) {          // 64-bit shuffle from lane (group base + u): two SHFL.32
                acc += __shfl_sync(0xffffffffu, acc, (grp + u) & 31);
            } else if (MODE == 1) {   // 64-bit shared load, one address per group of nine lanes
                acc += s[((it & 7) * 64 + (lane / 9) * 9 + u) & 2047];
            } else {                  // 32-bit shuffle
                float f = (float)acc;
                f += __shfl_sync(0xffffffffu, f, (grp + u) & 31);
                acc = f;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name) {
    const int blocks = 148, threads = 256, iters = 4096;
    double *out;
    long long *cyc;
    cudaMalloc(&out, blocks * threads * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    k<MODE><<<blocks, threads>>>(out, cyc, iters);
    k<MODE><<<blocks, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += h[i];
    mean /= blocks;
    printf("{\"op\": \"%s\", \"cycles_per_warp_op_per_sm\": %.3f}\n", name, mean / ((double)iters * 9 * (threads / 32)));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("64-bit value through __shfl_sync (2 SHFL.32), dependent DADD chain");
    run<1>("LDS.64, one address per group of nine lanes, dependent DADD chain");
    run<2>("32-bit __shfl_sync, dependent chain with conversions");
    return 0;
}
