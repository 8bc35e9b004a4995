// Shared-memory load cost as a function of width and address pattern (one B200; the question behind the sum-factorisation
// kernel's stage loops): warp-wide LDS.64 / LDS.128 with (a) one address for the whole warp (broadcast), (b) three addresses
// (groups of nine lanes), (c) nine addresses, (d) 32 consecutive words.  Cycles per load instruction and SM with 8 warps per
// SM issuing back to back.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/lds_bcast.cu -o lds_bcast && ./lds_bcast
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH, int PATTERN>
__global__ void k(double *out, long long *cycles, int iters) {
    __shared__ __align__(16) double s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i * 0.5;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int base;
    if (PATTERN == 0) base = 0;                       // one address
    else if (PATTERN == 1) base = (lane / 9) * 10;    // three (four) addresses, 80 bytes apart
    else if (PATTERN == 2) base = (lane % 9) * 4;     // nine addresses, 32 bytes apart
    else base = lane * (WIDTH / 8);                   // consecutive
    double acc = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int off = (base + ((it * 16 + u) & 63) * 32) & 4095 & ~1;
            if (WIDTH == 16) {
                const double2 v = *reinterpret_cast<const double2 *>(s + off);
                acc += v.x + v.y;
            } else {
                acc += s[off];
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WIDTH, int PATTERN>
void run(const char *name) {
    const int blocks = 148, threads = 256, iters = 4096;
    double *out;
    long long *cyc;
    cudaMalloc(&out, blocks * threads * sizeof(double));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    k<WIDTH, PATTERN><<<blocks, threads>>>(out, cyc, iters);
    k<WIDTH, PATTERN><<<blocks, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += h[i];
    mean /= blocks;
    const double loads_per_sm = (double)iters * 16 * (threads / 32);
    printf("{\"width_bytes\": %d, \"pattern\": \"%s\", \"cycles_per_warp_load_per_sm\": %.3f}\n", WIDTH, name, mean / loads_per_sm);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<8, 0>("one address");
    run<16, 0>("one address");
    run<8, 1>("four addresses 80 B apart");
    run<16, 1>("four addresses 80 B apart");
    run<8, 2>("nine addresses 32 B apart");
    run<16, 2>("nine addresses 32 B apart");
    run<8, 3>("consecutive");
    run<16, 3>("consecutive");
    return 0;
}
