// Measures the FP64 peaks of the device the roofline is quoted against ("measure DFMA
// and DMMA peaks on the box before quoting any roofline %"):
//   dfma : chains of independent fma.rn.f64 (16 per thread), all SMs, many warps
//   dmma : mma.sync.aligned.m8n8k4.row.col.f64 with 8 independent accumulator tiles per warp
//   redf64: throughput of red.global.add.f64 to distinct / shared addresses (scatter roofline)
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters, double a0, double b0) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every thread issues `per_thread` reds; addresses stride through a buffer of `span` doubles
__global__ void red_kernel(double *buf, size_t span, int per_thread, int mode) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < per_thread; k++) {
        size_t idx;
        if (mode == 0) idx = (tid + (size_t)k * nthreads) % span;              // coalesced, distinct
        else idx = ((tid * 2654435761u) + (size_t)k * 40503u * nthreads) % span;  // pseudo-random 8-byte scatter
        atomicAdd(buf + idx, 1.0);
    }
}

static float time_ms(void (*launch)(void *), void *arg, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(arg); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); launch(arg); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

struct Args { double *out; int grid, block, iters; double *buf; size_t span; int per_thread, mode; };
static void l_dfma(void *p) { Args *a = (Args *)p; dfma_kernel<<<a->grid, a->block>>>(a->out, a->iters, 1.0000001, 1e-9); }
static void l_dmma(void *p) { Args *a = (Args *)p; dmma_kernel<<<a->grid, a->block>>>(a->out, a->iters, 1.0000001, 1e-9); }
static void l_red(void *p) { Args *a = (Args *)p; red_kernel<<<a->grid, a->block>>>(a->buf, a->span, a->per_thread, a->mode); }

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int sms = prop.multiProcessorCount;
    Args a;
    cudaMalloc(&a.out, sizeof(double) * sms * 8 * 1024);
    a.block = 256; a.iters = 4096;
    double best_dfma = 0, best_dmma = 0;
    int bd = 0, bm = 0;
    for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
        a.grid = sms * per_sm;
        float ms = time_ms(l_dfma, &a, 5);
        double tf = 2.0 * 16 * (double)a.iters * a.block * a.grid / (ms * 1e-3) / 1e12;
        if (tf > best_dfma) { best_dfma = tf; bd = per_sm; }
        ms = time_ms(l_dmma, &a, 5);
        // one m8n8k4 = 8*8*4 = 256 FMA per warp
        tf = 2.0 * 256 * 8 * (double)a.iters * (a.block / 32) * a.grid / (ms * 1e-3) / 1e12;
        if (tf > best_dmma) { best_dmma = tf; bm = per_sm; }
    }
    // atomics: 64 MB span (L2 resident) and 2 GB span (HBM)
    double red_rates[4];
    size_t spans[2] = {(size_t)8 << 20, (size_t)256 << 20};
    int k = 0;
    for (int s = 0; s < 2; s++) {
        a.span = spans[s];
        cudaMalloc(&a.buf, a.span * sizeof(double));
        cudaMemset(a.buf, 0, a.span * sizeof(double));
        for (int mode = 0; mode < 2; mode++) {
            a.grid = sms * 8; a.block = 256; a.per_thread = 64; a.mode = mode;
            float ms = time_ms(l_red, &a, 3);
            red_rates[k++] = (double)a.grid * a.block * a.per_thread / (ms * 1e-3) / 1e9;
        }
        cudaFree(a.buf);
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"dfma_tflops\": %.2f, \"dfma_ctas_per_sm\": %d, "
           "\"dmma_tflops\": %.2f, \"dmma_ctas_per_sm\": %d, \"red_f64_Gops_L2_coalesced\": %.1f, "
           "\"red_f64_Gops_L2_scattered\": %.1f, \"red_f64_Gops_HBM_coalesced\": %.1f, \"red_f64_Gops_HBM_scattered\": %.1f}\n",
           prop.name, sms, prop.clockRate / 1000, best_dfma, bd, best_dmma, bm, red_rates[0], red_rates[1], red_rates[2], red_rates[3]);
    return 0;
}
