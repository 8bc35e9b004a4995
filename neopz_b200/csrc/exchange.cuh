// exchange.cuh — interface-row exchange of the row-sharded assembly over peer memory (NVLink / NVSwitch).
//
// every GPU owns a block of CSR rows; an element on GPU g may contribute to rows another GPU owns.  Those
// contributions are accumulated by the ordinary element kernels into a STAGING segment of g's own value array (rows the local
// pattern holds but g does not own) and then added straight into the owner's CSR values and load vector by the kernels below:
// plain stores / reductions on the peer's memory, mapped into this GPU's address space (cudaDeviceEnablePeerAccess inside a
// process, cudaIpcOpenMemHandle across the processes of a torchrun job).  Only doubles travel at assembly time; the positions
// in the owner's arrays were agreed at setup.  The elements that touch staging rows are stored FIRST in every group, so the push
// runs on a side stream while the interior elements are still being assembled (the reference has no counterpart: its only
// parallelism is threads over one shared matrix, StrMatrix/pzstrmatrixor.cpp:476-513).
//
// Ordering between the GPUs uses monotone step counters in device memory instead of host synchronisation:
//   owner:   zero A, rhs            -> signal ZEROED(step) into every pusher's flag block
//   pusher:  interface elements     -> wait ZEROED(step) from the owner -> push -> signal PUSHED(step) into the owner's flag block
//   owner:   own elements           -> wait PUSHED(step) from every pusher: the rows are complete
// Flags live on the WAITING device (signals are remote stores, polls are local loads).  A wait gives up after `timeout_ns`
// and raises an error word instead of hanging the device.
#pragma once
#include <cstdint>

namespace xch {

constexpr int MAX_LINKS = 16;  // peers per context (B200ASM_MAX_PEERS)

// dst_peer[dst[k]] += src[k]  (src = the contiguous staging segment of this GPU's CSR values).  Entries no element touched
// are zero and stay home.  System-scope reductions: the owner's own kernels add into the same rows concurrently.
__global__ void push_values_kernel(double *__restrict__ dst_peer, const int32_t *__restrict__ dst, const double *__restrict__ src,
                                   int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double v = src[k];
        if (v != 0.0) atomicAdd_system(dst_peer + dst[k], v);
    }
    __threadfence_system();
}

// load vector: dst_peer[dst[k]] += rhs[src[k]]
__global__ void push_gather_kernel(double *__restrict__ dst_peer, const int32_t *__restrict__ dst, const double *__restrict__ rhs,
                                   const int32_t *__restrict__ src, int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double v = rhs[src[k]];
        if (v != 0.0) atomicAdd_system(dst_peer + dst[k], v);
    }
    __threadfence_system();
}

__global__ void signal_kernel(unsigned long long *flag, unsigned long long value) {
    __threadfence_system();
    atomicExch_system(flag, value);
}

__global__ void wait_kernel(volatile unsigned long long *flag, unsigned long long value, unsigned long long timeout_ns, int *error) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag < value) {
        __nanosleep(256);
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            *error = 1;
            break;
        }
    }
    __threadfence_system();
}

}  // namespace xch
