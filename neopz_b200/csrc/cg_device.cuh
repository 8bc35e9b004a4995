// cg_device.cuh — conjugate gradients on the device-resident CSR (SURVEY.md §8 row f/N2): the step after assembly,
// so that the assembled matrix never has to leave the GPU.
//
// Follows the reference's algorithm statement by statement:
//   Solvers/LinearSolvers/cg.h:44-120        CG(A, x, b, M, residual, max_iter, tol, FromCurrent)
//   Matrix/pzsysmp.cpp:190-232               TPZSYsmpMatrix::MultAdd: z(row) += a*x(col); if (row != col) z(col) += a*x(row)
//   Matrix/pzysmp.cpp (TPZFYsmpMatrix::MultAdd)  plain CSR product
//   Solvers/TPZStepSolver (SetJacobi(1, 0, 0)) one Jacobi sweep from zero = z = D^-1 r;  TPZCopySolve = identity
// The reference's product is serial; here one warp owns a row: the row part is a warp-reduced dot product, the
// transposed part (symmetric storage) goes out as red.global.add.f64.  Dot products are reduced in two fixed stages
// (per-CTA partials, the last CTA sums them in index order), so a solve is reproducible run to run in full storage;
// symmetric storage inherits the order-dependence of the atomics in q = A p (last-ulp).
#pragma once

namespace cgdev {

constexpr int THREADS = 256;

struct Scalars {       // device-resident scalars of the iteration
    double rho, rho_prev, pq, rr;
    unsigned int counter;
};

// block reduction + "last CTA finishes": returns true in thread 0 of the last CTA with the total in `total`
__device__ __forceinline__ bool reduce_and_finish(double v, double *partials, unsigned int *counter, double &total) {
    __shared__ double s_warp[THREADS / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < THREADS / 32; w++) s += s_warp[w];
        partials[blockIdx.x] = s;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return false;
    // last CTA: fixed-order sum of the partials
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += THREADS) s += ((volatile double *)partials)[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    __syncthreads();
    if (lane == 0) s_warp[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < THREADS / 32; w++) t += s_warp[w];
        total = t;
        *counter = 0;
        return true;
    }
    return false;
}

// diag[row] = A(row,row): first entry of the row in symmetric storage, searched in full storage
__global__ void extract_diag_kernel(int64_t neq, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                    const double *__restrict__ a, int symmetric, double *__restrict__ diag) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < neq; row += (int64_t)gridDim.x * blockDim.x) {
        double d = 0.0;
        if (symmetric) {
            if (ia[row + 1] > ia[row] && ja[ia[row]] == row) d = a[ia[row]];
        } else {
            int64_t lo = ia[row], hi = ia[row + 1] - 1;
            while (lo <= hi) {
                const int64_t mid = (lo + hi) >> 1;
                const int64_t c = ja[mid];
                if (c == row) { d = a[mid]; break; }
                if (c < row) lo = mid + 1; else hi = mid - 1;
            }
        }
        diag[row] = d;
    }
}

// y += alpha * A x, one warp per row (y pre-initialised by the caller)
__global__ void __launch_bounds__(THREADS) spmv_kernel(int64_t neq, const int64_t *__restrict__ ia, const int32_t *__restrict__ ja,
                                                       const double *__restrict__ a, int symmetric, double alpha,
                                                       const double *__restrict__ x, double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (THREADS / 32);
    for (int64_t row = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); row < neq; row += nwarps) {
        const int64_t s = ia[row], e = ia[row + 1];
        const double xr = alpha * x[row];
        double sum = 0.0;
        for (int64_t k = s + lane; k < e; k += 32) {
            const int32_t c = ja[k];
            const double v = a[k];
            sum += v * x[c];
            if (symmetric && c != row) atomicAdd(y + c, v * xr);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
        if (lane == 0) {
            if (symmetric) atomicAdd(y + row, alpha * sum);
            else y[row] += alpha * sum;
        }
    }
}

// z = M^-1 r (Jacobi or identity), rho_prev = rho, rho = r.z
__global__ void __launch_bounds__(THREADS) precond_dot_kernel(int64_t n, const double *__restrict__ r, const double *__restrict__ diag,
                                                              double *__restrict__ z, double *partials, Scalars *sc) {
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        const double ri = r[i];
        const double zi = diag ? ri / diag[i] : ri;
        z[i] = zi;
        acc += ri * zi;
    }
    double total;
    if (reduce_and_finish(acc, partials, &sc->counter, total)) {
        sc->rho_prev = sc->rho;
        sc->rho = total;
    }
}

// p = z (first) or p = beta p + z with beta = rho / rho_prev  (TimesBetaPlusZ)
__global__ void update_p_kernel(int64_t n, int first, const double *__restrict__ z, double *__restrict__ p, const Scalars *sc) {
    const double beta = first ? 0.0 : sc->rho / sc->rho_prev;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = first ? z[i] : beta * p[i] + z[i];
}

// slot 0: pq = a.b ; slot 1: rr = a.b
__global__ void __launch_bounds__(THREADS) dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                                                      double *partials, Scalars *sc, int slot) {
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) acc += a[i] * b[i];
    double total;
    if (reduce_and_finish(acc, partials, &sc->counter, total)) {
        if (slot == 0) sc->pq = total; else sc->rr = total;
    }
}

// alpha = rho / pq;  x += alpha p;  r -= alpha q;  rr = r.r
__global__ void __launch_bounds__(THREADS) update_xr_kernel(int64_t n, const double *__restrict__ p, const double *__restrict__ q,
                                                            double *__restrict__ x, double *__restrict__ r, double *partials, Scalars *sc) {
    const double alpha = sc->rho / sc->pq;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;
    }
    double total;
    if (reduce_and_finish(acc, partials, &sc->counter, total)) sc->rr = total;
}

}  // namespace cgdev

// ---- row-sharded solve (several GPUs, b200asm_multi_cg_solve) -------------------------------------------------------------
// Every GPU holds the rows it owns (local rows [row_base, row_base + nrows)) with their complete pattern in a LOCAL numbering
// that also contains the equations other GPUs own but these rows couple to ("halo").  One iteration exchanges two halos over
// peer memory (NVLink): the owners' values of p are pulled into the halo before the product, and (symmetric storage) the
// transposed contributions that the product left in the halo of q are added into the owners' q.  Scalars are summed on the
// host in GPU order, so the iteration is the reference's statement by statement (Solvers/LinearSolvers/cg.h:44-120).
namespace cgdev {

// y += alpha * A x restricted to the owned rows
__global__ void __launch_bounds__(THREADS) spmv_rows_kernel(int64_t row_base, int64_t nrows, const int64_t *__restrict__ ia,
                                                            const int32_t *__restrict__ ja, const double *__restrict__ a, int symmetric,
                                                            double alpha, const double *__restrict__ x, double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (THREADS / 32);
    for (int64_t k0 = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); k0 < nrows; k0 += nwarps) {
        const int64_t row = row_base + k0;
        const int64_t s = ia[row], e = ia[row + 1];
        const double xr = alpha * x[row];
        double sum = 0.0;
        for (int64_t k = s + lane; k < e; k += 32) {
            const int32_t c = ja[k];
            const double v = a[k];
            sum += v * x[c];
            if (symmetric && c != row) atomicAdd(y + c, v * xr);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
        if (lane == 0) atomicAdd(y + row, alpha * sum);
    }
}

// halo of a vector: v[local[k]] = (vector of the owning GPU)[remote[k]]      (P2P loads)
__global__ void halo_pull_kernel(int64_t n, const int32_t *__restrict__ local, const int32_t *__restrict__ owner,
                                 const int32_t *__restrict__ remote, double *const *__restrict__ peer_vec, double *__restrict__ v) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        v[local[k]] = peer_vec[owner[k]][remote[k]];
}

// (vector of the owning GPU)[remote[k]] += v[local[k]]                          (P2P reductions)
__global__ void halo_push_kernel(int64_t n, const int32_t *__restrict__ local, const int32_t *__restrict__ owner,
                                 const int32_t *__restrict__ remote, double *const *__restrict__ peer_vec, const double *__restrict__ v) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[local[k]];
        if (x != 0.0) atomicAdd_system(peer_vec[owner[k]] + remote[k], x);
    }
    __threadfence_system();
}

// p = z (first) or p = beta p + z, beta from the host
__global__ void update_p_val_kernel(int64_t n, int first, double beta, const double *__restrict__ z, double *__restrict__ p) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = first ? z[i] : beta * p[i] + z[i];
}

// x += alpha p;  r -= alpha q;  rr (this GPU's part) = r.r
__global__ void __launch_bounds__(THREADS) update_xr_val_kernel(int64_t n, double alpha, const double *__restrict__ p,
                                                                const double *__restrict__ q, double *__restrict__ x,
                                                                double *__restrict__ r, double *partials, Scalars *sc) {
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;
    }
    double total;
    if (reduce_and_finish(acc, partials, &sc->counter, total)) sc->rr = total;
}

}  // namespace cgdev
