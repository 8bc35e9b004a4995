// b200asm — hand-written sm_100a kernels + the C ABI (include/b200asm.h) of the B200 assembly engine.
//
// What one b200asm_assemble() replaces (reference, all on the CPU, one element at a time):
//   StrMatrix/pzstrmatrixor.cpp:157-250   element loop: CalcStiff -> ComputeDestinationIndices -> AddKel -> AddFel
//   Mesh/pzinterpolationspace.cpp:404-473 quadrature loop (CalcStiffInternal)
//   Mesh/pzgeoel.cpp:1145-1356            Jacobian, determinant, inverse
//   Mesh/TPZCompElH1.cpp:140-149          dphix = jacinv^T dphi
//   Material/Poisson/TPZMatPoisson.cpp:19-42, Material/Elasticity/TPZElasticity3D.cpp:269-372  Contribute
//   Matrix/pzsysmp.cpp:370-411, Matrix/pzysmp.cpp:178-218, Matrix/pzfmatrix.cpp:285-299        scatter-add
//
// Design (DESIGN.md has the long version):
//   * elements are processed in batches of EPB per CTA by a persistent grid (multiple of the SM count);
//   * phase 1: per (element, point) Jacobian -> J^-1 and w|detJ| into shared memory;
//   * phase 2: per point q, x-gradients sqrt(w|detJ|) * J^-T grad(phi_i) of all shape functions into a
//     double-buffered shared-memory panel A (K rows x M columns per element);
//   * phase 3: the element matrix is the Gram matrix A^T A.  Only its local upper triangle is formed,
//     in TILE x TILE register tiles, one thread per (element, tile): each k-row costs 2*TILE shared
//     loads for TILE^2 DFMAs.  Poisson: M = nshape, 3 rows per point.  Elasticity3D: M = 3*nshape
//     (column 3*i+d), 1 row per point; the nine products per node pair are combined with C1,C2,C3
//     in registers afterwards (exactly the nine formulas of TPZElasticity3D.cpp:318-326);
//   * scatter: every tile entry has a precomputed CSR position (device-built scatter map, laid out so
//     that a warp reads consecutive ints) and is added with red.global.add.f64.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200asm.h"
#include "exchange.cuh"

// ------------------------------------------------------------------------------------------------
// device-side parameter blocks
// ------------------------------------------------------------------------------------------------
struct VolParams {
    int64_t nel;
    int64_t nbatch;
    int nq;
    int kind;  // B200ASM_POISSON / B200ASM_ELASTICITY3D
    const double *__restrict__ xyz;
    const int32_t *__restrict__ elnodes;  // [nel][NN]
    const int32_t *__restrict__ dest;     // [nel][M]
    const double *__restrict__ qw;        // [nq]
    const double *__restrict__ phi;       // [nq][N]
    const double *__restrict__ dphi;      // [nq][3][N]
    const double *__restrict__ dng;       // [nq][3][NN] gradients of the corner (geometry) functions
    const double *__restrict__ dng_t;     // [3][NN][nq]  same, point index fastest (DMMA kernels)
    const double *__restrict__ dphi_pad;  // [nq][3][NP]  rows padded to NP = 16*ceil(N/16) doubles (DMMA kernels)
    const double *__restrict__ phi_pad;   // [nq][NP]
    const double *__restrict__ force;     // optional [nel][nq][NS]
    const double *__restrict__ aux;       // kernel-specific tables of the group (affine_simplex.cuh: Ghat, cphi, cd)
    const double *__restrict__ aux2;      // sumfact_hex.cuh: the factor tables F1, F2 of the group's rule
    const int32_t *__restrict__ smap;     // [nbatch][TILE*TILE][EPB*NT]
    const int32_t *__restrict__ smapT;    // same, transposed entry (full storage only)
    double *__restrict__ a;
    double *__restrict__ rhs;
    int atomic;  // 1: red.global.add.f64; 0: plain read-modify-write (the launch covers one colour: no two
                 // elements of it share an equation, so no two threads touch the same entry)
    int rhs_only;  // 1: load vector only (TPZStrMatParInterface::Assemble(rhs)): no Gram products, no matrix scatter
    int debug;     // profiling aid (option "debug"): bit 0 = skip the matrix scatter (results are then WRONG)
    double coef[16];
};

template <int NN_, int N_, int NS_, int TILE_, int EPB_>
struct VolCfg {
    static constexpr int NN = NN_, N = N_, NS = NS_, TILE = TILE_, EPB = EPB_;
    static constexpr int M = N * NS;
    static constexpr int NTB = (M + TILE - 1) / TILE;
    static constexpr int MP = NTB * TILE;
    static constexpr int NT = NTB * (NTB + 1) / 2;
    static constexpr int KQ = NS == 1 ? 3 : 1;  // panel rows per integration point
    static constexpr int SLOTS = EPB * NT;      // (element, tile) work items per batch
    static constexpr int NTHREADS = ((SLOTS + 31) / 32) * 32;
    static constexpr int JS = 11;                       // J^-1 (9), w|detJ|, sqrt(w|detJ|)
    static constexpr int FPT = (EPB * M + NTHREADS - 1) / NTHREADS;  // rhs items per thread
    static constexpr int AS = KQ * MP;                  // panel doubles per element per buffer
    static size_t smem_bytes(int nq) { return sizeof(double) * ((size_t)EPB * NN * 3 + (size_t)EPB * nq * JS + 2 * (size_t)EPB * AS); }
};

__device__ __forceinline__ void red_add(double *addr, double v) { atomicAdd(addr, v); }
// scatter-add of one entry: atomic, or plain when the launch is conflict-free by colouring
__device__ __forceinline__ void scatter_add(double *addr, double v, int atomic) {
    if (atomic & 4) {  // option "drop_tiny": TPZSYsmpMatrix / TPZFYsmpMatrix::AddKel skip IsZero(value) entries (Matrix/pzsysmp.cpp:381)
        if (fabs(v) < 1e-12) return;
        atomic &= 3;
    }
    if (atomic == 1) atomicAdd(addr, v);
    else if (atomic == 0) *addr += v;
    // atomic == 2: profiling aid, the value is dropped (keeps the arithmetic, removes the memory traffic)
    else if (v == 1.2345e300) *addr = v;
}

// load-vector entry: destination -1 = equation removed by the equation filter (StrMatrix/TPZEquationFilter.h:120-141)
__device__ __forceinline__ void scatter_rhs(double *rhs, int32_t d, double v, int atomic) {
    if (d >= 0) scatter_add(rhs + d, v, atomic & 3);  // (TPZFMatrix::AddFel keeps every entry: no drop_tiny here)
}
// predicated reduction: no branch around the red (a divergent `if (pos >= 0) atomicAdd` costs BSSY/BRA/BSYNC per entry)
__device__ __forceinline__ void red_if_valid(double *a, int32_t pos, double v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\t@p red.global.add.f64 [%0], %2;\n\t}" ::"l"(a + pos), "r"(pos), "d"(v) : "memory");
}
// scatter-add of N register values through their precomputed CSR positions (-1 = no slot); the (uniform) scatter mode is
// tested once, not per entry
template <int N>
__device__ __forceinline__ void scatter_many(double *a, const int32_t (&pos)[N], const double (&val)[N], int atomic) {
    if (atomic == 1) {
#pragma unroll
        for (int k = 0; k < N; k++) red_if_valid(a, pos[k], val[k]);
    } else if (atomic == 0) {
#pragma unroll
        for (int k = 0; k < N; k++)
            if (pos[k] >= 0) a[pos[k]] += val[k];
    }
}

#include "gram_mma.cuh"
#include "gram_mma_team.cuh"
#include "affine_simplex.cuh"
#include "affine_hex.cuh"
#include "gather_rows.cuh"
#include "sumfact_hex.cuh"
#include "pattern_device.cuh"
#include "cg_device.cuh"

// decode the linear index of an upper-triangular tile into (bi, bj), bi <= bj
template <int NTB>
__device__ __forceinline__ void tile_coords(int t, int &bi, int &bj) {
    bi = 0;
#pragma unroll
    for (int row = 0; row < NTB - 1; row++) {
        const int cnt = NTB - row;
        if (bi == row && t >= cnt) {
            t -= cnt;
            bi = row + 1;
        }
    }
    bj = bi + t;
}

// ------------------------------------------------------------------------------------------------
// volume elements (hexahedra / tetrahedra)
// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NTHREADS) assemble_volume_kernel(const VolParams p) {
    constexpr int NN = C::NN, N = C::N, NS = C::NS, TILE = C::TILE, EPB = C::EPB, M = C::M, MP = C::MP;
    constexpr int NT = C::NT, KQ = C::KQ, NTHREADS = C::NTHREADS, JS = C::JS, FPT = C::FPT, AS = C::AS;
    extern __shared__ double smem[];
    double *Xs = smem;                              // [EPB][NN][3]
    double *JI = Xs + EPB * NN * 3;                 // [EPB][nq][JS]
    double *As = JI + (size_t)EPB * p.nq * JS;      // [2][EPB][KQ][MP]
    const int tid = threadIdx.x;
    const int nq = p.nq;

    for (int i = tid; i < 2 * EPB * AS; i += NTHREADS) As[i] = 0.0;  // padding columns stay zero for ever

    const bool has_tile = tid < C::SLOTS;
    const int el_t = tid / NT;
    int bi, bj;
    tile_coords<C::NTB>(tid % NT, bi, bj);
    const int I0 = bi * TILE, J0 = bj * TILE;

    for (int64_t batch = blockIdx.x; batch < p.nbatch; batch += gridDim.x) {
        const int64_t e0 = batch * EPB;
        const int nloc = (int)min((int64_t)EPB, p.nel - e0);
        __syncthreads();  // the previous batch no longer reads Xs / JI / As

        for (int i = tid; i < nloc * NN; i += NTHREADS) {
            const int64_t node = p.elnodes[e0 * NN + i];
            Xs[i * 3 + 0] = p.xyz[node * 3 + 0];
            Xs[i * 3 + 1] = p.xyz[node * 3 + 1];
            Xs[i * 3 + 2] = p.xyz[node * 3 + 2];
        }
        __syncthreads();

        // ---- phase 1: geometry per (element, point) ------------------------------------------
        for (int it = tid; it < nloc * nq; it += NTHREADS) {
            const int el = it / nq, q = it - el * nq;
            const double *X = Xs + el * NN * 3;
            const double *dn = p.dng + (size_t)q * 3 * NN;
            double j00 = 0, j01 = 0, j02 = 0, j10 = 0, j11 = 0, j12 = 0, j20 = 0, j21 = 0, j22 = 0;
#pragma unroll
            for (int a = 0; a < NN; a++) {  // gradx(j,k) += x_a[j] * dN_a/dxi_k   (Geom/TPZGeoCube.h:141-149)
                const double d0 = __ldg(dn + a), d1 = __ldg(dn + NN + a), d2 = __ldg(dn + 2 * NN + a);
                const double x = X[a * 3], y = X[a * 3 + 1], z = X[a * 3 + 2];
                j00 += x * d0; j01 += x * d1; j02 += x * d2;
                j10 += y * d0; j11 += y * d1; j12 += y * d2;
                j20 += z * d0; j21 += z * d1; j22 += z * d2;
            }
            // Mesh/pzgeoel.cpp:1309-1336
            double det = 0.0;
            det -= j02 * j11 * j20;
            det += j01 * j12 * j20;
            det += j02 * j10 * j21;
            det -= j00 * j12 * j21;
            det -= j01 * j10 * j22;
            det += j00 * j11 * j22;
            if (fabs(det) < 1.e-12) det = 1.e-12;
            const double id = 1.0 / det;
            double *o = JI + ((size_t)el * nq + q) * JS;
            o[0] = (-j12 * j21 + j11 * j22) * id;
            o[1] = (j02 * j21 - j01 * j22) * id;
            o[2] = (-j02 * j11 + j01 * j12) * id;
            o[3] = (j12 * j20 - j10 * j22) * id;
            o[4] = (-j02 * j20 + j00 * j22) * id;
            o[5] = (j02 * j10 - j00 * j12) * id;
            o[6] = (-j11 * j20 + j10 * j21) * id;
            o[7] = (j01 * j20 - j00 * j21) * id;
            o[8] = (-j01 * j10 + j00 * j11) * id;
            const double w = __ldg(p.qw + q) * fabs(det);  // weight *= fabs(detjac)  (pzinterpolationspace.cpp:468)
            o[9] = w;
            o[10] = sqrt(w);
        }

        double acc[TILE][TILE];
#pragma unroll
        for (int r = 0; r < TILE; r++)
#pragma unroll
            for (int c = 0; c < TILE; c++) acc[r][c] = 0.0;
        double facc[FPT];
#pragma unroll
        for (int k = 0; k < FPT; k++) facc[k] = 0.0;
        __syncthreads();

        for (int q = 0; q < nq; q++) {
            double *Ab = As + (size_t)(q & 1) * EPB * AS;
            // ---- phase 2: panel rows of point q: sqrt(w) * jacinv^T * dphi  (TPZCompElH1.cpp:147) ----
            for (int it = tid; it < nloc * N; it += NTHREADS) {
                const int el = it / N, i = it - el * N;
                const double *ji = JI + ((size_t)el * nq + q) * JS;
                const double *dp = p.dphi + (size_t)q * 3 * N + i;
                const double d0 = __ldg(dp), d1 = __ldg(dp + N), d2 = __ldg(dp + 2 * N);
                const double sw = ji[10];
                const double g0 = (ji[0] * d0 + ji[3] * d1 + ji[6] * d2) * sw;
                const double g1 = (ji[1] * d0 + ji[4] * d1 + ji[7] * d2) * sw;
                const double g2 = (ji[2] * d0 + ji[5] * d1 + ji[8] * d2) * sw;
                double *row = Ab + (size_t)el * AS;
                if (NS == 1) {
                    row[i] = g0;
                    row[MP + i] = g1;
                    row[2 * MP + i] = g2;
                } else {
                    row[3 * i] = g0;
                    row[3 * i + 1] = g1;
                    row[3 * i + 2] = g2;
                }
            }
            __syncthreads();
            // ---- phase 3: Gram update of this thread's register tile --------------------------
            if (has_tile && el_t < nloc && !p.rhs_only) {
                const double *row = Ab + (size_t)el_t * AS;
#pragma unroll
                for (int kr = 0; kr < KQ; kr++) {
                    double a[TILE], b[TILE];
#pragma unroll
                    for (int r = 0; r < TILE; r++) a[r] = row[kr * MP + I0 + r];
#pragma unroll
                    for (int c = 0; c < TILE; c++) b[c] = row[kr * MP + J0 + c];
#pragma unroll
                    for (int r = 0; r < TILE; r++)
#pragma unroll
                        for (int c = 0; c < TILE; c++) acc[r][c] = fma(a[r], b[c], acc[r][c]);
                }
            }
            // ---- load vector of point q ---------------------------------------------------------
#pragma unroll
            for (int k = 0; k < FPT; k++) {
                const int it = tid + k * NTHREADS;
                if (it < nloc * M) {
                    const int el = it / M, m = it - el * M;
                    const double *ji = JI + ((size_t)el * nq + q) * JS;
                    const double w = ji[9];
                    if (NS == 1) {
                        // ef(i) += weight*fScale*phi(i)*force   (TPZMatPoisson.cpp:39-40)
                        const double f = p.force ? p.force[((e0 + el) * nq + q)] : p.coef[1];
                        facc[k] += w * p.coef[0] * __ldg(p.phi + (size_t)q * N + m) * f;
                    } else {
                        // ef(3j+k) += weight*(force_k*phi_j - prestress_k*dphi(k,j))   (TPZElasticity3D.cpp:278)
                        const int j = m / 3, kd = m - 3 * j;
                        const double f = p.force ? p.force[((e0 + el) * nq + q) * 3 + kd] : p.coef[3 + kd];
                        const double wdphi = ji[10] * Ab[(size_t)el * AS + m];  // w * dphix(kd, j)
                        facc[k] += w * f * __ldg(p.phi + (size_t)q * N + j) - p.coef[6 + kd] * wdphi;
                    }
                }
            }
        }

        // ---- epilogue: scatter-add into CSR values and rhs -----------------------------------
        if (has_tile && el_t < nloc && !p.rhs_only) {
            if (NS == 3) {
                // nine sums S[v][u] per node pair -> the 3x3 block of ek  (TPZElasticity3D.cpp:318-326)
                const double C1 = p.coef[0], C2 = p.coef[1], C3 = p.coef[2];
#pragma unroll
                for (int il = 0; il < TILE / 3; il++)
#pragma unroll
                    for (int jl = 0; jl < TILE / 3; jl++) {
                        double S[3][3];
#pragma unroll
                        for (int v = 0; v < 3; v++)
#pragma unroll
                            for (int u = 0; u < 3; u++) S[v][u] = acc[3 * il + v][3 * jl + u];
#pragma unroll
                        for (int a = 0; a < 3; a++)
#pragma unroll
                            for (int b = 0; b < 3; b++) {
                                double e;
                                if (a == b) e = (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3;
                                else e = S[b][a] * C1 - S[a][b] * C2;
                                acc[3 * il + a][3 * jl + b] = e;
                            }
                    }
            } else {
                const double s = p.coef[0];
#pragma unroll
                for (int r = 0; r < TILE; r++)
#pragma unroll
                    for (int c = 0; c < TILE; c++) acc[r][c] *= s;
            }
            const int32_t *sm = p.smap + (size_t)batch * TILE * TILE * C::SLOTS + tid;
            const int32_t *smT = p.smapT ? p.smapT + (size_t)batch * TILE * TILE * C::SLOTS + tid : nullptr;
#pragma unroll
            for (int r = 0; r < TILE; r++)
#pragma unroll
                for (int c = 0; c < TILE; c++) {
                    const int32_t pos = sm[(r * TILE + c) * C::SLOTS];
                    if (pos >= 0) scatter_add(p.a + pos, acc[r][c], p.atomic);
                    if (smT) {
                        const int32_t posT = smT[(r * TILE + c) * C::SLOTS];
                        if (posT >= 0) scatter_add(p.a + posT, acc[r][c], p.atomic);
                    }
                }
        }
#pragma unroll
        for (int k = 0; k < FPT; k++) {
            const int it = tid + k * NTHREADS;
            if (it < nloc * M) scatter_rhs(p.rhs, p.dest[e0 * M + it], facc[k], p.atomic);
        }
    }
}

// interface exchange: dst[pos[k]] += val[k], positions are distinct (one per received CSR entry)
__global__ void scatter_add_kernel(double *__restrict__ dst, const int32_t *__restrict__ pos, const double *__restrict__ val, int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) dst[pos[k]] += val[k];
}

// scatter-map construction for a volume group: one thread per (batch, tile entry, slot)
template <class C>
__global__ void build_volume_smap_kernel(int64_t nel, int64_t nbatch, const int32_t *__restrict__ dest,
                                         const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, int symmetric,
                                         int32_t *__restrict__ smap, int32_t *__restrict__ smapT, int *__restrict__ missing) {
    constexpr int TILE = C::TILE, M = C::M, NT = C::NT, SLOTS = C::SLOTS;
    const int64_t total = nbatch * TILE * TILE * SLOTS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int slot = (int)(idx % SLOTS);
        const int rc = (int)((idx / SLOTS) % (TILE * TILE));
        const int64_t batch = idx / ((int64_t)SLOTS * TILE * TILE);
        const int el_l = slot / NT;
        int bi, bj;
        tile_coords<C::NTB>(slot % NT, bi, bj);
        const int r = rc / TILE, c = rc % TILE;
        const int i = bi * TILE + r, j = bj * TILE + c;
        const int64_t el = batch * C::EPB + el_l;
        int32_t pos = -1, posT = -1;
        if (el < nel && i < M && j < M && i <= j) {
            const int64_t di = dest[el * M + i], dj = dest[el * M + j];
            auto find = [&](int64_t row, int64_t col) -> int32_t {
                if (row < 0 || col < 0) return -1;  // equation removed by the TPZEquationFilter: no slot, not an error
                int64_t lo = ia[row], hi = ia[row + 1] - 1;
                while (lo <= hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    const int64_t v = ja[mid];
                    if (v == col) return (int32_t)mid;
                    if (v < col) lo = mid + 1; else hi = mid - 1;
                }
                atomicAdd(missing, 1);
                return -1;
            };
            if (symmetric) {
                pos = find(min(di, dj), max(di, dj));
            } else {
                pos = find(di, dj);
                if (i != j) posT = find(dj, di);
            }
        }
        smap[idx] = pos;
        if (smapT) smapT[idx] = posT;
    }
}

// ------------------------------------------------------------------------------------------------
// boundary faces (quadrilaterals / triangles): one thread per element
//   ek(ns*i+a, ns*j+b) += M[a][b]*phi_i*phi_j*w ,  ef(ns*i+a) += v[a]*phi_i*w
// with w = weight*|detjac| of the 2-D Gram-Schmidt branch of TPZGeoEl::Jacobian (pzgeoel.cpp:1228-1295)
// ------------------------------------------------------------------------------------------------
struct BcParams {
    int64_t nel;      // elements of the group (stride of the entry-major scatter map)
    int64_t el0, el1; // this launch covers elements [el0, el1) (one colour, or the whole group)
    int atomic;
    int rhs_only;
    int nq;
    const double *__restrict__ xyz;
    const int32_t *__restrict__ elnodes;
    const int32_t *__restrict__ dest;
    const double *__restrict__ qw;
    const double *__restrict__ phi;   // [nq][N]
    const double *__restrict__ dphi;  // [nq][2][N]   (plane domain elements only)
    const double *__restrict__ dng;   // [nq][fdim][NN]
    int kind;                         // plane domain elements: B200ASM_POISSON / B200ASM_ELASTICITY2D
    int fdim;                         // dimension of the element: 2 (faces, plane elements) or 1 (line elements)
    const double *__restrict__ force;   // generic volume kernel only: optional [nel][nq][NS] forcing-function table
    const int32_t *__restrict__ smap;   // [N*N*NS*NS][nel]
    const int32_t *__restrict__ smapT;
    double *__restrict__ a;
    double *__restrict__ rhs;
    double coef[16];
};

template <int NN, int N, int NS>
__global__ void __launch_bounds__(128) assemble_bc_kernel(const BcParams p) {
    const int64_t el = p.el0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (el >= p.el1) return;
    double X[NN][3];
#pragma unroll
    for (int a = 0; a < NN; a++) {
        const int64_t node = p.elnodes[el * NN + a];
        X[a][0] = p.xyz[node * 3];
        X[a][1] = p.xyz[node * 3 + 1];
        X[a][2] = p.xyz[node * 3 + 2];
    }
    double S[N][N];  // sum_q phi_i phi_j w   (upper part used)
    double T[N];     // sum_q phi_i w
#pragma unroll
    for (int i = 0; i < N; i++) {
        T[i] = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) S[i][j] = 0.0;
    }
    for (int q = 0; q < p.nq; q++) {
        const double *dn = p.dng + (size_t)q * 2 * NN;
        double v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < NN; a++) {
            const double d0 = __ldg(dn + a), d1 = __ldg(dn + NN + a);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                v1[k] += X[a][k] * d0;
                v2[k] += X[a][k] * d1;
            }
        }
        double n1 = 0, dot = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            n1 += v1[k] * v1[k];
            dot += v1[k] * v2[k];
        }
        n1 = sqrt(n1);
        double n2 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double v1t = v1[k] / n1;
            const double v2t = v2[k] - dot * v1t / n1;
            n2 += v2t * v2t;
        }
        n2 = sqrt(n2);
        double det = n1 * n2;
        if (fabs(det) < 1.e-12) det = 1.e-12;
        const double w = __ldg(p.qw + q) * fabs(det);
        double ph[N];
#pragma unroll
        for (int i = 0; i < N; i++) ph[i] = __ldg(p.phi + (size_t)q * N + i);
#pragma unroll
        for (int i = 0; i < N; i++) {
            T[i] += ph[i] * w;
#pragma unroll
            for (int j = i; j < N; j++) S[i][j] += ph[i] * ph[j] * w;
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = i; j < N; j++)
#pragma unroll
            for (int a = 0; a < NS; a++)
#pragma unroll
                for (int b = 0; b < NS; b++) {
                    if (p.rhs_only) continue;
                    if (i == j && b < a) continue;
                    const double mab = p.coef[a * 3 + b], mba = p.coef[b * 3 + a];
                    if (mab == 0.0 && mba == 0.0) continue;
                    const size_t idx = ((size_t)((i * N + j) * NS + a) * NS + b) * p.nel + el;
                    const int32_t pos = p.smap[idx];
                    if (pos >= 0 && mab != 0.0) scatter_add(p.a + pos, mab * S[i][j], p.atomic);
                    if (p.smapT) {
                        const int32_t posT = p.smapT[idx];
                        if (posT >= 0 && mba != 0.0) scatter_add(p.a + posT, mba * S[i][j], p.atomic);
                    }
                }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int a = 0; a < NS; a++) {
            const double v = p.coef[9 + a];
            if (v != 0.0) scatter_rhs(p.rhs, p.dest[el * (N * NS) + i * NS + a], v * T[i], p.atomic);
        }
}

// boundary faces of higher order (more than 9 shape functions): one WARP per face, runtime sizes.
// lanes <-> integration points for the surface Jacobian, then lanes <-> entries (i <= j) of the face mass matrix.
__global__ void __launch_bounds__(128) assemble_bc_warp_kernel(const BcParams p, int NN, int N, int NS) {
    extern __shared__ double bc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t el = p.el0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (el >= p.el1) return;
    double *W = bc_smem + (size_t)warp * 2 * p.nq;
    double *W2 = W + p.nq;  // w / jac(0,0)^2: weight of the first-axis derivatives (coef[12] != 0 only)
    for (int q = lane; q < p.nq; q += 32) {
        const double *dn = p.dng + (size_t)q * p.fdim * NN;
        double v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0};
        for (int a = 0; a < NN; a++) {
            const int64_t node = p.elnodes[el * NN + a];
            const double d0 = __ldg(dn + a), d1 = p.fdim == 2 ? __ldg(dn + NN + a) : 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double x = p.xyz[node * 3 + k];
                v1[k] += x * d0;
                v2[k] += x * d1;
            }
        }
        double n1 = 0, dot = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            n1 += v1[k] * v1[k];
            dot += v1[k] * v2[k];
        }
        n1 = sqrt(n1);
        double n2 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double v1t = v1[k] / n1;
            const double v2t = v2[k] - dot * v1t / n1;
            n2 += v2t * v2t;
        }
        n2 = sqrt(n2);
        double det = p.fdim == 2 ? n1 * n2 : n1;  // line elements: |dx/dxi|  (Mesh/pzgeoel.cpp:1185-1225)
        if (fabs(det) < 1.e-12) det = 1.e-12;
        W[q] = __ldg(p.qw + q) * fabs(det);
        // dphix(0,i) = jacinv(0,0) dphi(0,i) + jacinv(1,0) dphi(1,i) with jacinv(0,0) = 1 / |dx/dxi|, jacinv(1,0) = 0
        // (Gram-Schmidt Jacobian, Mesh/pzgeoel.cpp:1228-1295 / :1185-1225; dphix = jacinv^T dphi, Mesh/TPZCompElH1.cpp:140-149)
        W2[q] = W[q] / (n1 * n1);
    }
    __syncwarp();
    const double cgrad = p.coef[12];
    const int npair = p.rhs_only ? 0 : N * (N + 1) / 2;
    for (int idx = lane; idx < npair; idx += 32) {
        int i = 0, rem = idx;  // (i, j), i <= j, row-major over the upper triangle
        while (rem >= N - i) { rem -= N - i; i++; }
        const int j = i + rem;
        double S = 0.0;
        for (int q = 0; q < p.nq; q++) S += __ldg(p.phi + (size_t)q * N + i) * __ldg(p.phi + (size_t)q * N + j) * W[q];
        if (cgrad != 0.0) {
            // TPZMatPoisson boundary type 2 (Material/Poisson/TPZMatPoisson.cpp:104-118, one state variable):
            // ek(i,j) += BigNumber * Val1(0,0) * dphix(0,i) * dphix(0,j) * weight
            double S2 = 0.0;
            for (int q = 0; q < p.nq; q++)
                S2 += __ldg(p.dphi + ((size_t)q * p.fdim) * N + i) * __ldg(p.dphi + ((size_t)q * p.fdim) * N + j) * W2[q];
            const size_t sidx = ((size_t)((i * N + j) * NS) * NS) * p.nel + el;
            const int32_t pos = p.smap[sidx];
            if (pos >= 0) scatter_add(p.a + pos, cgrad * S2, p.atomic);
            if (p.smapT && i != j) {
                const int32_t posT = p.smapT[sidx];
                if (posT >= 0) scatter_add(p.a + posT, cgrad * S2, p.atomic);
            }
        }
        for (int a = 0; a < NS; a++)
            for (int b = 0; b < NS; b++) {
                if (i == j && b < a) continue;
                const double mab = p.coef[a * 3 + b], mba = p.coef[b * 3 + a];
                if (mab == 0.0 && mba == 0.0) continue;
                const size_t sidx = ((size_t)((i * N + j) * NS + a) * NS + b) * p.nel + el;
                const int32_t pos = p.smap[sidx];
                if (pos >= 0 && mab != 0.0) scatter_add(p.a + pos, mab * S, p.atomic);
                if (p.smapT) {
                    const int32_t posT = p.smapT[sidx];
                    if (posT >= 0 && mba != 0.0) scatter_add(p.a + posT, mba * S, p.atomic);
                }
            }
    }
    if (p.force) {
        // boundary data given by a function (TPZBndCondT::ForcingFunctionBC, e.g. TPZMatPoisson.cpp:62-64, TPZElasticity3D.cpp:637-662):
        // the host evaluated the coefficient of phi_i * weight in ef at every integration point: force[el][q][a]
        for (int i = lane; i < N; i += 32)
            for (int a = 0; a < NS; a++) {
                double t = 0.0;
                for (int q = 0; q < p.nq; q++) t += __ldg(p.phi + (size_t)q * N + i) * W[q] * p.force[((size_t)el * p.nq + q) * NS + a];
                scatter_rhs(p.rhs, p.dest[el * (N * NS) + i * NS + a], t, p.atomic);
            }
        return;
    }
    for (int i = lane; i < N; i += 32) {
        double T = 0.0;
        for (int q = 0; q < p.nq; q++) T += __ldg(p.phi + (size_t)q * N + i) * W[q];
        for (int a = 0; a < NS; a++) {
            const double v = p.coef[9 + a];
            if (v != 0.0) scatter_rhs(p.rhs, p.dest[el * (N * NS) + i * NS + a], v * T, p.atomic);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// plane (2-D) domain elements: quadrilaterals / triangles of a plane mesh with TPZMatPoisson(dim 2) or TPZElasticity2D
// (Material/Elasticity/TPZElasticity2D.cpp:86-203).  One WARP per element, runtime sizes.
//   lanes <-> integration points: Gram-Schmidt Jacobian, axes (Mesh/pzgeoel.cpp:1228-1295), dphix = jacinv^T dphi in the
//             element's axes (pzinterpolationspace.cpp:1688-1694), for elasticity rotated to x, y with the axes
//             (TPZElasticity2D.cpp:152-153), scaled by sqrt(w|detJ|)  -> shared memory G[q][i][2]
//   lanes <-> node pairs (in <= jn): S[v][u] = sum_q G[q][in][v] G[q][jn][u], then
//             Poisson:      ek(in,jn) = s (S00 + S11)
//             Elasticity2D: ek(2in+a,2jn+b) = a==b ? cA S_aa + cB S_a'a' : cC S_ab + cB S_ba
//             (plane strain: cA = F(1-nu), cB = F(1-2nu)/2, cC = F nu, F = E/((1+nu)(1-2nu)); plane stress: cA = E/(1-nu^2),
//              cB = E/(2(1+nu)), cC = nu E/(1-nu^2): the host passes the three constants)
// coef: Poisson [fScale, force]; Elasticity2D [cA, cB, cC, fx, fy, sxx, sxy, syy]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) assemble_plane_kernel(const BcParams p, int NN, int N, int NS) {
    extern __shared__ double pl_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t el = p.el0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (el >= p.el1) return;
    const int nq = p.nq;
    double *W = pl_smem + (size_t)warp * nq * (2 + 2 * N);  // [nq] w|detJ|
    double *SW = W + nq;                                     // [nq] sqrt(w|detJ|)
    double *G = SW + nq;                                     // [nq][N][2]
    for (int q = lane; q < nq; q += 32) {
        const double *dn = p.dng + (size_t)q * 2 * NN;
        double v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0};
        for (int a = 0; a < NN; a++) {
            const int64_t node = p.elnodes[el * NN + a];
            const double d0 = __ldg(dn + a), d1 = __ldg(dn + NN + a);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double x = p.xyz[node * 3 + k];
                v1[k] += x * d0;
                v2[k] += x * d1;
            }
        }
        double n1 = 0, dot = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            n1 += v1[k] * v1[k];
            dot += v1[k] * v2[k];
        }
        n1 = sqrt(n1);
        double a0[3], a1[3], n2 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            a0[k] = v1[k] / n1;
            a1[k] = v2[k] - dot * a0[k] / n1;
            n2 += a1[k] * a1[k];
        }
        n2 = sqrt(n2);
#pragma unroll
        for (int k = 0; k < 3; k++) a1[k] /= n2;
        const double j00 = n1, j01 = dot / n1, j11 = n2;
        double det = j00 * j11;
        const double i00 = j11 / det, i11 = j00 / det, i01 = -j01 / det;  // jacinv (i10 = 0)
        if (fabs(det) < 1.e-12) det = 1.e-12;
        const double w = __ldg(p.qw + q) * fabs(det);
        const double sw = sqrt(w);
        W[q] = w;
        SW[q] = sw;
        for (int i = 0; i < N; i++) {
            const double d0 = __ldg(p.dphi + ((size_t)q * 2 + 0) * N + i), d1 = __ldg(p.dphi + ((size_t)q * 2 + 1) * N + i);
            const double dx0 = i00 * d0;              // dphidx(0) = jacinv(0,0) dphi0 + jacinv(1,0) dphi1
            const double dx1 = i01 * d0 + i11 * d1;   // dphidx(1) = jacinv(0,1) dphi0 + jacinv(1,1) dphi1
            double g0 = dx0, g1 = dx1;
            if (p.kind == B200ASM_ELASTICITY2D) {     // du = dphix rotated to x, y
                g0 = dx0 * a0[0] + dx1 * a1[0];
                g1 = dx0 * a0[1] + dx1 * a1[1];
            }
            G[((size_t)q * N + i) * 2 + 0] = sw * g0;
            G[((size_t)q * N + i) * 2 + 1] = sw * g1;
        }
    }
    __syncwarp();
    const int npair = p.rhs_only ? 0 : N * (N + 1) / 2;
    for (int idx = lane; idx < npair; idx += 32) {
        int in = 0, rem = idx;
        while (rem >= N - in) { rem -= N - in; in++; }
        const int jn = in + rem;
        double S00 = 0, S01 = 0, S10 = 0, S11 = 0;
        for (int q = 0; q < nq; q++) {
            const double a0 = G[((size_t)q * N + in) * 2], a1 = G[((size_t)q * N + in) * 2 + 1];
            const double b0 = G[((size_t)q * N + jn) * 2], b1 = G[((size_t)q * N + jn) * 2 + 1];
            S00 += a0 * b0; S01 += a0 * b1; S10 += a1 * b0; S11 += a1 * b1;
        }
        double e[2][2];
        if (NS == 1) {
            e[0][0] = p.coef[0] * (S00 + S11);
        } else {
            const double cA = p.coef[0], cB = p.coef[1], cC = p.coef[2];
            e[0][0] = cA * S00 + cB * S11;
            e[0][1] = cC * S01 + cB * S10;
            e[1][0] = cC * S10 + cB * S01;
            e[1][1] = cA * S11 + cB * S00;
        }
        for (int a = 0; a < NS; a++)
            for (int b = 0; b < NS; b++) {
                if (in == jn && b < a) continue;
                const size_t sidx = ((size_t)((in * N + jn) * NS + a) * NS + b) * p.nel + el;
                const int32_t pos = p.smap[sidx];
                if (pos >= 0) scatter_add(p.a + pos, e[a][b], p.atomic);
                if (p.smapT) {
                    const int32_t posT = p.smapT[sidx];
                    if (posT >= 0) scatter_add(p.a + posT, e[a][b], p.atomic);
                }
            }
    }
    for (int i = lane; i < N; i += 32) {
        // tx, ty = sum_q w phi_i f(x_q) (the source: a constant, or the host-evaluated forcing function); sum_q w du_x ; sum_q w du_y
        double tx = 0, ty = 0, gx = 0, gy = 0;
        for (int q = 0; q < nq; q++) {
            const double wp = W[q] * __ldg(p.phi + (size_t)q * N + i);
            if (p.force) {
                const double *f = p.force + ((size_t)el * nq + q) * NS;
                tx += wp * f[0];
                if (NS == 2) ty += wp * f[1];
            } else {
                tx += wp * (NS == 1 ? p.coef[1] : p.coef[3]);
                ty += wp * p.coef[4];
            }
            gx += SW[q] * G[((size_t)q * N + i) * 2];
            gy += SW[q] * G[((size_t)q * N + i) * 2 + 1];
        }
        if (NS == 1) {
            scatter_rhs(p.rhs, p.dest[el * N + i], p.coef[0] * tx, p.atomic);
        } else {
            // ef(2i) += w (fx phi - du_x sxx - du_y sxy) ; ef(2i+1) += w (fy phi - du_x sxy - du_y syy)
            const double sxx = p.coef[5], sxy = p.coef[6], syy = p.coef[7];
            scatter_rhs(p.rhs, p.dest[el * (N * 2) + i * 2], tx - gx * sxx - gy * sxy, p.atomic);
            scatter_rhs(p.rhs, p.dest[el * (N * 2) + i * 2 + 1], ty - gx * sxy - gy * syy, p.atomic);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic volume kernel: ANY topology and ANY shape-function count (orders beyond the specialised kernels, elements whose
// sides carry different orders).  Runtime sizes, one CTA per element, thread <-> node pair (in <= jn):
//   phase 1  thread <-> point: Jacobian, inverse, w|detJ| into shared memory (Mesh/pzgeoel.cpp:1296-1344)
//   phase 2  per pair: for every point dphix = jacinv^T dphi of both functions (TPZCompElH1.cpp:140-149), the sums
//            S[v][u] = sum_q w dphix(v,in) dphix(u,jn), then the entries of TPZMatPoisson.cpp:31-38 / TPZElasticity3D.cpp:318-326
//   phase 3  thread <-> equation: load vector (TPZMatPoisson.cpp:39-40, TPZElasticity3D.cpp:278)
// Nothing is staged per element beyond 11 doubles per point, so the rule and the function count are only bounded by the
// entry-major scatter map (N^2 NS^2 int32 per element, the layout of the boundary / plane kernels).  It re-evaluates dphix per
// pair: ~3x the arithmetic of the tiled kernels, which is the price of having no compile-time size.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) assemble_volume_generic_kernel(const BcParams p, int NN, int N, int NS) {
    extern __shared__ double gv_smem[];
    const int nq = p.nq, tid = threadIdx.x;
    double *JI = gv_smem;                  // [nq][11]: jacinv (9), w|detJ|, unused
    double *X = JI + (size_t)nq * 11;      // [NN][3]
    const int M = N * NS;
    for (int64_t el = p.el0 + blockIdx.x; el < p.el1; el += gridDim.x) {
        __syncthreads();  // the previous element no longer reads JI / X
        for (int i = tid; i < NN * 3; i += blockDim.x) X[i] = p.xyz[(int64_t)p.elnodes[el * NN + i / 3] * 3 + i % 3];
        __syncthreads();
        for (int q = tid; q < nq; q += blockDim.x) {
            const double *dn = p.dng + (size_t)q * 3 * NN;
            double j00 = 0, j01 = 0, j02 = 0, j10 = 0, j11 = 0, j12 = 0, j20 = 0, j21 = 0, j22 = 0;
            for (int a = 0; a < NN; a++) {
                const double d0 = __ldg(dn + a), d1 = __ldg(dn + NN + a), d2 = __ldg(dn + 2 * NN + a);
                const double x = X[a * 3], y = X[a * 3 + 1], z = X[a * 3 + 2];
                j00 += x * d0; j01 += x * d1; j02 += x * d2;
                j10 += y * d0; j11 += y * d1; j12 += y * d2;
                j20 += z * d0; j21 += z * d1; j22 += z * d2;
            }
            double det = 0.0;
            det -= j02 * j11 * j20;
            det += j01 * j12 * j20;
            det += j02 * j10 * j21;
            det -= j00 * j12 * j21;
            det -= j01 * j10 * j22;
            det += j00 * j11 * j22;
            if (fabs(det) < 1.e-12) det = 1.e-12;
            const double id = 1.0 / det;
            double *o = JI + (size_t)q * 11;
            o[0] = (-j12 * j21 + j11 * j22) * id;
            o[1] = (j02 * j21 - j01 * j22) * id;
            o[2] = (-j02 * j11 + j01 * j12) * id;
            o[3] = (j12 * j20 - j10 * j22) * id;
            o[4] = (-j02 * j20 + j00 * j22) * id;
            o[5] = (j02 * j10 - j00 * j12) * id;
            o[6] = (-j11 * j20 + j10 * j21) * id;
            o[7] = (j01 * j20 - j00 * j21) * id;
            o[8] = (-j01 * j10 + j00 * j11) * id;
            o[9] = __ldg(p.qw + q) * fabs(det);
        }
        __syncthreads();
        const int npair = p.rhs_only ? 0 : N * (N + 1) / 2;
        for (int idx = tid; idx < npair; idx += blockDim.x) {
            int in = 0, rem = idx;
            while (rem >= N - in) { rem -= N - in; in++; }
            const int jn = in + rem;
            double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int q = 0; q < nq; q++) {
                const double *ji = JI + (size_t)q * 11;
                const double *dp = p.dphi + (size_t)q * 3 * N;
                const double a0 = __ldg(dp + in), a1 = __ldg(dp + N + in), a2 = __ldg(dp + 2 * N + in);
                const double b0 = __ldg(dp + jn), b1 = __ldg(dp + N + jn), b2 = __ldg(dp + 2 * N + jn);
                double gi[3], gj[3];
#pragma unroll
                for (int v = 0; v < 3; v++) {
                    gi[v] = ji[v] * a0 + ji[3 + v] * a1 + ji[6 + v] * a2;
                    gj[v] = ji[v] * b0 + ji[3 + v] * b1 + ji[6 + v] * b2;
                }
                const double w = ji[9];
                if (NS == 1) {
                    S[0][0] += w * (gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2]);
                } else {
#pragma unroll
                    for (int v = 0; v < 3; v++)
#pragma unroll
                        for (int u = 0; u < 3; u++) S[v][u] += w * gi[v] * gj[u];
                }
            }
            if (NS == 1) {
                const size_t sidx = (size_t)(in * N + jn) * p.nel + el;
                const double e = p.coef[0] * S[0][0];
                const int32_t pos = p.smap[sidx];
                if (pos >= 0) scatter_add(p.a + pos, e, p.atomic);
                if (p.smapT) {
                    const int32_t posT = p.smapT[sidx];
                    if (posT >= 0) scatter_add(p.a + posT, e, p.atomic);
                }
            } else {
                const double C1 = p.coef[0], C2 = p.coef[1], C3 = p.coef[2];
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++) {
                        if (in == jn && b < a) continue;
                        const double e = a == b ? (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3
                                                : S[b][a] * C1 - S[a][b] * C2;
                        const size_t sidx = ((size_t)((in * N + jn) * 3 + a) * 3 + b) * p.nel + el;
                        const int32_t pos = p.smap[sidx];
                        if (pos >= 0) scatter_add(p.a + pos, e, p.atomic);
                        if (p.smapT) {
                            const int32_t posT = p.smapT[sidx];
                            if (posT >= 0) scatter_add(p.a + posT, e, p.atomic);
                        }
                    }
            }
        }
        for (int m = tid; m < M; m += blockDim.x) {
            const int i = m / NS, a = m - i * NS;
            double t = 0.0;
            for (int q = 0; q < nq; q++) {
                const double *ji = JI + (size_t)q * 11;
                const double w = ji[9], ph = __ldg(p.phi + (size_t)q * N + i);
                if (NS == 1) {
                    const double f = p.force ? p.force[(size_t)el * nq + q] : p.coef[1];
                    t += w * p.coef[0] * ph * f;
                } else {
                    const double *dp = p.dphi + (size_t)q * 3 * N;
                    const double dx = ji[a] * __ldg(dp + i) + ji[3 + a] * __ldg(dp + N + i) + ji[6 + a] * __ldg(dp + 2 * N + i);
                    const double f = p.force ? p.force[((size_t)el * nq + q) * 3 + a] : p.coef[3 + a];
                    t += w * f * ph - p.coef[6 + a] * (w * dx);
                }
            }
            scatter_rhs(p.rhs, p.dest[el * M + m], t, p.atomic);
        }
    }
}

__global__ void build_bc_smap_kernel(int64_t nel, int n, int ns, const int32_t *__restrict__ dest,
                                     const int64_t *__restrict__ ia, const int32_t *__restrict__ ja, int symmetric,
                                     int32_t *__restrict__ smap, int32_t *__restrict__ smapT, int *__restrict__ missing) {
    const int m = n * ns;
    const int64_t total = (int64_t)n * n * ns * ns * nel;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx % nel;
        int e = (int)(idx / nel);
        const int b = e % ns; e /= ns;
        const int a = e % ns; e /= ns;
        const int j = e % n;
        const int i = e / n;
        int32_t pos = -1, posT = -1;
        if (i < j || (i == j && a <= b)) {
            const int64_t di = dest[el * m + i * ns + a], dj = dest[el * m + j * ns + b];
            auto find = [&](int64_t row, int64_t col) -> int32_t {
                if (row < 0 || col < 0) return -1;  // equation removed by the TPZEquationFilter: no slot, not an error
                int64_t lo = ia[row], hi = ia[row + 1] - 1;
                while (lo <= hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    const int64_t v = ja[mid];
                    if (v == col) return (int32_t)mid;
                    if (v < col) lo = mid + 1; else hi = mid - 1;
                }
                atomicAdd(missing, 1);
                return -1;
            };
            if (symmetric) {
                pos = find(min(di, dj), max(di, dj));
            } else {
                pos = find(di, dj);
                if (di != dj) posT = find(dj, di);
            }
        }
        smap[idx] = pos;
        if (smapT) smapT[idx] = posT;
    }
}

// ------------------------------------------------------------------------------------------------
// host side: context, groups, dispatch
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_create_error;

struct Group {
    int topology = 0, porder = 0, kind = 0, ns = 1, nn = 0, n = 0, nq = 0, m = 0, dim = 3;
    int64_t nel = 0, nbatch = 0;
    int64_t max_dest = -1;  // largest destination equation of the group
    bool plane = false;     // quadrilaterals / triangles as DOMAIN elements of a plane problem (kind POISSON / ELASTICITY2D)
    bool generic = false;   // volume group without a specialised kernel (other orders, non-uniform side orders): generic kernel
    bool uniform = true;    // nshape is the count of uniform order `porder` (false: the sides carry different orders)
    int cfg = -1;  // index into the dispatch table (register-tile kernels)
    int mma = -1;  // index into the DMMA dispatch table, -1: none
    int aff = -1;  // index into the closed-form table for parallelepiped hexahedra (affine_hex.cuh), -1: none
    int taff = -1;  // index into kTetAffS (tetrahedra p = 3, 4, option variant = 20), -1: none
    bool use_aff = false;      // every element of the group is a parallelepiped (measured): the closed-form kernel runs
    bool aff_checked = false;  // ... for the current node coordinates
    double coef[16];
    int32_t *d_elnodes = nullptr, *d_dest = nullptr, *d_smap = nullptr, *d_smapT = nullptr;
    double *d_qw = nullptr, *d_phi = nullptr, *d_dphi = nullptr, *d_dng = nullptr, *d_force = nullptr;
    double *d_dng_t = nullptr, *d_dphi_pad = nullptr, *d_phi_pad = nullptr, *d_aux = nullptr, *d_aux2 = nullptr;
    bool sumfact_ok = false;  // hexahedra p = 2, Poisson, the 27-point tensor rule: the sum-factorisation kernel applies
    // owner-computes assembly of the closed-form groups (gather_rows.cuh)
    bool use_gather = false;   // the group's matrix part runs gat::gather_rows_kernel (decided by choose_kernels)
    bool gather_bad = false;   // the setup found a requirement violated (node blocks not consecutive, filtered equations, long rows)
    int64_t ng = 0;            // node blocks of the group
    int gather_rl = 0;         // doubles per row buffer
    int32_t *d_grow = nullptr, *d_gptr = nullptr;
    gat::NodeRec *d_grec = nullptr;
    int gather_wpc = 8;        // warps per CTA of the gather kernel (as many as the row buffers leave room for)
    uint32_t *d_glist = nullptr;
    unsigned short *d_relpos = nullptr;
    double *d_fac = nullptr;
    std::vector<int64_t> gchunk, gchunk_min;  // node-block ranges launched apart (overlapped download) and their first rows
    size_t smap_len = 0;
    // element colouring (B200ASM_SCATTER_COLORED): elements are stored sorted by colour; seg = colour boundaries
    // ({0, nel} when not coloured); seg_smap = offset of every segment's scatter map (register-tile kernels)
    std::vector<int64_t> seg;
    std::vector<size_t> seg_smap;
    // overlapped download (b200asm_assemble with a host matrix): element chunks of a large uncoloured group and the
    // smallest destination equation of every chunk (rows below the minimum of all LATER launches are final)
    std::vector<int64_t> chunk;
    std::vector<int64_t> chunk_min;
    // row-sharded assembly: the elements [0, n_if) touch staging rows (options "staging_lo" / "staging_hi"); they are chunk 0,
    // launched before everything else so that their contributions travel while the rest is assembled
    int64_t n_if = 0;
    std::vector<int64_t> stored_order;  // groups with a force table: mesh index of every stored element (b200asm_set_group_force)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // "timing" option: events around this group's launches
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;  // ... around the interface prefix, when it is launched apart
    bool timed_prefix = false;
};

}  // namespace

struct b200asm_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double *d_xyz = nullptr;
    int64_t nnodes = 0;
    std::vector<Group> groups;
    int64_t neq = 0, nnz = 0;
    int symmetric = 1;
    bool have_pattern = false;  // IA/JA (and A, rhs) are on the device
    bool maps_valid = false;    // the scatter maps match the current groups and pattern (rebuilt lazily by assemble)
    int64_t *d_ia = nullptr;
    int32_t *d_ja = nullptr;  // column indices, resident (scatter maps, SpMV of the CG solver)
    double *d_a = nullptr, *d_rhs = nullptr;
    int *d_missing = nullptr;
    // conjugate-gradient workspace (b200asm_cg_solve), allocated on first use for cg_n equations
    int64_t cg_n = 0;
    double *d_cg = nullptr;       // x, r, p, z, q, diag, f: 7 vectors
    double *d_cg_part = nullptr;  // per-CTA partial sums
    cgdev::Scalars *d_cg_sc = nullptr;
    int64_t launches = 0, h2d = 0, d2h = 0;
    int scatter = B200ASM_SCATTER_ATOMIC;
    int engine = 1;  // 1: DMMA panel kernel where one exists, 0: register-tile DFMA kernels only
    int debug = 0;     // profiling aid, see VolParams::debug
    int gather = 0;     // option "gather" (off by default: measured 2-3x slower than the scatter kernels, profiles/r02_gather_vs_scatter.md):
                        // closed-form groups are assembled row by row in a fixed order (gather_rows.cuh) instead of scattered
    unsigned char *d_rowflag = nullptr;  // [neq] see gather_rows.cuh (valid with the scatter maps, when any group gathers)
    bool any_gather = false;
    int drop_tiny = 0;  // option "drop_tiny": the matrix scatter skips |value| < 1e-12 like the reference's AddKel (entry-major kernels only)
    int variant = 0;   // tuning alternative of the DMMA kernels (option "variant", before add_group)
    int rhs_only = 0;  // set while b200asm_assemble_rhs runs
    int timing = 0;  // 1: record CUDA events around every group's launches (b200asm_group_time_ms)
    int locality = 1;  // 1: volume groups are stored along a space-filling curve (within every overlap chunk), see add_group
    std::vector<double> h_xyz;  // host copy of the node coordinates (element centroids for the locality order)
    int affine = 1;   // 1: hexahedral groups whose elements are all parallelepipeds use the closed-form kernel
    int overlap = 1;  // 1: b200asm_assemble downloads the finished rows of A while later element chunks are assembled
    int64_t overlap_min_elements = 8192;      // smallest element chunk (option, before add_group)
    int64_t overlap_min_bytes = 32 << 20;     // smallest piece of A worth its own copy (option)
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    std::vector<int64_t> ov_upto;  // IA at the download frontiers of assemble_overlapped (valid while ov_valid)
    bool ov_valid = false;
    // ---- interface exchange with other GPUs (exchange.cuh) ----
    struct PeerLink {
        bool push = false;   // this context pushes its staged contributions into the peer (false: the peer pushes into this one)
        bool ipc = false;    // the pointers come from cudaIpcOpenMemHandle
        void *ipc_base[3] = {nullptr, nullptr, nullptr};
        double *a = nullptr, *rhs = nullptr;   // the peer's CSR values / load vector (push links)
        unsigned long long *flags = nullptr;   // the peer's flag block
        int slot_there = 0;                    // index of the link to this context in the PEER's table
        int64_t n_a = 0, a_src0 = 0, n_rhs = 0;
        int32_t *d_a_dst = nullptr, *d_rhs_src = nullptr, *d_rhs_dst = nullptr;
    };
    std::vector<PeerLink> links;
    unsigned long long *d_flags = nullptr;  // [2][xch::MAX_LINKS]: ZEROED(step) / PUSHED(step) written by the peer of link i
    int *d_xerr = nullptr;                  // raised by a wait that timed out
    unsigned long long xstep = 0;           // assemblies so far (every participating context counts the same)
    int64_t staging_lo = 0, staging_hi = 0; // local rows [lo, hi) are staging rows (owned by a peer)
    int64_t incoming_min_row = INT64_MAX;   // smallest local row a peer pushes into (caps the early download)
    int64_t xtimeout_ms = 30000;
    // window of the local arrays that goes to the host (options "download_a_count", "download_rhs_first",
    // "download_rhs_count"; -1 = everything): a context of a row-sharded system only returns the rows it owns
    int64_t dl_a = -1, dl_r0 = 0, dl_rn = -1;
    cudaStream_t xstream = nullptr;
    cudaEvent_t ev_if = nullptr, ev_push = nullptr;
    std::string err;
};

namespace {

int fail(b200asm_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, B200ASM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

template <class T>
int upload(b200asm_ctx *ctx, T **dptr, const T *host, size_t count) {
    CK(cudaMalloc((void **)dptr, std::max<size_t>(count, 1) * sizeof(T)));
    if (count) {
        CK(cudaMemcpyAsync(*dptr, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += (int64_t)(count * sizeof(T));
    }
    return 0;
}

// ---- dispatch tables ---------------------------------------------------------------------------
// volume configurations:            NN  N  NS TILE EPB
using HexP1Poisson = VolCfg<8, 8, 1, 8, 64>;
using HexP1Elast = VolCfg<8, 8, 3, 6, 12>;
using HexP2Poisson = VolCfg<8, 27, 1, 9, 16>;
using HexP2Elast = VolCfg<8, 27, 3, 9, 4>;
using HexP3Poisson = VolCfg<8, 64, 1, 8, 4>;
using HexP4Poisson = VolCfg<8, 125, 1, 9, 2>;
using HexP3Elast = VolCfg<8, 64, 3, 6, 1>;
using TetP1Poisson = VolCfg<4, 4, 1, 4, 128>;
using TetP1Elast = VolCfg<4, 4, 3, 6, 32>;
using TetP2Poisson = VolCfg<4, 10, 1, 5, 32>;
using TetP2Elast = VolCfg<4, 10, 3, 6, 8>;
using TetP3Poisson = VolCfg<4, 20, 1, 5, 8>;    // 20 functions: 4 + 6 x 2 (edges) + 4 (faces)
using TetP3Elast = VolCfg<4, 20, 3, 6, 2>;
using TetP4Poisson = VolCfg<4, 35, 1, 7, 4>;    // 35 functions: 4 + 6 x 3 + 4 x 3 + 1 (interior)
using TetP4Elast = VolCfg<4, 35, 3, 9, 1>;
// prisms (6 corner nodes; p=2: 6 + 9 edges + 3 quadrilateral faces = 18 functions) and pyramids (5 corner nodes; p=2: 5 + 8 edges +
// the base = 14 functions): same kernel, their own tables (rational corner functions of the pyramid included)
using PrismP1Poisson = VolCfg<6, 6, 1, 6, 64>;
using PrismP1Elast = VolCfg<6, 6, 3, 6, 16>;
using PrismP2Poisson = VolCfg<6, 18, 1, 6, 16>;
using PrismP2Elast = VolCfg<6, 18, 3, 9, 6>;
using PyrP1Poisson = VolCfg<5, 5, 1, 5, 32>;
using PyrP1Elast = VolCfg<5, 5, 3, 6, 16>;
using PyrP2Poisson = VolCfg<5, 14, 1, 7, 16>;
using PyrP2Elast = VolCfg<5, 14, 3, 6, 4>;

struct VolEntry {
    int topology, porder, ns;
    int epb, tile, slots, nthreads;
    size_t (*smem)(int nq);
    cudaError_t (*launch)(const VolParams &, int grid, size_t smem, cudaStream_t);
    cudaError_t (*launch_smap)(int64_t nel, int64_t nbatch, const int32_t *dest, const int64_t *ia, const int32_t *ja,
                               int symmetric, int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t);
    cudaError_t (*prepare)(size_t smem, int *ctas_per_sm);
};

template <class C>
cudaError_t launch_vol(const VolParams &p, int grid, size_t smem, cudaStream_t s) {
    assemble_volume_kernel<C><<<grid, C::NTHREADS, smem, s>>>(p);
    return cudaGetLastError();
}
template <class C>
cudaError_t launch_vol_smap(int64_t nel, int64_t nbatch, const int32_t *dest, const int64_t *ia, const int32_t *ja,
                            int symmetric, int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_volume_smap_kernel<C><<<grid, 256, 0, s>>>(nel, nbatch, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
template <class C>
cudaError_t prepare_vol(size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(assemble_volume_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_volume_kernel<C>, C::NTHREADS, smem);
}
template <class C>
VolEntry make_entry(int topology, int porder) {
    return VolEntry{topology, porder, C::NS, C::EPB, C::TILE, C::SLOTS, C::NTHREADS, &C::smem_bytes,
                    &launch_vol<C>, &launch_vol_smap<C>, &prepare_vol<C>};
}

const VolEntry kVol[] = {
    make_entry<HexP1Poisson>(B200ASM_HEX, 1), make_entry<HexP1Elast>(B200ASM_HEX, 1),
    make_entry<HexP2Poisson>(B200ASM_HEX, 2), make_entry<HexP2Elast>(B200ASM_HEX, 2),
    make_entry<HexP3Poisson>(B200ASM_HEX, 3), make_entry<HexP4Poisson>(B200ASM_HEX, 4), make_entry<HexP3Elast>(B200ASM_HEX, 3),
    make_entry<TetP1Poisson>(B200ASM_TET, 1), make_entry<TetP1Elast>(B200ASM_TET, 1),
    make_entry<TetP2Poisson>(B200ASM_TET, 2), make_entry<TetP2Elast>(B200ASM_TET, 2),
    make_entry<TetP3Poisson>(B200ASM_TET, 3), make_entry<TetP3Elast>(B200ASM_TET, 3),
    make_entry<TetP4Poisson>(B200ASM_TET, 4), make_entry<TetP4Elast>(B200ASM_TET, 4),
    make_entry<PrismP1Poisson>(B200ASM_PRISM, 1), make_entry<PrismP1Elast>(B200ASM_PRISM, 1),
    make_entry<PrismP2Poisson>(B200ASM_PRISM, 2), make_entry<PrismP2Elast>(B200ASM_PRISM, 2),
    make_entry<PyrP1Poisson>(B200ASM_PYRAMID, 1), make_entry<PyrP1Elast>(B200ASM_PYRAMID, 1),
    make_entry<PyrP2Poisson>(B200ASM_PYRAMID, 2), make_entry<PyrP2Elast>(B200ASM_PYRAMID, 2),
};
constexpr int kNumVol = sizeof(kVol) / sizeof(kVol[0]);

// DMMA panel kernels (gram_mma.cuh):  NN  N  warps/CTA  min CTAs/SM
using HexP2PoissonMma = MmaCfg<8, 27, 8, 2>;
using TetP2PoissonMma = MmaCfg<4, 10, 8, 2>;
using HexP1PoissonMma = MmaCfg<8, 8, 8, 4>;

struct MmaEntry {
    int variant;  // 0 = default; others are tuning alternatives selected with option "variant"
    int topology, porder, ns;
    int slots, nthreads, wpc;
    size_t (*smem)(int nq);
    cudaError_t (*launch)(const VolParams &, int grid, size_t smem, cudaStream_t);
    cudaError_t (*launch_smap)(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                               int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t);
    cudaError_t (*prepare)(size_t smem, int *ctas_per_sm);
    bool sumfact = false;  // sum-factorisation kernel: needs Group::sumfact_ok (the 3 x 3 x 3 tensor rule in the reference's point order)
    bool closed = false;   // closed-form kernel of affine elements (affine_simplex.cuh / affine_hex.cuh): the gather kernel can take over
};
template <class C>
cudaError_t launch_mma(const VolParams &p, int grid, size_t smem, cudaStream_t s) {
    assemble_gram_mma_kernel<C><<<grid, C::WPC * 32, smem, s>>>(p);
    return cudaGetLastError();
}
template <class C>
cudaError_t launch_mma_smap(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                            int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_mma_smap_kernel<C><<<grid, 256, 0, s>>>(nel, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
template <class C>
cudaError_t prepare_mma(size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(assemble_gram_mma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_gram_mma_kernel<C>, C::WPC * 32, smem);
}
template <class C>
MmaEntry make_mma_entry(int topology, int porder, int variant = 0) {
    return MmaEntry{variant, topology, porder, 1, C::SLOTS, C::WPC * 32, C::WPC, &C::smem_bytes, &launch_mma<C>, &launch_mma_smap<C>, &prepare_mma<C>};
}
// team kernels (gram_mma_team.cuh):   NN  N  NS  warps/element  elements/CTA  min CTAs/SM
using TetP2ElastTeamV2 = TeamCfg<4, 10, 3, 3, 4, 2>;
using HexP1ElastTeam = TeamCfg<8, 8, 3, 1, 8, 2>;
using HexP2ElastTeamV3 = TeamCfg<8, 27, 3, 10, 2, 1>;  // DEFAULT: two teams per CTA, 20 warps = 5 per scheduler (a 10-warp CTA leaves
                                                       // 3,3,2,2): 30.3 vs 29.0 M el/s at 64^3; the same change on p4 Poisson lost 3 %
// higher-order Poisson: one warp per 4x4 superblock of 8x8 tiles (p=3: 8x8 tiles -> 3 warps; p=4: 16x16 -> 10 warps)
using HexP3PoissonTeam = TeamCfg<8, 64, 1, 3, 2, 3, 4>;
using HexP4PoissonTeam = TeamCfg<8, 125, 1, 10, 1, 2, 4>;

template <class C>
cudaError_t launch_team(const VolParams &p, int grid, size_t smem, cudaStream_t s) {
    assemble_gram_team_kernel<C><<<grid, C::NTHREADS, smem, s>>>(p);
    return cudaGetLastError();
}
template <class C>
cudaError_t launch_team_smap(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                             int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_team_smap_kernel<C><<<grid, 256, 0, s>>>(nel, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
template <class C>
cudaError_t prepare_team(size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(assemble_gram_team_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_gram_team_kernel<C>, C::NTHREADS, smem);
}
template <class C>
MmaEntry make_team_entry(int topology, int porder, int variant = 0) {
    return MmaEntry{variant, topology, porder, C::NS, C::SLOTS, C::NTHREADS, C::EPC, &C::smem_bytes, &launch_team<C>, &launch_team_smap<C>, &prepare_team<C>};
}
// one warp (or a pair of warps) per element on the team kernel's panel / tiles / scatter map (gram_mma_team.cuh,
// assemble_gram_warp_elast_kernel): WPC warps per CTA, WPE warps per element, GI tile groups in flight per warp
template <class C, int WPC, int WPE, int GI>
size_t warp_elast_smem(int nq) { return WarpElastCfg<C>::smem_bytes(nq, WPC / WPE); }
template <class C, int WPC, int WPE, int GI>
cudaError_t launch_warp_elast(const VolParams &p, int grid, size_t smem, cudaStream_t s) {
    assemble_gram_warp_elast_kernel<C, WPC, WPE, GI><<<grid, WPC * 32, smem, s>>>(p);
    return cudaGetLastError();
}
template <class C, int WPC, int WPE, int GI>
cudaError_t prepare_warp_elast(size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(assemble_gram_warp_elast_kernel<C, WPC, WPE, GI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_gram_warp_elast_kernel<C, WPC, WPE, GI>, WPC * 32, smem);
}
template <class C, int WPC, int WPE, int GI>
MmaEntry make_warp_elast_entry(int topology, int porder, int variant) {
    return MmaEntry{variant, topology, porder, C::NS, C::SLOTS, WPC * 32, WPC / WPE, &warp_elast_smem<C, WPC, WPE, GI>, &launch_warp_elast<C, WPC, WPE, GI>,
                    &launch_team_smap<C>, &prepare_warp_elast<C, WPC, WPE, GI>};
}
// closed-form kernels for straight-sided tetrahedra (affine_simplex.cuh):  N  NS  warps/CTA  min CTAs/SM
using TetP1PoissonAff = AffCfg<4, 1, 8, 3>;
using TetP1ElastAff = AffCfg<4, 3, 8, 3>;
using TetP2PoissonAff = AffCfg<10, 1, 8, 3>;
using TetP2ElastAff = AffCfg<10, 3, 8, 2>;
// (measured on 64^3 x 5: <10,3,8,3> = 85 registers / 24 warps per SM spills and drops to 141 M el/s, <10,3,4,5> = 96 registers: 378)
template <class C>
cudaError_t launch_aff(const VolParams &p, int grid, size_t smem, cudaStream_t s) {
    assemble_affine_simplex_kernel<C><<<grid, C::WPC * 32, smem, s>>>(p);
    return cudaGetLastError();
}
template <class C>
cudaError_t launch_aff_smap(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                            int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_aff_smap_kernel<C><<<grid, 256, 0, s>>>(nel, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
template <class C>
cudaError_t prepare_aff(size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(assemble_affine_simplex_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_affine_simplex_kernel<C>, C::WPC * 32, smem);
}
template <class C>
MmaEntry make_aff_entry(int porder, int variant = 0) {
    MmaEntry e{variant, B200ASM_TET, porder, C::NS, C::SLOTS, C::WPC * 32, C::WPC, &C::smem_bytes, &launch_aff<C>, &launch_aff_smap<C>, &prepare_aff<C>};
    e.closed = true;
    return e;
}
// Ghat[e][f](in,jn) = sum_q w dphi(e,in) dphi(f,jn), cphi[j] = sum_q w phi_j, cd[e][j] = sum_q w dphi(e,j): the aux table
template <class C>
void aff_tables(int nq, const double *qw, const double *phi, const double *dphi, std::vector<double> &aux) {
    aux.assign(C::AUX_LEN, 0.0);
    for (int in = 0; in < C::N; in++)
        for (int jn = in; jn < C::N; jn++)
            for (int e = 0; e < 3; e++)
                for (int f = 0; f < 3; f++) {
                    double v = 0.0;
                    for (int q = 0; q < nq; q++) v += qw[q] * dphi[((size_t)q * 3 + e) * C::N + in] * dphi[((size_t)q * 3 + f) * C::N + jn];
                    aux[C::AUX_G + (e * 3 + f) * C::NPP + C::pair_index(in, jn)] = v;
                }
    for (int j = 0; j < C::N; j++) {
        double v = 0.0;
        for (int q = 0; q < nq; q++) v += qw[q] * phi[(size_t)q * C::N + j];
        aux[C::AUX_CPHI + j] = v;
        for (int e = 0; e < 3; e++) {
            double d = 0.0;
            for (int q = 0; q < nq; q++) d += qw[q] * dphi[((size_t)q * 3 + e) * C::N + j];
            aux[C::AUX_CD + e * C::N + j] = d;
        }
    }
}
// sum-factorisation kernel for hexahedra p = 2, Poisson (sumfact_hex.cuh): one CTA of 64 threads per element
template <int MINB, int PREFETCH>
cudaError_t launch_sumfact(const VolParams &p, int grid, size_t, cudaStream_t s) {
    assemble_sumfact_hex_p2_poisson_kernel<MINB, PREFETCH><<<grid, sf::NTHREADS, 0, s>>>(p);
    return cudaGetLastError();
}
inline cudaError_t launch_sumfact_smap(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                                       int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_sumfact_smap_kernel<0><<<grid, 256, 0, s>>>(nel, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
// the one-warp-per-element form (WPC warps = WPC elements per CTA)
template <int WPC, int MINB>
cudaError_t launch_sumfact_warp(const VolParams &p, int grid, size_t, cudaStream_t s) {
    assemble_sumfact_hex_p2_poisson_warp_kernel<WPC, MINB><<<grid, WPC * 32, 0, s>>>(p);
    return cudaGetLastError();
}
inline cudaError_t launch_sumfact_warp_smap(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                                            int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_sumfact_smap_kernel<1><<<grid, 256, 0, s>>>(nel, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
template <int WPC, int MINB>
cudaError_t prepare_sumfact_warp(size_t, int *ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_sumfact_hex_p2_poisson_warp_kernel<WPC, MINB>, WPC * 32, 0);
}
template <int MINB, int PREFETCH>
cudaError_t prepare_sumfact(size_t, int *ctas_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_sumfact_hex_p2_poisson_kernel<MINB, PREFETCH>,
                                                         sf::NTHREADS, 0);
}
inline size_t sumfact_smem(int) { return 0; }
template <int MINB, int PREFETCH>
MmaEntry make_sumfact_entry(int variant) {
    return MmaEntry{variant, B200ASM_HEX, 2, 1, sf::SLOTS, sf::NTHREADS, 1, &sumfact_smem, &launch_sumfact<MINB, PREFETCH>,
                    &launch_sumfact_smap, &prepare_sumfact<MINB, PREFETCH>, true};
}
template <int WPC, int MINB>
MmaEntry make_sumfact_warp_entry(int variant) {
    return MmaEntry{variant, B200ASM_HEX, 2, 1, sfw::SLOTS, WPC * 32, WPC, &sumfact_smem, &launch_sumfact_warp<WPC, MINB>,
                    &launch_sumfact_warp_smap, &prepare_sumfact_warp<WPC, MINB>, true};
}
// wpc = elements processed concurrently by one CTA; variant 0 = default (the first match wins).  Alternatives kept because a
// test or a profile refers to them: 7 = the DMMA Gram kernels for tetrahedra that the closed-form kernels replaced; for hexahedra
// p = 2 Poisson 20 = the default again (sum factorisation, one warp per element), 13 = sum factorisation with one CTA of 64
// threads per element (the default of the first round-2 passes), 16 = the one-warp DMMA Gram kernel that was the default in round
// 1 (it still runs every such group whose rule is not the 3 x 3 x 3 tensor rule); for hexahedra p = 2 elasticity 31 = the default
// again (a pair of warps per element), 30 = one warp per element, 34 = the team of ten warps per element that was the default
// before (64^3 perturbed grid, profiles/r02_elast_warp_variants.jsonl: 36.6 / 32.2 / 29.3 M elements/s).  Measured on a 96^3 perturbed grid
// (profiles/r02_sumfact_warp_variants.jsonl): 20: 286 M elements/s, 13: 219, 16: 159; the other alternatives lost and were removed.
const MmaEntry kMma[] = {make_aff_entry<TetP1PoissonAff>(1), make_aff_entry<TetP1ElastAff>(1),
                         make_aff_entry<TetP2PoissonAff>(2), make_aff_entry<TetP2ElastAff>(2),
                         make_sumfact_warp_entry<4, 6>(0),
                         make_mma_entry<HexP2PoissonMma>(B200ASM_HEX, 2), make_mma_entry<TetP2PoissonMma>(B200ASM_TET, 2, 7),
                         make_warp_elast_entry<HexP2ElastTeamV3, 8, 2, 1>(B200ASM_HEX, 2, 0), make_team_entry<HexP1ElastTeam>(B200ASM_HEX, 1),
                         make_team_entry<HexP3PoissonTeam>(B200ASM_HEX, 3), make_team_entry<HexP4PoissonTeam>(B200ASM_HEX, 4),
                         make_mma_entry<HexP1PoissonMma>(B200ASM_HEX, 1, 0),
                         make_mma_entry<HexP2PoissonMma>(B200ASM_HEX, 2, 16),
                         make_team_entry<TetP2ElastTeamV2>(B200ASM_TET, 2, 7),
                         make_sumfact_entry<8, 1>(13), make_sumfact_warp_entry<4, 6>(20),
                         make_warp_elast_entry<HexP2ElastTeamV3, 4, 1, 2>(B200ASM_HEX, 2, 30), make_warp_elast_entry<HexP2ElastTeamV3, 8, 2, 1>(B200ASM_HEX, 2, 31),
                         make_team_entry<HexP2ElastTeamV3>(B200ASM_HEX, 2, 34)};
// (tetrahedra p=2 elasticity, DMMA team kernel: 4 teams of 3 warps per CTA, 24 warps/SM: 208 M el/s vs 135 M el/s for the
//  register-tile kernel on a 64^3x5 mesh, although padding 10 shape functions to 16 wastes 60 % of every DMMA tile)
constexpr int kNumMma = sizeof(kMma) / sizeof(kMma[0]);

// closed-form kernels for parallelepiped hexahedra (affine_hex.cuh):  N  NS  warps/CTA  min CTAs/SM
using HexP1PoissonAff = AffHexCfg<8, 1, 8, 3>;
using HexP1ElastAff = AffHexCfg<8, 3, 8, 3>;
using HexP2PoissonAff = AffHexCfg<27, 1, 8, 3>;
using HexP2ElastAff = AffHexCfg<27, 3, 8, 3>;
template <class C>
cudaError_t launch_affhex(const VolParams &p, int grid, size_t smem, cudaStream_t s) {
    assemble_affine_hex_kernel<C><<<grid, C::WPC * 32, smem, s>>>(p);
    return cudaGetLastError();
}
template <class C>
cudaError_t launch_affhex_smap(int64_t nel, const int32_t *dest, const int64_t *ia, const int32_t *ja, int symmetric,
                               int32_t *smap, int32_t *smapT, int *missing, int grid, cudaStream_t s) {
    build_affhex_smap_kernel<C><<<grid, 256, 0, s>>>(nel, dest, ia, ja, symmetric, smap, smapT, missing);
    return cudaGetLastError();
}
template <class C>
cudaError_t prepare_affhex(size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(assemble_affine_hex_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, assemble_affine_hex_kernel<C>, C::WPC * 32, smem);
}
template <class C>
MmaEntry make_affhex_entry(int porder, int topology = B200ASM_HEX, int variant = 0) {
    MmaEntry e{variant, topology, porder, C::NS, C::SLOTS, C::WPC * 32, C::WPC, &C::smem_bytes, &launch_affhex<C>, &launch_affhex_smap<C>, &prepare_affhex<C>};
    e.closed = true;
    return e;
}
// the same closed-form kernel on straight-sided tetrahedra of order 3, 4 (table in shared memory): the default of these orders
// since round 2 (profiles/r02_time_tet_closed_form.jsonl: p3 Poisson 489 M elements/s against 108 M of the register-tile
// kernel, p3 elasticity 55 against 16, p4 20-74 against 3-14; parity profiles/r02_tet_closed_form_parity.jsonl).  Option
// variant = 21 keeps the register-tile kernel.
constexpr int kTetRegTileVariant = 21;
using TetP3PoissonAffS = AffHexCfg<20, 1, 8, 3, 4>;
using TetP3ElastAffS = AffHexCfg<20, 3, 8, 2, 4>;
using TetP4PoissonAffS = AffHexCfg<35, 1, 8, 2, 4>;
using TetP4ElastAffS = AffHexCfg<35, 3, 8, 1, 4>;
const MmaEntry kTetAffS[] = {make_affhex_entry<TetP3PoissonAffS>(3, B200ASM_TET), make_affhex_entry<TetP3ElastAffS>(3, B200ASM_TET),
                             make_affhex_entry<TetP4PoissonAffS>(4, B200ASM_TET), make_affhex_entry<TetP4ElastAffS>(4, B200ASM_TET)};
constexpr int kNumTetAffS = sizeof(kTetAffS) / sizeof(kTetAffS[0]);
const MmaEntry kAffHex[] = {make_affhex_entry<HexP1PoissonAff>(1), make_affhex_entry<HexP1ElastAff>(1),
                            make_affhex_entry<HexP2PoissonAff>(2), make_affhex_entry<HexP2ElastAff>(2)};
constexpr int kNumAffHex = sizeof(kAffHex) / sizeof(kAffHex[0]);
// volume group run by assemble_volume_generic_kernel (no specialised kernel, or engine 2 = the generic kernel everywhere)
bool runs_generic(const b200asm_ctx *ctx, const Group &g) { return g.dim == 3 && (g.generic || ctx->engine == 2 || ctx->drop_tiny); }
// groups whose scatter map is entry-major ([entry][element]): boundary elements, plane elements, generic volume groups
bool entry_major(const b200asm_ctx *ctx, const Group &g) { return g.kind == B200ASM_BC || g.plane || runs_generic(ctx, g); }
// the kernel that runs a volume group on the DMMA / closed-form engine (nullptr: register-tile kernel)
const MmaEntry *fast_entry(const b200asm_ctx *ctx, const Group &g) {
    if (ctx->engine != 1) return nullptr;
    if (g.taff >= 0) return &kTetAffS[g.taff];
    if (g.use_aff && g.aff >= 0) return &kAffHex[g.aff];
    return g.mma >= 0 ? &kMma[g.mma] : nullptr;
}

template <int NN, int N, int NS>
cudaError_t launch_bc(const BcParams &p, cudaStream_t s) {
    const int grid = (int)((p.el1 - p.el0 + 127) / 128);
    assemble_bc_kernel<NN, N, NS><<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}

// porder 0: the sides of the elements carry different orders (n functions): runtime-size warp kernel
cudaError_t dispatch_bc(int topology, int porder, int nn, int n, int ns, const BcParams &p, cudaStream_t s) {
    if (topology == B200ASM_LINE) {  // boundary of a plane problem
        const int grid = (int)((p.el1 - p.el0 + 3) / 4);
        assemble_bc_warp_kernel<<<grid, 128, 8 * (size_t)p.nq * sizeof(double), s>>>(p, 2, n, ns);
        return cudaGetLastError();
    }
    if (ns == 2) return cudaErrorInvalidValue;
    if (porder >= 3 || porder == 0 || p.force || p.coef[12] != 0.0) {  // (a table of boundary data, or the gradient term: the runtime-size kernel)
        if (topology != B200ASM_QUAD && topology != B200ASM_TRI) return cudaErrorInvalidValue;
        const int grid = (int)((p.el1 - p.el0 + 3) / 4);
        assemble_bc_warp_kernel<<<grid, 128, 8 * (size_t)p.nq * sizeof(double), s>>>(p, nn, n, ns);
        return cudaGetLastError();
    }
    if (topology == B200ASM_QUAD && porder == 1) return ns == 1 ? launch_bc<4, 4, 1>(p, s) : launch_bc<4, 4, 3>(p, s);
    if (topology == B200ASM_QUAD && porder == 2) return ns == 1 ? launch_bc<4, 9, 1>(p, s) : launch_bc<4, 9, 3>(p, s);
    if (topology == B200ASM_TRI && porder == 1) return ns == 1 ? launch_bc<3, 3, 1>(p, s) : launch_bc<3, 3, 3>(p, s);
    if (topology == B200ASM_TRI && porder == 2) return ns == 1 ? launch_bc<3, 6, 1>(p, s) : launch_bc<3, 6, 3>(p, s);
    return cudaErrorInvalidValue;
}

int nshape_of(int topology, int p) { return b200asm_nshape(topology, p); }
int ncorner_of(int topology) {
    if (topology == B200ASM_LINE) return 2;
    switch (topology) {
        case B200ASM_HEX: return 8;
        case B200ASM_TET: return 4;
        case B200ASM_QUAD: return 4;
        case B200ASM_TRI: return 3;
        case B200ASM_PRISM: return 6;
        case B200ASM_PYRAMID: return 5;
    }
    return -1;
}

// Greedy element colouring: two elements that share an equation get different colours (the idea of the
// reference's TPZStructMatrix::ComputeElementColors, StrMatrix/TPZStructMatrix.cpp:104-161, with a 64-bit colour
// mask per equation instead of its O(nel * ncolours) sweeps).  Returns the number of colours, -1 if more than 64.
int colour_elements(int64_t nel, int m, const int64_t *dest, int64_t neq_hint, std::vector<int32_t> &colour) {
    int64_t neq = neq_hint;
    for (int64_t k = 0; k < nel * m; k++) neq = std::max<int64_t>(neq, dest[k] + 1);
    std::vector<uint64_t> used((size_t)neq + 1, 0);
    uint64_t *usedp = used.data() + 1;  // index -1 (filtered equations) is a harmless scratch slot
    colour.assign((size_t)nel, 0);
    int ncol = 0;
    for (int64_t el = 0; el < nel; el++) {
        uint64_t u = 0;
        for (int k = 0; k < m; k++)
            if (dest[el * m + k] >= 0) u |= usedp[dest[el * m + k]];
        if (~u == 0) return -1;
        const int c = __builtin_ctzll(~u);
        colour[el] = c;
        ncol = std::max(ncol, c + 1);
        const uint64_t bit = 1ull << c;
        for (int k = 0; k < m; k++) usedp[dest[el * m + k]] |= bit;
    }
    return ncol;
}

void free_group(Group &g) {
    cudaFree(g.d_elnodes); cudaFree(g.d_dest); cudaFree(g.d_smap); cudaFree(g.d_smapT);
    cudaFree(g.d_qw); cudaFree(g.d_phi); cudaFree(g.d_dphi); cudaFree(g.d_dng); cudaFree(g.d_force);
    cudaFree(g.d_dng_t); cudaFree(g.d_dphi_pad); cudaFree(g.d_phi_pad); cudaFree(g.d_aux); cudaFree(g.d_aux2);
    cudaFree(g.d_grow); cudaFree(g.d_gptr); cudaFree(g.d_glist); cudaFree(g.d_relpos); cudaFree(g.d_fac); cudaFree(g.d_grec);
    if (g.ev0) cudaEventDestroy(g.ev0);
    if (g.ev1) cudaEventDestroy(g.ev1);
    if (g.ev2) cudaEventDestroy(g.ev2);
    if (g.ev3) cudaEventDestroy(g.ev3);
    g = Group();
}

// Hexahedral groups of order <= 2: measure (once per set of node coordinates) whether every element is a parallelepiped;
// those groups run the closed-form kernel of affine_hex.cuh.  A change of the decision changes the scatter-map layout.
int choose_kernels(b200asm_ctx *ctx) {
    for (Group &g : ctx->groups) {
        if (g.aff < 0 || g.aff_checked) continue;
        bool aff = false;
        if (ctx->affine && ctx->engine == 1 && !g.generic && g.nel > 0) {
            CK(cudaMemsetAsync(ctx->d_missing, 0, sizeof(int), ctx->stream));
            const int grid = (int)std::min<int64_t>((g.nel + 127) / 128, (int64_t)ctx->num_sms * 16);
            hex_affinity_kernel<<<grid, 128, 0, ctx->stream>>>(g.nel, g.d_elnodes, ctx->d_xyz, ctx->d_missing);
            CK(cudaGetLastError());
            ctx->launches++;
            int flag = 1;
            CK(cudaMemcpyAsync(&flag, ctx->d_missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            aff = flag == 0;
        }
        if (aff != g.use_aff) {
            g.use_aff = aff;
            ctx->maps_valid = false;
        }
        g.aff_checked = true;
    }
    // closed-form groups of order <= 2 are assembled row by row (gather_rows.cuh) when nothing else writes into their rows
    // concurrently: one GPU (no staging rows, no peers), atomic mode (the coloured mode keeps its own deterministic launches)
    for (Group &g : ctx->groups) {
        bool want = false;
        if (ctx->gather && ctx->engine == 1 && !ctx->drop_tiny && ctx->scatter == B200ASM_SCATTER_ATOMIC && ctx->links.empty() &&
            ctx->staging_lo == ctx->staging_hi && g.dim == 3 && !g.generic && g.uniform && g.porder <= 2 && g.nel > 0 &&
            g.nel < (1 << 27) && g.d_aux && !g.gather_bad && (g.topology == B200ASM_HEX || g.topology == B200ASM_TET)) {
            const MmaEntry *fe = fast_entry(ctx, g);
            want = fe && fe->closed;
        }
        if (want != g.use_gather) {
            g.use_gather = want;
            ctx->maps_valid = false;
        }
    }
    return 0;
}

// warps of gat::gather_rows_kernel<n, ns> whose buffers fit the shared memory of an SM next to the table (0: no such kernel)
template <int N, int NS>
int gather_warps_fit(int rl) {
    using C = gat::Cfg<N, NS>;
    const size_t avail = 226 * 1024, table = sizeof(double) * C::GT_LEN, per_warp = sizeof(double) * C::warp_doubles(rl);
    return table >= avail ? 0 : (int)((avail - table) / per_warp);
}
int gather_max_warps(int n, int ns, int rl) {
    if (n == 4) return ns == 1 ? gather_warps_fit<4, 1>(rl) : gather_warps_fit<4, 3>(rl);
    if (n == 8) return ns == 1 ? gather_warps_fit<8, 1>(rl) : gather_warps_fit<8, 3>(rl);
    if (n == 10) return ns == 1 ? gather_warps_fit<10, 1>(rl) : gather_warps_fit<10, 3>(rl);
    if (n == 27) return ns == 1 ? gather_warps_fit<27, 1>(rl) : gather_warps_fit<27, 3>(rl);
    return 0;
}

// Gather structures of one closed-form group (gather_rows.cuh).  Returns 1 when a requirement is violated (the group then
// keeps its scatter kernel), 0 when built, < 0 on errors.
int build_gather(b200asm_ctx *ctx, Group &g, const int32_t *d_ja) {
    auto drop = [&]() {
        cudaFree(g.d_grow); cudaFree(g.d_gptr); cudaFree(g.d_glist); cudaFree(g.d_relpos); cudaFree(g.d_fac); cudaFree(g.d_grec);
        g.d_grow = g.d_gptr = nullptr; g.d_glist = nullptr; g.d_relpos = nullptr; g.d_fac = nullptr; g.d_grec = nullptr;
    };
    drop();
    const int64_t neq = ctx->neq;
    const int64_t npairs = g.nel * g.n;
    cudaStream_t st = ctx->stream;
    int32_t *d_cnt = nullptr, *d_slot = nullptr, *d_flag = nullptr;
    int *d_misc = nullptr;  // [0] bad, [1] missing, [2] max row length
    void *d_tmp = nullptr;
    auto cleanup = [&]() { cudaFree(d_cnt); cudaFree(d_slot); cudaFree(d_flag); cudaFree(d_misc); cudaFree(d_tmp); };
#define CKG(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            cleanup();                                                                                   \
            return fail(ctx, B200ASM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                                \
    } while (0)
    CKG(cudaMalloc((void **)&d_cnt, (size_t)(neq + 1) * sizeof(int32_t)));
    CKG(cudaMalloc((void **)&d_slot, (size_t)(neq + 1) * sizeof(int32_t)));
    CKG(cudaMalloc((void **)&d_flag, (size_t)(neq + 1) * sizeof(int32_t)));
    CKG(cudaMalloc((void **)&d_misc, 3 * sizeof(int)));
    CKG(cudaMalloc((void **)&g.d_gptr, (size_t)(neq + 1) * sizeof(int32_t)));
    CKG(cudaMemsetAsync(d_cnt, 0, (size_t)(neq + 1) * sizeof(int32_t), st));
    CKG(cudaMemsetAsync(d_flag, 0, (size_t)(neq + 1) * sizeof(int32_t), st));
    CKG(cudaMemsetAsync(d_misc, 0, 3 * sizeof(int), st));
    const int grid = (int)std::min<int64_t>((npairs + 255) / 256, (int64_t)ctx->num_sms * 32);
    const int grid_eq = (int)std::max<int64_t>(1, std::min<int64_t>((neq + 255) / 256, (int64_t)ctx->num_sms * 32));
    gat::count_kernel<<<grid, 256, 0, st>>>(g.nel, g.n, g.ns, g.d_dest, d_cnt, d_misc);
    gat::flag_nonempty_kernel<<<grid_eq, 256, 0, st>>>(neq, d_cnt, d_flag);
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, g.d_gptr, (int)(neq + 1), st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_flag, d_slot, (int)(neq + 1), st);
    tmp_bytes = std::max(tmp_bytes, tmp2);
    CKG(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1)));
    CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, g.d_gptr, (int)(neq + 1), st));
    CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_slot, (int)(neq + 1), st));
    int misc[3] = {0, 0, 0};
    int32_t ng32 = 0;
    CKG(cudaMemcpyAsync(misc, d_misc, sizeof(misc), cudaMemcpyDeviceToHost, st));
    CKG(cudaMemcpyAsync(&ng32, d_slot + neq, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
    ctx->launches += 4;
    if (misc[0]) {  // node blocks not consecutive / filtered equations
        cleanup();
        drop();
        return 1;
    }
    g.ng = ng32;
    CKG(cudaMalloc((void **)&g.d_grow, std::max<size_t>((size_t)g.ng, 1) * sizeof(int32_t)));
    CKG(cudaMalloc((void **)&g.d_grec, std::max<size_t>((size_t)g.ng, 1) * sizeof(gat::NodeRec)));
    CKG(cudaMalloc((void **)&g.d_glist, std::max<size_t>((size_t)npairs, 1) * sizeof(uint32_t)));
    CKG(cudaMalloc((void **)&g.d_relpos, std::max<size_t>((size_t)npairs * g.ns * g.n, 1) * sizeof(unsigned short)));
    CKG(cudaMalloc((void **)&g.d_fac, std::max<size_t>((size_t)g.nel * (g.ns == 1 ? 6 : 10), 1) * sizeof(double)));
    gat::compact_keys_kernel<<<grid_eq, 256, 0, st>>>(neq, d_cnt, d_slot, g.d_grow);
    gat::node_records_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>((g.ng + 255) / 256, (int64_t)ctx->num_sms * 32)), 256, 0, st>>>(
        g.ng, g.d_grow, g.d_gptr, g.d_grec, d_misc);
    CKG(cudaMemsetAsync(d_flag, 0, (size_t)(neq + 1) * sizeof(int32_t), st));  // (reused as the fill cursor)
    gat::fill_kernel<<<grid, 256, 0, st>>>(g.nel, g.n, g.ns, g.d_dest, g.d_gptr, d_flag, g.d_glist);
    const int grid_g = (int)std::max<int64_t>(1, std::min<int64_t>((g.ng + 127) / 128, (int64_t)ctx->num_sms * 32));
    gat::sort_groups_kernel<<<grid_g, 128, 0, st>>>(g.ng, g.d_grow, g.d_gptr, g.d_glist);
    const int grid_w = (int)std::max<int64_t>(1, std::min<int64_t>((g.ng + gat::WPC - 1) / gat::WPC, (int64_t)ctx->num_sms * 16));
    gat::relpos_kernel<<<grid_w, gat::WPC * 32, 0, st>>>(g.ng, g.n, g.ns, ctx->symmetric, g.d_grow, g.d_gptr, g.d_glist, g.d_dest, ctx->d_ia, d_ja,
                                                          g.d_relpos, d_misc + 1, d_misc, d_misc + 2);
    CKG(cudaGetLastError());
    CKG(cudaMemcpyAsync(misc, d_misc, sizeof(misc), cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
    ctx->launches += 4;
    if (misc[1]) {
        cleanup();
        return fail(ctx, B200ASM_EPATTERN, "gather map: " + std::to_string(misc[1]) + " element entries have no position in the CSR pattern");
    }
    // per-warp row buffers: ns rows of the longest row for every node block a warp handles side by side; the table and at
    // least 4 warps must fit the shared memory of an SM (rows too long for that or for the 16-bit positions, nodes with more
    // than gat::MAXDEG elements: the group keeps its scatter kernel)
    g.gather_rl = ((misc[2] + 3) / 4) * 4 + 4;
    const int wmax = gather_max_warps(g.n, g.ns, g.gather_rl);
    if (misc[0] || wmax < 4) {
        cleanup();
        drop();
        return 1;
    }
    g.gather_wpc = std::min(wmax, 24);
    // node-block ranges for the overlapped download: the rows of a range are final when its launch has finished
    g.gchunk.assign(1, 0);
    if (g.nel >= 2 * ctx->overlap_min_elements) {
        const int nchunk = (int)std::min<int64_t>(16, g.nel / ctx->overlap_min_elements);
        for (int c = 1; c < nchunk; c++) {
            const int64_t b = g.ng * c / nchunk;
            if (b > g.gchunk.back() && b < g.ng) g.gchunk.push_back(b);
        }
    }
    g.gchunk.push_back(g.ng);
    g.gchunk_min.assign(g.gchunk.size() - 1, 0);
    for (size_t c = 0; c + 1 < g.gchunk.size(); c++) {
        int32_t key = 0;
        if (g.gchunk[c] < g.ng) CKG(cudaMemcpyAsync(&key, g.d_grow + g.gchunk[c], sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CKG(cudaStreamSynchronize(st));
        g.gchunk_min[c] = key;
    }
    cleanup();
    return 0;
#undef CKG
}

int build_smaps(b200asm_ctx *ctx, const int32_t *d_ja) {
    for (const Group &g : ctx->groups)
        if (g.max_dest >= ctx->neq)
            return fail(ctx, B200ASM_EINVAL, "a group's destination indices exceed the equations of the pattern (pattern and groups of different meshes?)");
    if (ctx->d_xyz) {  // the layout of a group's map follows the kernel that will run it
        const int rc = choose_kernels(ctx);
        if (rc) return rc;
    }
    // closed-form groups that gather: their row lists replace the scatter map; a group that does not qualify falls back
    ctx->any_gather = false;
    for (Group &g : ctx->groups) {
        if (!g.use_gather) continue;
        const int rc = build_gather(ctx, g, d_ja);
        if (rc < 0) return rc;
        if (rc == 1) {
            g.gather_bad = true;
            g.use_gather = false;
        } else {
            ctx->any_gather = true;
        }
    }
    cudaFree(ctx->d_rowflag);
    ctx->d_rowflag = nullptr;
    if (ctx->any_gather) {
        // rows another group contributes to are zeroed and added to; the rest of a gathering group's rows are stored once
        unsigned char *d_touched = nullptr;
        CK(cudaMalloc((void **)&ctx->d_rowflag, std::max<int64_t>(ctx->neq, 1)));
        CK(cudaMalloc((void **)&d_touched, std::max<int64_t>(ctx->neq, 1)));
        CK(cudaMemsetAsync(ctx->d_rowflag, 0, std::max<int64_t>(ctx->neq, 1), ctx->stream));
        CK(cudaMemsetAsync(d_touched, 0, std::max<int64_t>(ctx->neq, 1), ctx->stream));
        for (Group &g : ctx->groups) {
            if (g.use_gather || g.nel == 0) continue;
            const int64_t n = g.nel * g.m;
            gat::mark_rows_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 32), 256, 0, ctx->stream>>>(n, g.d_dest, d_touched);
            ctx->launches++;
        }
        for (Group &g : ctx->groups) {
            if (!g.use_gather) continue;
            const int grid_g = (int)std::max<int64_t>(1, std::min<int64_t>((g.ng + 127) / 128, (int64_t)ctx->num_sms * 32));
            gat::flag_rows_kernel<<<grid_g, 128, 0, ctx->stream>>>(g.ng, g.ns, g.d_grow, d_touched, ctx->d_rowflag);
            ctx->launches++;
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_touched);
    }
    CK(cudaMemsetAsync(ctx->d_missing, 0, sizeof(int), ctx->stream));
    for (Group &g : ctx->groups) {
        cudaFree(g.d_smap); cudaFree(g.d_smapT);
        g.d_smap = g.d_smapT = nullptr;
        if (g.use_gather) {
            g.smap_len = 0;
            continue;
        }
        if (entry_major(ctx, g)) {
            g.smap_len = (size_t)g.n * g.n * g.ns * g.ns * g.nel;
        } else if (const MmaEntry *me = fast_entry(ctx, g)) {
            g.smap_len = (size_t)g.nel * me->slots;
        } else {
            // batches never straddle a colour: every segment has its own batches and its own piece of the map
            const VolEntry &ve = kVol[g.cfg];
            g.seg_smap.assign(g.seg.size(), 0);
            for (size_t c = 0; c + 1 < g.seg.size(); c++) {
                const int64_t nb = (g.seg[c + 1] - g.seg[c] + ve.epb - 1) / ve.epb;
                g.seg_smap[c + 1] = g.seg_smap[c] + (size_t)nb * ve.tile * ve.tile * ve.slots;
            }
            g.smap_len = g.seg_smap.back();
        }
        CK(cudaMalloc((void **)&g.d_smap, std::max<size_t>(g.smap_len, 1) * sizeof(int32_t)));
        if (!ctx->symmetric) CK(cudaMalloc((void **)&g.d_smapT, std::max<size_t>(g.smap_len, 1) * sizeof(int32_t)));
        if (g.smap_len == 0) continue;
        const int grid = (int)std::min<size_t>((g.smap_len + 255) / 256, (size_t)ctx->num_sms * 32);
        if (entry_major(ctx, g)) {
            build_bc_smap_kernel<<<grid, 256, 0, ctx->stream>>>(g.nel, g.n, g.ns, g.d_dest, ctx->d_ia, d_ja, ctx->symmetric,
                                                                g.d_smap, g.d_smapT, ctx->d_missing);
            CK(cudaGetLastError());
        } else if (const MmaEntry *me = fast_entry(ctx, g)) {
            CK(me->launch_smap(g.nel, g.d_dest, ctx->d_ia, d_ja, ctx->symmetric, g.d_smap, g.d_smapT, ctx->d_missing, grid, ctx->stream));
        } else {
            const VolEntry &ve = kVol[g.cfg];
            for (size_t c = 0; c + 1 < g.seg.size(); c++) {
                const int64_t n = g.seg[c + 1] - g.seg[c];
                if (n == 0) continue;
                const int64_t nb = (n + ve.epb - 1) / ve.epb;
                CK(ve.launch_smap(n, nb, g.d_dest + g.seg[c] * g.m, ctx->d_ia, d_ja, ctx->symmetric, g.d_smap + g.seg_smap[c],
                                  g.d_smapT ? g.d_smapT + g.seg_smap[c] : nullptr, ctx->d_missing, grid, ctx->stream));
            }
        }
        ctx->launches++;
    }
    int missing = 0;
    CK(cudaMemcpyAsync(&missing, ctx->d_missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (missing)
        return fail(ctx, B200ASM_EPATTERN, "scatter map: " + std::to_string(missing) +
                                               " element entries have no position in the CSR pattern");
    ctx->maps_valid = true;
    ctx->ov_valid = false;
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" const char *b200asm_last_error(const b200asm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int b200asm_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return ndev;
}

extern "C" int b200asm_create(b200asm_ctx **out, int device) {
    b200asm_ctx *ctx = nullptr;
    if (!out) return fail(ctx, B200ASM_EINVAL, "b200asm_create: out == NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, B200ASM_ENODEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) +
                                                "): b200asm has no CPU path");
    if (device < 0 || device >= ndev) return fail(ctx, B200ASM_EINVAL, "b200asm_create: bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(ctx, B200ASM_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    b200asm_ctx *c = new b200asm_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&c->d_missing, sizeof(int)) != cudaSuccess) {
        delete c;
        return fail(ctx, B200ASM_ECUDA, "b200asm_create: stream/alloc failed");
    }
    c->own_stream = true;
    *out = c;
    return 0;
}

extern "C" void b200asm_destroy(b200asm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (Group &g : ctx->groups) free_group(g);
    b200asm_exchange_clear(ctx);
    cudaFree(ctx->d_flags); cudaFree(ctx->d_xerr);
    if (ctx->xstream) cudaStreamDestroy(ctx->xstream);
    if (ctx->ev_if) cudaEventDestroy(ctx->ev_if);
    if (ctx->ev_push) cudaEventDestroy(ctx->ev_push);
    cudaFree(ctx->d_xyz); cudaFree(ctx->d_ia); cudaFree(ctx->d_ja); cudaFree(ctx->d_a); cudaFree(ctx->d_rhs); cudaFree(ctx->d_missing);
    cudaFree(ctx->d_rowflag);
    cudaFree(ctx->d_cg); cudaFree(ctx->d_cg_part); cudaFree(ctx->d_cg_sc);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
    delete ctx;
}

extern "C" int b200asm_set_stream(b200asm_ctx *ctx, void *cuda_stream) {
    if (!ctx) return B200ASM_EINVAL;
    if (ctx->own_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return 0;
}

extern "C" int b200asm_set_option(b200asm_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return B200ASM_EINVAL;
    if (!strcmp(name, "scatter")) {
        if (value != B200ASM_SCATTER_ATOMIC && value != B200ASM_SCATTER_COLORED) return fail(ctx, B200ASM_EINVAL, "scatter: unknown mode");
        if (!ctx->groups.empty()) return fail(ctx, B200ASM_ESTATE, "scatter: set the mode before the first b200asm_add_group");
        ctx->scatter = (int)value;
        return 0;
    }
    if (!strcmp(name, "engine")) {
        if (value < 0 || value > 2) return fail(ctx, B200ASM_EINVAL, "engine: 0 (register tiles), 1 (DMMA / closed form where available) or 2 (generic runtime-size kernel)");
        ctx->engine = (int)value;
        ctx->maps_valid = false;  // the scatter-map layout depends on the kernel
        for (Group &g : ctx->groups) g.aff_checked = false;
        return 0;
    }
    if (!strcmp(name, "debug")) {
        ctx->debug = (int)value;
        return 0;
    }
    if (!strcmp(name, "gather")) {
        ctx->gather = value ? 1 : 0;
        ctx->maps_valid = false;
        for (Group &g : ctx->groups) g.aff_checked = false;
        return 0;
    }
    if (!strcmp(name, "drop_tiny")) {
        ctx->drop_tiny = value ? 1 : 0;
        ctx->maps_valid = false;  // every volume group moves to the generic (entry-major) kernel
        for (Group &g : ctx->groups) g.aff_checked = false;
        return 0;
    }
    if (!strcmp(name, "variant")) {
        ctx->variant = (int)value;
        return 0;
    }
    if (!strcmp(name, "timing")) {
        ctx->timing = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "overlap")) {
        ctx->overlap = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "locality")) {
        if (!ctx->groups.empty()) return fail(ctx, B200ASM_ESTATE, "locality: set it before the first b200asm_add_group");
        ctx->locality = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "affine")) {
        ctx->affine = value ? 1 : 0;
        for (Group &g : ctx->groups) g.aff_checked = false;
        return 0;
    }
    if (!strcmp(name, "staging_lo") || !strcmp(name, "staging_hi")) {
        if (value < 0) return fail(ctx, B200ASM_EINVAL, "staging rows: must not be negative");
        if (!ctx->groups.empty()) return fail(ctx, B200ASM_ESTATE, "staging rows: set them before the first b200asm_add_group");
        (name[8] == 'l' ? ctx->staging_lo : ctx->staging_hi) = value;
        return 0;
    }
    if (!strcmp(name, "download_a_count") || !strcmp(name, "download_rhs_first") || !strcmp(name, "download_rhs_count")) {
        if (value < -1) return fail(ctx, B200ASM_EINVAL, "download window: bad value");
        (name[9] == 'a' ? ctx->dl_a : (name[13] == 'f' ? ctx->dl_r0 : ctx->dl_rn)) = value;
        ctx->ov_valid = false;
        return 0;
    }
    if (!strcmp(name, "exchange_timeout_ms")) {
        if (value < 1) return fail(ctx, B200ASM_EINVAL, "exchange_timeout_ms: must be positive");
        ctx->xtimeout_ms = value;
        return 0;
    }
    if (!strcmp(name, "overlap_min_elements")) {
        if (value < 1) return fail(ctx, B200ASM_EINVAL, "overlap_min_elements: must be positive");
        if (!ctx->groups.empty()) return fail(ctx, B200ASM_ESTATE, "overlap_min_elements: set it before the first b200asm_add_group");
        ctx->overlap_min_elements = value;
        return 0;
    }
    if (!strcmp(name, "overlap_min_bytes")) {
        if (value < 0) return fail(ctx, B200ASM_EINVAL, "overlap_min_bytes: must not be negative");
        ctx->overlap_min_bytes = value;
        return 0;
    }
    return fail(ctx, B200ASM_EINVAL, std::string("unknown option ") + name);
}

extern "C" int b200asm_set_nodes(b200asm_ctx *ctx, int64_t nnodes, const double *xyz) {
    if (!ctx || nnodes < 0 || (nnodes && !xyz)) return fail(ctx, B200ASM_EINVAL, "b200asm_set_nodes: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (nnodes != ctx->nnodes || !ctx->d_xyz) {
        cudaFree(ctx->d_xyz);
        ctx->d_xyz = nullptr;
        CK(cudaMalloc((void **)&ctx->d_xyz, std::max<int64_t>(nnodes, 1) * 3 * sizeof(double)));
        ctx->nnodes = nnodes;
    }
    CK(cudaMemcpyAsync(ctx->d_xyz, xyz, (size_t)nnodes * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += nnodes * 3 * (int64_t)sizeof(double);
    for (Group &g : ctx->groups) g.aff_checked = false;  // parallelepiped or not is a property of the coordinates
    if (ctx->groups.empty()) ctx->h_xyz.assign(xyz, xyz + (size_t)nnodes * 3);  // (only add_group reads it)
    return 0;
}

extern "C" int b200asm_add_group(b200asm_ctx *ctx, const b200asm_group *gi) {
    if (!ctx || !gi) return B200ASM_EINVAL;
    CK(cudaSetDevice(ctx->device));
    Group g;
    // every early return below releases what has been uploaded so far
    struct Guard {
        Group &g;
        bool armed = true;
        ~Guard() { if (armed) free_group(g); }
    } guard{g};
    g.topology = gi->topology; g.porder = gi->porder; g.kind = gi->kind; g.ns = gi->nstate; g.nel = gi->nel;
    g.nn = ncorner_of(gi->topology);
    g.n = nshape_of(gi->topology, gi->porder);
    g.nq = gi->nqp;
    g.dim = (gi->topology == B200ASM_HEX || gi->topology == B200ASM_TET || gi->topology == B200ASM_PRISM || gi->topology == B200ASM_PYRAMID)
                ? 3 : (gi->topology == B200ASM_LINE ? 1 : 2);
    g.plane = g.dim == 2 && (gi->kind == B200ASM_POISSON || gi->kind == B200ASM_ELASTICITY2D);
    if (g.nn > 0 && g.n < 0 && gi->porder >= 1 && gi->nshape >= g.nn &&
        (gi->topology == B200ASM_PRISM || gi->topology == B200ASM_PYRAMID))
        g.n = gi->nshape;  // prisms / pyramids of order >= 3: no host formula for the count, the caller's tables define it
    if (g.nn < 0 || g.n < 0 || gi->porder < 1)
        return fail(ctx, B200ASM_EINVAL, "add_group: unsupported topology/order (H1, uniform p: hex/quad/tet/tri 1..4, prism/pyramid 1..2)");
    if (gi->nshape != g.n) {
        // the sides carry different orders (p-refined neighbours): the caller's tables define the functions, porder is the
        // largest order; runtime-size kernels run the group
        if (gi->nshape < g.nn || gi->nshape > g.n) return fail(ctx, B200ASM_EINVAL, "add_group: nshape does not fit topology/order");
        g.n = gi->nshape;
        g.uniform = false;
    }
    if (g.ns < 1 || g.ns > 3) return fail(ctx, B200ASM_EINVAL, "add_group: nstate must be 1, 2 or 3");
    if (g.nel < 0 || g.nq <= 0 || g.nq > 512) return fail(ctx, B200ASM_EINVAL, "add_group: bad nel/nqp");
    if (!gi->elnodes || !gi->dest || !gi->qpts || !gi->qwts || !gi->phi || !gi->dphi)
        return fail(ctx, B200ASM_EINVAL, "add_group: NULL table");
    const bool volume = g.dim == 3;
    if (volume) {
        if (gi->kind == B200ASM_POISSON && g.ns != 1) return fail(ctx, B200ASM_EINVAL, "add_group: Poisson has nstate 1");
        if (gi->kind == B200ASM_ELASTICITY3D && g.ns != 3) return fail(ctx, B200ASM_EINVAL, "add_group: Elasticity3D has nstate 3");
        if (gi->kind != B200ASM_POISSON && gi->kind != B200ASM_ELASTICITY3D)
            return fail(ctx, B200ASM_EINVAL, "add_group: volume elements need kind POISSON or ELASTICITY3D");
        for (int k = 0; k < kNumVol && g.uniform; k++)
            if (kVol[k].topology == g.topology && kVol[k].porder == g.porder && kVol[k].ns == g.ns) g.cfg = k;
        if (g.cfg < 0) g.generic = true;  // no specialised kernel: assemble_volume_generic_kernel (runtime sizes)
        if (g.topology == B200ASM_HEX && g.porder == 2 && g.ns == 1 && g.uniform && g.nq == 27) {
            // the rule is the 3 x 3 x 3 tensor rule with the point order q = q1 + 3 (q2 + 3 q3) (Integral/pzquad.cpp:268-284)
            g.sumfact_ok = true;
            for (int q = 0; q < 27 && g.sumfact_ok; q++)
                g.sumfact_ok = gi->qpts[3 * q] == gi->qpts[3 * (q % 3)] && gi->qpts[3 * q + 1] == gi->qpts[3 * ((q / 3) % 3)] &&
                               gi->qpts[3 * q + 2] == gi->qpts[3 * (q / 9)];
        }
        for (int k = 0; k < kNumMma && !g.generic; k++) {
            if (kMma[k].sumfact && !g.sumfact_ok) continue;
            if (kMma[k].topology == g.topology && kMma[k].porder == g.porder && kMma[k].ns == g.ns &&
                (kMma[k].variant == 0 ? g.mma < 0 : kMma[k].variant == ctx->variant))
                g.mma = k;
        }
    } else if (g.plane) {
        if (gi->porder > 4)
            return fail(ctx, B200ASM_EINVAL, "add_group: plane domain elements: p <= 4");
        if (gi->kind == B200ASM_POISSON && g.ns != 1) return fail(ctx, B200ASM_EINVAL, "add_group: Poisson has nstate 1");
        if (gi->kind == B200ASM_ELASTICITY2D && g.ns != 2) return fail(ctx, B200ASM_EINVAL, "add_group: Elasticity2D has nstate 2");
    } else if (gi->kind != B200ASM_BC) {
        return fail(ctx, B200ASM_EINVAL, "add_group: face / line elements need kind BC (or POISSON / ELASTICITY2D as plane domain elements)");
    } else if (g.dim == 1 && gi->porder > 4) {
        return fail(ctx, B200ASM_EINVAL, "add_group: line elements: p <= 4");
    }
    g.m = g.n * g.ns;
    memcpy(g.coef, gi->coef, sizeof(g.coef));

    // element order on the device: mesh order, or sorted by colour for the conflict-free scatter
    std::vector<int64_t> order((size_t)g.nel);
    for (int64_t e = 0; e < g.nel; e++) order[e] = e;
    g.seg = {0, g.nel};
    if (ctx->scatter == B200ASM_SCATTER_COLORED && g.nel > 0) {
        std::vector<int32_t> colour;
        const int ncol = colour_elements(g.nel, g.m, gi->dest, 0, colour);
        if (ncol < 0) return fail(ctx, B200ASM_EINVAL, "add_group: the mesh needs more than 64 element colours");
        std::vector<int64_t> count(ncol + 1, 0);
        for (int64_t e = 0; e < g.nel; e++) count[colour[e] + 1]++;
        for (int c = 0; c < ncol; c++) count[c + 1] += count[c];
        g.seg.assign(count.begin(), count.end());
        std::vector<int64_t> cursor(count.begin(), count.end() - 1);
        for (int64_t e = 0; e < g.nel; e++) order[cursor[colour[e]]++] = e;  // stable within a colour
    }
    // row-sharded assembly: elements with an equation among the staging rows come first (stable otherwise) and form chunk 0,
    // rounded up to the chunk alignment; coloured groups keep their colour order (their contributions are pushed after the
    // last launch instead of early)
    constexpr int64_t kAlign = 384;
    if (ctx->staging_hi > ctx->staging_lo && g.seg.size() == 2 && g.nel > 0) {
        std::vector<int64_t> first, rest;
        for (int64_t e = 0; e < g.nel; e++) {
            bool hit = false;
            for (int k = 0; k < g.m && !hit; k++) {
                const int64_t d = gi->dest[e * g.m + k];
                hit = d >= ctx->staging_lo && d < ctx->staging_hi;
            }
            (hit ? first : rest).push_back(e);
        }
        if (!first.empty()) {
            g.n_if = std::min<int64_t>(g.nel, ((int64_t)first.size() + kAlign - 1) / kAlign * kAlign);
            std::copy(first.begin(), first.end(), order.begin());
            std::copy(rest.begin(), rest.end(), order.begin() + (int64_t)first.size());
        }
    }
    // element chunks for the overlapped download: boundaries are multiples of every kernel's batch size
    {
        constexpr int64_t kMaxChunks = 16;
        const int64_t nrest = g.nel - g.n_if;
        int64_t nch = 1;
        if (volume && g.seg.size() == 2) nch = std::max<int64_t>(1, std::min<int64_t>(kMaxChunks, nrest / ctx->overlap_min_elements));
        g.chunk.assign(1, 0);
        if (g.n_if > 0 && g.n_if < g.nel) g.chunk.push_back(g.n_if);
        for (int64_t c = 1; c < nch; c++) {
            const int64_t b = g.n_if + (nrest * c / nch) / kAlign * kAlign;
            if (b > g.chunk.back() && b < g.nel) g.chunk.push_back(b);
        }
        g.chunk.push_back(g.nel);
    }
    // locality order (atomic scatter, node coordinates already known): inside every chunk the elements follow a Morton
    // curve through their centroids.  The persistent grid works on a window of a few thousand consecutive elements; in mesh
    // (lexicographic) order that window is a thin sheet and the CSR rows shared with the next sheet leave L2 before their
    // last contribution arrives (the read-modify-write of A then costs HBM traffic twice); along the curve it is a compact
    // block.  Chunks keep their element sets, so the download frontier of assemble_overlapped is unchanged.
    if (volume && ctx->locality && g.seg.size() == 2 && g.nel >= 4096 && (int64_t)ctx->h_xyz.size() == ctx->nnodes * 3) {
        std::vector<float> cen((size_t)g.nel * 3);
        float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        bool ok = true;
        for (int64_t e = 0; e < g.nel && ok; e++) {
            double c[3] = {0, 0, 0};
            for (int k = 0; k < g.nn; k++) {
                const int64_t node = gi->elnodes[e * g.nn + k];
                if (node < 0 || node >= ctx->nnodes) { ok = false; break; }
                for (int r = 0; r < 3; r++) c[r] += ctx->h_xyz[(size_t)node * 3 + r];
            }
            for (int r = 0; r < 3; r++) {
                const float v = (float)(c[r] / g.nn);
                cen[(size_t)e * 3 + r] = v;
                lo[r] = std::min(lo[r], v);
                hi[r] = std::max(hi[r], v);
            }
        }
        if (ok) {
            auto spread = [](uint64_t x) {  // 21 bits -> every third bit
                x &= 0x1fffff;
                x = (x | x << 32) & 0x1f00000000ffffull;
                x = (x | x << 16) & 0x1f0000ff0000ffull;
                x = (x | x << 8) & 0x100f00f00f00f00full;
                x = (x | x << 4) & 0x10c30c30c30c30c3ull;
                x = (x | x << 2) & 0x1249249249249249ull;
                return x;
            };
            // cells of one size in all directions (the longest extent spans 2^10 cells)
            const float ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-30f});
            const float scale = 1023.0f / ext;
            std::vector<std::pair<uint64_t, int64_t>> keyed;
            for (size_t c = 0; c + 1 < g.chunk.size(); c++) {
                keyed.clear();
                for (int64_t k = g.chunk[c]; k < g.chunk[c + 1]; k++) {
                    const int64_t e = order[k];  // (the interface prefix has already permuted the elements)
                    uint64_t key = 0;
                    for (int r = 0; r < 3; r++) key |= spread((uint64_t)((cen[(size_t)e * 3 + r] - lo[r]) * scale)) << r;
                    keyed.emplace_back(key, e);
                }
                std::sort(keyed.begin(), keyed.end());
                for (size_t k = 0; k < keyed.size(); k++) order[g.chunk[c] + (int64_t)k] = keyed[k].second;
            }
        }
    }
    // destination indices as int32 (the reference's own CSR loops are 32-bit: Matrix/pzsysmp.cpp:62,73)
    std::vector<int32_t> dest32((size_t)g.nel * g.m);
    std::vector<int32_t> elnodes((size_t)g.nel * g.nn);
    for (int64_t e = 0; e < g.nel; e++) {
        const int64_t src = order[e];
        for (int k = 0; k < g.m; k++) {
            const int64_t d = gi->dest[src * g.m + k];  // -1: equation removed by the equation filter
            if (d < -1 || d > 0x7fffffff) return fail(ctx, B200ASM_EINVAL, "add_group: destination index out of int32 range");
            dest32[(size_t)e * g.m + k] = (int32_t)d;
            g.max_dest = std::max(g.max_dest, d);
        }
        for (int k = 0; k < g.nn; k++) {
            const int32_t node = gi->elnodes[src * g.nn + k];
            if (node < 0 || (ctx->nnodes > 0 && node >= ctx->nnodes)) return fail(ctx, B200ASM_EINVAL, "add_group: corner node index outside the node table");
            elnodes[(size_t)e * g.nn + k] = node;
        }
    }
    // smallest destination equation of every chunk (chunks are sets of elements: the order inside does not matter)
    g.chunk_min.assign(g.chunk.size() - 1, INT64_MAX);
    for (size_t c = 0; c + 1 < g.chunk.size(); c++) {
        int32_t mn = INT32_MAX;
        const int32_t *d = dest32.data() + (size_t)g.chunk[c] * g.m, *dend = dest32.data() + (size_t)g.chunk[c + 1] * g.m;
        for (; d < dend; d++)
            if (*d >= 0 && *d < mn) mn = *d;
        if (mn != INT32_MAX) g.chunk_min[c] = mn;
    }
    // gradients of the geometric (corner) functions at the points = the p=1 shape gradients
    std::vector<double> gphi((size_t)g.nq * g.nn), dng((size_t)g.nq * g.dim * g.nn);
    if (b200asm_shape_tables(g.topology, 1, g.nq, gi->qpts, gphi.data(), dng.data()) != g.nn)
        return fail(ctx, B200ASM_EINVAL, "add_group: geometry table failed");
    int rc;
    if ((rc = upload(ctx, &g.d_elnodes, elnodes.data(), elnodes.size()))) return rc;
    if ((rc = upload(ctx, &g.d_dest, dest32.data(), dest32.size()))) return rc;
    if ((rc = upload(ctx, &g.d_qw, gi->qwts, (size_t)g.nq))) return rc;
    if ((rc = upload(ctx, &g.d_phi, gi->phi, (size_t)g.nq * g.n))) return rc;
    if ((rc = upload(ctx, &g.d_dphi, gi->dphi, (size_t)g.nq * g.dim * g.n))) return rc;
    if ((rc = upload(ctx, &g.d_dng, dng.data(), dng.size()))) return rc;
    std::vector<double> dng_t, dphi_pad, phi_pad;
    if (volume) {
        // table layouts of the DMMA kernels (gram_mma.cuh): coalesced / whole-line accesses
        const int np = ((g.n + 15) / 16) * 16;
        dng_t.assign((size_t)3 * g.nn * g.nq, 0.0);
        dphi_pad.assign((size_t)g.nq * 3 * np, 0.0);
        phi_pad.assign((size_t)g.nq * np, 0.0);
        for (int q = 0; q < g.nq; q++)
            for (int d = 0; d < 3; d++) {
                for (int a = 0; a < g.nn; a++) dng_t[((size_t)d * g.nn + a) * g.nq + q] = dng[((size_t)q * 3 + d) * g.nn + a];
                for (int i = 0; i < g.n; i++) dphi_pad[((size_t)q * 3 + d) * np + i] = gi->dphi[((size_t)q * 3 + d) * g.n + i];
            }
        for (int q = 0; q < g.nq; q++)
            for (int i = 0; i < g.n; i++) phi_pad[(size_t)q * np + i] = gi->phi[(size_t)q * g.n + i];
        if ((rc = upload(ctx, &g.d_dng_t, dng_t.data(), dng_t.size()))) return rc;
        if ((rc = upload(ctx, &g.d_dphi_pad, dphi_pad.data(), dphi_pad.size()))) return rc;
        if ((rc = upload(ctx, &g.d_phi_pad, phi_pad.data(), phi_pad.size()))) return rc;
    }
    std::vector<double> aux;  // (lives until the synchronize below)
    if (volume && g.topology == B200ASM_HEX && g.porder <= 2 && !g.generic) {
        for (int k = 0; k < kNumAffHex; k++)
            if (kAffHex[k].porder == g.porder && kAffHex[k].ns == g.ns) g.aff = k;
        if (g.n == 8 && g.ns == 1) aff_tables<HexP1PoissonAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else if (g.n == 8) aff_tables<HexP1ElastAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else if (g.ns == 1) aff_tables<HexP2PoissonAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else aff_tables<HexP2ElastAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        if ((rc = upload(ctx, &g.d_aux, aux.data(), aux.size()))) return rc;
    }
    if (volume && g.topology == B200ASM_TET && (g.porder == 3 || g.porder == 4) && g.uniform && ctx->variant != kTetRegTileVariant) {
        for (int k = 0; k < kNumTetAffS; k++)
            if (kTetAffS[k].porder == g.porder && kTetAffS[k].ns == g.ns) g.taff = k;
        if (g.porder == 3 && g.ns == 1) aff_tables<TetP3PoissonAffS>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else if (g.porder == 3) aff_tables<TetP3ElastAffS>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else if (g.ns == 1) aff_tables<TetP4PoissonAffS>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else aff_tables<TetP4ElastAffS>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        if ((rc = upload(ctx, &g.d_aux, aux.data(), aux.size()))) return rc;
    }
    if (volume && g.topology == B200ASM_TET && g.porder <= 2 && !g.generic) {  // (orders 3, 4 run the register-tile kernel)
        if (g.n == 4 && g.ns == 1) aff_tables<TetP1PoissonAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else if (g.n == 4) aff_tables<TetP1ElastAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else if (g.ns == 1) aff_tables<TetP2PoissonAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        else aff_tables<TetP2ElastAff>(g.nq, gi->qwts, gi->phi, gi->dphi, aux);
        if ((rc = upload(ctx, &g.d_aux, aux.data(), aux.size()))) return rc;
    }
    std::vector<double> sf_aux(sf::AUX_LEN), sf_f3(4 * 9 * 3);
    if (volume && g.sumfact_ok) {
        const double x1d[3] = {gi->qpts[0], gi->qpts[3], gi->qpts[6]};
        sf::build_tables(x1d, sf_aux.data(), sf_f3.data());
        if ((rc = upload(ctx, &g.d_aux2, sf_aux.data(), sf_aux.size()))) return rc;
        // (one table for every context: the values only depend on the reference's 3-point line rule)
        CK(cudaMemcpyToSymbolAsync(c_sfF3, sf_f3.data(), sf_f3.size() * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    }
    std::vector<double> force;
    if (gi->force) {
        const size_t per = (size_t)g.nq * g.ns;
        force.resize((size_t)g.nel * per);
        for (int64_t e = 0; e < g.nel; e++) memcpy(&force[(size_t)e * per], gi->force + (size_t)order[e] * per, per * sizeof(double));
        if ((rc = upload(ctx, &g.d_force, force.data(), force.size()))) return rc;
        g.stored_order = order;
    }
    CK(cudaStreamSynchronize(ctx->stream));  // dest32 / dng are stack-owned
    guard.armed = false;
    ctx->groups.push_back(g);
    ctx->maps_valid = false;  // scatter maps are rebuilt (on the resident pattern) by the next assembly
    return (int)ctx->groups.size() - 1;
}

extern "C" int b200asm_set_group_coef(b200asm_ctx *ctx, int group, const double coef[16]) {
    if (!ctx || group < 0 || group >= (int)ctx->groups.size() || !coef) return fail(ctx, B200ASM_EINVAL, "set_group_coef: bad arguments");
    memcpy(ctx->groups[group].coef, coef, sizeof(double) * 16);
    return 0;
}

extern "C" int b200asm_set_group_force(b200asm_ctx *ctx, int group, const double *force) {
    if (!ctx || group < 0 || group >= (int)ctx->groups.size()) return fail(ctx, B200ASM_EINVAL, "set_group_force: bad arguments");
    Group &g = ctx->groups[group];
    if ((force != nullptr) != (g.d_force != nullptr))
        return fail(ctx, B200ASM_EINVAL, "set_group_force: the group was added with / without a table (clear the groups to change that)");
    if (!force) return 0;
    CK(cudaSetDevice(ctx->device));
    const size_t per = (size_t)g.nq * g.ns;
    std::vector<double> tmp((size_t)g.nel * per);
    for (int64_t e = 0; e < g.nel; e++) memcpy(&tmp[(size_t)e * per], force + (size_t)g.stored_order[e] * per, per * sizeof(double));
    CK(cudaMemcpyAsync(g.d_force, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += (int64_t)(tmp.size() * sizeof(double));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int b200asm_clear_groups(b200asm_ctx *ctx) {
    if (!ctx) return B200ASM_EINVAL;
    cudaSetDevice(ctx->device);
    for (Group &g : ctx->groups) free_group(g);
    ctx->groups.clear();
    ctx->maps_valid = false;
    return 0;
}

namespace {
void drop_pattern(b200asm_ctx *ctx) {
    b200asm_exchange_clear(ctx);  // (peers map the arrays of the pattern that goes away)
    cudaFree(ctx->d_ia); cudaFree(ctx->d_ja); cudaFree(ctx->d_a); cudaFree(ctx->d_rhs);
    ctx->d_ia = nullptr; ctx->d_ja = nullptr; ctx->d_a = ctx->d_rhs = nullptr;
    ctx->have_pattern = false;
    ctx->maps_valid = false;
}
int grid_for(const b200asm_ctx *ctx, int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)ctx->num_sms * 16));
}
}  // namespace

extern "C" int b200asm_set_pattern(b200asm_ctx *ctx, int64_t neq, const int64_t *ia, const int64_t *ja, int symmetric) {
    if (!ctx || neq < 0 || !ia) return fail(ctx, B200ASM_EINVAL, "set_pattern: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int64_t nnz = ia[neq];
    if (nnz < 0 || nnz > 0x7fffffff) return fail(ctx, B200ASM_EINVAL, "set_pattern: nnz must fit int32 per device (shard the rows)");
    if (nnz && !ja) return fail(ctx, B200ASM_EINVAL, "set_pattern: ja == NULL");
    drop_pattern(ctx);
    ctx->neq = neq; ctx->nnz = nnz; ctx->symmetric = symmetric ? 1 : 0;
    int64_t *d_ja64 = nullptr;
    int rc;
    if ((rc = upload(ctx, &ctx->d_ia, ia, (size_t)neq + 1))) return rc;
    if ((rc = upload(ctx, &d_ja64, ja, (size_t)nnz))) return rc;
    CK(cudaMalloc((void **)&ctx->d_ja, std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    if (nnz) {
        patdev::narrow_kernel<<<grid_for(ctx, nnz, 256), 256, 0, ctx->stream>>>(nnz, d_ja64, ctx->d_ja);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_ja64);
    CK(cudaMalloc((void **)&ctx->d_a, std::max<int64_t>(nnz, 1) * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->d_rhs, std::max<int64_t>(neq, 1) * sizeof(double)));
    ctx->have_pattern = true;
    return build_smaps(ctx, ctx->d_ja);
}

extern "C" int b200asm_build_pattern_device(b200asm_ctx *ctx, int symmetric, int64_t nel, const int64_t *elgraphindex,
                                            const int64_t *elgraph, int64_t nblock, const int64_t *blockpos,
                                            const int64_t *blocksize, int64_t *neq_out, int64_t *nnz_out) {
    if (!ctx || nel < 0 || nblock < 0 || !elgraphindex || !blockpos || !blocksize || (nel && !elgraph))
        return fail(ctx, B200ASM_EINVAL, "build_pattern_device: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int64_t total = elgraphindex[nel];
    const int64_t neq = nblock ? blockpos[nblock - 1] + blocksize[nblock - 1] : 0;
    if (nblock > 0x7fffffff || neq > 0x7fffffff || nel > 0x7fffffff)
        return fail(ctx, B200ASM_EINVAL, "build_pattern_device: blocks / equations / elements must fit int32 per device");
    std::vector<int32_t> eg32((size_t)std::max<int64_t>(total, 1)), bs32((size_t)std::max<int64_t>(nblock, 1));
    for (int64_t k = 0; k < total; k++) {
        if (elgraph[k] < 0 || elgraph[k] >= nblock) return fail(ctx, B200ASM_EINVAL, "build_pattern_device: connect out of range");
        eg32[k] = (int32_t)elgraph[k];
    }
    for (int64_t b = 0; b < nblock; b++) bs32[b] = (int32_t)blocksize[b];
    drop_pattern(ctx);
    ctx->neq = neq; ctx->symmetric = symmetric ? 1 : 0;

    int64_t *d_egi = nullptr, *d_bpos = nullptr, *d_n2e_idx = nullptr, *d_rowlen = nullptr;
    int32_t *d_eg = nullptr, *d_bsize = nullptr, *d_cnt = nullptr, *d_n2e = nullptr;
    int *d_err = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_egi); cudaFree(d_bpos); cudaFree(d_n2e_idx); cudaFree(d_rowlen); cudaFree(d_eg); cudaFree(d_bsize);
        cudaFree(d_cnt); cudaFree(d_n2e); cudaFree(d_err); cudaFree(d_tmp);
    };
#define CKP(call)                                                                                      \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            cleanup();                                                                                   \
            return fail(ctx, B200ASM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
        }                                                                                                \
    } while (0)
    int rc;
    if ((rc = upload(ctx, &d_egi, elgraphindex, (size_t)nel + 1)) || (rc = upload(ctx, &d_eg, eg32.data(), (size_t)total)) ||
        (rc = upload(ctx, &d_bpos, blockpos, (size_t)nblock)) || (rc = upload(ctx, &d_bsize, bs32.data(), (size_t)nblock))) {
        cleanup();
        return rc;
    }
    cudaStream_t st = ctx->stream;
    // block -> elements: count, scan, fill
    CKP(cudaMalloc((void **)&d_cnt, (size_t)(nblock + 1) * sizeof(int32_t)));
    CKP(cudaMalloc((void **)&d_n2e_idx, (size_t)(nblock + 1) * sizeof(int64_t)));
    CKP(cudaMalloc((void **)&d_n2e, (size_t)std::max<int64_t>(total, 1) * sizeof(int32_t)));
    CKP(cudaMalloc((void **)&d_err, sizeof(int)));
    CKP(cudaMemsetAsync(d_cnt, 0, (size_t)(nblock + 1) * sizeof(int32_t), st));
    CKP(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    if (total) patdev::count_incidence_kernel<<<grid_for(ctx, total, 256), 256, 0, st>>>(total, d_eg, d_cnt);
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_n2e_idx, (int)(nblock + 1), st);
    CKP(cudaMalloc((void **)&d_rowlen, (size_t)(neq + 1) * sizeof(int64_t)));
    cub::DeviceScan::InclusiveSum(nullptr, tmp2, d_rowlen, d_rowlen, (int)(neq + 1), st);
    tmp_bytes = std::max(tmp_bytes, tmp2);
    CKP(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
    CKP(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_n2e_idx, (int)(nblock + 1), st));
    CKP(cudaMemsetAsync(d_cnt, 0, (size_t)(nblock + 1) * sizeof(int32_t), st));
    if (nel) patdev::fill_incidence_kernel<<<grid_for(ctx, nel, 256), 256, 0, st>>>(nel, d_egi, d_eg, d_n2e_idx, d_cnt, d_n2e);
    CKP(cudaGetLastError());
    // sweep 1: row lengths -> IA
    patdev::Params pp{};
    pp.symmetric = ctx->symmetric; pp.nel = nel; pp.nblock = nblock; pp.egi = d_egi; pp.eg = d_eg; pp.bpos = d_bpos;
    pp.bsize = d_bsize; pp.n2e_idx = d_n2e_idx; pp.n2e = d_n2e; pp.rowlen = d_rowlen; pp.error = d_err;
    CKP(cudaMemsetAsync(d_rowlen, 0, (size_t)(neq + 1) * sizeof(int64_t), st));
    const int pgrid = (int)std::max<int64_t>(1, std::min<int64_t>((nblock + patdev::WARPS - 1) / patdev::WARPS, (int64_t)ctx->num_sms * 64));
    patdev::pattern_kernel<1><<<pgrid, patdev::WARPS * 32, 0, st>>>(pp);
    CKP(cudaGetLastError());
    CKP(cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, d_rowlen, d_rowlen, (int)(neq + 1), st));
    int64_t nnz = 0;
    int err = 0;
    CKP(cudaMemcpyAsync(&nnz, d_rowlen + neq, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CKP(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    ctx->launches += 5;
    if (err) {
        cleanup();
        return fail(ctx, B200ASM_EINVAL, "build_pattern_device: a connect has more neighbour candidates than the device builder holds (use b200asm_build_pattern)");
    }
    if (nnz > 0x7fffffff) {
        cleanup();
        return fail(ctx, B200ASM_EINVAL, "build_pattern_device: nnz must fit int32 per device (shard the rows)");
    }
    // sweep 2: column indices
    ctx->nnz = nnz;
    ctx->d_ia = d_rowlen;
    d_rowlen = nullptr;
    CKP(cudaMalloc((void **)&ctx->d_ja, (size_t)std::max<int64_t>(nnz, 1) * sizeof(int32_t)));
    pp.ia = ctx->d_ia; pp.ja = ctx->d_ja;
    patdev::pattern_kernel<2><<<pgrid, patdev::WARPS * 32, 0, st>>>(pp);
    CKP(cudaGetLastError());
    ctx->launches++;
    CKP(cudaStreamSynchronize(st));
    cleanup();
#undef CKP
    CK(cudaMalloc((void **)&ctx->d_a, std::max<int64_t>(nnz, 1) * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->d_rhs, std::max<int64_t>(neq, 1) * sizeof(double)));
    if (neq_out) *neq_out = neq;
    if (nnz_out) *nnz_out = nnz;
    ctx->have_pattern = true;
    return build_smaps(ctx, ctx->d_ja);
}

extern "C" int b200asm_get_ja_range(b200asm_ctx *ctx, int64_t first, int64_t count, int64_t *ja_host) {
    if (!ctx || first < 0 || count < 0 || (count && !ja_host)) return fail(ctx, B200ASM_EINVAL, "get_ja_range: bad arguments");
    if (!ctx->d_ja) return fail(ctx, B200ASM_ESTATE, "get_ja_range: no pattern on the device");
    if (first + count > ctx->nnz) return fail(ctx, B200ASM_EINVAL, "get_ja_range: range exceeds nnz");
    if (count == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    int32_t *tmp = reinterpret_cast<int32_t *>(ja_host);
    CK(cudaMemcpyAsync(tmp, ctx->d_ja + first, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t k = count - 1; k >= 0; k--) ja_host[k] = tmp[k];
    ctx->d2h += count * (int64_t)sizeof(int32_t);
    return 0;
}

extern "C" int b200asm_get_pattern(b200asm_ctx *ctx, int64_t *ia_host, int64_t *ja_host) {
    if (!ctx) return B200ASM_EINVAL;
    if (!ctx->d_ia || !ctx->d_ja) return fail(ctx, B200ASM_ESTATE, "get_pattern: no pattern on the device");
    CK(cudaSetDevice(ctx->device));
    if (ia_host) {
        CK(cudaMemcpyAsync(ia_host, ctx->d_ia, (size_t)(ctx->neq + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h += (ctx->neq + 1) * (int64_t)sizeof(int64_t);
    }
    if (ja_host && ctx->nnz) {
        // int32 on the device; widened on the host in place (back to front)
        int32_t *tmp = reinterpret_cast<int32_t *>(ja_host);
        CK(cudaMemcpyAsync(tmp, ctx->d_ja, (size_t)ctx->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int64_t k = ctx->nnz - 1; k >= 0; k--) ja_host[k] = tmp[k];
        ctx->d2h += ctx->nnz * (int64_t)sizeof(int32_t);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

namespace {

// launches of one group; volume groups without colours may be restricted to the element range [r0, r1) (chunk boundaries
// of Group::chunk).  r0 < 0: the whole group.
int enqueue_group(b200asm_ctx *ctx, Group &g, int64_t r0, int64_t r1) {
    const int atomic = ctx->scatter == B200ASM_SCATTER_ATOMIC ? 1 : 0;
    const size_t nseg = g.seg.size() - 1;  // 1, or the number of colours
    if (entry_major(ctx, g)) {
        BcParams p;
        p.nel = g.nel; p.nq = g.nq; p.xyz = ctx->d_xyz; p.elnodes = g.d_elnodes; p.dest = g.d_dest;
        p.qw = g.d_qw; p.phi = g.d_phi; p.dphi = g.d_dphi; p.dng = g.d_dng; p.smap = g.d_smap; p.smapT = g.d_smapT;
        p.kind = g.kind; p.fdim = g.dim; p.force = g.d_force;
        p.a = ctx->d_a; p.rhs = ctx->d_rhs; p.atomic = atomic | (ctx->drop_tiny ? 4 : 0); p.rhs_only = ctx->rhs_only;
        memcpy(p.coef, g.coef, sizeof(p.coef));
        for (size_t c = 0; c < nseg; c++) {
            p.el0 = g.seg[c]; p.el1 = g.seg[c + 1];
            if (r0 >= 0) { p.el0 = r0; p.el1 = r1; }  // (element chunk of an uncoloured volume group: nseg == 1)
            if (p.el1 == p.el0) continue;
            if (g.dim == 3) {
                const size_t smem = ((size_t)g.nq * 11 + (size_t)g.nn * 3) * sizeof(double);
                const int grid = (int)std::min<int64_t>(p.el1 - p.el0, (int64_t)ctx->num_sms * 8);
                assemble_volume_generic_kernel<<<grid, 128, smem, ctx->stream>>>(p, g.nn, g.n, g.ns);
                CK(cudaGetLastError());
            } else if (g.plane) {
                const int grid = (int)((p.el1 - p.el0 + 3) / 4);
                const size_t smem = 4 * (size_t)g.nq * (2 + 2 * g.n) * sizeof(double);
                assemble_plane_kernel<<<grid, 128, smem, ctx->stream>>>(p, g.nn, g.n, g.ns);
                CK(cudaGetLastError());
            } else {
                CK(dispatch_bc(g.topology, g.uniform ? g.porder : 0, g.nn, g.n, g.ns, p, ctx->stream));
            }
            ctx->launches++;
        }
        return 0;
    }
    const MmaEntry *fast = fast_entry(ctx, g);
    const bool use_mma = fast != nullptr;
    size_t smem = 0;
    int per_sm = 1;
    if (g.use_gather && fast) {
        // load vector: the closed-form kernel in its load-vector-only mode (whole group, with the first node-block range);
        // matrix: node blocks [r0, r1) gathered row by row (r0 < 0: all)
        const int64_t b0 = r0 < 0 ? 0 : r0, b1 = r0 < 0 ? g.ng : r1;
        if (b0 == 0) {  // (once per assembly: a load-vector-only pass must not repeat it for every node-block range)
            smem = fast->smem(g.nq);
            CK(fast->prepare(smem, &per_sm));
            if (per_sm < 1) return fail(ctx, B200ASM_ECUDA, "assemble: kernel does not fit on an SM");
            VolParams p;
            p.nel = g.nel; p.nq = g.nq; p.kind = g.kind; p.atomic = 1; p.rhs_only = 1; p.debug = 0; p.nbatch = 0;
            p.xyz = ctx->d_xyz; p.elnodes = g.d_elnodes; p.dest = g.d_dest;
            p.qw = g.d_qw; p.phi = g.d_phi; p.dphi = g.d_dphi; p.dng = g.d_dng; p.force = g.d_force;
            p.dng_t = g.d_dng_t; p.dphi_pad = g.d_dphi_pad; p.phi_pad = g.d_phi_pad; p.aux = g.d_aux; p.aux2 = g.d_aux2;
            p.smap = nullptr; p.smapT = nullptr;
            p.a = ctx->d_a; p.rhs = ctx->d_rhs;
            memcpy(p.coef, g.coef, sizeof(p.coef));
            const int64_t want = (g.nel + fast->wpc - 1) / fast->wpc;
            CK(fast->launch(p, (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * per_sm), smem, ctx->stream));
            ctx->launches++;
        }
        if (ctx->rhs_only || b1 <= b0) return 0;
        if (b0 == 0) {  // Jacobian factors of every element of the group
            const int gridf = (int)std::min<int64_t>((g.nel + 127) / 128, (int64_t)ctx->num_sms * 16);
            const double scale = g.ns == 1 ? g.coef[0] : 1.0;
            if (g.nn == 4 && g.ns == 1) gat::factors_kernel<4, 1><<<gridf, 128, 0, ctx->stream>>>(g.nel, g.d_elnodes, ctx->d_xyz, scale, g.d_fac);
            else if (g.nn == 4) gat::factors_kernel<4, 3><<<gridf, 128, 0, ctx->stream>>>(g.nel, g.d_elnodes, ctx->d_xyz, scale, g.d_fac);
            else if (g.ns == 1) gat::factors_kernel<8, 1><<<gridf, 128, 0, ctx->stream>>>(g.nel, g.d_elnodes, ctx->d_xyz, scale, g.d_fac);
            else gat::factors_kernel<8, 3><<<gridf, 128, 0, ctx->stream>>>(g.nel, g.d_elnodes, ctx->d_xyz, scale, g.d_fac);
            CK(cudaGetLastError());
            ctx->launches++;
        }
        gat::Params gp;
        gp.g0 = b0; gp.g1 = b1; gp.symmetric = ctx->symmetric; gp.rl = g.gather_rl;
        gp.rec = g.d_grec; gp.glist = g.d_glist; gp.relpos = g.d_relpos; gp.rowflag = ctx->d_rowflag;
        gp.fac = g.d_fac; gp.aux = g.d_aux; gp.npp = ((g.n * (g.n + 1) / 2 + 31) / 32) * 32;
        gp.ia = ctx->d_ia; gp.a = ctx->d_a;
        gp.c1 = g.coef[0]; gp.c2 = g.coef[1]; gp.c3 = g.coef[2];
        // persistent grid: as many CTAs per SM as fit, every warp walks its tasks (SUBS node blocks each) with a grid stride
#define B200ASM_GATHER(N_, NS_)                                                                                                   \
    do {                                                                                                                          \
        using GC = gat::Cfg<N_, NS_>;                                                                                             \
        const size_t gsmem = GC::smem_bytes(g.gather_rl, g.gather_wpc);                                                           \
        int per = 0;                                                                                                              \
        CK(cudaFuncSetAttribute(gat::gather_rows_kernel<N_, NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));       \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, gat::gather_rows_kernel<N_, NS_>, g.gather_wpc * 32, gsmem));       \
        if (per < 1) return fail(ctx, B200ASM_ECUDA, "assemble: the gather kernel does not fit on an SM");                        \
        const int64_t tasks = (b1 - b0 + GC::SUBS - 1) / GC::SUBS;                                                                \
        const int ggrid = (int)std::max<int64_t>(1, std::min<int64_t>((tasks + g.gather_wpc - 1) / g.gather_wpc, (int64_t)ctx->num_sms * per)); \
        gat::gather_rows_kernel<N_, NS_><<<ggrid, g.gather_wpc * 32, gsmem, ctx->stream>>>(gp);                                    \
    } while (0)
        if (g.n == 4 && g.ns == 1) B200ASM_GATHER(4, 1);
        else if (g.n == 4) B200ASM_GATHER(4, 3);
        else if (g.n == 10 && g.ns == 1) B200ASM_GATHER(10, 1);
        else if (g.n == 10) B200ASM_GATHER(10, 3);
        else if (g.n == 8 && g.ns == 1) B200ASM_GATHER(8, 1);
        else if (g.n == 8) B200ASM_GATHER(8, 3);
        else if (g.n == 27 && g.ns == 1) B200ASM_GATHER(27, 1);
        else if (g.n == 27) B200ASM_GATHER(27, 3);
        else return fail(ctx, B200ASM_EINVAL, "assemble: no gather kernel for this group");
#undef B200ASM_GATHER
        CK(cudaGetLastError());
        ctx->launches++;
        return 0;
    }
    if (use_mma) {
        smem = fast->smem(g.nq);
        if (smem > 227 * 1024) return fail(ctx, B200ASM_EINVAL, "assemble: integration rule too large for shared memory");
        CK(fast->prepare(smem, &per_sm));
    } else {
        smem = kVol[g.cfg].smem(g.nq);
        if (smem > 227 * 1024) return fail(ctx, B200ASM_EINVAL, "assemble: integration rule too large for shared memory");
        CK(kVol[g.cfg].prepare(smem, &per_sm));
    }
    if (per_sm < 1) return fail(ctx, B200ASM_ECUDA, "assemble: kernel does not fit on an SM");
    for (size_t c = 0; c < nseg; c++) {
        int64_t e0 = g.seg[c], n = g.seg[c + 1] - g.seg[c];
        if (r0 >= 0) { e0 = r0; n = r1 - r0; }  // (nseg == 1)
        if (n == 0) continue;
        VolParams p;
        p.nel = n; p.nq = g.nq; p.kind = g.kind; p.atomic = (ctx->debug & 1) ? 2 : atomic; p.rhs_only = ctx->rhs_only; p.debug = ctx->debug;
        p.xyz = ctx->d_xyz; p.elnodes = g.d_elnodes + e0 * g.nn; p.dest = g.d_dest + e0 * g.m;
        p.qw = g.d_qw; p.phi = g.d_phi; p.dphi = g.d_dphi; p.dng = g.d_dng;
        p.force = g.d_force ? g.d_force + (size_t)e0 * g.nq * g.ns : nullptr;
        p.dng_t = g.d_dng_t; p.dphi_pad = g.d_dphi_pad; p.phi_pad = g.d_phi_pad; p.aux = g.d_aux; p.aux2 = g.d_aux2;
        p.a = ctx->d_a; p.rhs = ctx->d_rhs;
        memcpy(p.coef, g.coef, sizeof(p.coef));
        // persistent grid: SM count x resident CTAs per SM (registers / shared memory decide)
        if (use_mma) {
            const MmaEntry &me = *fast;
            const size_t off = (size_t)e0 * me.slots;
            p.nbatch = 0; p.smap = g.d_smap ? g.d_smap + off : nullptr; p.smapT = g.d_smapT ? g.d_smapT + off : nullptr;
            const int64_t want = (n + me.wpc - 1) / me.wpc;
            CK(me.launch(p, (int)std::min<int64_t>(want, (int64_t)ctx->num_sms * per_sm), smem, ctx->stream));
        } else {
            const VolEntry &ve = kVol[g.cfg];
            p.nbatch = (n + ve.epb - 1) / ve.epb;
            // (no maps yet in a pattern-less rhs-only run); chunk starts are multiples of the batch size
            size_t soff = c < g.seg_smap.size() ? g.seg_smap[c] : 0;
            if (r0 >= 0) soff += (size_t)(r0 / ve.epb) * ve.tile * ve.tile * ve.slots;
            p.smap = g.d_smap ? g.d_smap + soff : nullptr; p.smapT = g.d_smapT ? g.d_smapT + soff : nullptr;
            CK(ve.launch(p, (int)std::min<int64_t>(p.nbatch, (int64_t)ctx->num_sms * per_sm), smem, ctx->stream));
        }
        ctx->launches++;
    }
    return 0;
}

int begin_assembly(b200asm_ctx *ctx) {
    if (!ctx->have_pattern && !ctx->rhs_only) return fail(ctx, B200ASM_ESTATE, "assemble: call b200asm_set_pattern after the last add_group");
    if (!ctx->d_xyz) return fail(ctx, B200ASM_ESTATE, "assemble: call b200asm_set_nodes first");
    for (const Group &g : ctx->groups)
        if (g.max_dest >= ctx->neq) return fail(ctx, B200ASM_EINVAL, "assemble: a group's destination indices exceed the equations of the system");
    CK(cudaSetDevice(ctx->device));
    {
        const int rc = choose_kernels(ctx);
        if (rc) return rc;
    }
    if (!ctx->rhs_only && !ctx->maps_valid) {  // groups were added / the kernel family changed after the pattern was set
        const int rc = build_smaps(ctx, ctx->d_ja);
        if (rc) return rc;
    }
    // Matrix()->Zero() + rhs.Redim of Analysis/TPZLinearAnalysis.cpp:70-75.  With gathering groups only the rows that are not
    // stored outright start from zero (gather_rows.cuh): the stored rows are never read
    if (!ctx->rhs_only) {
        if (ctx->any_gather && ctx->d_rowflag) {
            gat::zero_rows_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->neq, ctx->d_ia, ctx->d_rowflag, ctx->d_a);
            CK(cudaGetLastError());
            ctx->launches++;
        } else {
            CK(cudaMemsetAsync(ctx->d_a, 0, std::max<int64_t>(ctx->nnz, 1) * sizeof(double), ctx->stream));
        }
    }
    CK(cudaMemsetAsync(ctx->d_rhs, 0, std::max<int64_t>(ctx->neq, 1) * sizeof(double), ctx->stream));
    return 0;
}

}  // namespace

namespace {

// One launch unit of an assembly: a whole group (r0 < 0) or the element range [r0, r1) of it.  Order: the interface prefixes of
// all groups (their contributions are pushed to the owning GPUs behind them), the small groups, then the element chunks of the
// large volume groups in element order (the order the overlapped download relies on).
struct Unit { int group; int64_t r0, r1, min_dest; bool prefix; };

void build_units(const b200asm_ctx *ctx, std::vector<Unit> &units, size_t &nprefix) {
    units.clear();
    for (size_t gi = 0; gi < ctx->groups.size(); gi++) {
        const Group &g = ctx->groups[gi];
        if (g.nel == 0 || g.n_if == 0 || g.use_gather) continue;
        units.push_back({(int)gi, 0, g.chunk[1], g.chunk_min[0], true});
    }
    nprefix = units.size();
    for (int pass = 0; pass < 2; pass++)
        for (size_t gi = 0; gi < ctx->groups.size(); gi++) {
            const Group &g = ctx->groups[gi];
            if (g.nel == 0) continue;
            if (g.use_gather) {  // node-block ranges in row order (they run last: the other groups add into zeroed rows first)
                if (pass == 1)
                    for (size_t c = 0; c + 1 < g.gchunk.size(); c++) units.push_back({(int)gi, g.gchunk[c], g.gchunk[c + 1], g.gchunk_min[c], false});
                continue;
            }
            const bool chunked = g.chunk.size() > 2;
            if (chunked != (pass == 1)) continue;
            if (!chunked) {
                if (g.n_if == 0) units.push_back({(int)gi, -1, -1, g.chunk_min.empty() ? 0 : g.chunk_min[0], false});
            } else {
                for (size_t c = g.n_if > 0 ? 1 : 0; c + 1 < g.chunk.size(); c++)
                    units.push_back({(int)gi, g.chunk[c], g.chunk[c + 1], g.chunk_min[c], false});
            }
        }
}

// after the memsets of begin_assembly: tell every GPU that pushes into this one that the arrays are zeroed
int exchange_signal_zeroed(b200asm_ctx *ctx) {
    for (const b200asm_ctx::PeerLink &l : ctx->links) {
        if (l.push) continue;
        xch::signal_kernel<<<1, 1, 0, ctx->stream>>>(l.flags + l.slot_there, ctx->xstep);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    return 0;
}

// behind the interface elements: push the staged contributions on the side stream (waits for the owner's ZEROED signal first)
int exchange_push(b200asm_ctx *ctx) {
    bool any = false;
    for (const b200asm_ctx::PeerLink &l : ctx->links) any = any || l.push;
    if (!any) return 0;
    if (!ctx->xstream) {
        CK(cudaStreamCreateWithFlags(&ctx->xstream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_if, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_push, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(ctx->ev_if, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->xstream, ctx->ev_if, 0));
    const unsigned long long timeout_ns = (unsigned long long)ctx->xtimeout_ms * 1000000ull;
    for (size_t i = 0; i < ctx->links.size(); i++) {
        const b200asm_ctx::PeerLink &l = ctx->links[i];
        if (!l.push) continue;
        xch::wait_kernel<<<1, 1, 0, ctx->xstream>>>(ctx->d_flags + i, ctx->xstep, timeout_ns, ctx->d_xerr);
        if (l.n_a && !ctx->rhs_only)
            xch::push_values_kernel<<<grid_for(ctx, l.n_a, 256), 256, 0, ctx->xstream>>>(l.a, l.d_a_dst, ctx->d_a + l.a_src0, l.n_a);
        if (l.n_rhs)
            xch::push_gather_kernel<<<grid_for(ctx, l.n_rhs, 256), 256, 0, ctx->xstream>>>(l.rhs, l.d_rhs_dst, ctx->d_rhs, l.d_rhs_src, l.n_rhs);
        xch::signal_kernel<<<1, 1, 0, ctx->xstream>>>(l.flags + xch::MAX_LINKS + l.slot_there, ctx->xstep);
        CK(cudaGetLastError());
        ctx->launches += 4;
    }
    CK(cudaEventRecord(ctx->ev_push, ctx->xstream));
    return 0;
}

// end of the step on the main stream: the own pushes are done (the staging rows may be zeroed again) and every peer that pushes
// into this context has delivered: the owned rows are complete
int exchange_finish(b200asm_ctx *ctx) {
    bool any = false;
    for (const b200asm_ctx::PeerLink &l : ctx->links) any = any || l.push;
    if (any) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_push, 0));
    const unsigned long long timeout_ns = (unsigned long long)ctx->xtimeout_ms * 1000000ull;
    for (size_t i = 0; i < ctx->links.size(); i++) {
        if (ctx->links[i].push) continue;
        xch::wait_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_flags + xch::MAX_LINKS + i, ctx->xstep, timeout_ns, ctx->d_xerr);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    return 0;
}

int64_t dl_nnz(const b200asm_ctx *ctx) { return ctx->dl_a < 0 ? ctx->nnz : std::min(ctx->dl_a, ctx->nnz); }
int64_t dl_rhs_first(const b200asm_ctx *ctx) { return std::min(std::max<int64_t>(ctx->dl_r0, 0), ctx->neq); }
int64_t dl_rhs_count(const b200asm_ctx *ctx) {
    const int64_t f = dl_rhs_first(ctx);
    return ctx->dl_rn < 0 ? ctx->neq - f : std::min(ctx->dl_rn, ctx->neq - f);
}

int check_exchange_error(b200asm_ctx *ctx) {
    if (ctx->links.empty()) return 0;
    int err = 0;
    CK(cudaMemcpyAsync(&err, ctx->d_xerr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (err) {
        CK(cudaMemsetAsync(ctx->d_xerr, 0, sizeof(int), ctx->stream));
        return fail(ctx, B200ASM_ECUDA, "interface exchange: a peer GPU did not signal within the timeout (do all contexts assemble the same number of times?)");
    }
    return 0;
}

// events of the "timing" option around the launches of a group (prefix and remainder apart: other groups run in between)
int time_mark(b200asm_ctx *ctx, Group &g, bool prefix, bool begin) {
    if (!g.ev0) {
        CK(cudaEventCreate(&g.ev0)); CK(cudaEventCreate(&g.ev1)); CK(cudaEventCreate(&g.ev2)); CK(cudaEventCreate(&g.ev3));
    }
    CK(cudaEventRecord(prefix ? (begin ? g.ev2 : g.ev3) : (begin ? g.ev0 : g.ev1), ctx->stream));
    if (prefix) g.timed_prefix = true;
    return 0;
}

}  // namespace

extern "C" int b200asm_assemble_async(b200asm_ctx *ctx) {
    if (!ctx) return B200ASM_EINVAL;
    int rc = begin_assembly(ctx);
    if (rc) return rc;
    ctx->xstep++;
    if ((rc = exchange_signal_zeroed(ctx))) return rc;
    std::vector<Unit> units;
    size_t nprefix = 0;
    build_units(ctx, units, nprefix);
    for (Group &g : ctx->groups) g.timed_prefix = false;
    bool pushed = false;
    for (size_t u = 0; u <= units.size(); u++) {
        if (u == nprefix && !ctx->links.empty() && nprefix > 0) {  // every interface element has been launched
            if ((rc = exchange_push(ctx))) return rc;
            pushed = true;
        }
        if (u == units.size()) break;
        Group &g = ctx->groups[units[u].group];
        const bool first_of_group = u == 0 || units[u - 1].group != units[u].group || units[u - 1].prefix != units[u].prefix;
        const bool last_of_group = u + 1 == units.size() || units[u + 1].group != units[u].group || units[u + 1].prefix != units[u].prefix;
        if (ctx->timing && first_of_group && (rc = time_mark(ctx, g, units[u].prefix, true))) return rc;
        if ((rc = enqueue_group(ctx, g, units[u].r0, units[u].r1))) return rc;
        if (ctx->timing && last_of_group && (rc = time_mark(ctx, g, units[u].prefix, false))) return rc;
    }
    if (!ctx->links.empty()) {
        if (!pushed && (rc = exchange_push(ctx))) return rc;  // (coloured groups / no prefix: push behind the last launch)
        if ((rc = exchange_finish(ctx))) return rc;
    }
    return 0;
}

extern "C" int b200asm_group_time_ms(b200asm_ctx *ctx, int group, double *ms) {
    if (!ctx || group < 0 || group >= (int)ctx->groups.size() || !ms) return fail(ctx, B200ASM_EINVAL, "group_time_ms: bad arguments");
    const Group &g = ctx->groups[group];
    if (!g.ev0) return fail(ctx, B200ASM_ESTATE, "group_time_ms: set option \"timing\" = 1 and assemble first");
    CK(cudaEventSynchronize(g.ev1));
    float t = 0.f, tp = 0.f;
    CK(cudaEventElapsedTime(&t, g.ev0, g.ev1));
    if (g.timed_prefix) {
        CK(cudaEventSynchronize(g.ev3));
        CK(cudaEventElapsedTime(&tp, g.ev2, g.ev3));
    }
    *ms = (double)t + (double)tp;
    return 0;
}

extern "C" int b200asm_synchronize(b200asm_ctx *ctx) {
    if (!ctx) return B200ASM_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_exchange_error(ctx);
}

extern "C" int b200asm_download(b200asm_ctx *ctx, double *a_host, double *rhs_host) {
    if (!ctx) return B200ASM_EINVAL;
    if (!ctx->d_a) return fail(ctx, B200ASM_ESTATE, "download: nothing assembled");
    CK(cudaSetDevice(ctx->device));
    if (a_host && dl_nnz(ctx)) {
        CK(cudaMemcpyAsync(a_host, ctx->d_a, (size_t)dl_nnz(ctx) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h += dl_nnz(ctx) * (int64_t)sizeof(double);
    }
    if (rhs_host && dl_rhs_count(ctx)) {
        CK(cudaMemcpyAsync(rhs_host, ctx->d_rhs + dl_rhs_first(ctx), (size_t)dl_rhs_count(ctx) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h += dl_rhs_count(ctx) * (int64_t)sizeof(double);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return check_exchange_error(ctx);
}

namespace {

// b200asm_assemble with a host matrix: the D2H copy of the CSR values (PCIe, several times longer than the assembly
// itself) starts while the kernels still run.  Launch units = whole small groups first, then the element chunks of the
// large volume groups in element order.  An entry of A lives in row min(dest_i, dest_j) (symmetric) or dest_i (full), so
// every row below the smallest destination equation of all LATER units is final once the current unit has finished:
// that prefix of A is copied on a second stream behind an event.  Atomic scatter only (the coloured mode keeps its
// deterministic launch order).
int assemble_overlapped(b200asm_ctx *ctx, double *a_host, double *rhs_host) {
    int rc = begin_assembly(ctx);
    if (rc) return rc;
    ctx->xstep++;
    if ((rc = exchange_signal_zeroed(ctx))) return rc;
    std::vector<Unit> units;
    size_t nprefix = 0;
    build_units(ctx, units, nprefix);
    for (Group &g : ctx->groups) g.timed_prefix = false;
    // frontier[u] = first row that a unit after u (or a peer GPU pushing into this one) may still touch
    std::vector<int64_t> frontier(units.size(), std::min(ctx->neq, ctx->incoming_min_row));
    for (int64_t u = (int64_t)units.size() - 2; u >= 0; u--)
        frontier[u] = std::min(frontier[u + 1], std::min<int64_t>(units[u + 1].min_dest, ctx->neq));
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // IA at the frontiers (the pattern is device-resident): read once per set of scatter maps
    if (!ctx->ov_valid || ctx->ov_upto.size() != units.size()) {
        ctx->ov_upto.assign(units.size(), 0);
        for (size_t u = 0; u + 1 < units.size(); u++)
            if (frontier[u] > 0)
                CK(cudaMemcpyAsync(&ctx->ov_upto[u], ctx->d_ia + frontier[u], sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
        ctx->ov_valid = true;
    }
    const int64_t kMinCopy = std::max<int64_t>(1, ctx->overlap_min_bytes / (int64_t)sizeof(double));
    // staging rows never go to the host early: their values are partial sums that belong to another GPU
    int64_t done = 0;                               // entries of A already queued for download
    size_t nev = 0;
    bool pushed = false;
    for (size_t u = 0; u < units.size(); u++) {
        if (u == nprefix && !ctx->links.empty() && nprefix > 0) {
            if ((rc = exchange_push(ctx))) return rc;
            pushed = true;
        }
        Group &g = ctx->groups[units[u].group];
        const bool first_of_group = u == 0 || units[u - 1].group != units[u].group || units[u - 1].prefix != units[u].prefix;
        const bool last_of_group = u + 1 == units.size() || units[u + 1].group != units[u].group || units[u + 1].prefix != units[u].prefix;
        if (ctx->timing && first_of_group && (rc = time_mark(ctx, g, units[u].prefix, true))) return rc;
        if ((rc = enqueue_group(ctx, g, units[u].r0, units[u].r1))) return rc;
        if (ctx->timing && last_of_group && (rc = time_mark(ctx, g, units[u].prefix, false))) return rc;
        if (u + 1 == units.size()) break;
        const int64_t upto = std::min(ctx->ov_upto[u], dl_nnz(ctx));
        if (upto - done < kMinCopy) continue;
        if (nev == ctx->copy_events.size()) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->copy_events.push_back(e);
        }
        CK(cudaEventRecord(ctx->copy_events[nev], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[nev], 0));
        nev++;
        CK(cudaMemcpyAsync(a_host + done, ctx->d_a + done, (size_t)(upto - done) * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        done = upto;
    }
    if (!ctx->links.empty()) {
        if (!pushed && (rc = exchange_push(ctx))) return rc;
        if ((rc = exchange_finish(ctx))) return rc;
    }
    // the rest of A and the load vector behind the last kernel (and the last incoming push)
    if (nev == ctx->copy_events.size()) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->copy_events.push_back(e);
    }
    CK(cudaEventRecord(ctx->copy_events[nev], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[nev], 0));
    if (dl_nnz(ctx) > done)
        CK(cudaMemcpyAsync(a_host + done, ctx->d_a + done, (size_t)(dl_nnz(ctx) - done) * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
    ctx->d2h += dl_nnz(ctx) * (int64_t)sizeof(double);
    if (rhs_host && dl_rhs_count(ctx)) {
        CK(cudaMemcpyAsync(rhs_host, ctx->d_rhs + dl_rhs_first(ctx), (size_t)dl_rhs_count(ctx) * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        ctx->d2h += dl_rhs_count(ctx) * (int64_t)sizeof(double);
    }
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_exchange_error(ctx);
}

}  // namespace

extern "C" int b200asm_assemble(b200asm_ctx *ctx, double *a_host, double *rhs_host) {
    if (!ctx) return B200ASM_EINVAL;
    if (a_host && ctx->overlap && ctx->scatter == B200ASM_SCATTER_ATOMIC && !ctx->rhs_only && ctx->have_pattern) {
        bool chunked = false;
        for (const Group &g : ctx->groups) chunked = chunked || (g.use_gather ? g.gchunk.size() > 2 : g.chunk.size() > 2);
        // pageable destinations make cudaMemcpyAsync block the launching thread: overlap only into pinned / registered memory
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, a_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (chunked && pinned) return assemble_overlapped(ctx, a_host, rhs_host);
    }
    int rc = b200asm_assemble_async(ctx);
    if (rc) return rc;
    return b200asm_download(ctx, a_host, rhs_host);
}

extern "C" int b200asm_pin_host(b200asm_ctx *ctx, void *ptr, size_t bytes) {
    if (!ctx || !ptr || !bytes) return fail(ctx, B200ASM_EINVAL, "pin_host: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

extern "C" int b200asm_unpin_host(b200asm_ctx *ctx, void *ptr) {
    if (!ctx || !ptr) return fail(ctx, B200ASM_EINVAL, "unpin_host: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(ptr));
    return 0;
}

extern "C" int b200asm_assemble_rhs(b200asm_ctx *ctx, double *rhs_host) {
    if (!ctx) return B200ASM_EINVAL;
    if (!ctx->have_pattern) {
        // no matrix pattern (TPZLinearAnalysis::AssembleResidual before any Assemble): size the load vector from the
        // destination indices of the groups
        int64_t neq = 0;
        for (const Group &g : ctx->groups) neq = std::max(neq, g.max_dest + 1);
        if (neq != ctx->neq || !ctx->d_rhs) {
            CK(cudaSetDevice(ctx->device));
            cudaFree(ctx->d_rhs);
            ctx->d_rhs = nullptr;
            CK(cudaMalloc((void **)&ctx->d_rhs, std::max<int64_t>(neq, 1) * sizeof(double)));
            ctx->neq = neq;
        }
    }
    ctx->rhs_only = 1;
    int rc = b200asm_assemble_async(ctx);
    ctx->rhs_only = 0;
    if (rc) return rc;
    if (rhs_host && dl_rhs_count(ctx)) {
        CK(cudaMemcpyAsync(rhs_host, ctx->d_rhs + dl_rhs_first(ctx), (size_t)dl_rhs_count(ctx) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h += dl_rhs_count(ctx) * (int64_t)sizeof(double);
        CK(cudaStreamSynchronize(ctx->stream));
        return check_exchange_error(ctx);
    }
    return 0;
}


// ---- multi-GPU: interface exchange over peer memory --------------------------------------------------------------
namespace {
void free_links(b200asm_ctx *ctx) {
    for (b200asm_ctx::PeerLink &l : ctx->links) {
        cudaFree(l.d_a_dst); cudaFree(l.d_rhs_src); cudaFree(l.d_rhs_dst);
        if (l.ipc)
            for (void *b : l.ipc_base)
                if (b) cudaIpcCloseMemHandle(b);
    }
    ctx->links.clear();
    ctx->incoming_min_row = INT64_MAX;
}
int ensure_flags(b200asm_ctx *ctx) {
    if (ctx->d_flags) return 0;
    CK(cudaMalloc((void **)&ctx->d_flags, 2 * xch::MAX_LINKS * sizeof(unsigned long long)));
    CK(cudaMalloc((void **)&ctx->d_xerr, sizeof(int)));
    CK(cudaMemset(ctx->d_flags, 0, 2 * xch::MAX_LINKS * sizeof(unsigned long long)));
    CK(cudaMemset(ctx->d_xerr, 0, sizeof(int)));
    return 0;
}
// cuMemGetAddressRange through the runtime's driver entry-point query (the library does not link libcuda: it must load on
// hosts without a driver for the host-side helpers)
int cuMemGetAddressRange_shim(void **base, size_t *size, void *ptr) {
    typedef int (*fn_t)(unsigned long long *, size_t *, unsigned long long);
    static fn_t fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
        fn = (fn_t)f;
    }
    unsigned long long b = 0;
    size_t sz = 0;
    if (fn(&b, &sz, (unsigned long long)(uintptr_t)ptr) != 0) return -1;
    *base = (void *)(uintptr_t)b;
    *size = sz;
    return 0;
}
int export_one(b200asm_ctx *ctx, void *ptr, b200asm_ipc_mem *out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(out->handle), "b200asm_ipc_mem::handle holds a cudaIpcMemHandle_t");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(out->handle, &h, sizeof(h));
    // the handle names the whole allocation the pointer lies in: the importer adds the offset
    void *base = nullptr;
    size_t size = 0;
    if (cuMemGetAddressRange_shim(&base, &size, ptr) != 0) return fail(ctx, B200ASM_ECUDA, "exchange_export: address range query failed");
    out->offset = (int64_t)((char *)ptr - (char *)base);
    return 0;
}
}  // namespace

extern "C" int b200asm_exchange_export(b200asm_ctx *ctx, b200asm_ipc_mem out[3]) {
    if (!ctx || !out) return B200ASM_EINVAL;
    if (!ctx->have_pattern || !ctx->d_a || !ctx->d_rhs) return fail(ctx, B200ASM_ESTATE, "exchange_export: create the pattern first (it allocates the arrays the peers map)");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_flags(ctx))) return rc;
    if ((rc = export_one(ctx, ctx->d_a, &out[0])) || (rc = export_one(ctx, ctx->d_rhs, &out[1])) || (rc = export_one(ctx, ctx->d_flags, &out[2]))) return rc;
    return 0;
}

extern "C" int b200asm_exchange_add_peer(b200asm_ctx *ctx, int push, int slot_there, const b200asm_ipc_mem mem[3], b200asm_ctx *peer,
                                         int64_t incoming_min_row) {
    if (!ctx || slot_there < 0 || slot_there >= xch::MAX_LINKS || (!mem == !peer)) return fail(ctx, B200ASM_EINVAL, "exchange_add_peer: bad arguments (give IPC handles or a context of this process)");
    if ((int)ctx->links.size() >= xch::MAX_LINKS) return fail(ctx, B200ASM_EINVAL, "exchange_add_peer: too many peers");
    if (!ctx->have_pattern) return fail(ctx, B200ASM_ESTATE, "exchange_add_peer: create the pattern first");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_flags(ctx))) return rc;
    b200asm_ctx::PeerLink l;
    l.push = push != 0;
    l.slot_there = slot_there;
    if (peer) {
        if (peer == ctx) return fail(ctx, B200ASM_EINVAL, "exchange_add_peer: a context cannot be its own peer");
        if (!peer->have_pattern) return fail(ctx, B200ASM_ESTATE, "exchange_add_peer: the peer has no pattern yet");
        if (peer->device != ctx->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->device, peer->device));
            if (!can) return fail(ctx, B200ASM_ECUDA, "exchange_add_peer: no peer access between the two devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, B200ASM_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        CK(cudaSetDevice(peer->device));
        if ((rc = ensure_flags(peer))) { cudaSetDevice(ctx->device); return fail(ctx, rc, peer->err); }
        CK(cudaSetDevice(ctx->device));
        l.a = peer->d_a; l.rhs = peer->d_rhs; l.flags = peer->d_flags;
    } else {
        l.ipc = true;
        void **dst[3] = {(void **)&l.a, (void **)&l.rhs, (void **)&l.flags};
        for (int k = 0; k < 3; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, mem[k].handle, sizeof(h));
            cudaError_t e = cudaIpcOpenMemHandle(&l.ipc_base[k], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int j = 0; j < k; j++) cudaIpcCloseMemHandle(l.ipc_base[j]);
                return fail(ctx, B200ASM_ECUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
            *dst[k] = (char *)l.ipc_base[k] + mem[k].offset;
        }
    }
    if (!l.push && incoming_min_row >= 0) ctx->incoming_min_row = std::min(ctx->incoming_min_row, incoming_min_row);
    ctx->links.push_back(l);
    ctx->ov_valid = false;
    for (const Group &g : ctx->groups)
        if (g.use_gather) ctx->maps_valid = false;  // another GPU adds into this context's rows: back to the scatter kernels
    return (int)ctx->links.size() - 1;
}

extern "C" int b200asm_exchange_set_map(b200asm_ctx *ctx, int link, int64_t n_a, int64_t a_src0, const int32_t *a_dst, int64_t n_rhs,
                                        const int32_t *rhs_src, const int32_t *rhs_dst) {
    if (!ctx || link < 0 || link >= (int)ctx->links.size() || n_a < 0 || n_rhs < 0 || (n_a && !a_dst) || (n_rhs && (!rhs_src || !rhs_dst)))
        return fail(ctx, B200ASM_EINVAL, "exchange_set_map: bad arguments");
    b200asm_ctx::PeerLink &l = ctx->links[link];
    if (!l.push) return fail(ctx, B200ASM_EINVAL, "exchange_set_map: the link does not push");
    if (a_src0 < 0 || a_src0 + n_a > ctx->nnz) return fail(ctx, B200ASM_EINVAL, "exchange_set_map: staging segment exceeds the CSR values");
    for (int64_t k = 0; k < n_rhs; k++)
        if (rhs_src[k] < 0 || rhs_src[k] >= ctx->neq) return fail(ctx, B200ASM_EINVAL, "exchange_set_map: rhs source out of range");
    CK(cudaSetDevice(ctx->device));
    cudaFree(l.d_a_dst); cudaFree(l.d_rhs_src); cudaFree(l.d_rhs_dst);
    l.d_a_dst = l.d_rhs_src = l.d_rhs_dst = nullptr;
    int rc;
    if ((rc = upload(ctx, &l.d_a_dst, a_dst, (size_t)n_a)) || (rc = upload(ctx, &l.d_rhs_src, rhs_src, (size_t)n_rhs)) ||
        (rc = upload(ctx, &l.d_rhs_dst, rhs_dst, (size_t)n_rhs)))
        return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    l.n_a = n_a; l.a_src0 = a_src0; l.n_rhs = n_rhs;
    return 0;
}

extern "C" int b200asm_exchange_clear(b200asm_ctx *ctx) {
    if (!ctx) return B200ASM_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->xstream) cudaStreamSynchronize(ctx->xstream);
    free_links(ctx);
    ctx->xstep = 0;
    if (ctx->d_flags) cudaMemset(ctx->d_flags, 0, 2 * xch::MAX_LINKS * sizeof(unsigned long long));
    return 0;
}

extern "C" int b200asm_scatter_add(b200asm_ctx *ctx, int target, const int32_t *positions_dev, const double *values_dev, int64_t n) {
    if (!ctx || n < 0 || (n && (!positions_dev || !values_dev)) || (target != 0 && target != 1))
        return fail(ctx, B200ASM_EINVAL, "scatter_add: bad arguments");
    double *dst = target == 0 ? ctx->d_a : ctx->d_rhs;
    if (!dst) return fail(ctx, B200ASM_ESTATE, "scatter_add: nothing assembled");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8);
    scatter_add_kernel<<<grid, 256, 0, ctx->stream>>>(dst, positions_dev, values_dev, n);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

extern "C" int b200asm_cg_solve(b200asm_ctx *ctx, int precond, int64_t max_iter, double tol, int from_current, const double *f_host,
                                double *x_host, int64_t *iters_out, double *resid_out) {
    if (!ctx || max_iter < 0 || (precond != 0 && precond != 1)) return fail(ctx, B200ASM_EINVAL, "cg_solve: bad arguments");
    if (!ctx->have_pattern || !ctx->d_a || !ctx->d_ja) return fail(ctx, B200ASM_ESTATE, "cg_solve: assemble a matrix first");
    if (from_current && !x_host && ctx->cg_n != ctx->neq) return fail(ctx, B200ASM_EINVAL, "cg_solve: from_current needs an initial guess");
    CK(cudaSetDevice(ctx->device));
    const int64_t n = ctx->neq;
    cudaStream_t st = ctx->stream;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + cgdev::THREADS - 1) / cgdev::THREADS, (int64_t)ctx->num_sms * 8));
    const int grid_rows = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)ctx->num_sms * 32));
    if (ctx->cg_n != n || !ctx->d_cg) {
        cudaFree(ctx->d_cg); cudaFree(ctx->d_cg_part); cudaFree(ctx->d_cg_sc);
        ctx->d_cg = ctx->d_cg_part = nullptr; ctx->d_cg_sc = nullptr; ctx->cg_n = 0;
        CK(cudaMalloc((void **)&ctx->d_cg, (size_t)std::max<int64_t>(n, 1) * 7 * sizeof(double)));
        CK(cudaMalloc((void **)&ctx->d_cg_part, (size_t)ctx->num_sms * 8 * sizeof(double)));
        CK(cudaMalloc((void **)&ctx->d_cg_sc, sizeof(cgdev::Scalars)));
        CK(cudaMemsetAsync(ctx->d_cg, 0, (size_t)std::max<int64_t>(n, 1) * 7 * sizeof(double), st));
        ctx->cg_n = n;
    }
    double *x = ctx->d_cg, *r = x + n, *p = r + n, *z = p + n, *q = z + n, *diag = q + n, *f = diag + n;
    cgdev::Scalars *sc = ctx->d_cg_sc;
    CK(cudaMemsetAsync(sc, 0, sizeof(cgdev::Scalars), st));
    // right-hand side: the host vector the caller passes (TPZMatrixSolver::Solve(F, ...)) or the assembled load vector
    if (f_host) {
        CK(cudaMemcpyAsync(f, f_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        ctx->h2d += n * (int64_t)sizeof(double);
    } else {
        CK(cudaMemcpyAsync(f, ctx->d_rhs, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    auto read_scalars = [&](cgdev::Scalars &h) -> cudaError_t {
        cudaError_t e = cudaMemcpyAsync(&h, sc, sizeof(h), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(st);
    };
    cgdev::Scalars h{};
    // normb = Norm(b)
    cgdev::dot_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, f, f, ctx->d_cg_part, sc, 1);
    CK(read_scalars(h));
    double normb = std::sqrt(h.rr);
    if (normb == 0.0) normb = 1.0;
    // r = b - A x  (FromCurrent)  or  x = 0, r = b
    CK(cudaMemcpyAsync(r, f, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (from_current) {
        if (x_host) {
            CK(cudaMemcpyAsync(x, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
            ctx->h2d += n * (int64_t)sizeof(double);
        }
        cgdev::spmv_kernel<<<grid_rows, cgdev::THREADS, 0, st>>>(n, ctx->d_ia, ctx->d_ja, ctx->d_a, ctx->symmetric, -1.0, x, r);
        ctx->launches++;
    } else {
        CK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), st));
    }
    if (precond == 1) {
        cgdev::extract_diag_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, ctx->d_ia, ctx->d_ja, ctx->d_a, ctx->symmetric, diag);
        ctx->launches++;
    }
    cgdev::dot_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, r, r, ctx->d_cg_part, sc, 1);
    ctx->launches += 2;
    CK(read_scalars(h));
    double resid = std::sqrt(h.rr) / normb;
    int64_t it = 0;
    if (resid > tol) {
        for (it = 1; it <= max_iter; it++) {
            cgdev::precond_dot_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, r, precond == 1 ? diag : nullptr, z, ctx->d_cg_part, sc);
            cgdev::update_p_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, it == 1, z, p, sc);
            CK(cudaMemsetAsync(q, 0, (size_t)n * sizeof(double), st));
            cgdev::spmv_kernel<<<grid_rows, cgdev::THREADS, 0, st>>>(n, ctx->d_ia, ctx->d_ja, ctx->d_a, ctx->symmetric, 1.0, p, q);
            cgdev::dot_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, p, q, ctx->d_cg_part, sc, 0);
            cgdev::update_xr_kernel<<<grid, cgdev::THREADS, 0, st>>>(n, p, q, x, r, ctx->d_cg_part, sc);
            ctx->launches += 5;
            CK(read_scalars(h));
            resid = std::sqrt(h.rr) / normb;
            if (!(resid == resid)) return fail(ctx, B200ASM_ECUDA, "cg_solve: the iteration produced NaN (singular or indefinite system)");
            if (resid <= tol) break;
        }
        if (it > max_iter) it = max_iter;
    }
    CK(cudaGetLastError());
    if (x_host) {
        CK(cudaMemcpyAsync(x_host, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->d2h += n * (int64_t)sizeof(double);
    }
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = resid;
    return 0;
}

// ---- row-sharded conjugate gradients (several GPUs of this process; driven by b200asm_multi_cg_solve in multi.cpp) ----------
#include "cg_sharded.h"
int b200asm_cg_sharded(int nshards, const b200asm_cg_shard *sh, int precond, int64_t max_iter, double tol, int from_current,
                       const double *f_host, double *x_host, int64_t *iters_out, double *resid_out, std::string &err) {
#define CKS(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                              \
            return B200ASM_ECUDA;                                                                  \
        }                                                                                          \
    } while (0)
    if (nshards < 1 || !sh || max_iter < 0 || (precond != 0 && precond != 1)) { err = "cg_sharded: bad arguments"; return B200ASM_EINVAL; }
    struct Dev {
        double *x, *r, *p, *z, *q, *diag, *f;
        int32_t *h_local = nullptr, *h_owner = nullptr, *h_remote = nullptr;
        double **peer_p = nullptr, **peer_q = nullptr;
        int grid = 1, grid_rows = 1, grid_halo = 1;
    };
    std::vector<Dev> dv(nshards);
    auto cleanup = [&]() {
        for (int k = 0; k < nshards; k++) {
            cudaSetDevice(sh[k].ctx->device);
            cudaFree(dv[k].h_local); cudaFree(dv[k].h_owner); cudaFree(dv[k].h_remote); cudaFree(dv[k].peer_p); cudaFree(dv[k].peer_q);
        }
    };
    // workspace per shard: 7 vectors of the LOCAL length (x, r, p, z, q, diag, f), the owned segment is authoritative
    for (int k = 0; k < nshards; k++) {
        b200asm_ctx *ctx = sh[k].ctx;
        if (!ctx->have_pattern || !ctx->d_a || !ctx->d_ja) { err = "cg_sharded: assemble a matrix first"; return B200ASM_ESTATE; }
        CKS(cudaSetDevice(ctx->device));
        const int64_t n = ctx->neq;
        if (ctx->cg_n != n || !ctx->d_cg) {
            if (from_current && !x_host) { err = "cg_sharded: from_current needs an initial guess"; return B200ASM_EINVAL; }
            cudaFree(ctx->d_cg); cudaFree(ctx->d_cg_part); cudaFree(ctx->d_cg_sc);
            ctx->d_cg = ctx->d_cg_part = nullptr; ctx->d_cg_sc = nullptr; ctx->cg_n = 0;
            CKS(cudaMalloc((void **)&ctx->d_cg, (size_t)std::max<int64_t>(n, 1) * 7 * sizeof(double)));
            CKS(cudaMalloc((void **)&ctx->d_cg_part, (size_t)ctx->num_sms * 8 * sizeof(double)));
            CKS(cudaMalloc((void **)&ctx->d_cg_sc, sizeof(cgdev::Scalars)));
            CKS(cudaMemsetAsync(ctx->d_cg, 0, (size_t)std::max<int64_t>(n, 1) * 7 * sizeof(double), ctx->stream));
            ctx->cg_n = n;
        }
        Dev &d = dv[k];
        d.x = ctx->d_cg; d.r = d.x + n; d.p = d.r + n; d.z = d.p + n; d.q = d.z + n; d.diag = d.q + n; d.f = d.diag + n;
        const int64_t no = std::max<int64_t>(sh[k].nown, 1);
        d.grid = (int)std::max<int64_t>(1, std::min<int64_t>((no + cgdev::THREADS - 1) / cgdev::THREADS, (int64_t)ctx->num_sms * 8));
        d.grid_rows = (int)std::max<int64_t>(1, std::min<int64_t>((no + 7) / 8, (int64_t)ctx->num_sms * 32));
        d.grid_halo = (int)std::max<int64_t>(1, std::min<int64_t>((sh[k].nhalo + 255) / 256, (int64_t)ctx->num_sms * 8));
    }
    // halo maps and the peers' vector addresses; peer access in both directions of every halo relation
    for (int k = 0; k < nshards; k++) {
        b200asm_ctx *ctx = sh[k].ctx;
        CKS(cudaSetDevice(ctx->device));
        Dev &d = dv[k];
        std::vector<double *> pp(nshards), pq(nshards);
        for (int j = 0; j < nshards; j++) { pp[j] = dv[j].p; pq[j] = dv[j].q; }
        CKS(cudaMalloc((void **)&d.peer_p, nshards * sizeof(double *)));
        CKS(cudaMalloc((void **)&d.peer_q, nshards * sizeof(double *)));
        CKS(cudaMemcpyAsync(d.peer_p, pp.data(), nshards * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream));
        CKS(cudaMemcpyAsync(d.peer_q, pq.data(), nshards * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream));
        if (sh[k].nhalo) {
            const size_t b = (size_t)sh[k].nhalo * sizeof(int32_t);
            CKS(cudaMalloc((void **)&d.h_local, b)); CKS(cudaMalloc((void **)&d.h_owner, b)); CKS(cudaMalloc((void **)&d.h_remote, b));
            CKS(cudaMemcpyAsync(d.h_local, sh[k].halo_local, b, cudaMemcpyHostToDevice, ctx->stream));
            CKS(cudaMemcpyAsync(d.h_owner, sh[k].halo_owner, b, cudaMemcpyHostToDevice, ctx->stream));
            CKS(cudaMemcpyAsync(d.h_remote, sh[k].halo_remote, b, cudaMemcpyHostToDevice, ctx->stream));
            std::vector<char> need(nshards, 0);
            for (int64_t i = 0; i < sh[k].nhalo; i++) need[sh[k].halo_owner[i]] = 1;
            for (int j = 0; j < nshards; j++) {
                if (!need[j] || sh[j].ctx->device == ctx->device) continue;
                int can = 0;
                CKS(cudaDeviceCanAccessPeer(&can, ctx->device, sh[j].ctx->device));
                if (!can) { err = "cg_sharded: no peer access between two devices"; cleanup(); return B200ASM_ECUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(sh[j].ctx->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); cleanup(); return B200ASM_ECUDA; }
                cudaGetLastError();
            }
        }
        CKS(cudaStreamSynchronize(ctx->stream));
    }
    auto sync_all = [&]() -> cudaError_t {
        for (int k = 0; k < nshards; k++) {
            cudaError_t e = cudaSetDevice(sh[k].ctx->device);
            if (e == cudaSuccess) e = cudaStreamSynchronize(sh[k].ctx->stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    // sum over the shards (in shard order) of one device scalar each
    std::vector<cgdev::Scalars> hs(nshards);
    auto gather_scalars = [&]() -> cudaError_t {
        for (int k = 0; k < nshards; k++) {
            cudaError_t e = cudaSetDevice(sh[k].ctx->device);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&hs[k], sh[k].ctx->d_cg_sc, sizeof(cgdev::Scalars), cudaMemcpyDeviceToHost, sh[k].ctx->stream);
            if (e != cudaSuccess) return e;
        }
        return sync_all();
    };
    // q = A p over all shards: halo of p in, product, halo of q out
    auto product = [&](int which /*0: p -> q ; 1: x -> q (initial residual)*/) -> cudaError_t {
        for (int k = 0; k < nshards; k++) {   // the owners' values must be complete before anybody pulls them
            b200asm_ctx *ctx = sh[k].ctx;
            cudaSetDevice(ctx->device);
            cudaError_t e = cudaMemsetAsync(dv[k].q, 0, (size_t)std::max<int64_t>(ctx->neq, 1) * sizeof(double), ctx->stream);
            if (e != cudaSuccess) return e;
            if (which == 1) {  // x rides in p's place for the halo pull
                e = cudaMemcpyAsync(dv[k].p + sh[k].own_first, dv[k].x + sh[k].own_first, (size_t)sh[k].nown * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
                if (e != cudaSuccess) return e;
            }
        }
        cudaError_t e = sync_all();
        if (e != cudaSuccess) return e;
        for (int k = 0; k < nshards; k++) {
            b200asm_ctx *ctx = sh[k].ctx;
            cudaSetDevice(ctx->device);
            if (sh[k].nhalo)
                cgdev::halo_pull_kernel<<<dv[k].grid_halo, 256, 0, ctx->stream>>>(sh[k].nhalo, dv[k].h_local, dv[k].h_owner, dv[k].h_remote, dv[k].peer_p, dv[k].p);
            if (sh[k].nown)
                cgdev::spmv_rows_kernel<<<dv[k].grid_rows, cgdev::THREADS, 0, ctx->stream>>>(sh[k].own_first, sh[k].nown, ctx->d_ia, ctx->d_ja, ctx->d_a,
                                                                                             ctx->symmetric, 1.0, dv[k].p, dv[k].q);
            ctx->launches += 2;
        }
        e = sync_all();   // every local product is done: the halo contributions may travel
        if (e != cudaSuccess) return e;
        for (int k = 0; k < nshards; k++) {
            b200asm_ctx *ctx = sh[k].ctx;
            if (!ctx->symmetric || !sh[k].nhalo) continue;
            cudaSetDevice(ctx->device);
            cgdev::halo_push_kernel<<<dv[k].grid_halo, 256, 0, ctx->stream>>>(sh[k].nhalo, dv[k].h_local, dv[k].h_owner, dv[k].h_remote, dv[k].peer_q, dv[k].q);
            ctx->launches++;
        }
        return sync_all();
    };
    // right-hand side, initial guess
    for (int k = 0; k < nshards; k++) {
        b200asm_ctx *ctx = sh[k].ctx;
        CKS(cudaSetDevice(ctx->device));
        Dev &d = dv[k];
        const int64_t o = sh[k].own_first, no = sh[k].nown;
        CKS(cudaMemsetAsync(ctx->d_cg_sc, 0, sizeof(cgdev::Scalars), ctx->stream));
        if (f_host) {
            CKS(cudaMemcpyAsync(d.f + o, f_host + sh[k].row0, (size_t)no * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d += no * (int64_t)sizeof(double);
        } else {
            CKS(cudaMemcpyAsync(d.f + o, ctx->d_rhs + o, (size_t)no * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (from_current && x_host) {
            CKS(cudaMemcpyAsync(d.x + o, x_host + sh[k].row0, (size_t)no * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d += no * (int64_t)sizeof(double);
        } else if (!from_current) {
            CKS(cudaMemsetAsync(d.x, 0, (size_t)std::max<int64_t>(ctx->neq, 1) * sizeof(double), ctx->stream));
        }
        CKS(cudaMemcpyAsync(d.r + o, d.f + o, (size_t)no * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        cgdev::dot_kernel<<<d.grid, cgdev::THREADS, 0, ctx->stream>>>(no, d.f + o, d.f + o, ctx->d_cg_part, ctx->d_cg_sc, 1);
        if (precond == 1) cgdev::extract_diag_kernel<<<d.grid, cgdev::THREADS, 0, ctx->stream>>>(ctx->neq, ctx->d_ia, ctx->d_ja, ctx->d_a, ctx->symmetric, d.diag);
        ctx->launches += 2;
    }
    CKS(gather_scalars());
    double normb2 = 0.0;
    for (int k = 0; k < nshards; k++) normb2 += hs[k].rr;
    double normb = std::sqrt(normb2);
    if (normb == 0.0) normb = 1.0;
    if (from_current) {  // r = b - A x
        CKS(product(1));
        for (int k = 0; k < nshards; k++) {
            b200asm_ctx *ctx = sh[k].ctx;
            CKS(cudaSetDevice(ctx->device));
            const int64_t o = sh[k].own_first, no = sh[k].nown;
            // r -= q with the x/r update kernel: alpha = 1 on (p = 0-vector is not at hand): do it as r = r - 1 * q through z as scratch
            cgdev::update_xr_val_kernel<<<dv[k].grid, cgdev::THREADS, 0, ctx->stream>>>(no, 1.0, dv[k].z + o, dv[k].q + o, dv[k].z + o, dv[k].r + o,
                                                                                        ctx->d_cg_part, ctx->d_cg_sc);
        }
    } else {
        for (int k = 0; k < nshards; k++) {
            b200asm_ctx *ctx = sh[k].ctx;
            CKS(cudaSetDevice(ctx->device));
            const int64_t o = sh[k].own_first, no = sh[k].nown;
            cgdev::dot_kernel<<<dv[k].grid, cgdev::THREADS, 0, ctx->stream>>>(no, dv[k].r + o, dv[k].r + o, ctx->d_cg_part, ctx->d_cg_sc, 1);
        }
    }
    CKS(gather_scalars());
    double rr = 0.0;
    for (int k = 0; k < nshards; k++) rr += hs[k].rr;
    double resid = std::sqrt(rr) / normb;
    int64_t it = 0;
    double rho = 0.0, rho_prev = 0.0;
    if (resid > tol) {
        for (it = 1; it <= max_iter; it++) {
            for (int k = 0; k < nshards; k++) {   // z = M^-1 r, rho = r.z
                b200asm_ctx *ctx = sh[k].ctx;
                cudaSetDevice(ctx->device);
                const int64_t o = sh[k].own_first;
                cgdev::precond_dot_kernel<<<dv[k].grid, cgdev::THREADS, 0, ctx->stream>>>(sh[k].nown, dv[k].r + o, precond == 1 ? dv[k].diag + o : nullptr,
                                                                                          dv[k].z + o, ctx->d_cg_part, ctx->d_cg_sc);
                ctx->launches++;
            }
            CKS(gather_scalars());
            rho_prev = rho;
            rho = 0.0;
            for (int k = 0; k < nshards; k++) rho += hs[k].rho;
            const double beta = it == 1 ? 0.0 : rho / rho_prev;
            for (int k = 0; k < nshards; k++) {
                b200asm_ctx *ctx = sh[k].ctx;
                cudaSetDevice(ctx->device);
                const int64_t o = sh[k].own_first;
                cgdev::update_p_val_kernel<<<dv[k].grid, cgdev::THREADS, 0, ctx->stream>>>(sh[k].nown, it == 1, beta, dv[k].z + o, dv[k].p + o);
                ctx->launches++;
            }
            CKS(product(0));
            for (int k = 0; k < nshards; k++) {
                b200asm_ctx *ctx = sh[k].ctx;
                cudaSetDevice(ctx->device);
                const int64_t o = sh[k].own_first;
                cgdev::dot_kernel<<<dv[k].grid, cgdev::THREADS, 0, ctx->stream>>>(sh[k].nown, dv[k].p + o, dv[k].q + o, ctx->d_cg_part, ctx->d_cg_sc, 0);
                ctx->launches++;
            }
            CKS(gather_scalars());
            double pq = 0.0;
            for (int k = 0; k < nshards; k++) pq += hs[k].pq;
            const double alpha = rho / pq;
            for (int k = 0; k < nshards; k++) {
                b200asm_ctx *ctx = sh[k].ctx;
                cudaSetDevice(ctx->device);
                const int64_t o = sh[k].own_first;
                cgdev::update_xr_val_kernel<<<dv[k].grid, cgdev::THREADS, 0, ctx->stream>>>(sh[k].nown, alpha, dv[k].p + o, dv[k].q + o, dv[k].x + o, dv[k].r + o,
                                                                                            ctx->d_cg_part, ctx->d_cg_sc);
                ctx->launches++;
            }
            CKS(gather_scalars());
            rr = 0.0;
            for (int k = 0; k < nshards; k++) rr += hs[k].rr;
            resid = std::sqrt(rr) / normb;
            if (!(resid == resid)) { err = "cg_sharded: the iteration produced NaN (singular or indefinite system)"; cleanup(); return B200ASM_ECUDA; }
            if (resid <= tol) break;
        }
        if (it > max_iter) it = max_iter;
    }
    if (x_host) {
        for (int k = 0; k < nshards; k++) {
            b200asm_ctx *ctx = sh[k].ctx;
            CKS(cudaSetDevice(ctx->device));
            CKS(cudaMemcpyAsync(x_host + sh[k].row0, dv[k].x + sh[k].own_first, (size_t)sh[k].nown * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h += sh[k].nown * (int64_t)sizeof(double);
        }
        CKS(sync_all());
    }
    cleanup();
    if (iters_out) *iters_out = it;
    if (resid_out) *resid_out = resid;
    return 0;
#undef CKS
}

extern "C" int b200asm_cg_solution_device(b200asm_ctx *ctx, double **x_dev) {
    if (!ctx || !x_dev) return B200ASM_EINVAL;
    *x_dev = ctx->d_cg;
    return ctx->d_cg ? 0 : fail(ctx, B200ASM_ESTATE, "cg_solution_device: no solve yet");
}

extern "C" int b200asm_device_pointers(b200asm_ctx *ctx, double **a_dev, double **rhs_dev) {
    if (!ctx) return B200ASM_EINVAL;
    if (a_dev) *a_dev = ctx->d_a;
    if (rhs_dev) *rhs_dev = ctx->d_rhs;
    return 0;
}

// kernel family of the matrix part of a group, as chosen for the LAST assembly / scatter-map build
extern "C" int b200asm_group_kernel(const b200asm_ctx *ctx, int group, char *name, int len) {
    if (!ctx || group < 0 || group >= (int)ctx->groups.size() || !name || len < 1) return B200ASM_EINVAL;
    const Group &g = ctx->groups[group];
    const char *k = "register_tile";
    if (g.kind == B200ASM_BC) k = "boundary";
    else if (g.plane) k = "plane";
    else if (runs_generic(ctx, g)) k = "generic";
    else if (g.use_gather) k = "gather_rows";
    else if (const MmaEntry *fe = fast_entry(ctx, g)) k = fe->closed ? "closed_form" : (fe->sumfact ? "sumfact" : "dmma");
    snprintf(name, (size_t)len, "%s", k);
    return 0;
}

extern "C" int b200asm_counters(const b200asm_ctx *ctx, int64_t *kernel_launches, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!ctx) return B200ASM_EINVAL;
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (h2d_bytes) *h2d_bytes = ctx->h2d;
    if (d2h_bytes) *d2h_bytes = ctx->d2h;
    return 0;
}
