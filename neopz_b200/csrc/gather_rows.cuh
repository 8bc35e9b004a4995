// gather_rows.cuh — owner-computes ("gather") assembly for the groups whose element matrices have a closed form
// (straight-sided tetrahedra, parallelepiped hexahedra of order <= 2: affine_simplex.cuh / affine_hex.cuh).  OPT-IN (option
// "gather" = 1): the deterministic alternative to the scatter kernels, not the fast path.
//
// Idea.  The scatter formulation of TPZSYsmpMatrix::AddKel (Matrix/pzsysmp.cpp:370-411) adds every element entry into the CSR
// values with a reduction: a memset of A, a read-modify-write of every line of A in L2 and one L2 reduction per element entry
// (~1.3 LSU cycles per lane: what bounds the closed-form kernels).  Here every CSR row is produced by
// exactly one sub-warp: the rows of a node (connect) are accumulated in shared memory from the elements that contain the node -
// each contribution recomputed from the element's Jacobian factors in closed form - and written to A ONCE, coalesced, with
// plain stores.  No memset, no atomics, no read of A; the summation order is fixed (elements in ascending order), so the rows
// that no other group adds to are bit-reproducible like the reference's ordered assembly (StrMatrix/pzstrmatrixor.cpp:714-717).
//
// Measured (one B200, profiles/r02_gather_vs_scatter.md): 2-3x SLOWER than the scatter kernels (hexahedra p2 Poisson 128^3 uniform:
// 146 against 407 M elements/s; hexahedra p2 elasticity 81^3: 25.9 against 45.0; tetrahedra p2 elasticity 113^3: 126 against 513).
// A (node, element) pair costs six / nine table loads, the staged factors and a shared-memory read-modify-write for at most N
// useful lanes (half of them idle under symmetric storage), and the pairs of a node are a serial chain: the gather trades the
// reductions for about as many shared-memory wavefronts and loses the parallelism of one warp per element.
//
// Data (built on the device when the scatter maps would be built):
//   rec[ng]             NodeRec {first equation ("key"), first pair, number of elements} of every node block, ascending keys
//   glist[nel * N]      (element << 5) | local node: the pairs of a node are consecutive, elements in ascending order
//   relpos[pair][a][jn] uint16: position, within row key + a, of the first stored column of local node jn of that element
//                       (0xFFFF: block not stored - symmetric storage keeps the columns >= row only)
//   rowflag[neq]        0 row without a gather group (zeroed, other groups add), 1 exclusive (stored, never zeroed),
//                       2 shared with another group (zeroed, the others add first, this kernel adds on top)
//   fac[nel][FS]        per assembly: Poisson s|detJ| (Jinv Jinv^T) (6), Elasticity3D Jinv (9) and |detJ|
// Requirements checked at setup (otherwise the group keeps its scatter kernel): the NS equations of a node are consecutive
// and none is filtered; rows shorter than 65535 entries; at most 2^27 elements and MAXDEG elements per node; one GPU, atomic mode.
#pragma once
#include <cstdint>

namespace gat {

constexpr unsigned short NOPOS = 0xFFFFu;
constexpr int WPC = 8;  // warps per CTA

// ---- setup ----------------------------------------------------------------------------------------------------------------
__global__ void count_kernel(int64_t nel, int N, int NS, const int32_t *__restrict__ dest, int32_t *__restrict__ cnt, int *__restrict__ bad) {
    const int64_t total = nel * N;
    const int M = N * NS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = idx / N;
        const int jn = (int)(idx - e * N);
        const int32_t k0 = dest[e * M + jn * NS];
        if (k0 < 0) { *bad = 1; continue; }
        for (int b = 1; b < NS; b++)
            if (dest[e * M + jn * NS + b] != k0 + b) *bad = 1;
        atomicAdd(cnt + k0, 1);
    }
}

__global__ void flag_nonempty_kernel(int64_t neq, const int32_t *__restrict__ cnt, int32_t *__restrict__ flag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < neq; i += (int64_t)gridDim.x * blockDim.x) flag[i] = cnt[i] > 0;
}

__global__ void compact_keys_kernel(int64_t neq, const int32_t *__restrict__ cnt, const int32_t *__restrict__ slot, int32_t *__restrict__ grow) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < neq; i += (int64_t)gridDim.x * blockDim.x)
        if (cnt[i] > 0) grow[slot[i]] = (int32_t)i;
}

__global__ void fill_kernel(int64_t nel, int N, int NS, const int32_t *__restrict__ dest, const int32_t *__restrict__ gptr,
                            int32_t *__restrict__ cursor, uint32_t *__restrict__ glist) {
    const int64_t total = nel * N;
    const int M = N * NS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = idx / N;
        const int jn = (int)(idx - e * N);
        const int32_t k0 = dest[e * M + jn * NS];
        if (k0 < 0) continue;
        const int32_t p = gptr[k0] + atomicAdd(cursor + k0, 1);
        glist[p] = ((uint32_t)e << 5) | (uint32_t)jn;
    }
}

// elements of a node in ascending order: fixed summation order, whatever order the atomics of fill_kernel produced
__global__ void sort_groups_kernel(int64_t ng, const int32_t *__restrict__ grow, const int32_t *__restrict__ gptr, uint32_t *__restrict__ glist) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
        const int32_t key = grow[g];
        const int32_t p0 = gptr[key], p1 = gptr[key + 1];
        for (int32_t i = p0 + 1; i < p1; i++) {
            const uint32_t v = glist[i];
            int32_t j = i - 1;
            while (j >= p0 && glist[j] > v) { glist[j + 1] = glist[j]; j--; }
            glist[j + 1] = v;
        }
    }
}

// rows another group contributes to: every destination of its elements (symmetric storage puts entry (i, j) in row min: any of them)
__global__ void mark_rows_kernel(int64_t n, const int32_t *__restrict__ dest, unsigned char *__restrict__ touched) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t d = dest[i];
        if (d >= 0) touched[d] = 1;
    }
}

// rowflag of the rows of this group: 1 exclusive, 2 shared (touched by another group, gather or not)
__global__ void flag_rows_kernel(int64_t ng, int NS, const int32_t *__restrict__ grow, const unsigned char *__restrict__ touched,
                                 unsigned char *__restrict__ rowflag) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
        const int32_t key = grow[g];
        for (int a = 0; a < NS; a++) rowflag[key + a] = (touched[key + a] || rowflag[key + a] != 0) ? 2 : 1;
    }
}

// positions of the column blocks: one warp per node block, lanes <-> local nodes of the element
__global__ void __launch_bounds__(WPC * 32) relpos_kernel(int64_t ng, int N, int NS, int symmetric, const int32_t *__restrict__ grow,
                                                          const int32_t *__restrict__ gptr, const uint32_t *__restrict__ glist,
                                                          const int32_t *__restrict__ dest, const int64_t *__restrict__ ia,
                                                          const int32_t *__restrict__ ja, unsigned short *__restrict__ relpos,
                                                          int *__restrict__ missing, int *__restrict__ bad, int *__restrict__ maxlen) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * WPC;
    const int M = N * NS;
    for (int64_t g = (int64_t)blockIdx.x * WPC + (threadIdx.x >> 5); g < ng; g += nwarps) {
        const int32_t key = grow[g];
        const int32_t p0 = gptr[key], p1 = gptr[key + 1];
        if (lane < NS) {
            const int64_t len = ia[key + lane + 1] - ia[key + lane];
            if (len >= NOPOS) *bad = 1;
            atomicMax(maxlen, (int)min(len, (int64_t)0x7fffffff));
        }
        for (int32_t p = p0; p < p1; p++) {
            const uint32_t code = glist[p];
            const int64_t e = code >> 5;
            const int li = code & 31;
            for (int jn = lane; jn < N; jn += 32) {
                const int32_t c0 = dest[e * M + jn * NS];
                for (int a = 0; a < NS; a++) {
                    const int64_t row = key + a;
                    unsigned short rel = NOPOS;
                    const bool diag = jn == li;
                    if (!symmetric || c0 >= key) {
                        const int32_t col = c0 + ((symmetric && diag) ? a : 0);
                        int64_t lo = ia[row], hi = ia[row + 1] - 1;
                        const int64_t base = lo;
                        bool found = false;
                        while (lo <= hi) {
                            const int64_t mid = (lo + hi) >> 1;
                            const int32_t v = ja[mid];
                            if (v == col) { rel = (unsigned short)(mid - base); found = true; break; }
                            if (v < col) lo = mid + 1; else hi = mid - 1;
                        }
                        if (!found) atomicAdd(missing, 1);
                    }
                    relpos[((size_t)p * NS + a) * N + jn] = rel;
                }
            }
        }
    }
}

// ---- per assembly ------------------------------------------------------------------------------------------------------
// Jacobian factors of every element (the single Jacobian of an affine element: affine_simplex.cuh / affine_hex.cuh)
template <int NN, int NS>
__global__ void factors_kernel(int64_t nel, const int32_t *__restrict__ elnodes, const double *__restrict__ xyz, double scale,
                               double *__restrict__ fac) {
    constexpr int FS = NS == 1 ? 6 : 10;
    for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < nel; el += (int64_t)gridDim.x * blockDim.x) {
        double j[3][3];
        if (NN == 4) {
            const int4 n0 = *reinterpret_cast<const int4 *>(elnodes + el * 4);
            const int32_t id[4] = {n0.x, n0.y, n0.z, n0.w};
            double x0[3];
#pragma unroll
            for (int r = 0; r < 3; r++) x0[r] = xyz[(int64_t)id[0] * 3 + r];
#pragma unroll
            for (int d = 0; d < 3; d++)
#pragma unroll
                for (int r = 0; r < 3; r++) j[r][d] = xyz[(int64_t)id[d + 1] * 3 + r] - x0[r];
        } else {
            const int4 n0 = *reinterpret_cast<const int4 *>(elnodes + el * 8), n1 = *reinterpret_cast<const int4 *>(elnodes + el * 8 + 4);
            const int32_t id[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
            for (int r = 0; r < 3; r++) j[r][0] = j[r][1] = j[r][2] = 0.0;
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const double x = xyz[(int64_t)id[a] * 3 + r];
#pragma unroll
                    for (int d = 0; d < 3; d++) j[r][d] += hex_sign(a, d) * x;
                }
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int d = 0; d < 3; d++) j[r][d] *= 0.125;
        }
        double det = 0.0;
        det -= j[0][2] * j[1][1] * j[2][0];
        det += j[0][1] * j[1][2] * j[2][0];
        det += j[0][2] * j[1][0] * j[2][1];
        det -= j[0][0] * j[1][2] * j[2][1];
        det -= j[0][1] * j[1][0] * j[2][2];
        det += j[0][0] * j[1][1] * j[2][2];
        if (fabs(det) < 1.e-12) det = 1.e-12;
        const double id_ = 1.0 / det, adet = fabs(det);
        double ji[3][3];
        ji[0][0] = (-j[1][2] * j[2][1] + j[1][1] * j[2][2]) * id_;
        ji[0][1] = (j[0][2] * j[2][1] - j[0][1] * j[2][2]) * id_;
        ji[0][2] = (-j[0][2] * j[1][1] + j[0][1] * j[1][2]) * id_;
        ji[1][0] = (j[1][2] * j[2][0] - j[1][0] * j[2][2]) * id_;
        ji[1][1] = (-j[0][2] * j[2][0] + j[0][0] * j[2][2]) * id_;
        ji[1][2] = (j[0][2] * j[1][0] - j[0][0] * j[1][2]) * id_;
        ji[2][0] = (-j[1][1] * j[2][0] + j[1][0] * j[2][1]) * id_;
        ji[2][1] = (j[0][1] * j[2][0] - j[0][0] * j[2][1]) * id_;
        ji[2][2] = (-j[0][1] * j[1][0] + j[0][0] * j[1][1]) * id_;
        double *f = fac + (size_t)el * FS;
        if (NS == 1) {
            constexpr int E[6] = {0, 1, 2, 0, 0, 1}, F[6] = {0, 1, 2, 1, 2, 2};
            const double sc = scale * adet;
#pragma unroll
            for (int q = 0; q < 6; q++) f[q] = (ji[E[q]][0] * ji[F[q]][0] + ji[E[q]][1] * ji[F[q]][1] + ji[E[q]][2] * ji[F[q]][2]) * sc;
        } else {
#pragma unroll
            for (int e = 0; e < 3; e++)
#pragma unroll
                for (int v = 0; v < 3; v++) f[e * 3 + v] = ji[e][v];
            f[9] = adet;
        }
    }
}

// rows that are not stored by a gather kernel start from zero (one warp per row)
__global__ void zero_rows_kernel(int64_t neq, const int64_t *__restrict__ ia, const unsigned char *__restrict__ rowflag, double *__restrict__ a) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < neq; row += nwarps) {
        if (rowflag[row] == 1) continue;
        for (int64_t k = ia[row] + lane; k < ia[row + 1]; k += 32) a[k] = 0.0;
    }
}

// one record per node block (built with the scatter maps): the kernel reads nothing else to find its work
struct NodeRec {
    int32_t key, p0, deg, rsv;  // first equation, first pair in glist / relpos, number of elements of the node
};
constexpr int MAXDEG = 64;  // elements per node the kernel stages (more: the group keeps its scatter kernel)
constexpr int PC = 4;       // pairs (node, element) staged at a time

__global__ void node_records_kernel(int64_t ng, const int32_t *__restrict__ grow, const int32_t *__restrict__ gptr, NodeRec *__restrict__ rec,
                                    int *__restrict__ bad) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < ng; g += (int64_t)gridDim.x * blockDim.x) {
        const int32_t key = grow[g];
        NodeRec r;
        r.key = key; r.p0 = gptr[key]; r.deg = gptr[key + 1] - r.p0; r.rsv = 0;
        if (r.deg > MAXDEG) *bad = 1;
        rec[g] = r;
    }
}

template <int N, int NS>
struct Cfg {
    static constexpr int SUBS = 32 / N > 0 ? 32 / N : 1;  // node blocks a warp handles side by side (lane = sub * N + jn)
    static constexpr int FS = NS == 1 ? 6 : 10;           // doubles per element in fac
    static constexpr int NT = NS == 1 ? 6 : 9;            // tables
    static constexpr int GT_LEN = (NT * N * N + 1) & ~1;  // doubles (even: what follows stays 16-byte aligned)
    static constexpr int REL_D = (SUBS * PC * NS * N * 2 + 7) / 8;  // doubles of the staged positions [SUBS][PC][NS][N] (uint16)
    // doubles per warp: row buffers [SUBS][NS][rl], staged factors [SUBS][PC][FS], pair codes [2][SUBS][MAXDEG] (uint32), positions
    __host__ __device__ static constexpr size_t warp_doubles(int rl) {
        return ((size_t)SUBS * NS * rl + (size_t)SUBS * PC * FS + (size_t)SUBS * MAXDEG + REL_D + 1) & ~(size_t)1;  // (even: 16-byte aligned warps)
    }
    static size_t smem_bytes(int rl, int wpc) { return sizeof(double) * (GT_LEN + warp_doubles(rl) * wpc); }
};

struct Params {
    int64_t g0, g1;     // node blocks [g0, g1) of the records
    int symmetric;
    int rl;             // doubles per row buffer (multiple of 4)
    const NodeRec *rec;
    const uint32_t *glist;
    const unsigned short *relpos;
    const unsigned char *rowflag;
    const double *fac;
    const double *aux;  // Ghat[9][NPP] of the group (pairs in <= jn), see AffCfg / AffHexCfg
    int npp;
    const int64_t *ia;
    double *a;
    double c1, c2, c3;  // TPZElasticity3D constants
};

__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ int64_t shfl64(int64_t v, int src) {
    const int lo = __shfl_sync(0xffffffffu, (int)(v & 0xffffffff), src), hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
    return ((int64_t)hi << 32) | (uint32_t)lo;
}

// One warp handles SUBS node blocks side by side (lane = sub * N + jn: the sub-warp of a node, one lane per local node of the
// element at hand); a sub-warp walks the elements of its node in ascending order, PC at a time: the Jacobian factors of the PC
// elements are staged in shared memory (one coalesced load per element), the positions of the column blocks come as 16-bit
// loads, both issued for the whole batch before the first use.  The records, pair codes, row starts and row flags of the warp's
// NEXT task are fetched while the current one is computed (records two tasks ahead), so a task waits for one round trip to
// memory per batch of pairs, not for a chain of four.  Sub-warps own disjoint row buffers: no conflicts, no atomics; the
// contributions of a node are added in ascending element order, so two assemblies agree bit for bit on the rows that no other
// group adds to.
template <int N, int NS>
__global__ void __launch_bounds__(768) gather_rows_kernel(const Params p) {
    using C = Cfg<N, NS>;
    constexpr int SUBS = C::SUBS, FS = C::FS, NN2 = N * N;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    double *Gt = smem;                                                            // [NT][N * N]
    double *buf = smem + C::GT_LEN + (size_t)warp * C::warp_doubles(p.rl);         // [SUBS][NS][rl]
    double *facs = buf + (size_t)SUBS * NS * p.rl;                                // [SUBS][PC][FS]
    uint32_t *codes = reinterpret_cast<uint32_t *>(facs + SUBS * PC * FS);        // [2][SUBS][MAXDEG]
    unsigned short *rels = reinterpret_cast<unsigned short *>(codes + 2 * SUBS * MAXDEG);  // [SUBS][PC][NS][N]
    // full (in, jn) table from the upper-pair table: Ghat[e][f](in,jn) = Ghat[f][e](jn,in)
    for (int idx = threadIdx.x; idx < NN2; idx += blockDim.x) {
        const int i = idx / N, j = idx - i * N;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        const int pr = lo * N - lo * (lo - 1) / 2 + (hi - lo);
        if (NS == 1) {
            Gt[0 * NN2 + idx] = __ldg(p.aux + 0 * p.npp + pr);
            Gt[1 * NN2 + idx] = __ldg(p.aux + 4 * p.npp + pr);
            Gt[2 * NN2 + idx] = __ldg(p.aux + 8 * p.npp + pr);
            Gt[3 * NN2 + idx] = __ldg(p.aux + 1 * p.npp + pr) + __ldg(p.aux + 3 * p.npp + pr);
            Gt[4 * NN2 + idx] = __ldg(p.aux + 2 * p.npp + pr) + __ldg(p.aux + 6 * p.npp + pr);
            Gt[5 * NN2 + idx] = __ldg(p.aux + 5 * p.npp + pr) + __ldg(p.aux + 7 * p.npp + pr);
        } else {
#pragma unroll
            for (int e = 0; e < 3; e++)
#pragma unroll
                for (int f = 0; f < 3; f++) Gt[(e * 3 + f) * NN2 + idx] = __ldg(p.aux + (i <= j ? e * 3 + f : f * 3 + e) * p.npp + pr);
        }
    }
    for (int k = lane; k < SUBS * NS * p.rl; k += 32) buf[k] = 0.0;
    __syncthreads();

    const int sub = lane / N, jn = lane - sub * N;
    const bool act = sub < SUBS;
    const int sb = act ? sub : SUBS - 1;
    const int64_t ntasks = (p.g1 - p.g0 + SUBS - 1) / SUBS;
    const int64_t tstep = (int64_t)gridDim.x * wpc;
    int64_t t = (int64_t)blockIdx.x * wpc + warp;
    auto record_of = [&](int64_t tt) -> int4 {
        int4 r = make_int4(0, 0, 0, 0);
        if (act && tt < ntasks) {
            const int64_t g = p.g0 + tt * SUBS + sub;
            if (g < p.g1) r = __ldg(reinterpret_cast<const int4 *>(p.rec) + g);
        }
        return r;
    };
    auto fetch_codes = [&](const int4 &r, int cb) {
        uint32_t *dst = codes + (cb * SUBS + sb) * MAXDEG;
        for (int k = jn; k < r.z; k += N) cp_async4(dst + k, p.glist + r.y + k);
    };
    int4 cur = record_of(t), nxt = record_of(t + tstep);
    int cb = 0;
    fetch_codes(cur, cb);
    int64_t ia_cur = 0;
    int fl_cur = 0;
    if (cur.z > 0 && jn <= NS) ia_cur = __ldg(p.ia + cur.x + jn);
    if (cur.z > 0 && jn < NS) fl_cur = __ldg(p.rowflag + cur.x + jn);

    for (; t < ntasks; t += tstep) {
        cp_async_wait_all();
        __syncwarp();
        // the next task of this warp: codes, row starts, row flags; the record of the one after
        const int4 nn = record_of(t + 2 * tstep);
        fetch_codes(nxt, cb ^ 1);
        int64_t ia_nxt = 0;
        int fl_nxt = 0;
        if (nxt.z > 0 && jn <= NS) ia_nxt = __ldg(p.ia + nxt.x + jn);
        if (nxt.z > 0 && jn < NS) fl_nxt = __ldg(p.rowflag + nxt.x + jn);

        const uint32_t *cc = codes + (cb * SUBS + sb) * MAXDEG;
        const int deg = cur.z;
        const int maxdeg = __reduce_max_sync(FULL, deg);
        double *mybuf = buf + (size_t)sb * NS * p.rl;
        double *myfac = facs + sb * PC * FS;
        unsigned short *myrel = rels + sb * PC * NS * N + jn;
        for (int c0 = 0; c0 < maxdeg; c0 += PC) {
            // stage: factors of the PC elements (16-byte pieces, lanes of the sub-warp side by side) and the positions of
            // this lane's column blocks; every load of the batch is in flight before the first store
            {
                unsigned short rel[PC][NS];
                constexpr int JC = (FS / 2 + N - 1) / N;
                double2 fv[PC][JC];
#pragma unroll
                for (int s = 0; s < PC; s++) {
                    if (c0 + s < deg) {
                        const double2 *f = reinterpret_cast<const double2 *>(p.fac + (size_t)(cc[c0 + s] >> 5) * FS);
#pragma unroll
                        for (int j = 0; j < JC; j++)
                            if (jn + j * N < FS / 2) fv[s][j] = __ldg(f + jn + j * N);
                        const unsigned short *rp = p.relpos + ((size_t)(cur.y + c0 + s) * NS) * N + jn;
#pragma unroll
                        for (int a = 0; a < NS; a++) rel[s][a] = __ldg(rp + a * N);
                    }
                }
#pragma unroll
                for (int s = 0; s < PC; s++) {
                    if (c0 + s < deg) {
#pragma unroll
                        for (int j = 0; j < JC; j++)
                            if (jn + j * N < FS / 2) reinterpret_cast<double2 *>(myfac + s * FS)[jn + j * N] = fv[s][j];
#pragma unroll
                        for (int a = 0; a < NS; a++) myrel[(s * NS + a) * N] = rel[s][a];
                    }
                }
            }
            __syncwarp();
#pragma unroll 1
            for (int s = 0; s < PC; s++) {
                // (a column block can be another local node of the next element: the additions of one element are complete - on
                //  every lane, whichever branch it took - before the next element's begin: compute-sanitizer racecheck)
                if (c0 + s < deg) {
                    const int li = cc[c0 + s] & 31;
                    const int tix = li * N + jn;
                    const double2 *f2 = reinterpret_cast<const double2 *>(myfac + s * FS);
                    if (NS == 1) {
                        const unsigned short r0 = myrel[s * N];
                        if (r0 != NOPOS) {
                            const double2 fa = f2[0], fb = f2[1], fc = f2[2];
                            double v = fa.x * Gt[0 * NN2 + tix];
                            v += fa.y * Gt[1 * NN2 + tix];
                            v += fb.x * Gt[2 * NN2 + tix];
                            v += fb.y * Gt[3 * NN2 + tix];
                            v += fc.x * Gt[4 * NN2 + tix];
                            v += fc.y * Gt[5 * NN2 + tix];
                            mybuf[r0] += v;
                        }
                    } else {
                        unsigned short rel[NS];
                        bool any = false;
#pragma unroll
                        for (int a = 0; a < NS; a++) {
                            rel[a] = myrel[(s * NS + a) * N];
                            any = any || rel[a] != NOPOS;
                        }
                        if (any) {
                            double ji[3][3];
                            const double2 f0 = f2[0], f1 = f2[1], f2_ = f2[2], f3 = f2[3], f4 = f2[4];
                            ji[0][0] = f0.x; ji[0][1] = f0.y; ji[0][2] = f1.x; ji[1][0] = f1.y; ji[1][1] = f2_.x;
                            ji[1][2] = f2_.y; ji[2][0] = f3.x; ji[2][1] = f3.y; ji[2][2] = f4.x;
                            const double adet = f4.y;
                            const double C1 = p.c1 * adet, C2 = p.c2 * adet, C3 = p.c3 * adet;
                            // S[v][u] = sum_{e,f} jacinv(e,v) jacinv(f,u) Ghat[e][f](li,jn), one table row e at a time
                            double S[3][3];
#pragma unroll
                            for (int e2 = 0; e2 < 3; e2++) {
                                const double g0 = Gt[(e2 * 3 + 0) * NN2 + tix], g1 = Gt[(e2 * 3 + 1) * NN2 + tix], g2 = Gt[(e2 * 3 + 2) * NN2 + tix];
                                double T[3];
#pragma unroll
                                for (int u = 0; u < 3; u++) T[u] = g0 * ji[0][u] + g1 * ji[1][u] + g2 * ji[2][u];
#pragma unroll
                                for (int v = 0; v < 3; v++)
#pragma unroll
                                    for (int u = 0; u < 3; u++) S[v][u] = e2 == 0 ? ji[0][v] * T[u] : S[v][u] + ji[e2][v] * T[u];
                            }
                            // the nine formulas of Material/Elasticity/TPZElasticity3D.cpp:318-326 (row node li, column node jn)
#pragma unroll
                            for (int a = 0; a < NS; a++) {
                                if (rel[a] == NOPOS) continue;
                                const int b0 = (p.symmetric && jn == li) ? a : 0;
                                double *row = mybuf + a * p.rl + rel[a] - b0;
#pragma unroll
                                for (int b = 0; b < NS; b++) {
                                    const double x = a == b ? (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3
                                                            : S[b][a] * C1 - S[a][b] * C2;
                                    if (b >= b0) row[b] += x;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
        // the rows of the node blocks leave once, coalesced (all lanes on every row); the buffers are zero again afterwards
#pragma unroll 1
        for (int s = 0; s < SUBS; s++) {
            if (__shfl_sync(FULL, deg, s * N) == 0) continue;
#pragma unroll
            for (int a = 0; a < NS; a++) {
                const int64_t r0 = shfl64(ia_cur, s * N + a), r1 = shfl64(ia_cur, s * N + a + 1);
                const int fl = __shfl_sync(FULL, fl_cur, s * N + a);
                const int len = (int)(r1 - r0);
                double *b = buf + (size_t)(s * NS + a) * p.rl;
                for (int k = lane; k < len; k += 32) {
                    const double v = b[k];
                    b[k] = 0.0;
                    if (fl == 2) p.a[r0 + k] += v; else __stcs(p.a + r0 + k, v);
                }
            }
        }
        __syncwarp();
        cur = nxt; nxt = nn; ia_cur = ia_nxt; fl_cur = fl_nxt; cb ^= 1;
    }
    cp_async_wait_all();
}

}  // namespace gat
