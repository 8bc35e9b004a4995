// pattern_device.cuh — the CSR pattern of the reference built ON THE DEVICE (SURVEY.md §8 row f/N1).
//
// Replaces, bit-exactly, what TPZSSpStructMatrix::Create / TPZSpStructMatrix::Create do on the host:
//   Mesh/pzcmesh.cpp:1223-1267           element graph (sequence numbers of every element's connects)  [input]
//   External/TPZRenumbering.cpp:30-110   NodeToElGraph + ConvertGraph: block -> neighbour blocks, ascending, self excluded
//   StrMatrix/TPZSSpStructMatrix.cpp:50-193   symmetric rows: own block columns >= row, then the LARGER neighbour blocks
//   StrMatrix/TPZSpStructMatrix.cpp:53-190    full rows: all neighbour blocks and the own block, sorted ascending
// The reference walks the blocks serially with a std::set per block (1.7 s for 36 k equations) and
// needs the whole pattern in host memory (impossible at 1e10 entries).  Here one WARP owns one block (connect):
//   gather   the connects of all elements around the block into shared memory (block -> element lists are built
//            with one counting pass + atomics),
//   sort     them with a warp-level bitonic network, drop duplicates / the block itself (= the std::set),
//   sweep 1  row lengths -> exclusive scan -> IA,
//   sweep 2  the same gather/sort, then every lane writes a strided part of the rows' column lists (coalesced).
// Column indices are kept as int32 on the device (nnz and neq < 2^31 per GPU; larger systems are row-sharded).
#pragma once

#include <cub/cub.cuh>

namespace patdev {

constexpr int CAP = 2048;       // neighbour candidates per block (hex p<=4: 8 elements x 27 connects = 216)
constexpr int UCAP = 1024;      // distinct neighbour blocks per block
constexpr int WARPS = 2;        // warps per CTA (32 KB of static shared memory)

struct Params {
    int symmetric;
    int64_t nel, nblock;
    const int64_t *egi;       // [nel+1]
    const int32_t *eg;        // [egi[nel]]
    const int64_t *bpos;      // [nblock]
    const int32_t *bsize;     // [nblock]
    const int64_t *n2e_idx;   // [nblock+1]
    const int32_t *n2e;       // [egi[nel]]
    int64_t *rowlen;          // sweep 1: [neq+1], rowlen[row+1] = stored columns of the row
    const int64_t *ia;        // sweep 2
    int32_t *ja;              // sweep 2
    int *error;               // != 0: a block exceeded CAP / UCAP
};

__global__ void count_incidence_kernel(int64_t total, const int32_t *__restrict__ eg, int32_t *__restrict__ cnt) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) atomicAdd(cnt + eg[k], 1);
}

__global__ void widen_kernel(int64_t n, const int32_t *__restrict__ in, int64_t *__restrict__ out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = in[k];
}
__global__ void narrow_kernel(int64_t n, const int64_t *__restrict__ in, int32_t *__restrict__ out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = (int32_t)in[k];
}

__global__ void fill_incidence_kernel(int64_t nel, const int64_t *__restrict__ egi, const int32_t *__restrict__ eg,
                                      const int64_t *__restrict__ n2e_idx, int32_t *__restrict__ cursor, int32_t *__restrict__ n2e) {
    for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < nel; el += (int64_t)gridDim.x * blockDim.x)
        for (int64_t k = egi[el]; k < egi[el + 1]; k++) {
            const int32_t b = eg[k];
            n2e[n2e_idx[b] + atomicAdd(cursor + b, 1)] = (int32_t)el;
        }
}

// Sorted, duplicate-free neighbour blocks of block i (self excluded; symmetric: only blocks > i) into uniq[0..nu),
// with off[k] = number of equations of uniq[0..k) (exclusive prefix).  Returns nu, or -1 on overflow.  One warp.
__device__ __forceinline__ int warp_neighbours(const Params &p, int64_t i, int32_t *buf, int32_t *uniq, int32_t *off, int lane,
                                               int &total_eq) {
    int n = 0;
    for (int64_t e = p.n2e_idx[i]; e < p.n2e_idx[i + 1]; e++) {
        const int64_t el = p.n2e[e];
        const int64_t k0 = p.egi[el];
        const int nc = (int)(p.egi[el + 1] - k0);
        if (n + nc > CAP) return -1;
        for (int k = lane; k < nc; k += 32) buf[n + k] = p.eg[k0 + k];
        n += nc;
    }
    int m = 32;
    while (m < n) m <<= 1;
    for (int k = n + lane; k < m; k += 32) buf[k] = 0x7fffffff;
    __syncwarp();
    // bitonic sort of buf[0..m)
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (m >> 1); t += 32) {
                const int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const int32_t a = buf[lo], b = buf[hi];
                if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
            }
            __syncwarp();
        }
    // compaction: first occurrences, not the block itself, (symmetric) larger blocks only
    int nu = 0, eq = 0;
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        bool keep = false;
        int32_t v = 0;
        if (k < n) {
            v = buf[k];
            keep = (k == 0 || buf[k - 1] != v) && v != (int32_t)i && (!p.symmetric || v > (int32_t)i) && p.bsize[v] > 0;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int sz = keep ? p.bsize[v] : 0;
        // exclusive prefix of sz over the warp
        int incl = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int slot = nu + __popc(mask & ((1u << lane) - 1));
        if (nu + __popc(mask) > UCAP) return -1;
        if (keep) {
            uniq[slot] = v;
            off[slot] = eq + incl - sz;
        }
        nu += __popc(mask);
        eq += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    total_eq = eq;
    return nu;
}

// SWEEP: 1 = row lengths, 2 = column indices
template <int SWEEP>
__global__ void __launch_bounds__(WARPS * 32) pattern_kernel(const Params p) {
    __shared__ int32_t s_buf[WARPS][CAP];
    __shared__ int32_t s_uniq[WARPS][UCAP];
    __shared__ int32_t s_off[WARPS][UCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *buf = s_buf[warp], *uniq = s_uniq[warp], *off = s_off[warp];
    for (int64_t i = (int64_t)blockIdx.x * WARPS + warp; i < p.nblock; i += (int64_t)gridDim.x * WARPS) {
        const int sz = p.bsize[i];
        if (sz == 0) continue;  // the connect carries no equation
        int neq_nb = 0;
        const int nu = warp_neighbours(p, i, buf, uniq, off, lane, neq_nb);
        if (nu < 0) {
            if (lane == 0) atomicExch(p.error, 1);
            continue;
        }
        const int64_t row0 = p.bpos[i];
        if (SWEEP == 1) {
            const int len0 = sz + neq_nb;
            for (int r = lane; r < sz; r += 32) p.rowlen[row0 + r + 1] = p.symmetric ? len0 - r : len0;
        } else if (p.symmetric) {
            // first-row list: own equations, then the larger neighbour blocks ascending; row r keeps list[r:]
            const int T = sz + neq_nb;
            for (int t = lane; t < T; t += 32) {
                int32_t col;
                if (t < sz) col = (int32_t)row0 + t;
                else {
                    const int tt = t - sz;
                    int lo = 0, hi = nu - 1;  // last k with off[k] <= tt
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (off[mid] <= tt) lo = mid; else hi = mid - 1;
                    }
                    col = (int32_t)p.bpos[uniq[lo]] + (tt - off[lo]);
                }
                const int rmax = min(sz - 1, t);
                for (int r = 0; r <= rmax; r++) p.ja[p.ia[row0 + r] + t - r] = col;
            }
        } else {
            // all neighbour blocks ascending with the own block at its sorted place; every row gets the same list
            int nbefore = 0;  // equations of the neighbour blocks smaller than i
            {
                int lo = 0, hi = nu;  // first k with uniq[k] > i
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (uniq[mid] > (int32_t)i) hi = mid; else lo = mid + 1;
                }
                nbefore = lo < nu ? off[lo] : neq_nb;
            }
            const int T = sz + neq_nb;
            for (int t = lane; t < T; t += 32) {
                int32_t col;
                if (t >= nbefore && t < nbefore + sz) col = (int32_t)row0 + (t - nbefore);
                else {
                    const int tt = t < nbefore ? t : t - sz;
                    int lo = 0, hi = nu - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (off[mid] <= tt) lo = mid; else hi = mid - 1;
                    }
                    col = (int32_t)p.bpos[uniq[lo]] + (tt - off[lo]);
                }
                for (int r = 0; r < sz; r++) p.ja[p.ia[row0 + r] + t] = col;
            }
        }
        __syncwarp();
    }
}

}  // namespace patdev
