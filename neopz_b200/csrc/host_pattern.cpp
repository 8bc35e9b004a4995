// Host-side CSR pattern builder (multi-threaded), bit-exact with the reference's
// TPZSSpStructMatrix::Create / TPZSpStructMatrix::Create:
//   element graph  -> Mesh/pzcmesh.cpp:1223-1267 (sequence numbers of each element's connects)
//   block adjacency-> External/TPZRenumbering.cpp:76-110 (ascending, self excluded)
//   rows           -> StrMatrix/TPZSSpStructMatrix.cpp:134-190 (diag block cols >= row, then larger
//                     neighbour blocks ascending) / TPZSpStructMatrix.cpp:142-178 (all blocks, sorted)
// The reference walks the blocks serially with a std::set per block; here every thread owns a
// contiguous range of blocks, rows are sized in a first sweep and filled in a second one, so no
// adjacency list is ever stored (memory = the CSR itself).
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/b200asm.h"

namespace {

struct PatternInput {
    int symmetric;
    int64_t nel;
    const int64_t *egi, *eg;
    int64_t nblock;
    const int64_t *bpos, *bsize;
    const int64_t *n2e_idx, *n2e;
};

// sorted unique neighbour blocks of block i (self excluded) into `nb`
inline void neighbours(const PatternInput &in, int64_t i, std::vector<int64_t> &nb) {
    nb.clear();
    for (int64_t e = in.n2e_idx[i]; e < in.n2e_idx[i + 1]; e++) {
        const int64_t el = in.n2e[e];
        for (int64_t k = in.egi[el]; k < in.egi[el + 1]; k++) nb.push_back(in.eg[k]);
    }
    std::sort(nb.begin(), nb.end());
    nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
    auto self = std::lower_bound(nb.begin(), nb.end(), i);
    if (self != nb.end() && *self == i) nb.erase(self);
}

// number of stored columns in the FIRST row of block i (later rows of the block: symmetric storage
// loses one diagonal-block column per row, full storage keeps the same length)
inline int64_t first_row_len(const PatternInput &in, int64_t i, const std::vector<int64_t> &nb) {
    int64_t len = in.bsize[i];
    for (int64_t col : nb) {
        if (in.symmetric && col < i) continue;
        len += in.bsize[col];
    }
    return len;
}

}  // namespace

extern "C" int64_t b200asm_build_pattern(int symmetric, int64_t nel, const int64_t *elgraphindex, const int64_t *elgraph,
                                         int64_t nblock, const int64_t *blockpos, const int64_t *blocksize, int64_t *ia,
                                         int64_t *ja, int nthreads) {
    if (nel < 0 || nblock < 0 || !elgraphindex || !blockpos || !blocksize || !ia) return B200ASM_EINVAL;
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(1, nblock / 1024));

    // block -> elements (counting sort, element order preserved like NodeToElGraph)
    std::vector<int64_t> n2e_idx(nblock + 1, 0);
    const int64_t last = elgraphindex[nel];
    for (int64_t k = 0; k < last; k++) {
        if (elgraph[k] < 0 || elgraph[k] >= nblock) return B200ASM_EINVAL;
        n2e_idx[elgraph[k] + 1]++;
    }
    for (int64_t b = 0; b < nblock; b++) n2e_idx[b + 1] += n2e_idx[b];
    std::vector<int64_t> n2e(last > 0 ? last : 1), cursor(n2e_idx.begin(), n2e_idx.end() - 1);
    for (int64_t el = 0; el < nel; el++)
        for (int64_t k = elgraphindex[el]; k < elgraphindex[el + 1]; k++) n2e[cursor[elgraph[k]]++] = el;

    PatternInput in{symmetric, nel, elgraphindex, elgraph, nblock, blockpos, blocksize, n2e_idx.data(), n2e.data()};
    auto range = [&](int t, int64_t &b0, int64_t &b1) {
        b0 = nblock * t / nthreads;
        b1 = nblock * (t + 1) / nthreads;
    };

    // sweep 1: row lengths -> ia (as counts at ia[row+1])
    const int64_t neq = nblock ? blockpos[nblock - 1] + blocksize[nblock - 1] : 0;
    ia[0] = 0;
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++)
            pool.emplace_back([&, t] {
                int64_t b0, b1;
                range(t, b0, b1);
                std::vector<int64_t> nb;
                for (int64_t i = b0; i < b1; i++) {
                    const int64_t sz = blocksize[i];
                    if (sz == 0) continue;  // NumActive == 0: the connect carries no equation
                    neighbours(in, i, nb);
                    const int64_t len0 = first_row_len(in, i, nb);
                    for (int64_t r = 0; r < sz; r++) ia[blockpos[i] + r + 1] = symmetric ? len0 - r : len0;
                }
            });
        for (auto &th : pool) th.join();
    }
    for (int64_t r = 0; r < neq; r++) ia[r + 1] += ia[r];
    const int64_t nnz = ia[neq];
    if (!ja) return nnz;

    // sweep 2: fill the column indices
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; t++)
            pool.emplace_back([&, t] {
                int64_t b0, b1;
                range(t, b0, b1);
                std::vector<int64_t> nb, cols;
                for (int64_t i = b0; i < b1; i++) {
                    const int64_t sz = blocksize[i];
                    if (sz == 0) continue;
                    neighbours(in, i, nb);
                    if (symmetric) {
                        // columns of the first row: own block, then larger neighbour blocks ascending
                        cols.clear();
                        for (int64_t j = 0; j < sz; j++) cols.push_back(blockpos[i] + j);
                        for (int64_t col : nb) {
                            if (col < i) continue;
                            for (int64_t j = 0; j < blocksize[col]; j++) cols.push_back(blockpos[col] + j);
                        }
                        for (int64_t r = 0; r < sz; r++) std::copy(cols.begin() + r, cols.end(), ja + ia[blockpos[i] + r]);
                    } else {
                        // all neighbour blocks; equations ascend with the block number, so inserting the
                        // own block at its sorted place equals the reference's per-row std::stable_sort
                        cols.clear();
                        bool own = false;
                        for (int64_t col : nb) {
                            if (!own && col > i) {
                                for (int64_t j = 0; j < sz; j++) cols.push_back(blockpos[i] + j);
                                own = true;
                            }
                            for (int64_t j = 0; j < blocksize[col]; j++) cols.push_back(blockpos[col] + j);
                        }
                        if (!own)
                            for (int64_t j = 0; j < sz; j++) cols.push_back(blockpos[i] + j);
                        for (int64_t r = 0; r < sz; r++) std::copy(cols.begin(), cols.end(), ja + ia[blockpos[i] + r]);
                    }
                }
            });
        for (auto &th : pool) th.join();
    }
    return nnz;
}
