// cg_sharded.h — internal interface between multi.cpp (partition, halo maps) and b200asm.cu (contexts, kernels): conjugate
// gradients on a row-sharded matrix that lives on several GPUs of one process (b200asm_multi_cg_solve).  Not part of the C ABI.
#pragma once
#include <cstdint>
#include <string>

struct b200asm_ctx;

struct b200asm_cg_shard {
    b200asm_ctx *ctx;
    int64_t own_first;   // local index of the first owned row
    int64_t nown;        // owned rows
    int64_t row0;        // global number of the first owned row
    int64_t nhalo;       // local equations owned by other GPUs that the owned rows couple to
    const int32_t *halo_local;   // [nhalo] local index
    const int32_t *halo_owner;   // [nhalo] index of the owning shard
    const int32_t *halo_remote;  // [nhalo] local index in the owning shard
};

// f_host / x_host: GLOBAL vectors (every shard copies its slice); f_host == NULL: the assembled load vector
int b200asm_cg_sharded(int nshards, const b200asm_cg_shard *shards, int precond, int64_t max_iter, double tol, int from_current,
                       const double *f_host, double *x_host, int64_t *iters_out, double *resid_out, std::string &err);
