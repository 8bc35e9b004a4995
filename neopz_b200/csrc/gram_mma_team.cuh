// gram_mma_team.cuh — DMMA Gram-panel kernel for a TEAM of warps per element (elasticity, larger p).
//
// Same formulation as gram_mma.cuh (element matrix = Gram matrix of the shared-memory panel
// P = sqrt(w|detJ|) * jacinv^T dphi, accumulated by mma.sync.aligned.m8n8k4.f64), generalised so that the
// 8x8 output tiles of one element are split over WPE warps:
//   * Poisson (NS=1): panel rows (q,d), columns = shape functions; work group = one upper tile (ib<=jb).
//   * Elasticity3D (NS=3): panel rows q, columns COMPONENT-MAJOR c = d*NPAD + i.  The nine sums
//     S[v][u] = sum_q w dphix(v,in) dphix(u,jn) of a node pair (in,jn) then sit in nine different tiles
//     (v*NBN+ib, u*NBN+jb) at the SAME in-tile position, i.e. in the same lane and register slot: the nine
//     formulas of TPZElasticity3D.cpp:318-326 are applied per lane with no data exchange.  Work group =
//     the nine tiles of a node-block pair (ib<=jb).
// Groups are dealt round-robin to the WPE warps of the team (compile-time tile lists, accumulators stay in
// registers); geometry and panel phases are shared by the team and separated by a named barrier.
#pragma once

// SB_ > 0 (Poisson, higher p): the upper triangle of tiles is cut into SB x SB "superblocks", ONE per warp of the
// team (WPE must equal the number of upper superblocks): a k-step then costs 2*SB fragment loads for SB*SB DMMAs
// (SB loads for the SB(SB+1)/2 DMMAs of a diagonal superblock) instead of 2 loads per DMMA.
template <int NN_, int N_, int NS_, int WPE_, int EPC_, int MINB_, int SB_ = 0>
struct TeamCfg {
    static constexpr int NN = NN_, N = N_, NS = NS_, WPE = WPE_, EPC = EPC_, MINB = MINB_, SB = SB_;
    static constexpr int NBN = (N + 7) / 8;              // node blocks of 8
    static constexpr int NPAD = 8 * NBN;
    static constexpr int MP = NS * NPAD;                 // panel columns
    static constexpr int LD = MP + 4;                    // bank-conflict-free fragment loads (see gram_mma.cuh)
    static constexpr int NP = ((N + 15) / 16) * 16;      // padded row length of the shape tables
    static constexpr int KQ = NS == 1 ? 3 : 1;           // panel rows per integration point
    static constexpr int QC = NS == 1 ? 4 : 8;           // points per chunk
    static constexpr int KC = QC * KQ;                   // panel rows per chunk (multiple of 4)
    static constexpr int NSB = SB > 0 ? NBN / SB : NBN;  // superblocks per direction (SB == 0: a group = a tile pair)
    static constexpr int NGROUPS = NSB * (NSB + 1) / 2;  // node-block pairs ib <= jb (SB > 0: superblock pairs)
    static constexpr int TPG = SB > 0 ? SB * SB : (NS == 1 ? 1 : 9);  // tiles per group
    static constexpr int GPW = (NGROUPS + WPE - 1) / WPE;  // groups per warp
    static_assert(SB == 0 || (NS == 1 && NBN % (SB > 0 ? SB : 1) == 0 && NGROUPS == WPE), "superblocks: Poisson, SB | NBN, one per warp");
    static constexpr int JS = 11;                        // sqrt(w|detJ|)*jacinv (9), w|detJ|, sqrt(w|detJ|)
    static constexpr int XSP = NN * 3 + ((NN * 3) & 1);
    static constexpr int TEAM_THREADS = WPE * 32;
    static constexpr int NTHREADS = EPC * TEAM_THREADS;
    static constexpr int M = N * NS;
    static constexpr int FPT = (M + TEAM_THREADS - 1) / TEAM_THREADS;  // load-vector items per thread
    static constexpr int SLOTS = WPE * GPW * TPG * 2 * 32;             // scatter-map entries per element
    __host__ __device__ static int qstride(int nq) { return ((nq + 31) / 32) * 32; }
    __host__ __device__ static int team_doubles(int nq) { int n = XSP + qstride(nq) * JS + 2 * KC * LD; return n + (n & 1); }  // panel double-buffered
    static size_t smem_bytes(int nq) { return sizeof(double) * (size_t)EPC * team_doubles(nq); }
    // group index -> (ib, jb), row-major over the upper triangle
    __host__ __device__ static constexpr int group_ib(int gidx) {
        int ib = 0, t = gidx;
        for (int row = 0; row < NSB - 1; row++) {
            const int cnt = NSB - row;
            if (ib == row && t >= cnt) { t -= cnt; ib = row + 1; }
        }
        return ib;
    }
    __host__ __device__ static constexpr int group_jb(int gidx) {
        int ib = 0, t = gidx;
        for (int row = 0; row < NSB - 1; row++) {
            const int cnt = NSB - row;
            if (ib == row && t >= cnt) { t -= cnt; ib = row + 1; }
        }
        return ib + t;
    }
};

template <int THREADS>
__device__ __forceinline__ void team_sync(int team) {
    if (THREADS == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(THREADS) : "memory");
}

// one chunk of KC panel rows into the accumulators of warp role W
template <class C, int W>
__device__ __forceinline__ void team_mma_chunk(const double *__restrict__ Pn, double (&acc)[C::GPW * C::TPG][2], int g, int tg) {
    constexpr int LD = C::LD, NPAD = C::NPAD;
#pragma unroll
    for (int s = 0; s < C::KC / 4; s++) {
        const double *row = Pn + (4 * s + tg) * LD + g;
#pragma unroll
        for (int gl = 0; gl < C::GPW; gl++) {
            constexpr int dummy = 0;
            (void)dummy;
            const int gidx = W + gl * C::WPE;  // compile-time after unrolling
            if (gidx < C::NGROUPS) {
                const int ib = C::group_ib(gidx), jb = C::group_jb(gidx);
                if (C::SB > 0) {
                    constexpr int SB = C::SB > 0 ? C::SB : 1;
                    double fi[SB], fj[SB];
#pragma unroll
                    for (int a = 0; a < SB; a++) fi[a] = row[8 * (ib * SB + a)];
#pragma unroll
                    for (int b = 0; b < SB; b++) fj[b] = (ib == jb) ? fi[b] : row[8 * (jb * SB + b)];
#pragma unroll
                    for (int a = 0; a < SB; a++)
#pragma unroll
                        for (int b = 0; b < SB; b++)
                            if (ib != jb || a <= b) dmma_m8n8k4(acc[gl * C::TPG + a * SB + b][0], acc[gl * C::TPG + a * SB + b][1], fi[a], fj[b]);
                } else if (C::NS == 1) {
                    const double fi = row[8 * ib];
                    const double fj = (ib == jb) ? fi : row[8 * jb];
                    dmma_m8n8k4(acc[gl][0], acc[gl][1], fi, fj);
                } else {
                    double fi[3], fj[3];
#pragma unroll
                    for (int v = 0; v < 3; v++) fi[v] = row[v * NPAD + 8 * ib];
#pragma unroll
                    for (int u = 0; u < 3; u++) fj[u] = (ib == jb) ? fi[u] : row[u * NPAD + 8 * jb];
#pragma unroll
                    for (int v = 0; v < 3; v++)
#pragma unroll
                        for (int u = 0; u < 3; u++) dmma_m8n8k4(acc[gl * 9 + v * 3 + u][0], acc[gl * 9 + v * 3 + u][1], fi[v], fj[u]);
                }
            }
        }
    }
}

// epilogue of warp role W: (elasticity: combine the nine sums,) scatter-add through the map
template <class C, int W>
__device__ __forceinline__ void team_epilogue(const VolParams &p, int64_t el, double (&acc)[C::GPW * C::TPG][2], int lane) {
    constexpr int TPG = C::TPG;
    const int32_t *sm = p.smap + (size_t)el * C::SLOTS + (size_t)W * C::GPW * TPG * 64 + lane;
    const int32_t *smT = p.smapT ? p.smapT + (size_t)el * C::SLOTS + (size_t)W * C::GPW * TPG * 64 + lane : nullptr;
#pragma unroll
    for (int gl = 0; gl < C::GPW; gl++) {
        const int gidx = W + gl * C::WPE;
        if (gidx < C::NGROUPS) {
            if (C::SB > 0) {
                // superblock: 8 entries at a time (positions first: the atomics would serialise the loads)
                constexpr int B = 8;
#pragma unroll
                for (int k0 = 0; k0 < TPG * 2; k0 += B) {
                    int32_t pos[B];
                    double val[B];
#pragma unroll
                    for (int k = 0; k < B; k++) pos[k] = __ldcs(sm + (gl * TPG * 2 + k0 + k) * 32);
#pragma unroll
                    for (int k = 0; k < B; k++) val[k] = p.coef[0] * acc[gl * TPG + ((k0 + k) >> 1)][(k0 + k) & 1];
                    scatter_many<B>(p.a, pos, val, p.atomic);
                    if (smT) {
#pragma unroll
                        for (int k = 0; k < B; k++) pos[k] = __ldcs(smT + (gl * TPG * 2 + k0 + k) * 32);
                        scatter_many<B>(p.a, pos, val, p.atomic);
                    }
                }
            } else {
                // positions of the whole group first: the atomics below would serialise the loads
                int32_t pos[TPG * 2];
#pragma unroll
                for (int k = 0; k < TPG * 2; k++) pos[k] = __ldcs(sm + (gl * TPG * 2 + k) * 32);
                double val[TPG * 2];
                if (C::NS == 1) {
                    val[0] = p.coef[0] * acc[gl][0];
                    val[1] = p.coef[0] * acc[gl][1];
                } else {
                    const double C1 = p.coef[0], C2 = p.coef[1], C3 = p.coef[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double S[3][3];
#pragma unroll
                        for (int v = 0; v < 3; v++)
#pragma unroll
                            for (int u = 0; u < 3; u++) S[v][u] = acc[gl * 9 + v * 3 + u][e];
#pragma unroll
                        for (int a = 0; a < 3; a++)
#pragma unroll
                            for (int b = 0; b < 3; b++) {
                                double x;
                                if (a == b) x = (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3;
                                else x = S[b][a] * C1 - S[a][b] * C2;
                                val[(a * 3 + b) * 2 + e] = x;
                            }
                    }
                }
                scatter_many<TPG * 2>(p.a, pos, val, p.atomic);
                if (smT) {
#pragma unroll
                    for (int k = 0; k < TPG * 2; k++) pos[k] = __ldcs(smT + (gl * TPG * 2 + k) * 32);
                    scatter_many<TPG * 2>(p.a, pos, val, p.atomic);
                }
            }
        }
    }
}

// Elasticity with ONE tile group per warp (GPW == 1): the node-block pair (ib, jb) of a warp is a RUNTIME value, all warps
// of the team run the same instructions.  (With compile-time roles the kernel carried WPE copies of the DMMA loop and of
// the epilogue: the hex p=2 kernel stalled 23 % of its issue slots on instruction fetch, profiles/r01_ncu_full_final_team_hexp2elast.csv.)
template <class C>
__device__ __forceinline__ void team_mma_chunk_rt(const double *__restrict__ Pn, double (&acc)[C::GPW * C::TPG][2], int g, int tg, int ib, int jb) {
    constexpr int LD = C::LD, NPAD = C::NPAD;
    const bool diag = ib == jb;
#pragma unroll
    for (int s = 0; s < C::KC / 4; s++) {
        const double *row = Pn + (4 * s + tg) * LD + g;
        double fi[3], fj[3];
#pragma unroll
        for (int v = 0; v < 3; v++) fi[v] = row[v * NPAD + 8 * ib];
        if (diag) {
#pragma unroll
            for (int u = 0; u < 3; u++) fj[u] = fi[u];
        } else {
#pragma unroll
            for (int u = 0; u < 3; u++) fj[u] = row[u * NPAD + 8 * jb];
        }
#pragma unroll
        for (int v = 0; v < 3; v++)
#pragma unroll
            for (int u = 0; u < 3; u++) dmma_m8n8k4(acc[v * 3 + u][0], acc[v * 3 + u][1], fi[v], fj[u]);
    }
}

template <class C>
__device__ __forceinline__ void team_epilogue_rt(const VolParams &p, int64_t el, double (&acc)[C::GPW * C::TPG][2], int lane, int w) {
    constexpr int TPG = C::TPG;
    const int32_t *sm = p.smap + (size_t)el * C::SLOTS + (size_t)w * TPG * 64 + lane;
    const int32_t *smT = p.smapT ? p.smapT + (size_t)el * C::SLOTS + (size_t)w * TPG * 64 + lane : nullptr;
    int32_t pos[TPG * 2];
#pragma unroll
    for (int k = 0; k < TPG * 2; k++) pos[k] = __ldcs(sm + k * 32);  // positions first: the atomics would serialise the loads
    double val[TPG * 2];
    const double C1 = p.coef[0], C2 = p.coef[1], C3 = p.coef[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        double S[3][3];
#pragma unroll
        for (int v = 0; v < 3; v++)
#pragma unroll
            for (int u = 0; u < 3; u++) S[v][u] = acc[v * 3 + u][e];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                double x;
                if (a == b) x = (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3;
                else x = S[b][a] * C1 - S[a][b] * C2;
                val[(a * 3 + b) * 2 + e] = x;
            }
    }
    scatter_many<TPG * 2>(p.a, pos, val, p.atomic);
    if (smT) {
#pragma unroll
        for (int k = 0; k < TPG * 2; k++) pos[k] = __ldcs(smT + k * 32);
        scatter_many<TPG * 2>(p.a, pos, val, p.atomic);
    }
}

template <class C, int W>
struct TeamRole {
    __device__ static __forceinline__ void mma(int w, const double *Pn, double (&acc)[C::GPW * C::TPG][2], int g, int tg) {
        if (w == W) team_mma_chunk<C, W>(Pn, acc, g, tg);
        else TeamRole<C, W + 1>::mma(w, Pn, acc, g, tg);
    }
    __device__ static __forceinline__ void epilogue(int w, const VolParams &p, int64_t el, double (&acc)[C::GPW * C::TPG][2], int lane) {
        if (w == W) team_epilogue<C, W>(p, el, acc, lane);
        else TeamRole<C, W + 1>::epilogue(w, p, el, acc, lane);
    }
};
template <class C>
struct TeamRole<C, C::WPE> {
    __device__ static __forceinline__ void mma(int, const double *, double (&)[C::GPW * C::TPG][2], int, int) {}
    __device__ static __forceinline__ void epilogue(int, const VolParams &, int64_t, double (&)[C::GPW * C::TPG][2], int) {}
};

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, C::MINB) assemble_gram_team_kernel(const VolParams p) {
    constexpr int NN = C::NN, N = C::N, NS = C::NS, LD = C::LD, QC = C::QC, KC = C::KC, JS = C::JS, NP = C::NP, NPAD = C::NPAD;
    constexpr int TT = C::TEAM_THREADS, M = C::M, FPT = C::FPT;
    extern __shared__ double smem[];
    const int team = threadIdx.x / TT;
    const int tt = threadIdx.x - team * TT;  // thread within the team
    const int lane = tt & 31, w = tt >> 5;
    const int nq = p.nq;
    const int QS = C::qstride(nq);
    double *Xs = smem + (size_t)team * C::team_doubles(nq);
    double *JI = Xs + C::XSP;
    double *Pn = JI + QS * JS;
    for (int i = tt; i < 2 * KC * LD; i += TT) Pn[i] = 0.0;  // padding columns stay zero (both panel buffers)
    team_sync<TT>(team);
    const int g = lane >> 2, tg = lane & 3;
    constexpr bool RT = C::NS == 3 && C::SB == 0 && C::GPW == 1;  // runtime warp roles (one tile group per warp)
    const int rt_ib = C::group_ib(w < C::NGROUPS ? w : 0), rt_jb = C::group_jb(w < C::NGROUPS ? w : 0);
    const int64_t nteams = (int64_t)gridDim.x * C::EPC;
    // the load vector needs the panel point by point only for a forcing-function table or a non-zero prestress
    const bool pointwise = p.force != nullptr || (NS == 3 && (p.coef[6] != 0.0 || p.coef[7] != 0.0 || p.coef[8] != 0.0));

    for (int64_t el = (int64_t)blockIdx.x * C::EPC + team; el < p.nel; el += nteams) {
        if (tt < NN) {
            const int64_t node = p.elnodes[el * NN + tt];
            Xs[tt * 3 + 0] = p.xyz[node * 3 + 0];
            Xs[tt * 3 + 1] = p.xyz[node * 3 + 1];
            Xs[tt * 3 + 2] = p.xyz[node * 3 + 2];
        }
        if (!p.rhs_only) {   // pull this element's scatter positions towards L2 while the arithmetic runs
            const char *base = (const char *)(p.smap + (size_t)el * C::SLOTS);
            for (int off = tt * 128; off < C::SLOTS * 4; off += TT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
        team_sync<TT>(team);
        // ---- phase 1: geometry at the integration points ----------------------------------------------
        for (int q = tt; q < nq; q += TT) {
            const double *dn = p.dng_t + q;
            double j00 = 0, j01 = 0, j02 = 0, j10 = 0, j11 = 0, j12 = 0, j20 = 0, j21 = 0, j22 = 0;
#pragma unroll
            for (int a = 0; a < NN; a++) {
                const double d0 = __ldg(dn + (size_t)a * nq), d1 = __ldg(dn + (size_t)(NN + a) * nq), d2 = __ldg(dn + (size_t)(2 * NN + a) * nq);
                const double x = Xs[a * 3], y = Xs[a * 3 + 1], z = Xs[a * 3 + 2];
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
            const double wq = __ldg(p.qw + q) * fabs(det);
            const double sid = sqrt(wq) * id;
            double *o = JI + q;
            o[0 * QS] = (-j12 * j21 + j11 * j22) * sid;
            o[1 * QS] = (j02 * j21 - j01 * j22) * sid;
            o[2 * QS] = (-j02 * j11 + j01 * j12) * sid;
            o[3 * QS] = (j12 * j20 - j10 * j22) * sid;
            o[4 * QS] = (-j02 * j20 + j00 * j22) * sid;
            o[5 * QS] = (j02 * j10 - j00 * j12) * sid;
            o[6 * QS] = (-j11 * j20 + j10 * j21) * sid;
            o[7 * QS] = (j01 * j20 - j00 * j21) * sid;
            o[8 * QS] = (-j01 * j10 + j00 * j11) * sid;
            o[9 * QS] = wq;
            o[10 * QS] = sqrt(wq);
        }
        team_sync<TT>(team);

        double acc[C::GPW * C::TPG][2];
#pragma unroll
        for (int t = 0; t < C::GPW * C::TPG; t++) acc[t][0] = acc[t][1] = 0.0;
        double facc[FPT];
#pragma unroll
        for (int k = 0; k < FPT; k++) facc[k] = 0.0;
        // ---- load vector with a constant force: ef(ns*j+k) = f_k * sum_q w_q phi_j(q)  (TPZMatPoisson.cpp:39-40,
        // TPZElasticity3D.cpp:278) — one pass over the points per dof, outside the panel loop
        if (!p.force) {
#pragma unroll
            for (int k = 0; k < FPT; k++) {
                const int m = tt + k * TT;
                if (m < M) {
                    const int j = NS == 1 ? m : m / 3;
                    double t = 0.0;
                    for (int q = 0; q < nq; q++) t += JI[9 * QS + q] * __ldg(p.phi_pad + (size_t)q * NP + j);
                    facc[k] = NS == 1 ? p.coef[0] * p.coef[1] * t : p.coef[3 + (m - 3 * j)] * t;
                }
            }
        }

        // ---- phase 2: panel rows of QC points into one of the two panel buffers; items (point, shape) over the team
        auto build_panel = [&](int q0, double *Pb) {
            for (int it = tt; it < QC * N; it += TT) {
                const int ql = it / N, i = it - ql * N;
                const int q = q0 + ql;
                double g0 = 0.0, g1 = 0.0, g2 = 0.0;
                if (q < nq) {
                    const double *dp = p.dphi_pad + (size_t)q * 3 * NP + i;
                    const double d0 = __ldg(dp), d1 = __ldg(dp + NP), d2 = __ldg(dp + 2 * NP);
                    const double *ji = JI + q;
                    g0 = ji[0 * QS] * d0 + ji[3 * QS] * d1 + ji[6 * QS] * d2;
                    g1 = ji[1 * QS] * d0 + ji[4 * QS] * d1 + ji[7 * QS] * d2;
                    g2 = ji[2 * QS] * d0 + ji[5 * QS] * d1 + ji[8 * QS] * d2;
                }
                if (NS == 1) {
                    double *row = Pb + (3 * ql) * LD + i;
                    row[0] = g0;
                    row[LD] = g1;
                    row[2 * LD] = g2;
                } else {
                    double *row = Pb + ql * LD + i;  // component-major columns
                    row[0] = g0;
                    row[NPAD] = g1;
                    row[2 * NPAD] = g2;
                }
            }
        };
        // software pipeline over the chunks: the panel of chunk c+1 is built (other buffer) by the same warps that then
        // run the DMMAs of chunk c, so the table loads / FMAs of the build overlap the tensor-pipe work and only ONE
        // team barrier per chunk remains
        build_panel(0, Pn);
        team_sync<TT>(team);
        int buf = 0;
        for (int q0 = 0; q0 < nq; q0 += QC, buf ^= 1) {
            const double *Pc = Pn + buf * (KC * LD);
            if (q0 + QC < nq) build_panel(q0 + QC, Pn + (buf ^ 1) * (KC * LD));
            // ---- phase 3: Gram update by every warp for its tile groups --------------------------------
            if (!p.rhs_only) {
                if constexpr (RT) {
                    if (w < C::NGROUPS) team_mma_chunk_rt<C>(Pc, acc, g, tg, rt_ib, rt_jb);
                } else {
                    TeamRole<C, 0>::mma(w, Pc, acc, g, tg);
                }
            }
            // ---- load vector, point-wise part (forcing-function table and/or prestress only) ------------------
            if (pointwise) {
#pragma unroll
                for (int k = 0; k < FPT; k++) {
                    const int m = tt + k * TT;
                    if (m < M) {
                        for (int ql = 0; ql < QC; ql++) {
                            const int q = q0 + ql;
                            if (q >= nq) break;
                            const double wq = JI[9 * QS + q];
                            if (NS == 1) {
                                facc[k] += wq * p.coef[0] * __ldg(p.phi_pad + (size_t)q * NP + m) * p.force[el * nq + q];
                            } else {
                                const int j = m / 3, kd = m - 3 * j;
                                if (p.force) facc[k] += wq * p.force[(el * nq + q) * 3 + kd] * __ldg(p.phi_pad + (size_t)q * NP + j);
                                // w*dphix(kd,j) = sqrt(w) * panel entry   (TPZElasticity3D.cpp:278)
                                facc[k] -= p.coef[6 + kd] * JI[10 * QS + q] * Pc[ql * LD + kd * NPAD + j];
                            }
                        }
                    }
                }
            }
            team_sync<TT>(team);
        }
#pragma unroll
        for (int k = 0; k < FPT; k++) {
            const int m = tt + k * TT;
            if (m < M) scatter_rhs(p.rhs, p.dest[el * M + m], facc[k], p.atomic);
        }
        if (!p.rhs_only) {
            if constexpr (RT) {
                if (w < C::NGROUPS) team_epilogue_rt<C>(p, el, acc, lane, w);
            } else {
                TeamRole<C, 0>::epilogue(w, p, el, acc, lane);
            }
        }
    }
}

// ---- TPZElasticity3D with ONE WARP PER ELEMENT ------------------------------------------------------------------------
// Same panel, tiles, epilogue and scatter map as the team kernel with GPW == 1 (C = its TeamCfg), but no warp waits for
// another one: a warp computes the geometry of its element, builds the WHOLE panel (all nq rows, its own shared-memory
// buffer) and then walks the NGROUPS tile groups one after the other - nine accumulator tiles, nq/4 DMMA steps, the nine
// formulas, the reductions - so the only synchronisation is __syncwarp.  Why: the team kernel spends its time in the named
// barriers that separate its phases (7.3 stall cycles per issued instruction, profiles/r01_ncu_full_final2_team_hexp2elast.csv)
// while the scatter (3321 reductions per element, ~1.3 LSU cycles each) is what bounds the configuration; here the
// reductions of one warp overlap the DMMAs of the others.  Shared memory per warp: 24 + 11 QS + KR LD doubles (25 KB for
// nq = 27), i.e. 8 warps per SM.
template <class C>
struct WarpElastCfg {
    __host__ __device__ static int krows(int nq) { return ((nq + 3) / 4) * 4; }
    __host__ __device__ static int warp_doubles(int nq) { return 24 + 11 * C::qstride(nq) + krows(nq) * C::LD; }
    static size_t smem_bytes(int nq, int teams) { return sizeof(double) * (size_t)teams * warp_doubles(nq); }
};

// WPE = 1: one warp per element, two tile groups at a time (GI = 2).  WPE = 2: a PAIR of warps shares the geometry and the
// panel of an element (same shared memory per element, twice the warps per SM: 16) and splits the tile groups round-robin;
// the pair meets at a 64-thread named barrier four times per element (coordinates, geometry, panel, end of the element).
template <class C, int WPC, int WPE, int GI>
__global__ void __launch_bounds__(WPC * 32, (8 * WPE / WPC > 0 ? 8 * WPE / WPC : 1)) assemble_gram_warp_elast_kernel(const VolParams p) {
    static_assert(C::NS == 3 && C::SB == 0 && C::GPW == 1 && C::NN == 8, "hexahedra, elasticity, one tile group per team warp");
    static_assert(WPC % WPE == 0 && C::NGROUPS % (WPE * GI) == 0, "whole teams per CTA, whole batches of groups per warp");
    constexpr int N = C::N, LD = C::LD, NP = C::NP, NPAD = C::NPAD, M = C::M, JS = C::JS, TT = 32 * WPE, EPC = WPC / WPE;
    extern __shared__ double smem[];
    const int team = threadIdx.x / TT, tt = threadIdx.x - team * TT;
    const int lane = tt & 31, w = tt >> 5;
    const int nq = p.nq;
    const int QS = C::qstride(nq), KR = WarpElastCfg<C>::krows(nq);
    double *Xs = smem + (size_t)team * WarpElastCfg<C>::warp_doubles(nq);
    double *JI = Xs + 24;
    double *Pn = JI + JS * QS;
    for (int i = tt; i < KR * LD; i += TT) Pn[i] = 0.0;  // padding columns and rows stay zero
    team_sync<TT>(team);
    const int g = lane >> 2, tg = lane & 3;
    const bool pointwise = p.force != nullptr || p.coef[6] != 0.0 || p.coef[7] != 0.0 || p.coef[8] != 0.0;
    const int64_t nteams = (int64_t)gridDim.x * EPC;
    int64_t el = (int64_t)blockIdx.x * EPC + team;
    // corner coordinates one element ahead, node ids two ahead (the dependent loads leave the critical path)
    double cnext = 0.0;
    int32_t node_next = 0;
    if (tt < 24 && el < p.nel) {
        cnext = p.xyz[(int64_t)p.elnodes[el * 8 + tt / 3] * 3 + tt % 3];
        if (el + nteams < p.nel) node_next = p.elnodes[(el + nteams) * 8 + tt / 3];
    }
    for (; el < p.nel; el += nteams) {
        if (!p.rhs_only) {
            const char *base = (const char *)(p.smap + (size_t)el * C::SLOTS);
            for (int off = tt * 128; off < C::SLOTS * 4; off += TT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
        if (tt < 24) {
            Xs[tt] = cnext;
            if (el + nteams < p.nel) {
                cnext = p.xyz[(int64_t)node_next * 3 + tt % 3];
                if (el + 2 * nteams < p.nel) node_next = p.elnodes[(el + 2 * nteams) * 8 + tt / 3];
            }
        }
        team_sync<TT>(team);
        // ---- geometry at the integration points (thread <-> point) ------------------------------------------------------
        for (int q = tt; q < nq; q += TT) {
            const double *dn = p.dng_t + q;
            double j00 = 0, j01 = 0, j02 = 0, j10 = 0, j11 = 0, j12 = 0, j20 = 0, j21 = 0, j22 = 0;
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const double d0 = __ldg(dn + (size_t)a * nq), d1 = __ldg(dn + (size_t)(8 + a) * nq), d2 = __ldg(dn + (size_t)(16 + a) * nq);
                const double x = Xs[a * 3], y = Xs[a * 3 + 1], z = Xs[a * 3 + 2];
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
            const double wq = __ldg(p.qw + q) * fabs(det);
            const double sid = sqrt(wq) * id;
            double *o = JI + q;
            o[0 * QS] = (-j12 * j21 + j11 * j22) * sid;
            o[1 * QS] = (j02 * j21 - j01 * j22) * sid;
            o[2 * QS] = (-j02 * j11 + j01 * j12) * sid;
            o[3 * QS] = (j12 * j20 - j10 * j22) * sid;
            o[4 * QS] = (-j02 * j20 + j00 * j22) * sid;
            o[5 * QS] = (j02 * j10 - j00 * j12) * sid;
            o[6 * QS] = (-j11 * j20 + j10 * j21) * sid;
            o[7 * QS] = (j01 * j20 - j00 * j21) * sid;
            o[8 * QS] = (-j01 * j10 + j00 * j11) * sid;
            o[9 * QS] = wq;
            o[10 * QS] = sqrt(wq);
        }
        team_sync<TT>(team);
        // ---- panel P[q][d * NPAD + i] = sqrt(w|detJ|) dphix(d, i) at point q, items (point, shape function) over the threads
#pragma unroll 2
        for (int it = tt; it < nq * N; it += TT) {
            const int q = it / N, i = it - q * N;
            const double *dp = p.dphi_pad + (size_t)q * 3 * NP + i;
            const double d0 = __ldg(dp), d1 = __ldg(dp + NP), d2 = __ldg(dp + 2 * NP);
            const double *ji = JI + q;
            double *row = Pn + q * LD + i;
            row[0] = ji[0 * QS] * d0 + ji[3 * QS] * d1 + ji[6 * QS] * d2;
            row[NPAD] = ji[1 * QS] * d0 + ji[4 * QS] * d1 + ji[7 * QS] * d2;
            row[2 * NPAD] = ji[2 * QS] * d0 + ji[5 * QS] * d1 + ji[8 * QS] * d2;
        }
        team_sync<TT>(team);
        // ---- load vector: ef(3j+k) = sum_q w (f_k phi_j - sigma0_k dphix(k,j))   (TPZElasticity3D.cpp:278), the last warp
        if (w == WPE - 1) {
            if (!pointwise) {
                // constant force: t_j = sum_q w phi_j(q) once per shape function (lane <-> j, three partial sums), then lane <-> dof
                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                if (lane < N) {
                    int q = 0;
                    for (; q + 2 < nq; q += 3) {
                        t0 += JI[9 * QS + q] * __ldg(p.phi_pad + (size_t)q * NP + lane);
                        t1 += JI[9 * QS + q + 1] * __ldg(p.phi_pad + (size_t)(q + 1) * NP + lane);
                        t2 += JI[9 * QS + q + 2] * __ldg(p.phi_pad + (size_t)(q + 2) * NP + lane);
                    }
                    for (; q < nq; q++) t0 += JI[9 * QS + q] * __ldg(p.phi_pad + (size_t)q * NP + lane);
                }
                const double tj = (t0 + t1) + t2;
                for (int m0 = 0; m0 < M; m0 += 32) {
                    const int m = m0 + lane, j = m < M ? m / 3 : 0;
                    const double t = __shfl_sync(0xffffffffu, tj, j);
                    if (m < M) scatter_rhs(p.rhs, p.dest[el * M + m], p.coef[3 + (m - 3 * j)] * t, p.atomic);
                }
            } else {
                for (int m = lane; m < M; m += 32) {
                    const int j = m / 3, kd = m - 3 * j;
                    double f = 0.0;
                    for (int q = 0; q < nq; q++) {
                        const double fq = p.force ? p.force[(el * nq + q) * 3 + kd] : p.coef[3 + kd];
                        f += JI[9 * QS + q] * fq * __ldg(p.phi_pad + (size_t)q * NP + j);
                        f -= p.coef[6 + kd] * JI[10 * QS + q] * Pn[q * LD + kd * NPAD + j];  // w dphix = sqrt(w) * panel entry
                    }
                    scatter_rhs(p.rhs, p.dest[el * M + m], f, p.atomic);
                }
            }
        }
        if (!p.rhs_only) {
            // ---- the tile groups (node-block pairs ib <= jb) of this warp, GI at a time: 9 GI independent accumulator tiles
            // keep the tensor pipe busy across the dependent k-steps, and the scatter positions are loaded before the DMMA
            // loop, so their latency is hidden behind it
            const double C1 = p.coef[0], C2 = p.coef[1], C3 = p.coef[2];
#pragma unroll 1
            for (int gidx = w * GI; gidx < C::NGROUPS; gidx += WPE * GI) {
                int ib[GI], jb[GI];
                int32_t pos[GI][18];
#pragma unroll
                for (int h = 0; h < GI; h++) {
                    ib[h] = C::group_ib(gidx + h);
                    jb[h] = C::group_jb(gidx + h);
                    const int32_t *sm = p.smap + (size_t)el * C::SLOTS + (size_t)(gidx + h) * 9 * 64 + lane;
#pragma unroll
                    for (int k = 0; k < 18; k++) pos[h][k] = __ldcs(sm + k * 32);
                }
                double acc[GI][9][2];
#pragma unroll
                for (int h = 0; h < GI; h++)
#pragma unroll
                    for (int t = 0; t < 9; t++) acc[h][t][0] = acc[h][t][1] = 0.0;
                const double *col = Pn + tg * LD + g;
#pragma unroll 1
                for (int s = 0; s < KR / 4; s++) {
                    const double *r = col + 4 * s * LD;
#pragma unroll
                    for (int h = 0; h < GI; h++) {
                        double fi[3], fj[3];
#pragma unroll
                        for (int v = 0; v < 3; v++) fi[v] = r[v * NPAD + 8 * ib[h]];
#pragma unroll
                        for (int u = 0; u < 3; u++) fj[u] = r[u * NPAD + 8 * jb[h]];
#pragma unroll
                        for (int v = 0; v < 3; v++)
#pragma unroll
                            for (int u = 0; u < 3; u++) dmma_m8n8k4(acc[h][v * 3 + u][0], acc[h][v * 3 + u][1], fi[v], fj[u]);
                    }
                }
#pragma unroll
                for (int h = 0; h < GI; h++) {
                    double val[18];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double S[3][3];
#pragma unroll
                        for (int v = 0; v < 3; v++)
#pragma unroll
                            for (int u = 0; u < 3; u++) S[v][u] = acc[h][v * 3 + u][e];
#pragma unroll
                        for (int a = 0; a < 3; a++)
#pragma unroll
                            for (int b = 0; b < 3; b++)
                                val[(a * 3 + b) * 2 + e] = a == b ? (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3
                                                                  : S[b][a] * C1 - S[a][b] * C2;
                    }
                    scatter_many<18>(p.a, pos[h], val, p.atomic);
                    if (p.smapT) {
                        const int32_t *smT = p.smapT + (size_t)el * C::SLOTS + (size_t)(gidx + h) * 9 * 64 + lane;
                        int32_t posT[18];
#pragma unroll
                        for (int k = 0; k < 18; k++) posT[k] = __ldcs(smT + k * 32);
                        scatter_many<18>(p.a, posT, val, p.atomic);
                    }
                }
            }
        }
        team_sync<TT>(team);  // Xs, JI and the panel are rewritten for the next element
    }
}

// scatter map of the team kernel: slot ((((W*GPW+gl)*TPG + t)*2 + e)*32 + lane) of element el
template <class C>
__global__ void build_team_smap_kernel(int64_t nel, const int32_t *__restrict__ dest, const int64_t *__restrict__ ia,
                                       const int32_t *__restrict__ ja, int symmetric, int32_t *__restrict__ smap,
                                       int32_t *__restrict__ smapT, int *__restrict__ missing) {
    constexpr int N = C::N, NS = C::NS, SLOTS = C::SLOTS, TPG = C::TPG, M = C::M;
    const int64_t total = nel * SLOTS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx / SLOTS;
        int slot = (int)(idx - el * SLOTS);
        const int lane = slot & 31; slot >>= 5;
        const int e = slot & 1; slot >>= 1;
        const int t = slot % TPG; slot /= TPG;
        const int gl = slot % C::GPW;
        const int W = slot / C::GPW;
        const int gidx = W + gl * C::WPE;
        int32_t pos = -1, posT = -1;
        if (gidx < C::NGROUPS) {
            int ib = C::group_ib(gidx), jb = C::group_jb(gidx);
            bool tile_ok = true;
            if (C::SB > 0) {  // tile t = a*SB + b of superblock (ib, jb)
                constexpr int SB = C::SB > 0 ? C::SB : 1;
                const int ta = t / SB, tb = t % SB;
                tile_ok = ib != jb || ta <= tb;
                ib = ib * SB + ta;
                jb = jb * SB + tb;
            }
            const int in = 8 * ib + (lane >> 2), jn = 8 * jb + 2 * (lane & 3) + e;
            const int a = (NS == 1 || C::SB > 0) ? 0 : t / 3, b = (NS == 1 || C::SB > 0) ? 0 : t % 3;
            if (tile_ok && in < N && jn < N && (in < jn || (in == jn && a <= b))) {
                const int i = in * NS + a, j = jn * NS + b;
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
        }
        smap[idx] = pos;
        if (smapT) smapT[idx] = posT;
    }
}
