// affine_simplex.cuh — straight-sided tetrahedra (TPZGeoTetrahedra with a linear map): closed-form element matrices.
//
// On an affine element the Jacobian is the same at every integration point (Geom/pzgeotetrahedra.h:106-151: gradx =
// sum_a x_a (x) dN_a with constant dN_a), so the quadrature loop of CalcStiff (Mesh/pzinterpolationspace.cpp:404-473)
// factors out of the geometry:
//     S[v][u](in,jn) := sum_q w_q |detJ| dphix(v,in;q) dphix(u,jn;q)
//                     = |detJ| sum_{e,f} jacinv(e,v) jacinv(f,u) Ghat[e][f](in,jn),
//     Ghat[e][f](in,jn) = sum_q w_q dphi(e,in;q) dphi(f,jn;q)          (reference element only: ONE table per group),
// and the weak forms become
//     TPZMatPoisson (Material/Poisson/TPZMatPoisson.cpp:31-38):      ek(in,jn) = s |detJ| sum_{e,f} (Jinv Jinv^T)[e][f] Ghat[e][f]
//     TPZElasticity3D (Material/Elasticity/TPZElasticity3D.cpp:286-326): the nine formulas on S = Jinv^T Ghat Jinv.
// That is 54 FMAs per node pair instead of 9*nq, no square roots and no panel: the kernel is bound by the scatter, not by
// the FP64 pipe.  (Hexahedra are trilinear, their Jacobian varies from point to point: they keep the Gram/DMMA kernels.)
//
// Mapping: ONE WARP PER ELEMENT, persistent grid.
//   * all lanes compute the (single) Jacobian, its inverse and |detJ| redundantly from broadcast loads;
//   * lane <-> node pair (in <= jn), ROUNDS pairs per lane; the nine Ghat values of a lane's pairs live in REGISTERS for
//     the whole kernel (they do not depend on the element);
//   * the NS x NS block of every pair goes to per-warp shared memory, then lane <-> entry (i <= j) in row-major order of
//     the upper triangle, so that consecutive lanes hit consecutive CSR entries of one row wherever the mesh numbering
//     allows (the NS equations of a node are consecutive columns): warp-wide reds touch ~3x fewer sectors than a
//     lane <-> node-pair scatter (tools/red_pattern.cu: cost is proportional to the sectors touched);
//   * scatter map: SLOTS = 32*ceil(M(M+1)/2 / 32) int32 per element, read as whole 128-byte lines.
#pragma once

template <int N_, int NS_, int WPC_, int MINB_>
struct AffCfg {
    static constexpr int NN = 4, N = N_, NS = NS_, WPC = WPC_, MINB = MINB_;
    static constexpr int M = N * NS;
    static constexpr int NPAIR = N * (N + 1) / 2;
    static constexpr int ROUNDS = (NPAIR + 31) / 32;
    static constexpr int NPP = ROUNDS * 32;        // padded pair count
    static constexpr int KPB = NS * NS;            // values per node pair
    static constexpr int NENT = M * (M + 1) / 2;   // entries (i <= j) of the element matrix
    static constexpr int NK = (NENT + 31) / 32;    // scatter instructions per lane
    static constexpr int SLOTS = NK * 32;          // scatter-map entries per element
    // aux table of the group (device, doubles): Ghat[9][NPP], cphi[N] = sum_q w phi, cd[3][N] = sum_q w dphi
    static constexpr int AUX_G = 0, AUX_CPHI = 9 * NPP, AUX_CD = AUX_CPHI + N, AUX_LEN = AUX_CD + 3 * N;
    __host__ __device__ static constexpr int pair_index(int in, int jn) { return in * N - in * (in - 1) / 2 + (jn - in); }
    static size_t smem_bytes(int) { return sizeof(double) * (size_t)WPC * KPB * NPP + sizeof(int) * SLOTS; }
};

// entry s (row-major upper triangle of the M x M element matrix) -> (i, j), i <= j
template <int M>
__host__ __device__ __forceinline__ void aff_entry(int s, int &i, int &j) {
    i = 0;
    while (s >= M - i) { s -= M - i; i++; }
    j = i + s;
}

template <class C>
__global__ void __launch_bounds__(C::WPC * 32, C::MINB) assemble_affine_simplex_kernel(const VolParams p) {
    constexpr int N = C::N, NS = C::NS, M = C::M, NPP = C::NPP, ROUNDS = C::ROUNDS, KPB = C::KPB, NK = C::NK, SLOTS = C::SLOTS;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *vals = smem + (size_t)warp * KPB * NPP;                 // [KPB][NPP] blocks of this warp's element (measured: faster than
                                                                    // the pair-major layout, 485 vs 445 M el/s on 64^3 x 5 p2 elasticity)
    int *tab = reinterpret_cast<int *>(smem + (size_t)C::WPC * KPB * NPP);  // entry -> index into vals (-1: padding)
    for (int s = threadIdx.x; s < SLOTS; s += blockDim.x) {
        int off = -1;
        if (s < C::NENT) {
            int i, j;
            aff_entry<M>(s, i, j);
            const int in = i / NS, a = i - in * NS, jn = j / NS, b = j - jn * NS;
            off = (a * NS + b) * NPP + C::pair_index(in, jn);
        }
        tab[s] = off;
    }
    __syncthreads();
    // element-independent data of this lane: its node pairs and their Ghat blocks
    double G[ROUNDS][9];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int pr = r * 32 + lane;
#pragma unroll
        for (int k = 0; k < 9; k++) G[r][k] = pr < C::NPAIR ? __ldg(p.aux + C::AUX_G + k * NPP + pr) : 0.0;
    }
    unsigned voff2[(NK + 1) / 2];  // index into vals of this lane's entries, two 16-bit values per register (0xffff: padding)
#pragma unroll
    for (int k = 0; k < (NK + 1) / 2; k++) {
        const unsigned lo = (unsigned)tab[(2 * k) * 32 + lane] & 0xffffu;
        const unsigned hi = 2 * k + 1 < NK ? (unsigned)tab[(2 * k + 1) * 32 + lane] & 0xffffu : 0xffffu;
        voff2[k] = lo | (hi << 16);
    }
    const bool pointwise_force = p.force != nullptr;

    const int64_t nwarps = (int64_t)gridDim.x * C::WPC;
    const int64_t el_first = (int64_t)blockIdx.x * C::WPC + warp;
    // software prefetch: the corner coordinates of the next element are in flight while this one is computed
    double X[4][3];
    if (el_first < p.nel) {
        const int4 nd = *reinterpret_cast<const int4 *>(p.elnodes + el_first * 4);
        const int32_t id[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int k = 0; k < 3; k++) X[a][k] = p.xyz[(int64_t)id[a] * 3 + k];
    }
    for (int64_t el = el_first; el < p.nel; el += nwarps) {
        // ---- geometry: gradx, det, inverse (Mesh/pzgeoel.cpp:1309-1336), identical at every point ------------------
        // gradx(r,d) = sum_a x_a[r] dN_a/dxi_d with dN = (-1,-1,-1), (1,0,0), (0,1,0), (0,0,1) (Topology/tpztetrahedron.cpp:239-260,
        // Geom/pzgeotetrahedra.h:141-149): -x_0 + x_{d+1} plus exact zeros, i.e. bitwise x_{d+1} - x_0
        double j[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int d = 0; d < 3; d++) j[r][d] = X[d + 1][r] - X[0][r];
        const int64_t nxt = el + nwarps;
        if (nxt < p.nel) {
            if (!p.rhs_only) {  // pull the next element's scatter positions towards L2
                const char *base = (const char *)(p.smap + (size_t)nxt * SLOTS);
                for (int off = lane * 128; off < SLOTS * 4; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
            const int4 nd = *reinterpret_cast<const int4 *>(p.elnodes + nxt * 4);
            const int32_t id[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int k = 0; k < 3; k++) X[a][k] = p.xyz[(int64_t)id[a] * 3 + k];
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
        double ji[3][3];  // jacinv(e, v)
        ji[0][0] = (-j[1][2] * j[2][1] + j[1][1] * j[2][2]) * id_;
        ji[0][1] = (j[0][2] * j[2][1] - j[0][1] * j[2][2]) * id_;
        ji[0][2] = (-j[0][2] * j[1][1] + j[0][1] * j[1][2]) * id_;
        ji[1][0] = (j[1][2] * j[2][0] - j[1][0] * j[2][2]) * id_;
        ji[1][1] = (-j[0][2] * j[2][0] + j[0][0] * j[2][2]) * id_;
        ji[1][2] = (j[0][2] * j[1][0] - j[0][0] * j[1][2]) * id_;
        ji[2][0] = (-j[1][1] * j[2][0] + j[1][0] * j[2][1]) * id_;
        ji[2][1] = (j[0][1] * j[2][0] - j[0][0] * j[2][1]) * id_;
        ji[2][2] = (-j[0][1] * j[1][0] + j[0][0] * j[1][1]) * id_;

        // ---- load vector: lane <-> equation ---------------------------------------------------------------------
        if (lane < M) {
            const int jn = lane / NS, k = lane - jn * NS;
            double f;
            if (!pointwise_force) {
                f = (NS == 1 ? p.coef[0] * p.coef[1] : p.coef[3 + k]) * __ldg(p.aux + C::AUX_CPHI + jn);
            } else {
                f = 0.0;
                for (int q = 0; q < p.nq; q++)
                    f += __ldg(p.qw + q) * __ldg(p.phi + (size_t)q * N + jn) * p.force[((size_t)el * p.nq + q) * NS + k];
                if (NS == 1) f *= p.coef[0];
            }
            if (NS == 3) {  // prestress: - sigma0_k sum_q w dphix(k, jn)   (TPZElasticity3D.cpp:278)
                double gk = 0.0;
#pragma unroll
                for (int e = 0; e < 3; e++) gk += (k == 0 ? ji[e][0] : (k == 1 ? ji[e][1] : ji[e][2])) * __ldg(p.aux + C::AUX_CD + e * N + jn);
                f -= p.coef[6 + k] * gk;
            }
            scatter_rhs(p.rhs, p.dest[el * M + lane], f * adet, p.atomic);
        }
        if (p.rhs_only) continue;

        // ---- node-pair blocks --------------------------------------------------------------------------------------
        if (NS == 1) {
            double mm[9];  // (Jinv Jinv^T)[e][f] s |detJ|
            const double sc = p.coef[0] * adet;
#pragma unroll
            for (int e = 0; e < 3; e++)
#pragma unroll
                for (int f = 0; f < 3; f++) mm[e * 3 + f] = (ji[e][0] * ji[f][0] + ji[e][1] * ji[f][1] + ji[e][2] * ji[f][2]) * sc;
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 9; k++) v += mm[k] * G[r][k];
                vals[r * 32 + lane] = v;
            }
        } else {
            const double C1 = p.coef[0] * adet, C2 = p.coef[1] * adet, C3 = p.coef[2] * adet;
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                double T[3][3], S[3][3];
#pragma unroll
                for (int e = 0; e < 3; e++)
#pragma unroll
                    for (int u = 0; u < 3; u++) T[e][u] = G[r][e * 3 + 0] * ji[0][u] + G[r][e * 3 + 1] * ji[1][u] + G[r][e * 3 + 2] * ji[2][u];
#pragma unroll
                for (int v = 0; v < 3; v++)
#pragma unroll
                    for (int u = 0; u < 3; u++) S[v][u] = ji[0][v] * T[0][u] + ji[1][v] * T[1][u] + ji[2][v] * T[2][u];
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int b = 0; b < 3; b++) {
                        double x;
                        if (a == b) x = (S[(a + 1) % 3][(a + 1) % 3] + S[(a + 2) % 3][(a + 2) % 3]) * C1 + S[a][a] * C3;
                        else x = S[b][a] * C1 - S[a][b] * C2;
                        vals[(a * 3 + b) * NPP + r * 32 + lane] = x;
                    }
            }
        }
        __syncwarp();
        // ---- scatter: lane <-> entry, positions first (the reds would serialise the loads) ---------------------------
        const int32_t *sm = p.smap + (size_t)el * SLOTS + lane;
        const int32_t *smT = p.smapT ? p.smapT + (size_t)el * SLOTS + lane : nullptr;
        constexpr int B = 8;
#pragma unroll
        for (int k0 = 0; k0 < NK; k0 += B) {
            int32_t pos[B];
            double val[B];
#pragma unroll
            for (int k = 0; k < B; k++) pos[k] = k0 + k < NK ? __ldcs(sm + (k0 + k) * 32) : -1;
#pragma unroll
            for (int k = 0; k < B; k++) {
                const int kk = k0 + k < NK ? k0 + k : 0;
                const unsigned o = (voff2[kk >> 1] >> ((kk & 1) * 16)) & 0xffffu;  // (padding slots carry position -1)
                val[k] = (k0 + k < NK && o != 0xffffu) ? vals[o] : 0.0;
            }
            scatter_many<B>(p.a, pos, val, p.atomic);
            if (smT) {
#pragma unroll
                for (int k = 0; k < B; k++) pos[k] = k0 + k < NK ? __ldcs(smT + (k0 + k) * 32) : -1;
                scatter_many<B>(p.a, pos, val, p.atomic);
            }
        }
        __syncwarp();  // vals is rewritten by the next element
    }
}

// scatter map of the affine kernel: slot s = entry (i <= j) in row-major order of the upper triangle
template <class C>
__global__ void build_aff_smap_kernel(int64_t nel, const int32_t *__restrict__ dest, const int64_t *__restrict__ ia,
                                      const int32_t *__restrict__ ja, int symmetric, int32_t *__restrict__ smap,
                                      int32_t *__restrict__ smapT, int *__restrict__ missing) {
    constexpr int M = C::M, SLOTS = C::SLOTS;
    const int64_t total = nel * SLOTS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx / SLOTS;
        const int s = (int)(idx - el * SLOTS);
        int32_t pos = -1, posT = -1;
        if (s < C::NENT) {
            int i, j;
            aff_entry<M>(s, i, j);
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
