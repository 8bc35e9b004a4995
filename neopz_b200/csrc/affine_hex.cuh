// affine_hex.cuh — hexahedra whose trilinear map (Geom/TPZGeoCube.h:106-151) is AFFINE (parallelepipeds: every cell of
// TPZGeoMeshTools::CreateGeoMeshOnGrid, Mesh/TPZGeoMeshTools.cpp:105-219, and of any sheared / stretched lattice).
//
// Same closed form as affine_simplex.cuh: with a constant Jacobian the quadrature loop of CalcStiff factors out,
//     S[v][u](in,jn) = |detJ| sum_{e,f} jacinv(e,v) jacinv(f,u) Ghat[e][f](in,jn),   Ghat = sum_q w dphi(e,in) dphi(f,jn),
// TPZMatPoisson: ek(in,jn) = s |detJ| sum_{e,f} (Jinv Jinv^T)[e][f] Ghat[e][f];  TPZElasticity3D: the nine formulas of
// Material/Elasticity/TPZElasticity3D.cpp:318-326 on S.  9 (Poisson) or 54 + 21 (elasticity) FMAs per node pair instead
// of 9*nq / 57*nq: the kernel is bound by the scatter (LSU wavefronts / L2 reductions / HBM), not by the FP64 pipe.
//
// Whether a group qualifies is MEASURED on the device (hex_affinity_kernel) every time the node coordinates change:
// the four non-affine coefficient vectors of the trilinear map must vanish to kAffineTol relative to the shortest edge
// vector, for every element of the group; otherwise the group keeps its Gram / DMMA kernel.  The Jacobian of a qualifying
// element differs from point to point only by the rounding noise of the node coordinates themselves.
//
// Mapping: one warp per element, persistent grid.  Ghat lives in shared memory ([9][NPP], conflict-free: consecutive lanes
// read consecutive pairs).  A ROUND is 32 node pairs (in <= jn, row-major): lane <-> pair computes the NS x NS block, the
// blocks go through a per-warp staging buffer so that lane <-> (pair, a, b) with b fastest for the scatter: the NS
// equations of a node are consecutive columns of a CSR row, so a warp-wide red touches runs of NS consecutive doubles.
// Scatter map: [el][round][a*NS+b][32] int32, read as whole 128-byte lines.
#pragma once

// NN_ = 8: parallelepiped hexahedra; NN_ = 4: straight-sided tetrahedra of the orders whose table no longer fits the registers of
// affine_simplex.cuh (p = 3, 4: 210 / 630 node pairs) - same kernel, the (single) Jacobian from the four corners
template <int N_, int NS_, int WPC_, int MINB_, int NN_ = 8>
struct AffHexCfg {
    static constexpr int NN = NN_, N = N_, NS = NS_, WPC = WPC_, MINB = MINB_;
    static constexpr int M = N * NS;
    static constexpr int NPAIR = N * (N + 1) / 2;
    static constexpr int ROUNDS = (NPAIR + 31) / 32;
    static constexpr int NPP = ROUNDS * 32;
    static constexpr int KPB = NS * NS;
    static constexpr int SLOTS = ROUNDS * KPB * 32;
    static constexpr int STAGE = KPB > 1 ? 2 * 32 * KPB : 0;  // doubles per warp (double-buffered)
    static constexpr int NG = NS == 1 ? 6 : 9;  // table rows in shared memory (Poisson: symmetrised, see the kernel)
    static constexpr int AUX_G = 0, AUX_CPHI = 9 * NPP, AUX_CD = AUX_CPHI + N, AUX_LEN = AUX_CD + 3 * N;
    __host__ __device__ static constexpr int pair_index(int in, int jn) { return in * N - in * (in - 1) / 2 + (jn - in); }
    static size_t smem_bytes(int) { return sizeof(double) * ((size_t)NG * NPP + (size_t)WPC * STAGE); }
};

constexpr double kAffineTol = 1e-13;

// corner signs of the reference cube (Topology/tpzcube.cpp:367-417): node a sits at (sx, sy, sz) in {-1,+1}^3
__device__ __forceinline__ double hex_sign(int a, int d) {
    // a: 0 (-,-,-) 1 (+,-,-) 2 (+,+,-) 3 (-,+,-) 4 (-,-,+) 5 (+,-,+) 6 (+,+,+) 7 (-,+,+)
    const unsigned sx = 0x66u, sy = 0xCCu, sz = 0xF0u;  // bit a set <=> coordinate is +1
    const unsigned m = d == 0 ? sx : (d == 1 ? sy : sz);
    return ((m >> a) & 1u) ? 1.0 : -1.0;
}

// flag != 0 when some element of the group is not a parallelepiped (to kAffineTol)
__global__ void hex_affinity_kernel(int64_t nel, const int32_t *__restrict__ elnodes, const double *__restrict__ xyz, int *flag) {
    for (int64_t el = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; el < nel; el += (int64_t)gridDim.x * blockDim.x) {
        double c[8][3];  // coefficients of 1, xi, eta, zeta, xi eta, xi zeta, eta zeta, xi eta zeta (times 8)
#pragma unroll
        for (int k = 0; k < 8; k++) c[k][0] = c[k][1] = c[k][2] = 0.0;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int64_t node = elnodes[el * 8 + a];
            const double sx = hex_sign(a, 0), sy = hex_sign(a, 1), sz = hex_sign(a, 2);
            const double w[8] = {1.0, sx, sy, sz, sx * sy, sx * sz, sy * sz, sx * sy * sz};
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double x = xyz[node * 3 + r];
#pragma unroll
                for (int k = 0; k < 8; k++) c[k][r] += w[k] * x;
            }
        }
        double edge = 1e300, dev = 0.0;
#pragma unroll
        for (int k = 1; k < 4; k++) edge = fmin(edge, c[k][0] * c[k][0] + c[k][1] * c[k][1] + c[k][2] * c[k][2]);
#pragma unroll
        for (int k = 4; k < 8; k++) dev = fmax(dev, c[k][0] * c[k][0] + c[k][1] * c[k][1] + c[k][2] * c[k][2]);
        if (!(dev <= kAffineTol * kAffineTol * edge)) *flag = 1;
    }
}

template <class C>
__global__ void __launch_bounds__(C::WPC * 32, C::MINB) assemble_affine_hex_kernel(const VolParams p) {
    constexpr int N = C::N, NS = C::NS, M = C::M, NPP = C::NPP, ROUNDS = C::ROUNDS, KPB = C::KPB, SLOTS = C::SLOTS;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *Gs = smem;                                              // [NG][NPP]
    double *stage = smem + C::NG * NPP + (size_t)warp * C::STAGE;   // 2 x [32][KPB] blocks of the current / next round
    if (NS == 1) {
        // Poisson contracts Ghat with the SYMMETRIC matrix Jinv Jinv^T: only Ghat[e][e] and Ghat[e][f] + Ghat[f][e] are needed
        for (int i = threadIdx.x; i < NPP; i += blockDim.x) {
            const double *g = p.aux + C::AUX_G + i;
            Gs[0 * NPP + i] = __ldg(g + 0 * NPP);
            Gs[1 * NPP + i] = __ldg(g + 4 * NPP);
            Gs[2 * NPP + i] = __ldg(g + 8 * NPP);
            Gs[3 * NPP + i] = __ldg(g + 1 * NPP) + __ldg(g + 3 * NPP);
            Gs[4 * NPP + i] = __ldg(g + 2 * NPP) + __ldg(g + 6 * NPP);
            Gs[5 * NPP + i] = __ldg(g + 5 * NPP) + __ldg(g + 7 * NPP);
        }
    } else {
        for (int i = threadIdx.x; i < 9 * NPP; i += blockDim.x) Gs[i] = __ldg(p.aux + C::AUX_G + i);
    }
    __syncthreads();
    const bool pointwise_force = p.force != nullptr;
    const int64_t nwarps = (int64_t)gridDim.x * C::WPC;

    for (int64_t el = (int64_t)blockIdx.x * C::WPC + warp; el < p.nel; el += nwarps) {
        double j[3][3];
        if (C::NN == 4) {
            // straight-sided tetrahedron (Geom/pzgeotetrahedra.h:106-151): gradx(r,d) = x_{d+1}(r) - x_0(r)
            const int4 n0 = *reinterpret_cast<const int4 *>(p.elnodes + el * 4);
            const int32_t id[4] = {n0.x, n0.y, n0.z, n0.w};
            double x0[3];
#pragma unroll
            for (int r = 0; r < 3; r++) x0[r] = p.xyz[(int64_t)id[0] * 3 + r];
#pragma unroll
            for (int d = 0; d < 3; d++)
#pragma unroll
                for (int r = 0; r < 3; r++) j[r][d] = p.xyz[(int64_t)id[d + 1] * 3 + r] - x0[r];
        } else {
        // corner coordinates: broadcast loads (every lane needs the whole Jacobian)
        const int4 n0 = *reinterpret_cast<const int4 *>(p.elnodes + el * 8), n1 = *reinterpret_cast<const int4 *>(p.elnodes + el * 8 + 4);
        const int32_t id[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
        // gradx at the centre: sum_a x_a dN_a(0), dN_a(0) = sign_a / 8  (constant over a parallelepiped)
#pragma unroll
        for (int r = 0; r < 3; r++) j[r][0] = j[r][1] = j[r][2] = 0.0;
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double x = p.xyz[(int64_t)id[a] * 3 + r];
#pragma unroll
                for (int d = 0; d < 3; d++) j[r][d] += hex_sign(a, d) * x;
            }
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int d = 0; d < 3; d++) j[r][d] *= 0.125;
        }
        // next element: node coordinates and scatter positions towards L2
        const int64_t nxt = el + nwarps;
        if (nxt < p.nel) {
            if (lane < C::NN) {
                const int64_t node = p.elnodes[nxt * C::NN + lane];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.xyz + node * 3));
            }
            if (!p.rhs_only) {
                const char *base = (const char *)(p.smap + (size_t)nxt * SLOTS);
                for (int off = lane * 128; off < SLOTS * 4; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
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
        for (int m = lane; m < M; m += 32) {
            const int jn = m / NS, k = m - jn * NS;
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
            scatter_rhs(p.rhs, p.dest[el * M + m], f * adet, p.atomic);
        }
        if (p.rhs_only) continue;

        const int32_t *sm = p.smap + (size_t)el * SLOTS + lane;
        const int32_t *smT = p.smapT ? p.smapT + (size_t)el * SLOTS + lane : nullptr;
        if (NS == 1) {
            double mm[6];  // (Jinv Jinv^T)[e][f] s |detJ|: 00, 11, 22, 01, 02, 12
            const double sc = p.coef[0] * adet;
            {
                constexpr int E[6] = {0, 1, 2, 0, 0, 1}, F[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
                for (int q = 0; q < 6; q++) mm[q] = (ji[E[q]][0] * ji[F[q]][0] + ji[E[q]][1] * ji[F[q]][1] + ji[E[q]][2] * ji[F[q]][2]) * sc;
            }
            constexpr int B = 4;  // rounds per batch: positions first, then the arithmetic, then the reds
#pragma unroll
            for (int r0 = 0; r0 < ROUNDS; r0 += B) {
                int32_t pos[B];
                double val[B];
#pragma unroll
                for (int k = 0; k < B; k++) pos[k] = r0 + k < ROUNDS ? __ldcs(sm + (r0 + k) * 32) : -1;
#pragma unroll
                for (int k = 0; k < B; k++) {
                    double v = 0.0;
                    if (r0 + k < ROUNDS) {
#pragma unroll
                        for (int q = 0; q < 6; q++) v += mm[q] * Gs[q * NPP + (r0 + k) * 32 + lane];
                    }
                    val[k] = v;
                }
                scatter_many<B>(p.a, pos, val, p.atomic);
                if (smT) {
#pragma unroll
                    for (int k = 0; k < B; k++) pos[k] = r0 + k < ROUNDS ? __ldcs(smT + (r0 + k) * 32) : -1;
                    scatter_many<B>(p.a, pos, val, p.atomic);
                }
            }
        } else {
            const double C1 = p.coef[0] * adet, C2 = p.coef[1] * adet, C3 = p.coef[2] * adet;
            // software pipeline over the rounds: the positions of round r+1 are in flight while round r computes and
            // scatters; the staging buffer alternates, so one __syncwarp per round is enough
            int32_t pnext[KPB];
#pragma unroll
            for (int k = 0; k < KPB; k++) pnext[k] = __ldcs(sm + k * 32);
            for (int r = 0; r < ROUNDS; r++) {
                int32_t pos[KPB];
                double val[KPB];
#pragma unroll
                for (int k = 0; k < KPB; k++) pos[k] = pnext[k];
                if (r + 1 < ROUNDS) {
#pragma unroll
                    for (int k = 0; k < KPB; k++) pnext[k] = __ldcs(sm + ((r + 1) * KPB + k) * 32);
                }
                double *st = stage + (r & 1) * (32 * KPB);
                double G[9];
#pragma unroll
                for (int q = 0; q < 9; q++) G[q] = Gs[q * NPP + r * 32 + lane];
                double T[3][3], S[3][3];
#pragma unroll
                for (int e = 0; e < 3; e++)
#pragma unroll
                    for (int u = 0; u < 3; u++) T[e][u] = G[e * 3 + 0] * ji[0][u] + G[e * 3 + 1] * ji[1][u] + G[e * 3 + 2] * ji[2][u];
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
                        st[lane * KPB + a * 3 + b] = x;
                    }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < KPB; k++) val[k] = st[k * 32 + lane];
                scatter_many<KPB>(p.a, pos, val, p.atomic);
                if (smT) {
#pragma unroll
                    for (int k = 0; k < KPB; k++) pos[k] = __ldcs(smT + (r * KPB + k) * 32);
                    scatter_many<KPB>(p.a, pos, val, p.atomic);
                }
            }
            __syncwarp();  // (the next element starts with buffer 0 again)
        }
    }
}

// scatter map: slot ((r*KPB + k)*32 + lane) of element el <-> idx = k*32 + lane within round r, local pair idx / KPB,
// component pair (a, b) = idx % KPB; entry (in*NS + a, jn*NS + b) of the pair r*32 + idx/KPB  (diagonal pairs: a <= b only)
template <class C>
__global__ void build_affhex_smap_kernel(int64_t nel, const int32_t *__restrict__ dest, const int64_t *__restrict__ ia,
                                         const int32_t *__restrict__ ja, int symmetric, int32_t *__restrict__ smap,
                                         int32_t *__restrict__ smapT, int *__restrict__ missing) {
    constexpr int N = C::N, NS = C::NS, M = C::M, SLOTS = C::SLOTS, KPB = C::KPB;
    const int64_t total = nel * SLOTS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx / SLOTS;
        const int slot = (int)(idx - el * SLOTS);
        const int r = slot / (KPB * 32), w = slot - r * (KPB * 32);  // w = k*32 + lane
        int pair = r * 32 + w / KPB;
        const int kk = w % KPB, a = kk / NS, b = kk % NS;
        int32_t pos = -1, posT = -1;
        if (pair < C::NPAIR) {
            int in = 0;
            while (pair >= N - in) { pair -= N - in; in++; }
            const int jn = in + pair;
            if (in < jn || a <= b) {
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
