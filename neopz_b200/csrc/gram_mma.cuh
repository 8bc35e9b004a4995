// gram_mma.cuh — FP64 tensor-core (DMMA) element kernel: the element matrix as the Gram matrix of a
// shared-memory panel, accumulated with mma.sync.aligned.m8n8k4.f64.
//
// Why DMMA here (measured on this pool's B200, profiles/r01_fp64_peak.json and
// profiles/r01_ncu_full_assemble_volume_hexp2poisson_v1.csv): DMMA peaks at 37.1 TFLOP/s vs 33.7 for DFMA,
// and the register-tile DFMA kernel (b200asm.cu, v1) is bound by shared-memory wavefronts (2 loads per 9
// FMAs, 254 registers -> 6 warps/SM, FP64 pipe 14 % busy).  One m8n8k4 does 256 FMAs from ONE 8-byte
// shared load per operand and lane, and a lane keeps only 2 accumulators per 8x8 tile, so a warp holds a
// whole 32x32 upper triangle in 20 registers pairs and 16 warps fit per SM.
//
// Mapping (one warp per element, no block-level synchronisation at all):
//   phase 1  lane q: Jacobian, inverse, w|detJ| of integration point q            -> per-warp smem
//   phase 2  per chunk of QC=4 points: lane <-> shape function computes the panel rows
//            P[(q,d)][i] = sqrt(w|detJ|) * sum_e jacinv(e,d) dphi(e,i)            -> per-warp smem [12][LD]
//   phase 3  3 k-steps of 4 rows: one fragment load per 8-column block, then one DMMA per upper tile
//            (fragment of block b: lane l reads P[4s+(l&3)][8b+(l>>2)]; it is the A operand of tile
//            rows b and the B operand of tile columns b, PTX ISA "mma.m8n8k4" f64 layouts)
//   epilogue lane l holds C[8bi+(l>>2)][8bj+2(l&3)+{0,1}] of every tile: scatter through the map.
// LD = 8*NB + 4 doubles keeps the 16 lanes of a half-warp (4 k-rows x 4 columns) in 16 distinct banks.
#pragma once

template <int NN_, int N_, int WPC_, int MINB_>
struct MmaCfg {
    static constexpr int NN = NN_, N = N_, WPC = WPC_, MINB = MINB_;
    static constexpr int NB = (N + 7) / 8;
    static constexpr int MP = 8 * NB;
    static constexpr int LD = MP + 4;   // 4 k-rows x 4 column groups of a half-warp fall into 16 distinct 8-byte banks
    static constexpr int NP = ((N + 15) / 16) * 16;  // padded row length of the shape tables (whole 128-byte lines)
    static constexpr int NTILES = NB * (NB + 1) / 2;
    static constexpr int QC = 4, KC = 12;
    static constexpr int JS = 10;       // sqrt(w|detJ|)*jacinv (9) and w|detJ|
    static constexpr int XSP = NN * 3 + ((NN * 3) & 1);
    static constexpr int SLOTS = NTILES * 2 * 32;  // scatter-map entries per element
    __host__ __device__ static int qstride(int nq) { return ((nq + 31) / 32) * 32; }  // JI is [JS][qstride]
    __host__ __device__ static int warp_doubles(int nq) { int n = XSP + qstride(nq) * JS + KC * LD; return n + (n & 1); }
    static size_t smem_bytes(int nq) { return sizeof(double) * (size_t)WPC * warp_doubles(nq); }
};

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <class C>
__global__ void __launch_bounds__(C::WPC * 32, C::MINB) assemble_gram_mma_kernel(const VolParams p) {
    constexpr int NN = C::NN, N = C::N, NB = C::NB, LD = C::LD, NTILES = C::NTILES, QC = C::QC, KC = C::KC, JS = C::JS, NP = C::NP;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nq = p.nq;
    double *Xs = smem + (size_t)warp * C::warp_doubles(nq);
    double *JI = Xs + C::XSP;
    const int QS = C::qstride(nq);
    double *Pn = JI + QS * JS;
    for (int i = lane; i < KC * LD; i += 32) Pn[i] = 0.0;  // padding columns stay zero
    __syncwarp();
    const int g = lane >> 2, tg = lane & 3;
    const int64_t nwarps = (int64_t)gridDim.x * C::WPC;

    const int64_t el_first = (int64_t)blockIdx.x * C::WPC + warp;
    int32_t next_node = (lane < NN && el_first < p.nel) ? p.elnodes[el_first * NN + lane] : 0;
    for (int64_t el = el_first; el < p.nel; el += nwarps) {
        // the scatter positions of this element are fetched now (HBM latency hidden behind the arithmetic):
        // the atomics of the epilogue would otherwise serialise these loads
        int32_t pos[NTILES * 2];
        if (!p.rhs_only) {
            const int32_t *sm = p.smap + (size_t)el * C::SLOTS + lane;
#pragma unroll
            for (int k = 0; k < NTILES * 2; k++) pos[k] = __ldcs(sm + k * 32);
        }
        if (lane < NN) {
            const int64_t node = next_node;
            Xs[lane * 3 + 0] = p.xyz[node * 3 + 0];
            Xs[lane * 3 + 1] = p.xyz[node * 3 + 1];
            Xs[lane * 3 + 2] = p.xyz[node * 3 + 2];
            if (el + nwarps < p.nel) next_node = p.elnodes[(el + nwarps) * NN + lane];
        }
        __syncwarp();
        // ---- phase 1: geometry at the integration points (Geom/TPZGeoCube.h:141-149, Mesh/pzgeoel.cpp:1309-1336)
        for (int q = lane; q < nq; q += 32) {
            const double *dn = p.dng_t + q;  // [3][NN][nq]: consecutive lanes read consecutive doubles
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
            const double w = __ldg(p.qw + q) * fabs(det);  // weight *= fabs(detjac)
            const double sid = sqrt(w) * id;
            double *o = JI + q;  // field k of point q at JI[k*QS + q]: conflict-free stores, broadcast reads
            o[0 * QS] = (-j12 * j21 + j11 * j22) * sid;
            o[1 * QS] = (j02 * j21 - j01 * j22) * sid;
            o[2 * QS] = (-j02 * j11 + j01 * j12) * sid;
            o[3 * QS] = (j12 * j20 - j10 * j22) * sid;
            o[4 * QS] = (-j02 * j20 + j00 * j22) * sid;
            o[5 * QS] = (j02 * j10 - j00 * j12) * sid;
            o[6 * QS] = (-j11 * j20 + j10 * j21) * sid;
            o[7 * QS] = (j01 * j20 - j00 * j21) * sid;
            o[8 * QS] = (-j01 * j10 + j00 * j11) * sid;
            o[9 * QS] = w;
        }
        __syncwarp();

        double acc[NTILES][2];
#pragma unroll
        for (int t = 0; t < NTILES; t++) acc[t][0] = acc[t][1] = 0.0;

        for (int q0 = 0; q0 < (p.rhs_only ? 0 : nq); q0 += QC) {
            // ---- phase 2: panel rows of QC points (Mesh/TPZCompElH1.cpp:147): lane <-> shape function,
            // the table loads of all QC points are issued before the first use
            for (int i = lane; i < N; i += 32) {
                double d[QC][3];
#pragma unroll
                for (int ql = 0; ql < QC; ql++) {
                    const int q = q0 + ql;
                    const double *dp = p.dphi_pad + (size_t)q * 3 * NP + i;  // rows padded to whole 128-byte lines
                    d[ql][0] = q < nq ? __ldg(dp) : 0.0;
                    d[ql][1] = q < nq ? __ldg(dp + NP) : 0.0;
                    d[ql][2] = q < nq ? __ldg(dp + 2 * NP) : 0.0;
                }
#pragma unroll
                for (int ql = 0; ql < QC; ql++) {
                    const int q = min(q0 + ql, nq - 1);  // tail rows: d == 0 -> zero rows
                    const double *ji = JI + q;
                    const double g0 = ji[0 * QS] * d[ql][0] + ji[3 * QS] * d[ql][1] + ji[6 * QS] * d[ql][2];
                    const double g1 = ji[1 * QS] * d[ql][0] + ji[4 * QS] * d[ql][1] + ji[7 * QS] * d[ql][2];
                    const double g2 = ji[2 * QS] * d[ql][0] + ji[5 * QS] * d[ql][1] + ji[8 * QS] * d[ql][2];
                    double *row = Pn + (3 * ql) * LD + i;
                    row[0] = g0;
                    row[LD] = g1;
                    row[2 * LD] = g2;
                }
            }
            __syncwarp();
            // ---- phase 3: Gram update, 3 k-steps of 4 panel rows -----------------------------------
#pragma unroll
            for (int s = 0; s < KC / 4; s++) {
                double fr[NB];
#pragma unroll
                for (int b = 0; b < NB; b++) fr[b] = Pn[(4 * s + tg) * LD + 8 * b + g];
                int t = 0;
#pragma unroll
                for (int bi = 0; bi < NB; bi++)
#pragma unroll
                    for (int bj = bi; bj < NB; bj++) {
                        dmma_m8n8k4(acc[t][0], acc[t][1], fr[bi], fr[bj]);
                        t++;
                    }
            }
            __syncwarp();
        }

        // ---- load vector: ef(i) += weight*fScale*phi(i)*force (TPZMatPoisson.cpp:39-40) -----------
        if (lane < N) {
            double f = 0.0;
            for (int q = 0; q < nq; q++) {
                const double fq = p.force ? p.force[el * nq + q] : p.coef[1];
                f += JI[9 * QS + q] * p.coef[0] * __ldg(p.phi_pad + (size_t)q * NP + lane) * fq;
            }
            scatter_add(p.rhs + p.dest[el * N + lane], f, p.atomic);
        }
        // ---- scatter-add of the upper triangle ---------------------------------------------------
        if (p.rhs_only) continue;
        const double s = p.coef[0];
        double val[NTILES * 2];
#pragma unroll
        for (int k = 0; k < NTILES * 2; k++) val[k] = s * acc[k >> 1][k & 1];
        if (p.smapT) {
            const int32_t *smT = p.smapT + (size_t)el * C::SLOTS + lane;
            int32_t posT[NTILES * 2];
#pragma unroll
            for (int k = 0; k < NTILES * 2; k++) posT[k] = __ldcs(smT + k * 32);
            scatter_many<NTILES * 2>(p.a, posT, val, p.atomic);
        }
        scatter_many<NTILES * 2>(p.a, pos, val, p.atomic);
    }
}

// scatter map of the DMMA kernel: entry (el, tile t, e, lane) -> CSR position of
// (row 8bi+(lane>>2), col 8bj+2(lane&3)+e) of the element matrix, -1 for padding / lower triangle
template <class C>
__global__ void build_mma_smap_kernel(int64_t nel, const int32_t *__restrict__ dest, const int64_t *__restrict__ ia,
                                      const int32_t *__restrict__ ja, int symmetric, int32_t *__restrict__ smap,
                                      int32_t *__restrict__ smapT, int *__restrict__ missing) {
    constexpr int N = C::N, NB = C::NB, SLOTS = C::SLOTS;
    const int64_t total = nel * SLOTS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx / SLOTS;
        const int slot = (int)(idx - el * SLOTS);
        const int lane = slot & 31, e = (slot >> 5) & 1;
        int t = slot >> 6;
        int bi = 0;
        for (int row = 0; row < NB - 1; row++) {
            const int cnt = NB - row;
            if (bi == row && t >= cnt) {
                t -= cnt;
                bi = row + 1;
            }
        }
        const int bj = bi + t;
        const int i = 8 * bi + (lane >> 2), j = 8 * bj + 2 * (lane & 3) + e;
        int32_t pos = -1, posT = -1;
        if (i < N && j < N && i <= j) {
            const int64_t di = dest[el * N + i], dj = dest[el * N + j];
            auto find = [&](int64_t row, int64_t col) -> int32_t {
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
