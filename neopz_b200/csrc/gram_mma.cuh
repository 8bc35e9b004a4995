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
    static constexpr int JS = 14;       // doubles per point: sqrt(w|detJ|)*jacinv (9), w|detJ|, padding (even stride: 128-bit
                                        // broadcast loads; 14 -> the stores of a half-warp fall into 2-way conflicts only)
    static constexpr int XSP = NN * 3 + ((NN * 3) & 1);
    static constexpr int SLOTS = NTILES * 2 * 32;  // scatter-map entries per element
    __host__ __device__ static int qstride(int nq) { return ((nq + 31) / 32) * 32; }  // JI is [JS][qstride]
    __host__ __device__ static int warp_doubles(int nq) { int n = XSP + nq * JS + 2 * KC * LD; return n + (n & 1); }  // two panel buffers
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
    double *Pn = JI + nq * JS;
    for (int i = lane; i < 2 * KC * LD; i += 32) Pn[i] = 0.0;  // padding columns stay zero (both panel buffers)
    __syncwarp();
    const int g = lane >> 2, tg = lane & 3;
    const int64_t nwarps = (int64_t)gridDim.x * C::WPC;

    const int64_t el_first = (int64_t)blockIdx.x * C::WPC + warp;
    // software prefetch across elements: corner-node ids two elements ahead, their coordinates one element ahead
    int32_t next_node = 0;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    if (lane < NN && el_first < p.nel) {
        const int64_t node = p.elnodes[el_first * NN + lane];
        cx = p.xyz[node * 3 + 0]; cy = p.xyz[node * 3 + 1]; cz = p.xyz[node * 3 + 2];
        if (el_first + nwarps < p.nel) next_node = p.elnodes[(el_first + nwarps) * NN + lane];
    }
    for (int64_t el = el_first; el < p.nel; el += nwarps) {
        // pull this element's scatter positions towards L2 now; they are loaded into registers only after the last
        // DMMA has been issued (20 registers less during the main loop)
        if (!p.rhs_only) {
            const char *base = (const char *)(p.smap + (size_t)el * C::SLOTS);
            for (int off = lane * 128; off < C::SLOTS * 4; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
        if (lane < NN) {
            Xs[lane * 3 + 0] = cx;
            Xs[lane * 3 + 1] = cy;
            Xs[lane * 3 + 2] = cz;
            if (el + nwarps < p.nel) {
                const int64_t node = next_node;
                cx = p.xyz[node * 3 + 0]; cy = p.xyz[node * 3 + 1]; cz = p.xyz[node * 3 + 2];
                if (el + 2 * nwarps < p.nel) next_node = p.elnodes[(el + 2 * nwarps) * NN + lane];
            }
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
            double *o = JI + q * JS;  // point-major: the nine factors of a point are read back with 128-bit broadcast loads
            o[0] = (-j12 * j21 + j11 * j22) * sid;
            o[1] = (j02 * j21 - j01 * j22) * sid;
            o[2] = (-j02 * j11 + j01 * j12) * sid;
            o[3] = (j12 * j20 - j10 * j22) * sid;
            o[4] = (-j02 * j20 + j00 * j22) * sid;
            o[5] = (j02 * j10 - j00 * j12) * sid;
            o[6] = (-j11 * j20 + j10 * j21) * sid;
            o[7] = (j01 * j20 - j00 * j21) * sid;
            o[8] = (-j01 * j10 + j00 * j11) * sid;
            o[9] = w;
        }
        __syncwarp();

        // ---- load vector: ef(i) += weight*fScale*phi(i)*force (TPZMatPoisson.cpp:39-40); three independent chains
        if (lane < N) {
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            for (int q = 0; q < nq; q += 3) {
                const double a0 = JI[q * JS + 9] * __ldg(p.phi_pad + (size_t)q * NP + lane);
                const double a1 = q + 1 < nq ? JI[(q + 1) * JS + 9] * __ldg(p.phi_pad + (size_t)(q + 1) * NP + lane) : 0.0;
                const double a2 = q + 2 < nq ? JI[(q + 2) * JS + 9] * __ldg(p.phi_pad + (size_t)(q + 2) * NP + lane) : 0.0;
                if (p.force) {
                    f0 += a0 * p.force[el * nq + q];
                    if (q + 1 < nq) f1 += a1 * p.force[el * nq + q + 1];
                    if (q + 2 < nq) f2 += a2 * p.force[el * nq + q + 2];
                } else {
                    f0 += a0; f1 += a1; f2 += a2;
                }
            }
            const double f = (f0 + f1 + f2) * p.coef[0] * (p.force ? 1.0 : p.coef[1]);
            scatter_rhs(p.rhs, p.dest[el * N + lane], f, p.atomic);
        }

        double acc[NTILES][2];
#pragma unroll
        for (int t = 0; t < NTILES; t++) acc[t][0] = acc[t][1] = 0.0;

        // ---- phase 2: panel rows of QC points (Mesh/TPZCompElH1.cpp:147): lane <-> shape function,
        // the table loads of all QC points are issued before the first use
        auto build_panel = [&](int q0, double *Pb) {
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
                    const double2 *ji = reinterpret_cast<const double2 *>(JI + q * JS);  // 5 x LDS.128, same address in all lanes
                    const double2 j01 = ji[0], j23 = ji[1], j45 = ji[2], j67 = ji[3], j8w = ji[4];
                    const double g0 = j01.x * d[ql][0] + j23.y * d[ql][1] + j67.x * d[ql][2];
                    const double g1 = j01.y * d[ql][0] + j45.x * d[ql][1] + j67.y * d[ql][2];
                    const double g2 = j23.x * d[ql][0] + j45.y * d[ql][1] + j8w.x * d[ql][2];
                    double *row = Pb + (3 * ql) * LD + i;
                    row[0] = g0;
                    row[LD] = g1;
                    row[2 * LD] = g2;
                }
            }
        };
        // software pipeline: the panel of chunk c+1 is built into the other buffer before the DMMAs of chunk c are
        // issued, so its table loads / FMAs run while the tensor pipe drains; one __syncwarp per chunk
        const int nq_run = p.rhs_only ? 0 : nq;
        if (nq_run) build_panel(0, Pn);
        __syncwarp();
        int buf = 0;
        for (int q0 = 0; q0 < nq_run; q0 += QC, buf ^= 1) {
            const double *Pc = Pn + buf * (KC * LD);
            if (q0 + QC < nq_run) build_panel(q0 + QC, Pn + (buf ^ 1) * (KC * LD));
            // ---- phase 3: Gram update, 3 k-steps of 4 panel rows -----------------------------------
#pragma unroll
            for (int s = 0; s < KC / 4; s++) {
                double fr[NB];
#pragma unroll
                for (int b = 0; b < NB; b++) fr[b] = Pc[(4 * s + tg) * LD + 8 * b + g];
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

        // ---- scatter-add of the upper triangle ---------------------------------------------------
        if (p.rhs_only) continue;
        // positions in batches of HB entries (loads first: the reds would serialise them); values are scaled on the fly
        constexpr int HB = NTILES;  // two batches
        const double s = p.coef[0];
        const int32_t *sm = p.smap + (size_t)el * C::SLOTS + lane;
        const int32_t *smT = p.smapT ? p.smapT + (size_t)el * C::SLOTS + lane : nullptr;
#pragma unroll
        for (int k0 = 0; k0 < NTILES * 2; k0 += HB) {
            int32_t pos[HB];
            double val[HB];
#pragma unroll
            for (int k = 0; k < HB; k++) pos[k] = __ldcs(sm + (k0 + k) * 32);
#pragma unroll
            for (int k = 0; k < HB; k++) val[k] = s * acc[(k0 + k) >> 1][(k0 + k) & 1];
            scatter_many<HB>(p.a, pos, val, p.atomic);
            if (smT) {
#pragma unroll
                for (int k = 0; k < HB; k++) pos[k] = __ldcs(smT + (k0 + k) * 32);
                scatter_many<HB>(p.a, pos, val, p.atomic);
            }
        }
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
