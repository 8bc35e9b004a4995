// sumfact_hex.cuh — TPZMatPoisson on general (trilinear) hexahedra of order 2 by SUM FACTORISATION.
//
// The reference forms ek(i,j) = s * sum_q w|detJ| grad_x phi_i . grad_x phi_j point by point
// (Material/Poisson/TPZMatPoisson.cpp:31-38 inside the quadrature loop of Mesh/pzinterpolationspace.cpp:404-473): 27 points x
// 27 x 27 pairs.  On a hexahedron of uniform order <= 2 every shape function is a product of one-dimensional functions
// (Shape/pzshapecube.cpp:36-152 after the higher-side corrections collapse: b_0 = (1-t)/2, b_1 = (1+t)/2, b_2 = 4 b_0 b_1) and the
// rule is a tensor rule (Integral/pzquad.cpp:268-284, point q = q1 + 3 (q2 + 3 q3)), so with
//     M^{ef}(q) = s w_q |detJ_q| sum_v jacinv_q(e,v) jacinv_q(f,v)            (3 x 3, symmetric, per point)
//     ek(i,j)  = sum_{e,f} sum_{q3} F3^{ef}(i3,j3,q3) sum_{q2} F2^{ef}(i2,j2,q2) sum_{q1} F1^{ef}(i1,j1,q1) M^{ef}(q1,q2,q3)
// where Fd^{ef}(a,b,k) = X(a,k) Y(b,k), X = D (derivative table) if e == d else B (value table), Y likewise with f.
// Per (e,f): 54*3 + 54*9 + 54*27 multiply-adds instead of 27*378: 21 k DFMA per element (geometry included) against ~55 k
// DMMA-equivalent of the Gram kernel, on a GPU whose FP64 pipe is the bound of this configuration.
//
// Mapping: one CTA of 64 threads per element (persistent grid), 54 work items (p1, p2):
//   p1 = unordered pair (i1 <= j1) of one-dimensional functions in direction 1 (6), p2 = ordered pair (i2, j2) (9);
//   a work item accumulates the nine entries (i3, j3) in registers.  Restricting direction 1 to i1 <= j1 computes 486 of the
//   729 entries; the 378 of the upper triangle are among them exactly once when, for i1 == j1, only the canonical order
//   (side index i <= j) is kept (the scatter map carries -1 for the rest).
//   stage 1  thread (p1, q2q3):  S1[p1][q2q3] = sum_q1 F1[p1][q1] M(q1, q2q3)         F1 of the thread in registers, M in smem
//   stage 2  thread (p1, p2):    s2[q3]       = sum_q2 F2[p2][q2] S1[p1][q2 + 3 q3]   F2 of the thread in registers
//   stage 3                      acc[i3j3]   += s2[q3] * F3[i3j3][q3]                 F3 uniform: __constant__ operand of the DFMA
//   S1 is double buffered over the nine (e,f): one __syncthreads per (e,f).
// The phases are written as functions of the thread index so that tools/sumfact_emu.cpp runs the identical arithmetic on the
// CPU (thread loops instead of threads) against the reference's element matrices.
#pragma once

#ifndef SF_HD
#define SF_HD __host__ __device__ __forceinline__
#endif

namespace sf {

constexpr int NITEM = 54, NTHREADS = 64, SLOTS = 9 * NTHREADS;  // scatter-map entries per element
// aux table of the group (doubles): F1[4][6][3] (variant, unordered pair, point), F2[4][9][3] (variant, ordered pair, point)
constexpr int AUX_F1 = 0, AUX_F2 = 4 * 6 * 3, AUX_LEN = AUX_F2 + 4 * 9 * 3;

// unordered pairs (a <= b) of the one-dimensional functions
SF_HD void pair6(int p, int &a, int &b) {
    a = p < 3 ? 0 : (p < 5 ? 1 : 2);
    b = p < 3 ? p : (p < 5 ? p - 2 : 2);
}
// side index (shape-function number) of the tensor index (a0, a1, a2): inverse of the table kHexSide of host_tables.cpp
// (Topology/tpzcube.cpp:30-80: vertices, edges, faces, interior; 0 -> b_0, 1 -> b_1, 2 -> bubble)
SF_HD int side_index(int a0, int a1, int a2) {
    constexpr signed char inv[27] = {0, 1, 8, 3, 2, 10, 11, 9, 20, 4, 5, 16, 7, 6, 18, 19, 17, 25, 12, 13, 21, 15, 14, 23, 24, 22, 26};
    return inv[a0 + 3 * a1 + 9 * a2];
}
// variant of direction d for the pair (e, f): 0 B*B, 1 B*D, 2 D*B, 3 D*D
SF_HD constexpr int variant(int d, int e, int f) { return (e == d ? 2 : 0) + (f == d ? 1 : 0); }
// index of M^{ef} in the symmetric storage (00, 01, 02, 11, 12, 22)
SF_HD constexpr int msym(int e, int f) { return e <= f ? (e == 0 ? f : (e == 1 ? 2 + f : 5)) : (f == 0 ? e : (f == 1 ? 2 + e : 5)); }

// host: the one-dimensional tables of the group's rule and the factor tables.  x[3] = the three points of the line rule in the
// reference's order (Integral/tpzgaussrule.cpp:231-238).  F3 is [4][9][3] (variant, ordered pair, point).
inline void build_tables(const double x[3], double *aux, double *F3) {
    double B[3][3], D[3][3];
    for (int k = 0; k < 3; k++) {
        const double l0 = (1. - x[k]) / 2., l1 = (1. + x[k]) / 2.;
        B[0][k] = l0; B[1][k] = l1; B[2][k] = 4. * l0 * l1;
        D[0][k] = -0.5; D[1][k] = 0.5; D[2][k] = 4. * (-0.5 * l1 + l0 * 0.5);
    }
    for (int v = 0; v < 4; v++) {
        const double(*X)[3] = (v & 2) ? D : B;
        const double(*Y)[3] = (v & 1) ? D : B;
        for (int p = 0; p < 6; p++) {
            int a = 0, b = 0;
            a = p < 3 ? 0 : (p < 5 ? 1 : 2);
            b = p < 3 ? p : (p < 5 ? p - 2 : 2);
            // the unordered pair stands for (a,b) with a <= b: K(i,j) with i1 = a, j1 = b
            for (int k = 0; k < 3; k++) aux[AUX_F1 + (v * 6 + p) * 3 + k] = X[a][k] * Y[b][k];
        }
        for (int p = 0; p < 9; p++)
            for (int k = 0; k < 3; k++) {
                aux[AUX_F2 + (v * 9 + p) * 3 + k] = X[p / 3][k] * Y[p % 3][k];
                F3[(v * 9 + p) * 3 + k] = X[p / 3][k] * Y[p % 3][k];
            }
    }
}

// ---- phases (thread index t) -------------------------------------------------------------------------------------------
// geometry at point q = t < 27: Msm[6][27] = s w|detJ| jacinv jacinv^T (symmetric part), Wd[27] = w|detJ|
// X: corner coordinates [8][3]; gradients of the corner functions: dN_a/dxi_k at point q = dn[(k * 8 + a) * ds] (dn = dng + 24 q,
// ds = 1 for the table [27][3][8]; dn = dng_t + q, ds = 27 for the transposed table [3][8][27], whose reads are contiguous over
// the lanes q of a warp: 2 wavefronts per load instead of 27); wq = weight of the point.
// Layout of M: Msm[m * MS + MT * (q / 3) + q % 3] - (MS, MT) = (27, 3) dense; (36, 4) pads the q1-triples to 32 bytes so that
// stage 1 can read a triple with one 16-byte + one 8-byte load: measured SLOWER (259 against 287 M elements/s,
// profiles/r02_sumfact_warp_variants.jsonl: a 16-byte shared load of nine distinct addresses costs more wavefronts than two
// 8-byte loads), so the kernels use the dense layout
template <int MS = 27, int MT = 3>
SF_HD void geometry(int q, const double *X, const double *dn, int ds, double wq, double scale, double *Msm, double *Wd) {
    double j00 = 0, j01 = 0, j02 = 0, j10 = 0, j11 = 0, j12 = 0, j20 = 0, j21 = 0, j22 = 0;
#pragma unroll
    for (int a = 0; a < 8; a++) {  // gradx(j,k) += x_a[j] * dN_a/dxi_k   (Geom/TPZGeoCube.h:141-149)
        const double d0 = dn[a * ds], d1 = dn[(8 + a) * ds], d2 = dn[(16 + a) * ds];
        const double x = X[a * 3], y = X[a * 3 + 1], z = X[a * 3 + 2];
        j00 += x * d0; j01 += x * d1; j02 += x * d2;
        j10 += y * d0; j11 += y * d1; j12 += y * d2;
        j20 += z * d0; j21 += z * d1; j22 += z * d2;
    }
    double det = 0.0;  // Mesh/pzgeoel.cpp:1309-1336
    det -= j02 * j11 * j20;
    det += j01 * j12 * j20;
    det += j02 * j10 * j21;
    det -= j00 * j12 * j21;
    det -= j01 * j10 * j22;
    det += j00 * j11 * j22;
    if (fabs(det) < 1.e-12) det = 1.e-12;
    const double id = 1.0 / det;
    double ji[9];  // jacinv(e,v) = ji[3e+v]
    ji[0] = (-j12 * j21 + j11 * j22) * id;
    ji[1] = (j02 * j21 - j01 * j22) * id;
    ji[2] = (-j02 * j11 + j01 * j12) * id;
    ji[3] = (j12 * j20 - j10 * j22) * id;
    ji[4] = (-j02 * j20 + j00 * j22) * id;
    ji[5] = (j02 * j10 - j00 * j12) * id;
    ji[6] = (-j11 * j20 + j10 * j21) * id;
    ji[7] = (j01 * j20 - j00 * j21) * id;
    ji[8] = (-j01 * j10 + j00 * j11) * id;
    const double w = wq * fabs(det);  // weight *= fabs(detjac)  (pzinterpolationspace.cpp:468)
    Wd[q] = w;
    const double sw = scale * w;
    int m = 0;
#pragma unroll
    for (int e = 0; e < 3; e++)
#pragma unroll
        for (int f = e; f < 3; f++) {
            Msm[m * MS + MT * (q / 3) + q % 3] = sw * (ji[3 * e] * ji[3 * f] + ji[3 * e + 1] * ji[3 * f + 1] + ji[3 * e + 2] * ji[3 * f + 2]);
            m++;
        }
}

// stage 1 of the pair (e,f): thread t = (p1, q2q3); F1v = the thread's F1[variant(0,e,f)][3].  S1buf[GS * (t / 9) + t % 9]:
// GS = 9 dense, GS = 10 the nine values of a p1 group start on a 16-byte boundary (stage 2 then reads them with 16-byte loads:
// measured to cost 250 bytes of spills in the one-warp kernel - aligned register quads -, so GS = 9 is what the kernels use)
template <int MS = 27, int MT = 3, int GS = 9>
SF_HD void stage1(int t, int e, int f, const double *F1v, const double *Msm, double *S1buf) {
    const double *Mq = Msm + msym(e, f) * MS + MT * (t % 9);
    double m0, m1, m2;
    if constexpr (MT == 4) {
        const double2 m01 = *reinterpret_cast<const double2 *>(Mq);
        m0 = m01.x; m1 = m01.y; m2 = Mq[2];
    } else {
        m0 = Mq[0]; m1 = Mq[1]; m2 = Mq[2];
    }
    S1buf[GS * (t / 9) + t % 9] = F1v[0] * m0 + F1v[1] * m1 + F1v[2] * m2;
}

}  // namespace sf

#ifdef __CUDACC__
__constant__ double c_sfF3[4 * 9 * 3];  // [variant][ordered pair i3 j3][q3]: uniform operand of the stage-3 DFMAs
#define SF_F3(idx) c_sfF3[idx]
#else
extern double h_sfF3[4 * 9 * 3];        // (CPU emulation, tools/sumfact_emu.cpp)
#define SF_F3(idx) h_sfF3[idx]
#endif

namespace sf {
// stages 2 and 3 of the pair (e,f) for work item t = (p1, p2): F2v = the thread's F2[variant(1,e,f)][3]
template <int GS = 9>
SF_HD void stage23(int t, int e, int f, const double *F2v, const double *S1buf, double *acc) {
    const double *s1 = S1buf + GS * (t / 9);  // [q2 + 3 q3]
    const int v3 = variant(2, e, f);
#pragma unroll
    for (int q3 = 0; q3 < 3; q3++) {
        double a, b, c;
        if constexpr (GS == 10) {  // 16-byte loads where the alignment allows: (0,1) 2 | 3 (4,5) | (6,7) 8
            if (q3 == 0) {
                const double2 x = *reinterpret_cast<const double2 *>(s1);
                a = x.x; b = x.y; c = s1[2];
            } else if (q3 == 1) {
                const double2 x = *reinterpret_cast<const double2 *>(s1 + 4);
                a = s1[3]; b = x.x; c = x.y;
            } else {
                const double2 x = *reinterpret_cast<const double2 *>(s1 + 6);
                a = x.x; b = x.y; c = s1[8];
            }
        } else {
            a = s1[3 * q3]; b = s1[3 * q3 + 1]; c = s1[3 * q3 + 2];
        }
        const double s2 = F2v[0] * a + F2v[1] * b + F2v[2] * c;
#pragma unroll
        for (int k = 0; k < 9; k++) acc[k] = fma(s2, SF_F3((v3 * 9 + k) * 3 + q3), acc[k]);
    }
}
// the entry (i, j) of the element matrix that work item t holds in acc[k]; false: idle thread or second copy
SF_HD bool entry_of(int t, int k, int &i, int &j) {
    if (t >= NITEM) return false;
    int i1, j1;
    pair6(t / 9, i1, j1);
    const int p2 = t % 9;
    i = side_index(i1, p2 / 3, k / 3);
    j = side_index(j1, p2 % 3, k % 3);
    return i1 < j1 || i <= j;
}
}  // namespace sf

#ifdef __CUDACC__

// The round-2 default until the one-warp-per-element kernel below replaced it (kept as variant 13: profiles refer to it).
// PREFETCH = 1: the corner coordinates of the CTA's next element are loaded (and the node ids of the one after) while the current
// element is computed, and the scatter positions are loaded before the factor stages instead of after them: the dependent global
// loads (ids -> coordinates, ~2 DRAM latencies) and the position loads leave the critical path of every element
template <int MINB, int PREFETCH = 0>
__global__ void __launch_bounds__(sf::NTHREADS, MINB) assemble_sumfact_hex_p2_poisson_kernel(const VolParams p) {
    __shared__ double Xs[24];
    __shared__ double Msm[6 * 27];
    __shared__ double Wd[27];
    __shared__ double S1[2][sf::NITEM + 2];
    __shared__ double Dn[24 * 27];  // dng_t [3][8][27]: the geometry phase reads it with the point index on the lanes
    const int t = threadIdx.x;
    for (int i = t; i < 24 * 27; i += sf::NTHREADS) Dn[i] = __ldg(p.dng_t + i);
    const bool active = t < sf::NITEM;
    const int p1 = active ? t / 9 : 0, p2 = active ? t % 9 : 0;
    double F1[4][3], F2[4][3];  // the thread's rows of the factor tables, all four variants
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            F1[v][k] = __ldg(p.aux2 + sf::AUX_F1 + (v * 6 + p1) * 3 + k);
            F2[v][k] = __ldg(p.aux2 + sf::AUX_F2 + (v * 9 + p2) * 3 + k);
        }
    const int64_t stride = gridDim.x;
    double cnext = 0.0;
    int32_t node_next = 0;
    if (PREFETCH && t < 24 && (int64_t)blockIdx.x < p.nel) {
        cnext = p.xyz[(int64_t)p.elnodes[(int64_t)blockIdx.x * 8 + t / 3] * 3 + t % 3];
        if (blockIdx.x + stride < p.nel) node_next = p.elnodes[(blockIdx.x + stride) * 8 + t / 3];
    }
    for (int64_t el = blockIdx.x; el < p.nel; el += stride) {
        __syncthreads();  // the previous element no longer reads Xs / Msm / Wd / S1
        if (!p.rhs_only) {
            const char *base = (const char *)(p.smap + (size_t)el * sf::SLOTS);
            if (t < (sf::SLOTS * 4 + 127) / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + t * 128));
        }
        if (t < 24) {
            if (PREFETCH) {
                Xs[t] = cnext;
                if (el + stride < p.nel) {
                    cnext = p.xyz[(int64_t)node_next * 3 + t % 3];
                    if (el + 2 * stride < p.nel) node_next = p.elnodes[(el + 2 * stride) * 8 + t / 3];
                }
            } else {
                Xs[t] = p.xyz[(int64_t)p.elnodes[el * 8 + t / 3] * 3 + t % 3];
            }
        }
        __syncthreads();
        if (t < 27) sf::geometry(t, Xs, Dn + t, 27, __ldg(p.qw + t), p.coef[0], Msm, Wd);
        __syncthreads();
        // ---- load vector: ef(i) += weight*fScale*phi(i)*force (TPZMatPoisson.cpp:39-40), threads 32..58 (the second warp)
        if (t >= 32 && t < 32 + 27) {
            const int i = t - 32;
            double f = 0.0;
            for (int q = 0; q < 27; q++) {
                const double a = Wd[q] * __ldg(p.phi + (size_t)q * 27 + i);
                f += p.force ? a * p.force[el * 27 + q] : a;
            }
            f *= p.coef[0] * (p.force ? 1.0 : p.coef[1]);
            scatter_rhs(p.rhs, p.dest[el * 27 + i], f, p.atomic);
        }
        if (p.rhs_only) continue;
        const int32_t *sm = p.smap + (size_t)el * sf::SLOTS + t;
        int32_t pos[9];
        if (PREFETCH) {
#pragma unroll
            for (int k = 0; k < 9; k++) pos[k] = __ldcs(sm + k * sf::NTHREADS);
        }
        double acc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) acc[k] = 0.0;
#pragma unroll
        for (int c = 0; c < 9; c++) {
            const int e = c / 3, f = c % 3;
            if (active) sf::stage1(t, e, f, F1[sf::variant(0, e, f)], Msm, S1[c & 1]);
            __syncthreads();
            if (active) sf::stage23(t, e, f, F2[sf::variant(1, e, f)], S1[c & 1], acc);
        }
        // ---- scatter-add: entry k = (i3, j3) of work item t
        if (!PREFETCH) {
#pragma unroll
            for (int k = 0; k < 9; k++) pos[k] = __ldcs(sm + k * sf::NTHREADS);
        }
        scatter_many<9>(p.a, pos, acc, p.atomic);
        if (p.smapT) {
            const int32_t *smT = p.smapT + (size_t)el * sf::SLOTS + t;
#pragma unroll
            for (int k = 0; k < 9; k++) pos[k] = __ldcs(smT + k * sf::NTHREADS);
            scatter_many<9>(p.a, pos, acc, p.atomic);
        }
    }
}

// ---- one warp per element ---------------------------------------------------------------------------------------------
// The same three stages with NO block-wide barrier: a warp owns an element and walks its 54 work items in two passes of 27
// (pass ps: p1 = 3 ps + lane / 9, p2 = lane % 9, i.e. item 27 ps + lane; the nine S1 values an item needs are produced by
// the nine lanes of its own p1 group, so __syncwarp orders stage 1 against stage 2).  Stage 1 runs for all nine (e,f) first
// (S1[9][28] per warp), then stages 2 + 3: two __syncwarp per pass instead of nine __syncthreads per element, and the warps
// of a CTA never wait for each other (the CTA only shares the read-only tables).  Scatter map: [el][ps][k][32 lanes].
namespace sfw {
constexpr int SLOTS = 2 * 9 * 32;
}

template <int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB) assemble_sumfact_hex_p2_poisson_warp_kernel(const VolParams p) {
    __shared__ double Dn[24 * 27];       // dng_t [3][8][27]
    __shared__ double Ft[sf::AUX_LEN];   // F1[4][6][3], F2[4][9][3]
    __shared__ double Xs[WPC][24];
    __shared__ double Msm[WPC][6 * 27];
    __shared__ double Wd[WPC][28];
    __shared__ double S1[WPC][9][28];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 24 * 27; i += WPC * 32) Dn[i] = __ldg(p.dng_t + i);
    for (int i = threadIdx.x; i < sf::AUX_LEN; i += WPC * 32) Ft[i] = __ldg(p.aux2 + i);
    __syncthreads();
    const bool active = lane < 27;
    const int l = active ? lane : 0;
    const int g1 = l / 9, p2 = l % 9;
    double F2[4][3];
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
        for (int k = 0; k < 3; k++) F2[v][k] = Ft[sf::AUX_F2 + (v * 9 + p2) * 3 + k];
    const double myqw = active ? __ldg(p.qw + l) : 0.0;
    const int64_t nwarps = (int64_t)gridDim.x * WPC;
    int64_t el = (int64_t)blockIdx.x * WPC + warp;
    double cnext = 0.0;
    int32_t node_next = 0;
    if (lane < 24 && el < p.nel) {
        cnext = p.xyz[(int64_t)p.elnodes[el * 8 + lane / 3] * 3 + lane % 3];
        if (el + nwarps < p.nel) node_next = p.elnodes[(el + nwarps) * 8 + lane / 3];
    }
    for (; el < p.nel; el += nwarps) {
        if (!p.rhs_only && lane < (sfw::SLOTS * 4 + 127) / 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)(p.smap + (size_t)el * sfw::SLOTS) + lane * 128));
        if (lane < 24) {
            Xs[warp][lane] = cnext;
            if (el + nwarps < p.nel) {
                cnext = p.xyz[(int64_t)node_next * 3 + lane % 3];
                if (el + 2 * nwarps < p.nel) node_next = p.elnodes[(el + 2 * nwarps) * 8 + lane / 3];
            }
        }
        __syncwarp();
        if (active) sf::geometry(lane, Xs[warp], Dn + lane, 27, myqw, p.coef[0], Msm[warp], Wd[warp]);
        __syncwarp();
        // ---- load vector: ef(i) += weight*fScale*phi(i)*force (TPZMatPoisson.cpp:39-40), lane <-> i
        if (active) {
            double f = 0.0;
            for (int q = 0; q < 27; q++) {
                const double a = Wd[warp][q] * __ldg(p.phi + (size_t)q * 27 + lane);
                f += p.force ? a * p.force[el * 27 + q] : a;
            }
            f *= p.coef[0] * (p.force ? 1.0 : p.coef[1]);
            scatter_rhs(p.rhs, p.dest[el * 27 + lane], f, p.atomic);
        }
        if (p.rhs_only) {
            __syncwarp();
            continue;
        }
#pragma unroll
        for (int ps = 0; ps < 2; ps++) {
            const int32_t *sm = p.smap + (size_t)el * sfw::SLOTS + ps * 9 * 32 + lane;
            int32_t pos[9];
#pragma unroll
            for (int k = 0; k < 9; k++) pos[k] = __ldcs(sm + k * 32);
            {
                double F1[4][3];
                const int p1 = 3 * ps + g1;
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int k = 0; k < 3; k++) F1[v][k] = Ft[sf::AUX_F1 + (v * 6 + p1) * 3 + k];
#pragma unroll
                for (int c = 0; c < 9; c++)
                    if (active) sf::stage1(l, c / 3, c % 3, F1[sf::variant(0, c / 3, c % 3)], Msm[warp], S1[warp][c]);
            }
            __syncwarp();
            double acc[9];
#pragma unroll
            for (int k = 0; k < 9; k++) acc[k] = 0.0;
#pragma unroll
            for (int c = 0; c < 9; c++)
                if (active) sf::stage23<9>(l, c / 3, c % 3, F2[sf::variant(1, c / 3, c % 3)], S1[warp][c], acc);
            scatter_many<9>(p.a, pos, acc, p.atomic);
            if (p.smapT) {
                const int32_t *smT = p.smapT + (size_t)el * sfw::SLOTS + ps * 9 * 32 + lane;
#pragma unroll
                for (int k = 0; k < 9; k++) pos[k] = __ldcs(smT + k * 32);
                scatter_many<9>(p.a, pos, acc, p.atomic);
            }
            __syncwarp();  // S1 is rewritten by the next pass / Xs, Msm, Wd by the next element
        }
    }
}

// scatter map: entry (el, k = i3*3+j3, t = p1*9+p2) -> CSR position of ek(i, j), i = (i1,i2,i3), j = (j1,j2,j3), (i1,j1) the
// unordered pair p1 (i1 <= j1), (i2,j2) = (p2/3, p2%3); -1 for idle threads and for the second copy of an entry (i1 == j1, i > j)
// WARP = 1: the layout of the one-warp-per-element kernel, slot (ps * 9 + k) * 32 + lane <-> work item 27 ps + lane (lane < 27)
template <int WARP>
__global__ void build_sumfact_smap_kernel(int64_t nel, const int32_t *__restrict__ dest, const int64_t *__restrict__ ia,
                                          const int32_t *__restrict__ ja, int symmetric, int32_t *__restrict__ smap,
                                          int32_t *__restrict__ smapT, int *__restrict__ missing) {
    static_assert(sf::SLOTS == sfw::SLOTS, "both layouts have 576 slots per element");
    const int64_t total = nel * sf::SLOTS;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx / sf::SLOTS;
        const int slot = (int)(idx - el * sf::SLOTS);
        int t, k;
        if (WARP) {
            const int lane = slot % 32, pk = slot / 32;
            k = pk % 9;
            t = lane < 27 ? 27 * (pk / 9) + lane : sf::NITEM;  // (NITEM: idle lane)
        } else {
            t = slot % sf::NTHREADS;
            k = slot / sf::NTHREADS;
        }
        int32_t pos = -1, posT = -1;
        int i, j;
        if (sf::entry_of(t, k, i, j)) {
            {
                const int64_t di = dest[el * 27 + i], dj = dest[el * 27 + j];
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
#endif  // __CUDACC__
