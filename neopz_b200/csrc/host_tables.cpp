// Host-side tables of the B200 assembly engine: quadrature rules and H1 shape tables.
//
// These are what the flattener uploads once per element batch.  For uniform order p<=2 no side
// of a NeoPZ H1 element carries more than one shape function, so the tables are independent of
// the element's node ids (Shape/TPZShapeH1.cpp:71,77 skip the orientation transforms) and the
// hierarchical basis has a closed form:
//   hex / quad : tensor product of the 1-D set { l0=(1-t)/2, l1=(1+t)/2, b=4*l0*l1 }
//                (vertex + "blend" functions of Shape/pzshapecube.cpp:36-152, pzshapequad.cpp:36-93
//                 after the higher-side correction :131-143 collapses)
//   tet / tri  : barycentric l_a and 4*l_a*l_b on the edges (Shape/pzshapetetra.cpp:53-164,
//                pzshapetriang.cpp:34-81)
// Shape order = side order of the reference topology (Topology/tpzcube.cpp:30-80 etc.), which is
// also the connect order and the local dof order of TPZElementMatrix.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/b200asm.h"

namespace {

// 1-D factor index per direction for every hexahedron side: 0 -> l0, 1 -> l1, 2 -> b
const int kHexSide[27][3] = {
    {0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1},  // vertices
    {2, 0, 0}, {1, 2, 0}, {2, 1, 0}, {0, 2, 0},                                              // edges 8-11 (z=-1)
    {0, 0, 2}, {1, 0, 2}, {1, 1, 2}, {0, 1, 2},                                              // edges 12-15 (vertical)
    {2, 0, 1}, {1, 2, 1}, {2, 1, 1}, {0, 2, 1},                                              // edges 16-19 (z=+1)
    {2, 2, 0}, {2, 0, 2}, {1, 2, 2}, {2, 1, 2}, {0, 2, 2}, {2, 2, 1},                        // faces 20-25
    {2, 2, 2}};                                                                              // interior
const int kQuadSide[9][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}, {2, 0}, {1, 2}, {2, 1}, {0, 2}, {2, 2}};
const int kTetEdge[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
const int kTriEdge[3][2] = {{0, 1}, {1, 2}, {2, 0}};

inline void factors1d(double t, double f[3], double df[3]) {
    f[0] = (1. - t) / 2.;
    f[1] = (1. + t) / 2.;
    f[2] = 4. * f[0] * f[1];
    df[0] = -0.5;
    df[1] = 0.5;
    df[2] = 4. * (df[0] * f[1] + f[0] * df[1]);
}

}  // namespace

// Gauss-Legendre points by Newton iteration on P_n in extended precision, stored as the reference
// stores them (Integral/tpzgaussrule.cpp:231-238: symmetric pairs (-z,+z), pair after pair).  Kept
// in long double because the reference multiplies the 1-D weights in extended precision before
// rounding (Integral/pzquad.cpp:167,283 with long double W()).
static int gauss_legendre_ld(int order, long double *loc, long double *w) {
    int n = (int)(0.51 * (order + 2));
    if (n < 1) n = 1;
    if (n > 64) return -1;
    const long double pi = 3.14159265358979323846264338327950288L;
    const int m = (n + 1) / 2;
    for (int i = 0; i < m; i++) {
        long double z = cosl(pi * ((long double)i + 0.75L) / ((long double)n + 0.5L));
        long double dp = 1.0L, pk = 1.0L, pkm1 = 0.0L;
        for (int it = 0; it < 200; it++) {
            pk = 1.0L;
            pkm1 = 0.0L;
            for (int j = 0; j < n; j++) {
                const long double pkm2 = pkm1;
                pkm1 = pk;
                pk = ((2.0L * j + 1.0L) * z * pkm1 - (long double)j * pkm2) / ((long double)j + 1.0L);
            }
            dp = (long double)n * (z * pk - pkm1) / (z * z - 1.0L);
            const long double zprev = z;
            z = zprev - pk / dp;
            if (fabsl(z - zprev) <= 1.0842021724855044e-19L) break;
        }
        const long double wt = 2.0L / ((1.0L - z * z) * dp * dp);
        loc[2 * i] = -z;
        w[2 * i] = wt;
        if (2 * i + 1 < n) {
            loc[2 * i + 1] = z;
            w[2 * i + 1] = wt;
        }
    }
    return n;
}

extern "C" int b200asm_gauss_legendre(int order, double *loc, double *w) {
    long double l[64], ww[64];
    const int n = gauss_legendre_ld(order, l, ww);
    if (n < 0) return B200ASM_EINVAL;
    for (int i = 0; i < n; i++) {
        loc[i] = (double)l[i];
        w[i] = (double)ww[i];
    }
    return n;
}

extern "C" int b200asm_tensor_rule(int topology, int order, double *qpts, double *qw) {
    long double l[64], w[64];
    const int n = gauss_legendre_ld(order, l, w);
    if (n < 0) return B200ASM_EINVAL;
    if (topology == B200ASM_HEX) {
        // point index = ik + n*(ie + n*iz): ksi fastest
        for (int iz = 0; iz < n; iz++)
            for (int ie = 0; ie < n; ie++)
                for (int ik = 0; ik < n; ik++) {
                    const int ip = ik + n * (ie + n * iz);
                    qpts[3 * ip + 0] = (double)l[ik];
                    qpts[3 * ip + 1] = (double)l[ie];
                    qpts[3 * ip + 2] = (double)l[iz];
                    qw[ip] = (double)(w[ik] * w[ie] * w[iz]);
                }
        return n * n * n;
    }
    if (topology == B200ASM_QUAD) {
        // point index = ie + n*ik: ksi slowest
        for (int ik = 0; ik < n; ik++)
            for (int ie = 0; ie < n; ie++) {
                const int ip = ie + n * ik;
                qpts[2 * ip + 0] = (double)l[ik];
                qpts[2 * ip + 1] = (double)l[ie];
                qw[ip] = (double)(w[ik] * w[ie]);
            }
        return n * n;
    }
    return B200ASM_EINVAL;
}

extern "C" int b200asm_shape_tables(int topology, int porder, int nqp, const double *qpts, double *phi, double *dphi) {
    if (porder < 1 || porder > 2) return B200ASM_EINVAL;
    int n = 0;
    switch (topology) {
        case B200ASM_HEX: n = porder == 1 ? 8 : 27; break;
        case B200ASM_QUAD: n = porder == 1 ? 4 : 9; break;
        case B200ASM_TET: n = porder == 1 ? 4 : 10; break;
        case B200ASM_TRI: n = porder == 1 ? 3 : 6; break;
        default: return B200ASM_EINVAL;
    }
    const int dim = (topology == B200ASM_HEX || topology == B200ASM_TET) ? 3 : 2;
    for (int q = 0; q < nqp; q++) {
        const double *pt = qpts + (size_t)q * dim;
        double *ph = phi + (size_t)q * n;
        double *dp = dphi + (size_t)q * dim * n;  // [d][i]
        if (topology == B200ASM_HEX || topology == B200ASM_QUAD) {
            double f[3][3], df[3][3];
            for (int d = 0; d < dim; d++) factors1d(pt[d], f[d], df[d]);
            for (int s = 0; s < n; s++) {
                if (dim == 3) {
                    const int *a = kHexSide[s];
                    ph[s] = f[0][a[0]] * f[1][a[1]] * f[2][a[2]];
                    dp[0 * n + s] = df[0][a[0]] * f[1][a[1]] * f[2][a[2]];
                    dp[1 * n + s] = f[0][a[0]] * df[1][a[1]] * f[2][a[2]];
                    dp[2 * n + s] = f[0][a[0]] * f[1][a[1]] * df[2][a[2]];
                } else {
                    const int *a = kQuadSide[s];
                    ph[s] = f[0][a[0]] * f[1][a[1]];
                    dp[0 * n + s] = df[0][a[0]] * f[1][a[1]];
                    dp[1 * n + s] = f[0][a[0]] * df[1][a[1]];
                }
            }
        } else {
            const int nc = dim + 1;
            double lam[4], dlam[4][3];
            lam[0] = 1.0;
            for (int d = 0; d < dim; d++) {
                lam[0] -= pt[d];
                lam[d + 1] = pt[d];
            }
            for (int a = 0; a < nc; a++)
                for (int d = 0; d < dim; d++) dlam[a][d] = (a == 0) ? -1.0 : (a == d + 1 ? 1.0 : 0.0);
            for (int a = 0; a < nc; a++) {
                ph[a] = lam[a];
                for (int d = 0; d < dim; d++) dp[d * n + a] = dlam[a][d];
            }
            if (porder == 2) {
                const int ne = (dim == 3) ? 6 : 3;
                for (int e = 0; e < ne; e++) {
                    const int a = dim == 3 ? kTetEdge[e][0] : kTriEdge[e][0];
                    const int b = dim == 3 ? kTetEdge[e][1] : kTriEdge[e][1];
                    ph[nc + e] = 4.0 * (lam[a] * lam[b]);
                    for (int d = 0; d < dim; d++) dp[d * n + nc + e] = 4.0 * (dlam[a][d] * lam[b] + lam[a] * dlam[b][d]);
                }
            }
        }
    }
    return n;
}
