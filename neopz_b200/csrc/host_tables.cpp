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
//   prism      : (triangle set) x (line set)  (Shape/pzshapeprism.cpp:42-205)
//   pyramid    : rational corner functions and their products (Shape/pzshapepiram.cpp:47-119,331-392)
// Shape order = side order of the reference topology (Topology/tpzcube.cpp:30-80 etc.), which is
// also the connect order and the local dof order of TPZElementMatrix.
//
// Order p >= 3 (hexahedra / quadrilaterals): a side with more than one function multiplies its blend
// function by Chebyshev polynomials of the side's own parametric coordinates (Shape/TPZShapeH1.cpp:64-113,
// pzshapelinear.cpp:15-33, pzshapequad.cpp:317-337, pzshapecube.cpp:460-485).  Those coordinates are the
// element coordinates of the side's free axes, permuted and reflected according to the GLOBAL indices of
// the side's corner nodes (transform ids: Topology/tpzcube.cpp:1059-1111, tpzquadrilateral.cpp:591-618;
// matrices: pzshapequad.cpp:23-32, tpzcube.cpp:583-655), so that neighbouring elements agree on the
// functions of a shared side.  Here the whole orientation of an element is packed into one integer key
// (b200asm_orientation_keys: 1 bit per edge, 3 bits per face); elements with equal keys share their tables
// and are assembled as one group.
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

// corner pairs of the edges and corner cycles of the faces (side order of Topology/tpzcube.cpp:30-80)
const int kHexEdge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}};
const int kHexFace[6][4] = {{0, 1, 2, 3}, {0, 1, 5, 4}, {1, 2, 6, 5}, {3, 2, 6, 7}, {0, 3, 7, 4}, {4, 5, 6, 7}};
// natural parametrisation of a side inside the element: free axis and its sign per side coordinate
// (TransformElementToSide: edges 10,11,18,19 run against their axis)
const int kHexEdgeAxis[12] = {0, 1, 0, 1, 2, 2, 2, 2, 0, 1, 0, 1};
const int kHexEdgeSign[12] = {1, 1, -1, -1, 1, 1, 1, 1, 1, 1, -1, -1};
const int kHexFaceAxes[6][2] = {{0, 1}, {0, 2}, {1, 2}, {0, 2}, {1, 2}, {0, 1}};
const int kQuadEdgeAxis[4] = {0, 1, 0, 1};
const int kQuadEdgeSign[4] = {1, 1, -1, -1};
// the eight symmetries of the square as (source coordinate, sign) per output coordinate (gTrans2dQ)
const int kSquareSrc[8][2] = {{0, 1}, {1, 0}, {1, 0}, {0, 1}, {0, 1}, {1, 0}, {1, 0}, {0, 1}};
const int kSquareSgn[8][2] = {{1, 1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}, {-1, -1}, {-1, 1}, {1, -1}};

// which of the eight square symmetries maps the face onto its canonical orientation: start at the corner with the
// smallest global index, walk towards the smaller of its two neighbours
inline int square_symmetry(int64_t a, int64_t b, int64_t c, int64_t d) {
    const int64_t v[4] = {a, b, c, d};
    int m = 0;
    for (int k = 1; k < 4; k++)
        if (v[k] < v[m]) m = k;
    const bool ccw = v[(m + 1) & 3] < v[(m + 3) & 3];
    return 2 * m + (ccw ? 0 : 1);
}

// triangular sides: corner triples of the tetrahedron faces (Topology/tpztetrahedron.h:280) and the symmetry class
const int kTetFace[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};
inline int triangle_symmetry(int64_t a, int64_t b, int64_t c) {
    const int64_t v[3] = {a, b, c};
    int m = 0;
    for (int k = 1; k < 3; k++)
        if (v[k] < v[m]) m = k;
    const bool ccw = v[(m + 1) % 3] < v[(m + 2) % 3];
    return 2 * m + (ccw ? 0 : 1);
}
// With mu = (1 - u - v, u, v) the barycentric coordinates of a triangular side, symmetry class t maps the side coordinates to
// (u', v') = (mu[kTriPerm[t][0]], mu[kTriPerm[t][1]])  (the six affine maps gTrans2dT / gVet2dT of Shape/pzshapetriang.cpp:18-29)
const int kTriPerm[6][2] = {{1, 2}, {2, 1}, {2, 0}, {0, 2}, {0, 1}, {1, 0}};
// side coordinates (u, v) of the tetrahedron faces as affine functions c0 + cx x + cy y + cz z of the element coordinates
// (Topology/tpztetrahedron.cpp:513-538: faces 10, 11, 13 drop one coordinate, face 12 projects along (1,1,1)/3)
const double kTetFaceUV[4][2][4] = {{{0, 1, 0, 0}, {0, 0, 1, 0}},
                                    {{0, 1, 0, 0}, {0, 0, 0, 1}},
                                    {{1.0 / 3.0, -1.0 / 3.0, 2.0 / 3.0, -1.0 / 3.0}, {1.0 / 3.0, -1.0 / 3.0, -1.0 / 3.0, 2.0 / 3.0}},
                                    {{0, 0, 1, 0}, {0, 0, 0, 1}}};

struct SideParam {  // side coordinate k = sign[k] * x[axis[k]]
    int sdim;
    int axis[3];
    int sign[3];
};

// decode the orientation key of an element into the parametrisation of side `side`
inline SideParam side_param(int topology, int side, int64_t key) {
    SideParam sp;
    if (topology == B200ASM_HEX) {
        if (side < 20) {
            const int e = side - 8;
            sp.sdim = 1;
            sp.axis[0] = kHexEdgeAxis[e];
            sp.sign[0] = kHexEdgeSign[e] * (((key >> e) & 1) ? -1 : 1);
        } else if (side < 26) {
            const int f = side - 20;
            const int t = (int)((key >> (12 + 3 * f)) & 7);
            sp.sdim = 2;
            for (int k = 0; k < 2; k++) {
                sp.axis[k] = kHexFaceAxes[f][kSquareSrc[t][k]];
                sp.sign[k] = kSquareSgn[t][k];
            }
        } else {
            sp.sdim = 3;
            for (int k = 0; k < 3; k++) { sp.axis[k] = k; sp.sign[k] = 1; }
        }
    } else if (topology == B200ASM_LINE) {
        sp.sdim = 1;
        sp.axis[0] = 0;
        sp.sign[0] = (key & 1) ? -1 : 1;
    } else {
        if (side < 8) {
            const int e = side - 4;
            sp.sdim = 1;
            sp.axis[0] = kQuadEdgeAxis[e];
            sp.sign[0] = kQuadEdgeSign[e] * (((key >> e) & 1) ? -1 : 1);
        } else {
            const int t = (int)((key >> 4) & 7);
            sp.sdim = 2;
            for (int k = 0; k < 2; k++) {
                sp.axis[k] = kSquareSrc[t][k];
                sp.sign[k] = kSquareSgn[t][k];
            }
        }
    }
    return sp;
}

// T_0..T_{num-1} and derivatives (three-term recurrence, Shape/pzshapelinear.cpp:15-33)
inline void chebyshev_T(double x, int num, double *t, double *dt) {
    if (num <= 0) return;
    t[0] = 1.0; dt[0] = 0.0;
    if (num == 1) return;
    t[1] = x; dt[1] = 1.0;
    for (int k = 2; k < num; k++) {
        t[k] = 2.0 * x * t[k - 1] - t[k - 2];
        dt[k] = 2.0 * x * dt[k - 1] + 2.0 * t[k - 1] - dt[k - 2];
    }
}

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
    if (topology == B200ASM_LINE) {  // TPZInt1d: the 1-D rule itself
        for (int ik = 0; ik < n; ik++) {
            qpts[ik] = (double)l[ik];
            qw[ik] = (double)w[ik];
        }
        return n;
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

// TPZIntPrism3D (Integral/pzquad.cpp:408-436): point ip = (triangle point ip % ntri, line point ip / ntri), weight = line weight
// (rounded to double) times triangle weight.  The triangle rule is a table of the reference and comes from the caller.
extern "C" int b200asm_prism_rule(int order, int ntri, const double *tripts, const double *triw, double *qpts, double *qw) {
    long double l[64], w[64];
    const int n = gauss_legendre_ld(order, l, w);
    if (n < 0 || ntri <= 0 || !tripts || !triw) return B200ASM_EINVAL;
    for (int iz = 0; iz < n; iz++)
        for (int it = 0; it < ntri; it++) {
            const int ip = iz * ntri + it;
            qpts[3 * ip + 0] = tripts[2 * it];
            qpts[3 * ip + 1] = tripts[2 * it + 1];
            qpts[3 * ip + 2] = (double)l[iz];
            qw[ip] = (double)w[iz] * triw[it];
        }
    return n * ntri;
}

namespace {

// Prism, p <= 2: tensor product of the triangle set { lam_a, 4 lam_a lam_b } with the line set { l0, l1, b = 4 l0 l1 }.
// After the higher-side corrections of Shape/pzshapeprism.cpp:178-193 collapse (l0 + l1 = 1, lam0 + lam1 + lam2 = 1) the
// functions of the sides are: bottom / top edges 4 lam_a lam_b l0 / l1, vertical edges lam_a b, quadrilateral faces
// 4 lam_a lam_b b; the triangular faces and the interior carry none at p = 2 (NConnectShapeF :761-775).
// Side order of Topology/tpzprism.h:275-278.
void prism_shapes(int porder, const double *pt, int n, double *ph, double *dp) {
    const double lam[3] = {1.0 - pt[0] - pt[1], pt[0], pt[1]};
    static const double dlam[3][2] = {{-1.0, -1.0}, {1.0, 0.0}, {0.0, 1.0}};
    double T[6], dT[6][2];  // triangle functions and their (xi, eta) gradients
    for (int a = 0; a < 3; a++) {
        T[a] = lam[a];
        dT[a][0] = dlam[a][0];
        dT[a][1] = dlam[a][1];
    }
    for (int e = 0; e < 3; e++) {
        const int a = kTriEdge[e][0], b = kTriEdge[e][1];
        T[3 + e] = 4.0 * (lam[a] * lam[b]);
        for (int d = 0; d < 2; d++) dT[3 + e][d] = 4.0 * (dlam[a][d] * lam[b] + lam[a] * dlam[b][d]);
    }
    double f[3], df[3];
    factors1d(pt[2], f, df);
    // (triangle function, line factor) of every shape function, side by side
    static const int kTri[18] = {0, 1, 2, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 3, 4, 5};
    static const int kLin[18] = {0, 0, 0, 1, 1, 1, 0, 0, 0, 2, 2, 2, 1, 1, 1, 2, 2, 2};
    (void)porder;
    for (int s = 0; s < n; s++) {
        const int t = kTri[s], l = kLin[s];
        ph[s] = T[t] * f[l];
        dp[0 * n + s] = dT[t][0] * f[l];
        dp[1 * n + s] = dT[t][1] * f[l];
        dp[2 * n + s] = T[t] * df[l];
    }
}

// Pyramid, p <= 2 (Shape/pzshapepiram.cpp:331-392, :47-119): the corner functions are the rational functions
// phi_a = (1 - z -+ x)(1 - z -+ y) / (4 (1 - z)), phi_4 = z; edge functions 4 phi_a phi_b, to which the base edges add the base
// function phi_0 phi_2 before the scaling; base 16 phi_0 phi_2.  Not defined at the apex (never an integration point).
void pyramid_shapes(int porder, const double *pt, int n, double *ph, double *dp) {
    const double x = pt[0], y = pt[1], z = pt[2];
    const double h = 1.0 - z;
    const double sx[4] = {-1.0, 1.0, 1.0, -1.0}, sy[4] = {-1.0, -1.0, 1.0, 1.0};
    double v[14], g[14][3];
    for (int a = 0; a < 4; a++) {
        const double u = h + sx[a] * x, w = h + sy[a] * y;  // (1 - z +- x), (1 - z +- y)
        v[a] = 0.25 * u * w / h;
        g[a][0] = 0.25 * sx[a] * w / h;
        g[a][1] = 0.25 * sy[a] * u / h;
        g[a][2] = -0.25 * (u + w - u * w / h) / h;
    }
    v[4] = z;
    g[4][0] = 0.0; g[4][1] = 0.0; g[4][2] = 1.0;
    if (porder >= 2) {
        auto product = [&](int s, int a, int b) {
            v[s] = v[a] * v[b];
            for (int d = 0; d < 3; d++) g[s][d] = g[a][d] * v[b] + v[a] * g[b][d];
        };
        product(13, 0, 2);
        for (int e = 0; e < 4; e++) {
            product(5 + e, e, (e + 1) & 3);
            v[5 + e] += v[13];
            for (int d = 0; d < 3; d++) g[5 + e][d] += g[13][d];
            product(9 + e, e, 4);
        }
        for (int s = 5; s < 14; s++) {
            const double scale = s < 13 ? 4.0 : 16.0;
            v[s] *= scale;
            for (int d = 0; d < 3; d++) g[s][d] *= scale;
        }
    }
    for (int s = 0; s < n; s++) {
        ph[s] = v[s];
        for (int d = 0; d < 3; d++) dp[d * n + s] = g[s][d];
    }
}

}  // namespace

extern "C" int b200asm_shape_tables(int topology, int porder, int nqp, const double *qpts, double *phi, double *dphi) {
    if (porder < 1 || porder > 2) return B200ASM_EINVAL;
    int n = 0;
    if (topology == B200ASM_PRISM || topology == B200ASM_PYRAMID) {
        n = topology == B200ASM_PRISM ? (porder == 1 ? 6 : 18) : (porder == 1 ? 5 : 14);
        for (int q = 0; q < nqp; q++) {
            if (topology == B200ASM_PRISM) prism_shapes(porder, qpts + (size_t)q * 3, n, phi + (size_t)q * n, dphi + (size_t)q * 3 * n);
            else pyramid_shapes(porder, qpts + (size_t)q * 3, n, phi + (size_t)q * n, dphi + (size_t)q * 3 * n);
        }
        return n;
    }
    switch (topology) {
        case B200ASM_HEX: n = porder == 1 ? 8 : 27; break;
        case B200ASM_QUAD: n = porder == 1 ? 4 : 9; break;
        case B200ASM_TET: n = porder == 1 ? 4 : 10; break;
        case B200ASM_TRI: n = porder == 1 ? 3 : 6; break;
        case B200ASM_LINE: n = porder + 1; break;
        default: return B200ASM_EINVAL;
    }
    const int dim = (topology == B200ASM_HEX || topology == B200ASM_TET) ? 3 : (topology == B200ASM_LINE ? 1 : 2);
    for (int q = 0; q < nqp; q++) {
        const double *pt = qpts + (size_t)q * dim;
        double *ph = phi + (size_t)q * n;
        double *dp = dphi + (size_t)q * dim * n;  // [d][i]
        if (topology == B200ASM_LINE) {  // TPZShapeLinear (Shape/pzshapelinear.cpp:269-292): l0, l1, 4 l0 l1
            double f[3], df[3];
            factors1d(pt[0], f, df);
            for (int s = 0; s < n; s++) {
                ph[s] = f[s];
                dp[s] = df[s];
            }
        } else if (topology == B200ASM_HEX || topology == B200ASM_QUAD) {
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


// ------------------------------------------------------------------------------------------------
// orientation keys and tables of arbitrary order (hexahedra, quadrilaterals)
// ------------------------------------------------------------------------------------------------
extern "C" int b200asm_nshape(int topology, int porder) {
    if (porder < 1) return B200ASM_EINVAL;
    const int p = porder;
    switch (topology) {
        case B200ASM_HEX: return (p + 1) * (p + 1) * (p + 1);
        case B200ASM_QUAD: return (p + 1) * (p + 1);
        case B200ASM_TET: return (p + 1) * (p + 2) * (p + 3) / 6;
        case B200ASM_TRI: return (p + 1) * (p + 2) / 2;
        case B200ASM_LINE: return p + 1;
        case B200ASM_PRISM: return p <= 2 ? (p == 1 ? 6 : 18) : B200ASM_EINVAL;
        case B200ASM_PYRAMID: return p <= 2 ? (p == 1 ? 5 : 14) : B200ASM_EINVAL;
    }
    return B200ASM_EINVAL;
}

extern "C" int b200asm_orientation_keys(int topology, int64_t nel, const int32_t *elnodes, int64_t *keys) {
    if (nel < 0 || (nel && (!elnodes || !keys))) return B200ASM_EINVAL;
    if (topology == B200ASM_LINE) {
        for (int64_t e = 0; e < nel; e++) keys[e] = elnodes[2 * e] < elnodes[2 * e + 1] ? 0 : 1;
        return 0;
    }
    if (topology == B200ASM_PRISM || topology == B200ASM_PYRAMID) {  // p <= 2 only: no orientation dependence
        for (int64_t e = 0; e < nel; e++) keys[e] = 0;
        return 0;
    }
    if (topology == B200ASM_TET || topology == B200ASM_TRI) {
        // 1 bit per edge (runs from the larger to the smaller global index), 3 bits per triangular side (which of the six
        // symmetries of the triangle makes it start at its smallest corner and turn towards the smaller neighbour:
        // Topology/tpztriangle.cpp:599-622); triangles: edges in bits 0-2, the element itself in bits 3-5;
        // tetrahedra: edges in bits 0-5, faces in bits 6-17
        const bool tet = topology == B200ASM_TET;
        const int nc = tet ? 4 : 3, ne = tet ? 6 : 3;
        for (int64_t e = 0; e < nel; e++) {
            const int32_t *id = elnodes + e * nc;
            int64_t key = 0;
            for (int k = 0; k < ne; k++) {
                const int a = tet ? kTetEdge[k][0] : kTriEdge[k][0], b = tet ? kTetEdge[k][1] : kTriEdge[k][1];
                if (!(id[a] < id[b])) key |= (int64_t)1 << k;
            }
            if (tet) {
                for (int f = 0; f < 4; f++)
                    key |= (int64_t)triangle_symmetry(id[kTetFace[f][0]], id[kTetFace[f][1]], id[kTetFace[f][2]]) << (6 + 3 * f);
            } else {
                key |= (int64_t)triangle_symmetry(id[0], id[1], id[2]) << 3;
            }
            keys[e] = key;
        }
        return 0;
    }
    if (topology != B200ASM_HEX && topology != B200ASM_QUAD) return B200ASM_EINVAL;
    const int nc = topology == B200ASM_HEX ? 8 : 4;
    for (int64_t e = 0; e < nel; e++) {
        const int32_t *id = elnodes + e * nc;
        int64_t key = 0;
        if (topology == B200ASM_HEX) {
            for (int k = 0; k < 12; k++)
                if (!(id[kHexEdge[k][0]] < id[kHexEdge[k][1]])) key |= (int64_t)1 << k;
            for (int f = 0; f < 6; f++) {
                const int *c = kHexFace[f];
                key |= (int64_t)square_symmetry(id[c[0]], id[c[1]], id[c[2]], id[c[3]]) << (12 + 3 * f);
            }
        } else {
            for (int k = 0; k < 4; k++)
                if (!(id[k] < id[(k + 1) & 3])) key |= (int64_t)1 << k;
            key |= (int64_t)square_symmetry(id[0], id[1], id[2], id[3]) << 4;
        }
        keys[e] = key;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// simplices of order >= 3 (Shape/TPZShapeH1.cpp:42-116 with pzshapetetra.cpp / pzshapetriang.cpp): every side carries its blend
// function B (edges 4 l_a l_b, triangular sides 27 l_a l_b l_c, tetrahedron interior 54 l_0 l_1 l_2 l_3) and the products of B
// with Chebyshev polynomials of the side's own coordinates:
//   edge (a,b):        T_i(s), s = +-(l_b - l_a), i = 1..p-2                 (sign: orientation bit of the edge)
//   triangular side:   T_i(2u'-1) T_j(2v'-1), i + j <= p-3, by (i+j, j)      ((u',v') = two barycentric coordinates of the side,
//                                                                              chosen by its symmetry class)
//   tetrahedron:       T_i(2x-1) T_j(2y-1) T_k(2z-1), i + j + k <= p-4, lexicographic in (i,j,k)
// ------------------------------------------------------------------------------------------------
static int simplex_tables_oriented(int topology, int porder, int64_t key, int nqp, const double *qpts, double *phi, double *dphi) {
    const bool tet = topology == B200ASM_TET;
    const int dim = tet ? 3 : 2, nc = dim + 1, ne = tet ? 6 : 3, nf = tet ? 4 : 1;
    const int n = b200asm_nshape(topology, porder);
    const int p = porder;
    if (p > 9) return B200ASM_EINVAL;
    for (int q = 0; q < nqp; q++) {
        const double *pt = qpts + (size_t)q * dim;
        double *ph = phi + (size_t)q * n;
        double *dp = dphi + (size_t)q * dim * n;
        double lam[4], dlam[4][3];
        lam[0] = 1.0;
        for (int d = 0; d < dim; d++) {
            lam[0] -= pt[d];
            lam[d + 1] = pt[d];
        }
        for (int a = 0; a < nc; a++)
            for (int d = 0; d < dim; d++) dlam[a][d] = (a == 0) ? -1.0 : (a == d + 1 ? 1.0 : 0.0);
        int shape = 0;
        auto emit = [&](double v, const double *g) {
            ph[shape] = v;
            for (int d = 0; d < dim; d++) dp[d * n + shape] = g[d];
            shape++;
        };
        for (int a = 0; a < nc; a++) emit(lam[a], dlam[a]);
        double T[3][16], dT[3][16];
        // edges
        for (int e = 0; e < ne; e++) {
            const int a = tet ? kTetEdge[e][0] : kTriEdge[e][0], b = tet ? kTetEdge[e][1] : kTriEdge[e][1];
            const double B = 4.0 * (lam[a] * lam[b]);
            double dB[3], ds[3];
            const double sgn = ((key >> e) & 1) ? -1.0 : 1.0;
            for (int d = 0; d < dim; d++) {
                dB[d] = 4.0 * (dlam[a][d] * lam[b] + lam[a] * dlam[b][d]);
                ds[d] = sgn * (dlam[b][d] - dlam[a][d]);
            }
            emit(B, dB);
            chebyshev_T(sgn * (lam[b] - lam[a]), p - 1, T[0], dT[0]);
            for (int i = 1; i < p - 1; i++) {
                double g[3];
                for (int d = 0; d < dim; d++) g[d] = dB[d] * T[0][i] + B * (dT[0][i] * ds[d]);
                emit(B * T[0][i], g);
            }
        }
        // triangular sides
        for (int f = 0; f < nf; f++) {
            const int *c = tet ? kTetFace[f] : kTetFace[0];
            const double B = 27.0 * (lam[c[0]] * lam[c[1]] * lam[c[2]]);
            double dB[3];
            for (int d = 0; d < dim; d++)
                dB[d] = 27.0 * (dlam[c[0]][d] * lam[c[1]] * lam[c[2]] + lam[c[0]] * dlam[c[1]][d] * lam[c[2]] + lam[c[0]] * lam[c[1]] * dlam[c[2]][d]);
            if (p >= 3) emit(B, dB);
            const int nin = (p - 2) * (p - 1) / 2;
            if (nin <= 1) continue;
            const int t = (int)((key >> (tet ? 6 + 3 * f : 3)) & 7);
            // side coordinates and their gradients, then mu = (1-u-v, u, v)
            double mu[3], dmu[3][3];
            for (int k = 0; k < 2; k++) {
                double v, g[3] = {0, 0, 0};
                if (tet) {
                    const double *cf = kTetFaceUV[f][k];
                    v = cf[0] + cf[1] * pt[0] + cf[2] * pt[1] + cf[3] * pt[2];
                    for (int d = 0; d < 3; d++) g[d] = cf[1 + d];
                } else {
                    v = pt[k];
                    g[k] = 1.0;
                }
                mu[k + 1] = v;
                for (int d = 0; d < 3; d++) dmu[k + 1][d] = g[d];
            }
            mu[0] = 1.0 - mu[1] - mu[2];
            for (int d = 0; d < 3; d++) dmu[0][d] = -dmu[1][d] - dmu[2][d];
            const int k0 = kTriPerm[t][0], k1 = kTriPerm[t][1];
            chebyshev_T(2.0 * mu[k0] - 1.0, p - 2, T[0], dT[0]);
            chebyshev_T(2.0 * mu[k1] - 1.0, p - 2, T[1], dT[1]);
            for (int s = 1; s <= p - 3; s++)       // i + j = s  (s = 0 is the blend function itself)
                for (int j = 0; j <= s; j++) {
                    const int i = s - j;
                    const double val = T[0][i] * T[1][j];
                    double g[3];
                    for (int d = 0; d < dim; d++)
                        g[d] = dB[d] * val + B * (2.0 * dT[0][i] * T[1][j] * dmu[k0][d] + 2.0 * T[0][i] * dT[1][j] * dmu[k1][d]);
                    emit(B * val, g);
                }
        }
        // interior of the tetrahedron
        if (tet && p >= 4) {
            const double B = 54.0 * (lam[0] * lam[1] * lam[2] * lam[3]);
            double dB[3];
            for (int d = 0; d < 3; d++)
                dB[d] = 54.0 * (dlam[0][d] * lam[1] * lam[2] * lam[3] + lam[0] * dlam[1][d] * lam[2] * lam[3] +
                                lam[0] * lam[1] * dlam[2][d] * lam[3] + lam[0] * lam[1] * lam[2] * dlam[3][d]);
            const int ord = p - 3;
            for (int k = 0; k < 3; k++) chebyshev_T(2.0 * pt[k] - 1.0, ord, T[k], dT[k]);
            for (int i = 0; i < ord; i++)
                for (int j = 0; j < ord; j++)
                    for (int k = 0; k < ord; k++) {
                        if (i + j + k >= ord) continue;
                        const double val = T[0][i] * T[1][j] * T[2][k];
                        double g[3];
                        g[0] = dB[0] * val + B * (2.0 * dT[0][i] * T[1][j] * T[2][k]);
                        g[1] = dB[1] * val + B * (2.0 * T[0][i] * dT[1][j] * T[2][k]);
                        g[2] = dB[2] * val + B * (2.0 * T[0][i] * T[1][j] * dT[2][k]);
                        emit(B * val, g);
                    }
        }
        if (shape != n) return B200ASM_EINVAL;
    }
    return n;
}

extern "C" int b200asm_shape_tables_oriented(int topology, int porder, int64_t key, int nqp, const double *qpts,
                                             double *phi, double *dphi) {
    if (porder < 1 || porder > 8 || nqp < 0) return B200ASM_EINVAL;
    if (porder <= 2 || topology == B200ASM_PRISM || topology == B200ASM_PYRAMID)
        return b200asm_shape_tables(topology, porder, nqp, qpts, phi, dphi);
    if (topology == B200ASM_TET || topology == B200ASM_TRI) return simplex_tables_oriented(topology, porder, key, nqp, qpts, phi, dphi);
    if (topology != B200ASM_HEX && topology != B200ASM_QUAD && topology != B200ASM_LINE) return B200ASM_EINVAL;
    const int dim = topology == B200ASM_HEX ? 3 : (topology == B200ASM_LINE ? 1 : 2);
    const int nsides = topology == B200ASM_HEX ? 27 : (topology == B200ASM_LINE ? 3 : 9);
    const int nc = topology == B200ASM_HEX ? 8 : (topology == B200ASM_LINE ? 2 : 4);
    static const int kLineSide[3][2] = {{0, 0}, {1, 0}, {2, 0}};
    const int n = b200asm_nshape(topology, porder);
    const int m = porder - 1;  // Chebyshev functions per side direction
    for (int q = 0; q < nqp; q++) {
        const double *pt = qpts + (size_t)q * dim;
        double *ph = phi + (size_t)q * n;
        double *dp = dphi + (size_t)q * dim * n;
        double f[3][3], df[3][3];
        for (int d = 0; d < dim; d++) factors1d(pt[d], f[d], df[d]);
        int shape = 0;
        for (int side = 0; side < nsides; side++) {
            // blend function of the side and its gradient
            const int *a = dim == 3 ? kHexSide[side] : (dim == 1 ? kLineSide[side] : kQuadSide[side]);
            double B, dB[3] = {0, 0, 0};
            if (dim == 1) {
                B = f[0][a[0]];
                dB[0] = df[0][a[0]];
            } else if (dim == 3) {
                B = f[0][a[0]] * f[1][a[1]] * f[2][a[2]];
                dB[0] = df[0][a[0]] * f[1][a[1]] * f[2][a[2]];
                dB[1] = f[0][a[0]] * df[1][a[1]] * f[2][a[2]];
                dB[2] = f[0][a[0]] * f[1][a[1]] * df[2][a[2]];
            } else {
                B = f[0][a[0]] * f[1][a[1]];
                dB[0] = df[0][a[0]] * f[1][a[1]];
                dB[1] = f[0][a[0]] * df[1][a[1]];
            }
            ph[shape] = B;
            for (int d = 0; d < dim; d++) dp[d * n + shape] = dB[d];
            shape++;
            if (side < nc) continue;
            const SideParam sp = side_param(topology, side, key);
            int count = m;
            for (int k = 1; k < sp.sdim; k++) count *= m;
            double T[3][8], dT[3][8];
            for (int k = 0; k < sp.sdim; k++) chebyshev_T(sp.sign[k] * pt[sp.axis[k]], m, T[k], dT[k]);
            for (int idx = 1; idx < count; idx++) {
                int i[3] = {0, 0, 0};
                if (sp.sdim == 1) i[0] = idx;
                else if (sp.sdim == 2) { i[0] = idx / m; i[1] = idx % m; }
                else { i[0] = idx / (m * m); i[1] = (idx / m) % m; i[2] = idx % m; }
                double val = T[0][i[0]];
                for (int k = 1; k < sp.sdim; k++) val *= T[k][i[k]];
                double g[3] = {0, 0, 0};  // gradient of the Chebyshev product w.r.t. the element coordinates
                for (int k = 0; k < sp.sdim; k++) {
                    double t = 1.0;
                    bool first = true;
                    for (int l = 0; l < sp.sdim; l++) {
                        const double fac = l == k ? dT[l][i[l]] : T[l][i[l]];
                        t = first ? fac : t * fac;
                        first = false;
                    }
                    g[sp.axis[k]] = sp.sign[k] * t;
                }
                ph[shape] = B * val;
                for (int d = 0; d < dim; d++) dp[d * n + shape] = dB[d] * val + B * g[d];
                shape++;
            }
        }
        if (shape != n) return B200ASM_EINVAL;
    }
    return n;
}
