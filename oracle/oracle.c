/*
 * TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's global-assembly path.
 *
 * This file restates, in plain C, the algorithm of labmec/NeoPZ for
 *   TPZLinearAnalysis::Assemble -> TPZStructMatrixOR::Serial_Assemble -> TPZCompElH1::CalcStiff
 *   -> TPZMatPoisson / TPZElasticity3D Contribute(BC) -> TPZSYsmpMatrix / TPZFYsmpMatrix AddKel + AddFel
 * following the reference's arithmetic ORDER (so results agree to the last few ulps), each function
 * citing the reference file:line it follows.  It is a checker only: nothing under neopz_b200/ may
 * import, link or call it (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against fixtures
 * produced by the compiled, unmodified reference (oracle/refdriver.cpp -> tests/golden/ *.npz) and
 * against the reference's own known-answer vector UnitTest_PZ/TestMaterial/CubeStiffMatrix.txt.
 *
 * Scope: H1 elements of uniform order p<=2 on hexahedra / tetrahedra and their quadrilateral /
 * triangular boundary faces (for p<=2 no side has more than one shape function, hence no
 * orientation transforms: Shape/TPZShapeH1.cpp:71,77), and of uniform order 3..6 on hexahedra /
 * quadrilaterals (side orientation from the global corner-node indices, orc_shape_ids), and of uniform order p<=2 on prisms
 * and pyramids (mixed quadrilateral / triangular faces).  Simplex and pyramid quadrature tables are data of the
 * reference (Integral/tpzintrulet.cpp, tpzintrulet3d.cpp) and are passed in by the caller.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle/oracle.c -o oracle/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_HEX 0
#define ORC_TET 1
#define ORC_QUAD 2
#define ORC_TRI 3
#define ORC_LINE 4
#define ORC_PRISM 5 /* EPrisma   (TPZShapePrism, TPZGeoPrism) */
#define ORC_PYR 6   /* EPiramide (TPZShapePiram, TPZGeoPyramid) */

#define ORC_POISSON 0
#define ORC_ELAST3D 1
#define ORC_POISSON_BC 2
#define ORC_ELAST3D_BC 3
#define ORC_ELAST2D 4    /* TPZElasticity2D on quadrilaterals / triangles (plane meshes) */
#define ORC_ELAST2D_BC 5 /* its boundary conditions on line elements */

/* ------------------------------------------------------------------------------------------
 * Quadrature.  Integral/tpzgaussrule.cpp:171-243 (Gauss-Legendre by Newton iteration in
 * long double, points stored interleaved -z,+z), :172 npts = (int)(0.51*(order+2)).
 * ------------------------------------------------------------------------------------------ */
static long double orc_machine_precision(void) {
    /* Integral/tpzgaussrule.cpp:552-563 */
    long double value = 1.0L;
    while (1.0L < (long double)(1.0L + value)) value = value / 2.0L;
    value = 2.0L * value;
    return value;
}

int orc_gauss1d_ld(int order, long double *loc, long double *w) {
    int npoints = (int)(0.51 * (order + 2));
    if (npoints < 1) npoints = 1;
    const long double tol = orc_machine_precision();
    const int m = (npoints + 1) / 2;
    for (int i = 0; i < m; i++) {
        long double p1 = ((long double)i) + 0.75L;
        long double p2 = ((long double)npoints) + 0.5L;
        long double z = cosl((M_PI * p1) / p2);
        long double z1, pp, p3, dif, den;
        long iteration = 0;
        do {
            iteration++;
            p1 = 1.0L;
            p2 = 0.0L;
            for (int j = 0; j < npoints; j++) {
                p3 = p2;
                p2 = p1;
                p1 = ((2.0L * ((long double)j) + 1.0L) * z * p2 - (((long double)j) * p3)) / (((long double)j) + 1.0L);
            }
            den = (z * z) - 1.0L;
            if (fabsl(den) < 1.e-16L) z = 0.5L;
            pp = ((long double)npoints) * (z * p1 - p2) / den;
            z1 = z;
            if (fabsl(pp) < 1.e-16L) z = 0.5L;
            else z = z1 - p1 / pp;
            dif = fabsl(z - z1);
        } while (fabsl(dif) > tol && iteration < 100000);
        long double weight = 2.0L / ((1.0L - z * z) * pp * pp);
        loc[2 * i] = -z;
        w[2 * i] = weight;
        if ((2 * i + 1) < npoints) {
            loc[2 * i + 1] = z;
            w[2 * i + 1] = weight;
        }
    }
    return npoints;
}

/* hexahedron: Integral/pzquad.cpp:268-284, ik fastest; weight product formed in long double */
int orc_rule_hex(int order, double *pts, double *w) {
    long double l[64], ww[64];
    const int n = orc_gauss1d_ld(order, l, ww);
    for (int ip = 0; ip < n * n * n; ip++) {
        const int ik = ip % n, ie = (ip % (n * n)) / n, iz = ip / (n * n);
        pts[3 * ip + 0] = (double)l[ik];
        pts[3 * ip + 1] = (double)l[ie];
        pts[3 * ip + 2] = (double)l[iz];
        w[ip] = (double)(ww[ik] * ww[ie] * ww[iz]);
    }
    return n * n * n;
}

/* quadrilateral: Integral/pzquad.cpp:153-169, ik = ip / nEta (ksi slowest) */
/* TPZInt1d: the 1-D Gauss-Legendre rule itself (Integral/pzquad.cpp, TPZGaussRule) */
int orc_rule_line(int order, double *pts, double *w) {
    long double l[64], ww[64];
    const int n = orc_gauss1d_ld(order, l, ww);
    for (int ip = 0; ip < n; ip++) { pts[ip] = (double)l[ip]; w[ip] = (double)ww[ip]; }
    return n;
}

int orc_rule_quad(int order, double *pts, double *w) {
    long double l[64], ww[64];
    const int n = orc_gauss1d_ld(order, l, ww);
    for (int ip = 0; ip < n * n; ip++) {
        const int ik = ip / n, ie = ip - (ip / n) * n;
        pts[2 * ip + 0] = (double)l[ik];
        pts[2 * ip + 1] = (double)l[ie];
        w[ip] = (double)(ww[ik] * ww[ie]);
    }
    return n * n;
}

/* ------------------------------------------------------------------------------------------
 * Topology tables (combinatorial facts of the reference elements; Topology/tpzcube.cpp:30-80,
 * Topology/tpztetrahedron.h:283, Topology/tpzquadrilateral.cpp, Topology/tpztriangle.cpp).
 * ------------------------------------------------------------------------------------------ */
static const int cube_edge_nodes[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {7, 4}};
/* first and third corner of each face (ContainedSideLocId(face,0) and (face,2)), tpzcube.cpp:31-33 */
static const int cube_face_n02[6][2] = {{0, 2}, {0, 5}, {1, 6}, {3, 6}, {0, 7}, {4, 6}};
static const int cube_highsides[27][7] = {
    {8, 11, 12, 20, 21, 24, 26}, {8, 9, 13, 20, 21, 22, 26}, {9, 10, 14, 20, 22, 23, 26}, {10, 11, 15, 20, 23, 24, 26},
    {12, 16, 19, 21, 24, 25, 26}, {13, 16, 17, 21, 22, 25, 26}, {14, 17, 18, 22, 23, 25, 26}, {15, 18, 19, 23, 24, 25, 26},
    {20, 21, 26}, {20, 22, 26}, {20, 23, 26}, {20, 24, 26}, {21, 24, 26}, {21, 22, 26}, {22, 23, 26}, {23, 24, 26},
    {21, 25, 26}, {22, 25, 26}, {23, 25, 26}, {24, 25, 26}, {26}, {26}, {26}, {26}, {26}, {26}, {-1}};
static const int cube_nhigh[27] = {7, 7, 7, 7, 7, 7, 7, 7, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1, 1, 1, 1, 1, 1, 0};
static const int tet_edge_nodes[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};

/* ------------------------------------------------------------------------------------------
 * H1 shape functions, p <= 2.  Shape/TPZShapeH1.cpp:42-116 with
 *   hex : Shape/pzshapecube.cpp:36-85 (ShapeCorner), :92-152 (ShapeGenerating)
 *   quad: Shape/pzshapequad.cpp:36-59, :67-93
 *   tet : Shape/pzshapetetra.cpp:53-71, :101-164
 *   tri : Shape/pzshapetriang.cpp:34-45, :53-81
 * phi[n], dphi[dim][n] (row = direction), n returned.
 * NConnectShapeF at p=2: edge 1, quad face 1, hex interior 1, tri face 0, tet interior 0.
 * ------------------------------------------------------------------------------------------ */
static void hex_corner(const double *pt, double *phi, double (*d)[27]) {
    double x[2], dx[2], y[2], dy[2], z[2], dz[2];
    x[0] = (1. - pt[0]) / 2.; x[1] = (1. + pt[0]) / 2.; dx[0] = -0.5; dx[1] = 0.5;
    y[0] = (1. - pt[1]) / 2.; y[1] = (1. + pt[1]) / 2.; dy[0] = -0.5; dy[1] = 0.5;
    z[0] = (1. - pt[2]) / 2.; z[1] = (1. + pt[2]) / 2.; dz[0] = -0.5; dz[1] = 0.5;
    static const int s[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    for (int a = 0; a < 8; a++) {
        phi[a] = x[s[a][0]] * y[s[a][1]] * z[s[a][2]];
        d[0][a] = dx[s[a][0]] * y[s[a][1]] * z[s[a][2]];
        d[1][a] = x[s[a][0]] * dy[s[a][1]] * z[s[a][2]];
        d[2][a] = x[s[a][0]] * y[s[a][1]] * dz[s[a][2]];
    }
}

static int shape_hex(int p, const double *pt, double *phi, double *dphi_out) {
    double ph[27], d[3][27];
    hex_corner(pt, ph, d);
    if (p == 1) {
        for (int a = 0; a < 8; a++) { phi[a] = ph[a]; for (int k = 0; k < 3; k++) dphi_out[k * 8 + a] = d[k][a]; }
        return 8;
    }
    for (int is = 8; is < 27; is++) {
        int is1, is2;
        if (is < 20) { is1 = cube_edge_nodes[is - 8][0]; is2 = cube_edge_nodes[is - 8][1]; }
        else if (is < 26) { is1 = cube_face_n02[is - 20][0]; is2 = cube_face_n02[is - 20][1]; }
        else { is1 = 0; is2 = 6; }
        ph[is] = ph[is1] * ph[is2];
        for (int k = 0; k < 3; k++) d[k][is] = d[k][is1] * ph[is2] + ph[is1] * d[k][is2];
    }
    for (int is = 8; is < 27; is++) {
        for (int h = 0; h < cube_nhigh[is]; h++) {
            const int hs = cube_highsides[is][h];
            ph[is] += ph[hs];
            for (int k = 0; k < 3; k++) d[k][is] += d[k][hs];
        }
        const int mult = is < 20 ? 4 : (is < 26 ? 16 : 64);
        ph[is] *= mult;
        for (int k = 0; k < 3; k++) d[k][is] *= mult;
    }
    for (int a = 0; a < 27; a++) { phi[a] = ph[a]; for (int k = 0; k < 3; k++) dphi_out[k * 27 + a] = d[k][a]; }
    return 27;
}

static int shape_quad(int p, const double *pt, double *phi, double *dphi_out) {
    double ph[9], d[2][9];
    double x[2], dx[2], y[2], dy[2];
    x[0] = (1. - pt[0]) / 2.; x[1] = (1. + pt[0]) / 2.; dx[0] = -0.5; dx[1] = 0.5;
    y[0] = (1. - pt[1]) / 2.; y[1] = (1. + pt[1]) / 2.; dy[0] = -0.5; dy[1] = 0.5;
    static const int s[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
    for (int a = 0; a < 4; a++) {
        ph[a] = x[s[a][0]] * y[s[a][1]];
        d[0][a] = dx[s[a][0]] * y[s[a][1]];
        d[1][a] = x[s[a][0]] * dy[s[a][1]];
    }
    int n = 4;
    if (p >= 2) {
        for (int is = 4; is < 8; is++) {
            const int a = is % 4, b = (is + 1) % 4;
            ph[is] = ph[a] * ph[b];
            for (int k = 0; k < 2; k++) d[k][is] = d[k][a] * ph[b] + ph[a] * d[k][b];
        }
        ph[8] = ph[0] * ph[2];
        for (int k = 0; k < 2; k++) d[k][8] = d[k][0] * ph[2] + ph[0] * d[k][2];
        for (int is = 4; is < 8; is++) {
            ph[is] += ph[8];
            d[0][is] += d[0][8]; d[1][is] += d[1][8];
            ph[is] *= 4.; d[0][is] *= 4.; d[1][is] *= 4.;
        }
        ph[8] *= 16.; d[0][8] *= 16.; d[1][8] *= 16.;
        n = 9;
    }
    for (int a = 0; a < n; a++) { phi[a] = ph[a]; for (int k = 0; k < 2; k++) dphi_out[k * n + a] = d[k][a]; }
    return n;
}

static int shape_tet(int p, const double *pt, double *phi, double *dphi_out) {
    double ph[10], d[3][10];
    ph[0] = 1 - pt[0] - pt[1] - pt[2]; ph[1] = pt[0]; ph[2] = pt[1]; ph[3] = pt[2];
    static const double dc[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int a = 0; a < 4; a++) for (int k = 0; k < 3; k++) d[k][a] = dc[a][k];
    int n = 4;
    if (p >= 2) {
        for (int e = 0; e < 6; e++) {
            const int a = tet_edge_nodes[e][0], b = tet_edge_nodes[e][1], is = 4 + e;
            ph[is] = ph[a] * ph[b];
            for (int k = 0; k < 3; k++) d[k][is] = d[k][a] * ph[b] + ph[a] * d[k][b];
        }
        for (int is = 4; is < 10; is++) { ph[is] *= 4.; for (int k = 0; k < 3; k++) d[k][is] *= 4.; }
        n = 10;
    }
    for (int a = 0; a < n; a++) { phi[a] = ph[a]; for (int k = 0; k < 3; k++) dphi_out[k * n + a] = d[k][a]; }
    return n;
}

static int shape_tri(int p, const double *pt, double *phi, double *dphi_out) {
    double ph[6], d[2][6];
    ph[0] = 1. - pt[0] - pt[1]; ph[1] = pt[0]; ph[2] = pt[1];
    d[0][0] = -1.; d[1][0] = -1.; d[0][1] = 1.; d[1][1] = 0.; d[0][2] = 0.; d[1][2] = 1.;
    int n = 3;
    if (p >= 2) {
        for (int is = 3; is < 6; is++) {
            const int a = is % 3, b = (is + 1) % 3;
            ph[is] = ph[a] * ph[b];
            for (int k = 0; k < 2; k++) d[k][is] = d[k][a] * ph[b] + ph[a] * d[k][b];
        }
        for (int is = 3; is < 6; is++) { ph[is] *= 4.; d[0][is] *= 4.; d[1][is] *= 4.; }
        n = 6;
    }
    for (int a = 0; a < n; a++) { phi[a] = ph[a]; for (int k = 0; k < 2; k++) dphi_out[k * n + a] = d[k][a]; }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Prism, p <= 2.  Shape/pzshapeprism.cpp:42-74 (ShapeCorner), :115-205 (ShapeGenerating), :761-775 (NConnectShapeF:
 * edges p-1, triangular faces (p-2)(p-1)/2, quadrilateral faces (p-1)^2, interior (p-2)(p-1)^2/2); side tables
 * Topology/tpzprism.h:275-278, higher-dimension sides Topology/tpzprism.cpp:38-60.  At p = 2: 6 + 9 + 3 = 18 functions
 * (sides 6..14 and 16, 17, 18).
 * ------------------------------------------------------------------------------------------ */
static const int prism_edge_nodes[9][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 4}, {2, 5}, {3, 4}, {4, 5}, {5, 3}};
static const int prism_face_n02[5][2] = {{0, 2}, {0, 4}, {1, 5}, {0, 5}, {3, 5}}; /* FaceNodes[f][0], [f][2] */
/* quadrilateral faces among the higher-dimension sides of every edge, in the order of highsides[] */
static const int prism_edge_quads[9][2] = {{16, -1}, {17, -1}, {18, -1}, {16, 18}, {16, 17}, {17, 18}, {16, -1}, {17, -1}, {18, -1}};

static void prism_corner(const double *pt, double *phi, double (*d)[27]) {
    phi[0] = .5 * (1. - pt[0] - pt[1]) * (1. - pt[2]);
    phi[1] = .5 * pt[0] * (1. - pt[2]);
    phi[2] = .5 * pt[1] * (1. - pt[2]);
    phi[3] = .5 * (1. - pt[0] - pt[1]) * (1. + pt[2]);
    phi[4] = .5 * pt[0] * (1. + pt[2]);
    phi[5] = .5 * pt[1] * (1. + pt[2]);
    d[0][0] = -.5 * (1. - pt[2]); d[1][0] = -.5 * (1. - pt[2]); d[2][0] = -.5 * (1. - pt[0] - pt[1]);
    d[0][1] = .5 * (1. - pt[2]);  d[1][1] = .0;                 d[2][1] = -.5 * pt[0];
    d[0][2] = .0;                 d[1][2] = .5 * (1. - pt[2]);  d[2][2] = -.5 * pt[1];
    d[0][3] = -.5 * (1. + pt[2]); d[1][3] = -.5 * (1. + pt[2]); d[2][3] = .5 * (1. - pt[0] - pt[1]);
    d[0][4] = .5 * (1. + pt[2]);  d[1][4] = .0;                 d[2][4] = .5 * pt[0];
    d[0][5] = .0;                 d[1][5] = .5 * (1. + pt[2]);  d[2][5] = .5 * pt[1];
}

static int shape_prism(int p, const double *pt, double *phi, double *dphi_out) {
    double ph[27], d[3][27];
    prism_corner(pt, ph, d);
    if (p == 1) {
        for (int a = 0; a < 6; a++) { phi[a] = ph[a]; for (int k = 0; k < 3; k++) dphi_out[k * 6 + a] = d[k][a]; }
        return 6;
    }
    for (int is = 6; is < 19; is++) {
        if (is == 15) continue; /* triangular faces and the interior carry no function at p = 2 */
        int is1, is2;
        if (is < 15) { is1 = prism_edge_nodes[is - 6][0]; is2 = prism_edge_nodes[is - 6][1]; }
        else { is1 = prism_face_n02[is - 15][0]; is2 = prism_face_n02[is - 15][1]; }
        ph[is] = ph[is1] * ph[is2];
        for (int k = 0; k < 3; k++) d[k][is] = d[k][is1] * ph[is2] + ph[is1] * d[k][is2];
    }
    for (int is = 6; is < 15; is++)
        for (int h = 0; h < 2; h++) {
            const int hs = prism_edge_quads[is - 6][h];
            if (hs < 0) continue;
            ph[is] += ph[hs];
            for (int k = 0; k < 3; k++) d[k][is] += d[k][hs];
        }
    int n = 6;
    for (int is = 6; is < 19; is++) {
        if (is == 15) continue;
        const double mult = is < 15 ? 4. : 16.;
        phi[n] = ph[is] * mult;
        for (int k = 0; k < 3; k++) d[k][is] *= mult;
        n++;
    }
    for (int a = 0; a < 6; a++) phi[a] = ph[a];
    n = 0;
    for (int is = 0; is < 19; is++) {
        if (is == 15) continue;
        for (int k = 0; k < 3; k++) dphi_out[k * 18 + n] = d[k][is];
        n++;
    }
    return 18;
}

/* ------------------------------------------------------------------------------------------
 * Pyramid, p <= 2.  Shape/pzshapepiram.cpp:331-392 (ShapeCorner: rational corner functions; the apex branch is never
 * taken at an integration point), :47-119 (ShapeGenerating), :679-696 (NConnectShapeF: edges p-1, base (p-1)^2,
 * triangular faces (p-2)(p-1)/2, interior sum i(i+1)/2); ContainedSideLocId Topology/tpzpyramid.cpp:898-925.
 * At p = 2: 5 + 8 + 1 = 14 functions (sides 5..12 and 13).
 * ------------------------------------------------------------------------------------------ */
static void pyr_corner(const double *pt, double *phi, double (*d)[27]) {
    const double T0xz = .5 * (1. - pt[2] - pt[0]) / (1. - pt[2]);
    const double T0yz = .5 * (1. - pt[2] - pt[1]) / (1. - pt[2]);
    const double T1xz = .5 * (1. - pt[2] + pt[0]) / (1. - pt[2]);
    const double T1yz = .5 * (1. - pt[2] + pt[1]) / (1. - pt[2]);
    const double lmez = (1. - pt[2]);
    phi[0] = T0xz * T0yz * lmez;
    phi[1] = T1xz * T0yz * lmez;
    phi[2] = T1xz * T1yz * lmez;
    phi[3] = T0xz * T1yz * lmez;
    phi[4] = pt[2];
    const double lmexmez = 1. - pt[0] - pt[2];
    const double lmeymez = 1. - pt[1] - pt[2];
    const double lmaxmez = 1. + pt[0] - pt[2];
    const double lmaymez = 1. + pt[1] - pt[2];
    d[0][0] = -.25 * lmeymez / lmez;
    d[1][0] = -.25 * lmexmez / lmez;
    d[2][0] = -.25 * (lmeymez + lmexmez - lmexmez * lmeymez / lmez) / lmez;
    d[0][1] = .25 * lmeymez / lmez;
    d[1][1] = -.25 * lmaxmez / lmez;
    d[2][1] = -.25 * (lmeymez + lmaxmez - lmaxmez * lmeymez / lmez) / lmez;
    d[0][2] = .25 * lmaymez / lmez;
    d[1][2] = .25 * lmaxmez / lmez;
    d[2][2] = -.25 * (lmaymez + lmaxmez - lmaxmez * lmaymez / lmez) / lmez;
    d[0][3] = -.25 * lmaymez / lmez;
    d[1][3] = .25 * lmexmez / lmez;
    d[2][3] = -.25 * (lmaymez + lmexmez - lmexmez * lmaymez / lmez) / lmez;
    d[0][4] = 0.0;
    d[1][4] = 0.0;
    d[2][4] = 1.0;
}

static int shape_pyr(int p, const double *pt, double *phi, double *dphi_out) {
    double ph[27], d[3][27];
    pyr_corner(pt, ph, d);
    if (p == 1) {
        for (int a = 0; a < 5; a++) { phi[a] = ph[a]; for (int k = 0; k < 3; k++) dphi_out[k * 5 + a] = d[k][a]; }
        return 5;
    }
    for (int is = 5; is < 14; is++) {
        int is1, is2;
        if (is < 9) { is1 = is - 5; is2 = (is - 5 + 1) % 4; }
        else if (is < 13) { is1 = is - 9; is2 = 4; }
        else { is1 = 0; is2 = 2; } /* ShapeFaceId[0][0], [0][2] */
        ph[is] = ph[is1] * ph[is2];
        for (int k = 0; k < 3; k++) d[k][is] = d[k][is1] * ph[is2] + ph[is1] * d[k][is2];
    }
    for (int is = 5; is < 9; is++) { /* the base edges take the base-face function */
        ph[is] += ph[13];
        for (int k = 0; k < 3; k++) d[k][is] += d[k][13];
    }
    for (int is = 5; is < 14; is++) {
        const double scale = is < 13 ? 4. : 16.;
        ph[is] *= scale;
        for (int k = 0; k < 3; k++) d[k][is] *= scale;
    }
    for (int a = 0; a < 14; a++) { phi[a] = ph[a]; for (int k = 0; k < 3; k++) dphi_out[k * 14 + a] = d[k][a]; }
    return 14;
}

/* TPZShapeLinear, p <= 2: Shape/pzshapelinear.cpp:269-283 (corner), :285-292 (generating: phi0*phi1*4) */
static int shape_line(int p, const double *pt, double *phi, double *dphi_out) {
    const int n = p == 1 ? 2 : 3;
    phi[0] = (1. - pt[0]) / 2.; phi[1] = (1. + pt[0]) / 2.;
    dphi_out[0] = -0.5; dphi_out[1] = 0.5;
    if (p >= 2) {
        phi[2] = phi[0] * phi[1];
        dphi_out[2] = dphi_out[0] * phi[1] + phi[0] * dphi_out[1];
        phi[2] *= 4.; dphi_out[2] *= 4.;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * H1 shape functions of arbitrary order on hexahedra / quadrilaterals (p >= 3 needs the side
 * orientations).  Shape/TPZShapeH1.cpp:42-116:
 *   first function of a side = blend function (ShapeGenerating); the others = blend * phin(i), i >= 1, where
 *   phin = TSHAPE::ShapeInternal(side, T pt) (Chebyshev products, Shape/pzshapelinear.cpp:15-33,306-312,
 *   Shape/pzshapequad.cpp:317-337, Shape/pzshapecube.cpp:460-485) and T = ParametricTransform(transform id)
 *   * TransformElementToSide(side) (Shape/pzgenericshape.cpp:13-55; Topology/tpzcube.cpp:583-655,
 *   Topology/tpzquadrilateral.cpp:296-332; table gTrans2dQ Shape/pzshapequad.cpp:23-32).  The volume side of a
 *   3-D element keeps the untransformed point (pzgenericshape.cpp:40-43).
 *   Transform ids from the global corner-node indices: edges Topology/tpzcube.cpp:1073-1089, quadrilateral
 *   faces Topology/tpzquadrilateral.cpp:591-618.
 *   dphi of a side function: dphiblend*phin + phiblend*(T.Mult^T dphin)  (TPZShapeH1.cpp:86-101).
 * ------------------------------------------------------------------------------------------ */
#define ORC_MAXP 6
#define ORC_MAXSHAPE ((ORC_MAXP + 1) * (ORC_MAXP + 1) * (ORC_MAXP + 1))

static const double gTrans2dQ[8][2][2] = {{{1., 0.}, {0., 1.}},  {{0., 1.}, {1., 0.}},   {{0., 1.}, {-1., 0.}}, {{-1., 0.}, {0., 1.}},
                                          {{-1., 0.}, {0., -1.}}, {{0., -1.}, {-1., 0.}}, {{0., -1.}, {1., 0.}}, {{1., 0.}, {0., -1.}}};
static const int cube_face_nodes[6][4] = {{0, 1, 2, 3}, {0, 1, 5, 4}, {1, 2, 6, 5}, {3, 2, 6, 7}, {0, 3, 7, 4}, {4, 5, 6, 7}};

/* Topology/tpzquadrilateral.cpp:591-618 */
static int quad_transform_id(const int64_t *id) {
    int id0, id1, minid;
    id0 = (id[0] < id[1]) ? 0 : 1;
    id1 = (id[2] < id[3]) ? 2 : 3;
    minid = (id[id0] < id[id1]) ? id0 : id1;
    id0 = (minid + 1) % 4;
    id1 = (minid + 3) % 4;
    const int64_t minglob = id[minid];
    if (id[id0] < id[id1]) {
        if (minglob == id[0]) return 0;
        if (minglob == id[1]) return 2;
        if (minglob == id[2]) return 4;
        if (minglob == id[3]) return 6;
    } else {
        if (minglob == id[0]) return 1;
        if (minglob == id[1]) return 3;
        if (minglob == id[2]) return 5;
        if (minglob == id[3]) return 7;
    }
    return 0;
}

/* Shape/pzshapelinear.cpp:15-33 */
static void chebyshev(double x, int num, double *phi, double *dphi) {
    if (num <= 0) return;
    phi[0] = 1.0; dphi[0] = 0.0;
    if (num == 1) return;
    phi[1] = x; dphi[1] = 1.0;
    for (int ord = 2; ord < num; ord++) {
        phi[ord] = 2.0 * x * phi[ord - 1] - phi[ord - 2];
        dphi[ord] = 2.0 * x * dphi[ord - 1] + 2.0 * phi[ord - 1] - dphi[ord - 2];
    }
}

/* TSHAPE::TransformElementToSide(side): Mult (sidedim x dim) and Sum (zero for cube and quad; the simplices project
 * affinely: Topology/tpztetrahedron.cpp:460-545, Topology/tpztriangle.cpp:443-480) */
static void element_to_side(int topo, int side, int *sidedim, double E[3][3], double Esum[3]) {
    memset(E, 0, sizeof(double) * 9);
    Esum[0] = Esum[1] = Esum[2] = 0.;
    if (topo == ORC_TET) {
        *sidedim = side < 10 ? 1 : (side < 14 ? 2 : 3);
        switch (side) {
            case 4: E[0][0] = 2.0; E[0][1] = 1.0; E[0][2] = 1.0; Esum[0] = -1.0; break;
            case 5: E[0][0] = -1.0; E[0][1] = 1.0; break;
            case 6: E[0][0] = -1.0; E[0][1] = -2.0; E[0][2] = -1.0; Esum[0] = 1.0; break;
            case 7: E[0][0] = 1.0; E[0][1] = 1.0; E[0][2] = 2.0; Esum[0] = -1.0; break;
            case 8: E[0][0] = -1.0; E[0][2] = 1.0; break;
            case 9: E[0][1] = -1.0; E[0][2] = 1.0; break;
            case 10: E[0][0] = 1.0; E[1][1] = 1.0; break;
            case 11: E[0][0] = 1.0; E[1][2] = 1.0; break;
            case 12:
                E[0][0] = -1.0 / 3.0; E[0][1] = 2.0 / 3.0; E[0][2] = -1.0 / 3.0;
                E[1][0] = -1.0 / 3.0; E[1][1] = -1.0 / 3.0; E[1][2] = 2.0 / 3.0;
                Esum[0] = 1.0 / 3.0; Esum[1] = 1.0 / 3.0;
                break;
            case 13: E[0][1] = 1.0; E[1][2] = 1.0; break;
            default: E[0][0] = E[1][1] = E[2][2] = 1.0; break; /* 14 */
        }
    } else if (topo == ORC_TRI) {
        *sidedim = side < 6 ? 1 : 2;
        switch (side) {
            case 3: E[0][0] = 2.0; E[0][1] = 1.0; Esum[0] = -1.0; break;
            case 4: E[0][0] = -1.0; E[0][1] = 1.0; break;
            case 5: E[0][0] = -1.0; E[0][1] = -2.0; Esum[0] = 1.0; break;
            default: E[0][0] = 1.0; E[1][1] = 1.0; break; /* 6 */
        }
    } else if (topo == ORC_HEX) {
        if (side < 20) {
            *sidedim = 1;
            switch (side) {
                case 8: case 16: E[0][0] = 1.; break;
                case 9: case 17: E[0][1] = 1.; break;
                case 10: case 18: E[0][0] = -1.; break;
                case 11: case 19: E[0][1] = -1.; break;
                default: E[0][2] = 1.; break; /* 12..15 */
            }
        } else if (side < 26) {
            *sidedim = 2;
            switch (side) {
                case 20: case 25: E[0][0] = 1.; E[1][1] = 1.; break;
                case 21: case 23: E[0][0] = 1.; E[1][2] = 1.; break;
                default: E[0][1] = 1.; E[1][2] = 1.; break; /* 22, 24 */
            }
        } else {
            *sidedim = 3;
            E[0][0] = E[1][1] = E[2][2] = 1.;
        }
    } else if (topo == ORC_LINE) { /* Topology/tpzline.cpp:285-300: the line itself */
        *sidedim = 1;
        E[0][0] = 1.;
    } else { /* ORC_QUAD */
        if (side < 8) {
            *sidedim = 1;
            switch (side) {
                case 4: E[0][0] = 1.; break;
                case 5: E[0][1] = 1.; break;
                case 6: E[0][0] = -1.; break;
                default: E[0][1] = -1.; break;
            }
        } else {
            *sidedim = 2;
            E[0][0] = E[1][1] = 1.;
        }
    }
}

/* Shape/pzshapetriang.cpp:18-27 */
static const double gTrans2dT[6][2][2] = {{{1., 0.}, {0., 1.}},   {{0., 1.}, {1., 0.}},   {{0., 1.}, {-1., -1.}},
                                          {{-1., -1.}, {0., 1.}}, {{-1., -1.}, {1., 0.}}, {{1., 0.}, {-1., -1.}}};
static const double gVet2dT[6][2] = {{0., 0.}, {0., 0.}, {0., 1.}, {1., 0.}, {1., 0.}, {0., 1.}};
static const int tet_face_nodes[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};

/* Topology/tpztriangle.cpp:599-622 */
static int tri_transform_id(const int64_t *id) {
    int id0, id1, minid;
    id0 = (id[0] < id[1]) ? 0 : 1;
    minid = (id[2] < id[id0]) ? 2 : id0;
    id0 = (minid + 1) % 3;
    id1 = (minid + 2) % 3;
    if (id[id0] < id[id1]) return 2 * minid;
    return 2 * minid + 1;
}

/* GetSideTransform (pzgenericshape.cpp:13-55): Mult = P.Mult * E.Mult, Sum = P.Mult * E.Sum + P.Sum, each product formed as
 * TPZFMatrix::MultAdd does (Matrix/pzfmatrix.cpp:596-606: z = 0, then z += 1 * a(i,k) * x(k,j) for k ascending); the volume
 * side of a 3-D element keeps TransformElementToSide. */
static void side_transform(int topo, int side, const int64_t *ids, int *sidedim, double T[3][3], double Tsum[3]) {
    double E[3][3], Esum[3];
    const int dim = (topo == ORC_HEX || topo == ORC_TET) ? 3 : (topo == ORC_LINE ? 1 : 2);
    element_to_side(topo, side, sidedim, E, Esum);
    if ((topo == ORC_HEX && side == 26) || (topo == ORC_TET && side == 14)) {
        memcpy(T, E, sizeof(E));
        memcpy(Tsum, Esum, sizeof(Esum));
        return;
    }
    double P[3][3], Psum[3] = {0., 0., 0.};
    memset(P, 0, sizeof(P));
    if (*sidedim == 1) {
        int a, b;
        if (topo == ORC_HEX) { a = cube_edge_nodes[side - 8][0]; b = cube_edge_nodes[side - 8][1]; }
        else if (topo == ORC_TET) { a = tet_edge_nodes[side - 4][0]; b = tet_edge_nodes[side - 4][1]; }
        else if (topo == ORC_TRI) { a = side - 3; b = (side - 2) % 3; }
        else if (topo == ORC_LINE) { a = 0; b = 1; }
        else { a = side - 4; b = (side - 3) % 4; }
        P[0][0] = ids[a] < ids[b] ? 1. : -1.;
    } else if (topo == ORC_TET || topo == ORC_TRI) {
        int64_t loc[3];
        for (int i = 0; i < 3; i++) loc[i] = topo == ORC_TET ? ids[tet_face_nodes[side - 10][i]] : ids[i];
        const int tid = tri_transform_id(loc);
        for (int i = 0; i < 2; i++) {
            for (int j = 0; j < 2; j++) P[i][j] = gTrans2dT[tid][i][j];
            Psum[i] = gVet2dT[tid][i];
        }
    } else {
        int64_t loc[4];
        for (int i = 0; i < 4; i++) loc[i] = topo == ORC_HEX ? ids[cube_face_nodes[side - 20][i]] : ids[i];
        const int tid = quad_transform_id(loc);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) P[i][j] = gTrans2dQ[tid][i][j];
    }
    for (int i = 0; i < *sidedim; i++) {
        for (int j = 0; j < dim; j++) {
            double v = 0.;
            for (int k = 0; k < *sidedim; k++) v += 1. * P[i][k] * E[k][j];
            T[i][j] = v;
        }
        double v = 0.;
        for (int k = 0; k < *sidedim; k++) v += 1. * P[i][k] * Esum[k];
        Tsum[i] = v + Psum[i];
    }
}

/* blend (generating) functions of all sides: the p = 2 tables of shape_hex / shape_quad */
static int shape_hq_general(int topo, int p, const int64_t *ids, const double *pt, double *phi, double *dphi_out) {
    const int dim = topo == ORC_HEX ? 3 : (topo == ORC_LINE ? 1 : 2);
    const int nc = topo == ORC_HEX ? 8 : (topo == ORC_LINE ? 2 : 4), nsides = topo == ORC_HEX ? 27 : (topo == ORC_LINE ? 3 : 9);
    double bphi[27], bd[3 * 27];
    if (topo == ORC_HEX) shape_hex(2, pt, bphi, bd); else if (topo == ORC_LINE) shape_line(2, pt, bphi, bd); else shape_quad(2, pt, bphi, bd);
    int n = nc;
    for (int side = nc; side < nsides; side++) {
        int sd;
        if (topo == ORC_HEX) sd = side < 20 ? 1 : (side < 26 ? 2 : 3); else if (topo == ORC_LINE) sd = 1; else sd = side < 8 ? 1 : 2;
        int ns = p - 1;
        if (sd == 2) ns *= (p - 1);
        if (sd == 3) ns *= (p - 1) * (p - 1);
        n += ns;
    }
    /* output arrays are [dim][n] */
    for (int a = 0; a < nc; a++) { phi[a] = bphi[a]; for (int k = 0; k < dim; k++) dphi_out[k * n + a] = bd[k * nsides + a]; }
    int shape = nc;
    for (int side = nc; side < nsides; side++) {
        int sidedim;
        double T[3][3], Tsum[3];
        side_transform(topo, side, ids, &sidedim, T, Tsum);
        const int ord1 = p - 1;
        int numshape = ord1;
        if (sidedim == 2) numshape = ord1 * ord1;
        if (sidedim == 3) numshape = ord1 * ord1 * ord1;
        if (numshape == 0) continue;
        phi[shape] = bphi[side];
        for (int k = 0; k < dim; k++) dphi_out[k * n + shape] = bd[k * nsides + side];
        shape++;
        if (numshape == 1) continue;
        double out[3] = {0, 0, 0};
        for (int i = 0; i < sidedim; i++) { /* TPZTransform::Apply, Matrix/pztrnsform.cpp:119-135 */
            double v = Tsum[i];
            for (int j = 0; j < dim; j++) v += T[i][j] * pt[j];
            out[i] = v;
        }
        double c[3][ORC_MAXP], dc[3][ORC_MAXP];
        for (int i = 0; i < sidedim; i++) chebyshev(out[i], ord1, c[i], dc[i]);
        for (int idx = 1; idx < numshape; idx++) {
            double pn, dn[3] = {0, 0, 0};
            if (sidedim == 1) { pn = c[0][idx]; dn[0] = dc[0][idx]; }
            else if (sidedim == 2) {
                const int i = idx / ord1, j = idx % ord1;
                pn = c[0][i] * c[1][j];
                dn[0] = dc[0][i] * c[1][j];
                dn[1] = c[0][i] * dc[1][j];
            } else {
                const int i = idx / (ord1 * ord1), j = (idx / ord1) % ord1, k = idx % ord1;
                pn = c[0][i] * c[1][j] * c[2][k];
                dn[0] = dc[0][i] * c[1][j] * c[2][k];
                dn[1] = c[0][i] * dc[1][j] * c[2][k];
                dn[2] = c[0][i] * c[1][j] * dc[2][k];
            }
            phi[shape] = bphi[side] * pn;
            for (int xj = 0; xj < dim; xj++) {
                double aux;
                if (sidedim < 3) {
                    aux = 0.;
                    for (int s2 = 0; s2 < sidedim; s2++) aux += T[s2][xj] * dn[s2];
                } else aux = dn[xj];
                dphi_out[xj * n + shape] = bd[xj * nsides + side] * pn + bphi[side] * aux;
            }
            shape++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Tetrahedra / triangles of order >= 3 (TPZShapeH1<TSHAPE>::Shape, Shape/TPZShapeH1.cpp:42-116):
 *   blends: ShapeGenerating of ALL sides, Shape/pzshapetetra.cpp:101-164 (edges x4, faces phi_a phi_b phi_c x27, interior
 *   phi_0 phi_1 phi_2 phi_3 x54), Shape/pzshapetriang.cpp:53-81 (edges x4, interior x27);
 *   NConnectShapeF: Shape/pzshapetetra.cpp:451-468, pzshapetriang.cpp:452-468;
 *   internal functions at the transformed point: edges Chebyshev T_i(x) (pzshapelinear.cpp:306-312), triangular sides
 *   T_i(2 x0 - 1) T_j(2 x1 - 1), i + j < p - 2 ordered by (i + j, j) (pzshapetriang.cpp:279-311), tetrahedron interior
 *   T_i T_j T_k at 2 x - 1, i + j + k < p - 3 (pzshapetetra.cpp:345-370).
 * ------------------------------------------------------------------------------------------ */
static int simplex_nconnect(int topo, int side, int p) {
    if (topo == ORC_TET) {
        if (side < 4) return 1;
        if (side < 10) return p - 1;
        if (side < 14) return (p - 2) * (p - 1) / 2;
        int tot = 0;
        for (int i = 1; i < p - 2; i++) tot += i * (i + 1) / 2;
        return tot;
    }
    if (side < 3) return 1;
    if (side < 6) return p - 1;
    return (p - 2) < 0 ? 0 : ((p - 2) * (p - 1)) / 2;
}

static int shape_simplex_general(int topo, int p, const int64_t *ids, const double *pt, double *phi, double *dphi_out) {
    const int dim = topo == ORC_TET ? 3 : 2, nc = dim + 1, nsides = topo == ORC_TET ? 15 : 7;
    double bphi[15], bd[3][15];
    memset(bd, 0, sizeof(bd));
    bphi[0] = 1.;
    for (int k = 0; k < dim; k++) bphi[0] -= pt[k]; /* 1 - pt[0] - pt[1] (- pt[2]) */
    for (int k = 0; k < dim; k++) { bphi[k + 1] = pt[k]; bd[k][0] = -1.; bd[k][k + 1] = 1.; }
    if (topo == ORC_TET) {
        for (int is = 4; is < 10; is++) {
            const int a = tet_edge_nodes[is - 4][0], b = tet_edge_nodes[is - 4][1];
            bphi[is] = bphi[a] * bphi[b];
            for (int k = 0; k < 3; k++) bd[k][is] = bd[k][a] * bphi[b] + bphi[a] * bd[k][b];
        }
        for (int is = 10; is < 14; is++) {
            const int a = tet_face_nodes[is - 10][0], b = tet_face_nodes[is - 10][1], c = tet_face_nodes[is - 10][2];
            bphi[is] = bphi[a] * bphi[b] * bphi[c];
            for (int k = 0; k < 3; k++)
                bd[k][is] = bd[k][a] * bphi[b] * bphi[c] + bphi[a] * bd[k][b] * bphi[c] + bphi[a] * bphi[b] * bd[k][c];
        }
        bphi[14] = bphi[0] * bphi[1] * bphi[2] * bphi[3];
        for (int k = 0; k < 3; k++)
            bd[k][14] = bd[k][0] * bphi[1] * bphi[2] * bphi[3] + bphi[0] * bd[k][1] * bphi[2] * bphi[3] +
                        bphi[0] * bphi[1] * bd[k][2] * bphi[3] + bphi[0] * bphi[1] * bphi[2] * bd[k][3];
        for (int is = 4; is < 15; is++) {
            const double mult = is < 10 ? 4. : (is < 14 ? 27. : 54.);
            bphi[is] *= mult;
            for (int k = 0; k < 3; k++) bd[k][is] *= mult;
        }
    } else {
        for (int is = 3; is < 6; is++) {
            const int a = is % 3, b = (is + 1) % 3;
            bphi[is] = bphi[a] * bphi[b];
            for (int k = 0; k < 2; k++) bd[k][is] = bd[k][a] * bphi[b] + bphi[a] * bd[k][b];
        }
        bphi[6] = bphi[0] * bphi[1] * bphi[2];
        for (int k = 0; k < 2; k++)
            bd[k][6] = bd[k][0] * bphi[1] * bphi[2] + bphi[0] * bd[k][1] * bphi[2] + bphi[0] * bphi[1] * bd[k][2];
        for (int is = 3; is < 7; is++) {
            const double mult = is < 6 ? 4. : 27.;
            bphi[is] *= mult;
            for (int k = 0; k < 2; k++) bd[k][is] *= mult;
        }
    }
    int n = nc;
    for (int side = nc; side < nsides; side++) n += simplex_nconnect(topo, side, p);
    for (int a = 0; a < nc; a++) { phi[a] = bphi[a]; for (int k = 0; k < dim; k++) dphi_out[k * n + a] = bd[k][a]; }
    int shape = nc;
    for (int side = nc; side < nsides; side++) {
        const int numshape = simplex_nconnect(topo, side, p);
        if (numshape == 0) continue;
        phi[shape] = bphi[side];
        for (int k = 0; k < dim; k++) dphi_out[k * n + shape] = bd[k][side];
        shape++;
        if (numshape == 1) continue;
        int sidedim;
        double T[3][3], Tsum[3];
        side_transform(topo, side, ids, &sidedim, T, Tsum);
        double out[3] = {0, 0, 0};
        for (int i = 0; i < sidedim; i++) {
            double v = Tsum[i];
            for (int j = 0; j < dim; j++) v += T[i][j] * pt[j];
            out[i] = v;
        }
        static __thread double pn[ORC_MAXSHAPE], dn[3][ORC_MAXSHAPE];
        double c[3][64], dc[3][64];
        if (sidedim == 1) {
            chebyshev(out[0], p - 1, c[0], dc[0]);
            for (int i = 0; i < p - 1; i++) { pn[i] = c[0][i]; dn[0][i] = dc[0][i]; }
        } else if (sidedim == 2) {
            const int ns2 = ((p - 2) * (p - 1)) / 2;
            if (ns2 > 64) return -1;
            chebyshev(2. * out[0] - 1., ns2, c[0], dc[0]);
            chebyshev(2. * out[1] - 1., ns2, c[1], dc[1]);
            int index = 0;
            for (int iplusj = 0; iplusj < p - 2; iplusj++)
                for (int j = 0; j <= iplusj; j++) {
                    const int i = iplusj - j;
                    pn[index] = c[0][i] * c[1][j];
                    dn[0][index] = 2.0 * dc[0][i] * c[1][j];
                    dn[1][index] = 2.0 * c[0][i] * dc[1][j];
                    index++;
                }
        } else {
            const int ord = p - 3;
            for (int k = 0; k < 3; k++) chebyshev(2. * out[k] - 1., ord, c[k], dc[k]);
            int index = 0;
            for (int i = 0; i < ord; i++)
                for (int j = 0; j < ord; j++)
                    for (int k = 0; k < ord; k++)
                        if (i + j + k < ord) {
                            pn[index] = c[0][i] * c[1][j] * c[2][k];
                            dn[0][index] = 2. * dc[0][i] * c[1][j] * c[2][k];
                            dn[1][index] = 2. * c[0][i] * dc[1][j] * c[2][k];
                            dn[2][index] = 2. * c[0][i] * c[1][j] * dc[2][k];
                            index++;
                        }
        }
        for (int idx = 1; idx < numshape; idx++) {
            phi[shape] = bphi[side] * pn[idx];
            for (int xj = 0; xj < dim; xj++) {
                double aux;
                if (sidedim < 3) {
                    aux = 0.;
                    for (int s2 = 0; s2 < sidedim; s2++) aux += T[s2][xj] * dn[s2][idx];
                } else aux = dn[xj][idx];
                dphi_out[xj * n + shape] = bd[xj][side] * pn[idx] + bphi[side] * aux;
            }
            shape++;
        }
    }
    return n;
}

/* ids: global corner-node indices (gel->NodeIndex, Mesh/TPZCompElH1.cpp:110); may be NULL for p <= 2 */
int orc_shape_ids(int topo, int p, const int64_t *ids, const double *pt, double *phi, double *dphi) {
    if (p < 1) return -1;
    if (p <= 2) {
        switch (topo) {
            case ORC_HEX: return shape_hex(p, pt, phi, dphi);
            case ORC_TET: return shape_tet(p, pt, phi, dphi);
            case ORC_QUAD: return shape_quad(p, pt, phi, dphi);
            case ORC_TRI: return shape_tri(p, pt, phi, dphi);
            case ORC_LINE: return shape_line(p, pt, phi, dphi);
            case ORC_PRISM: return shape_prism(p, pt, phi, dphi);
            case ORC_PYR: return shape_pyr(p, pt, phi, dphi);
        }
        return -1;
    }
    if (p > ORC_MAXP || !ids) return -1;
    if (topo == ORC_HEX || topo == ORC_QUAD || topo == ORC_LINE) return shape_hq_general(topo, p, ids, pt, phi, dphi);
    if (topo == ORC_TET || topo == ORC_TRI) return shape_simplex_general(topo, p, ids, pt, phi, dphi);
    return -1; /* prisms / pyramids of order >= 3: not restated */
}

int orc_shape(int topo, int p, const double *pt, double *phi, double *dphi) {
    if (p < 1 || p > 2) return -1;
    return orc_shape_ids(topo, p, 0, pt, phi, dphi);
}

/* ------------------------------------------------------------------------------------------
 * Geometry: gradient of the (multi)linear map.  coords[node][3].
 *   hex : Geom/TPZGeoCube.h:123-151 with Topology/tpzcube.cpp:367-417 (TShape)
 *   quad: Geom/pzgeoquad.h:144-172 with Topology/tpzquadrilateral.cpp:153-173
 *   tet : Geom/pzgeotetrahedra.h:106-151 ; tri: Geom/pzgeotriangle.h:150-179
 * gradx[3][dim]
 * ------------------------------------------------------------------------------------------ */
static void gradx_of(int topo, const double *coords, const double *pt, double gradx[3][3]) {
    double ph[8], d[3][27];
    int nn = 0, dim = 3;
    memset(d, 0, sizeof(d));
    if (topo == ORC_HEX) {
        hex_corner(pt, ph, d);
        nn = 8;
    } else if (topo == ORC_QUAD) {
        const double qsi = pt[0], eta = pt[1];
        d[0][0] = 0.25 * (eta - 1.); d[1][0] = 0.25 * (qsi - 1.);
        d[0][1] = 0.25 * (1. - eta); d[1][1] = -0.25 * (1. + qsi);
        d[0][2] = 0.25 * (1. + eta); d[1][2] = 0.25 * (1. + qsi);
        d[0][3] = -0.25 * (1. + eta); d[1][3] = 0.25 * (1. - qsi);
        nn = 4; dim = 2;
    } else if (topo == ORC_TET) {
        static const double dc[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int a = 0; a < 4; a++) for (int k = 0; k < 3; k++) d[k][a] = dc[a][k];
        nn = 4;
    } else if (topo == ORC_PRISM) { /* Geom/pzgeoprism.h:132-158 with Topology/tpzprism.cpp:323-351 (TShape) */
        prism_corner(pt, ph, d);
        nn = 6;
    } else if (topo == ORC_PYR) { /* Geom/pzgeopyramid.h:127-156 with Topology/tpzpyramid.cpp:282-340 (TShape: the same
                                     gradients as ShapeCorner away from the apex) */
        pyr_corner(pt, ph, d);
        nn = 5;
    } else if (topo == ORC_LINE) { /* Geom/pzgeolinear.h: x = sum_a x_a phi_a, phi = (1 -+ xi)/2 */
        d[0][0] = -0.5; d[0][1] = 0.5;
        nn = 2; dim = 1;
    } else {
        d[0][0] = -1.; d[1][0] = -1.; d[0][1] = 1.; d[1][1] = 0.; d[0][2] = 0.; d[1][2] = 1.;
        nn = 3; dim = 2;
    }
    for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) gradx[j][k] = 0.;
    for (int i = 0; i < nn; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < dim; k++) gradx[j][k] += coords[3 * i + j] * d[k][i];
}

/* data.x of a boundary face: TPZGeoEl::X (Mesh/pzinterpolationspace.cpp:278) = sum_j phi_j(qsi) x_j with the corner functions
 * TShape of the face, accumulated from 0 in node order (Geom/pzgeoquad.h:127-141 with Topology/tpzquadrilateral.cpp:153-160;
 * Geom/pzgeotriangle.h with Topology/tpztriangle.cpp:25-29).  coords[node][3] -> x[3].  Quadrilaterals / triangles only. */
int orc_point_x(int topo, const double *coords, const double *pt, double *x) {
    double phi[4];
    int nn;
    if (topo == ORC_QUAD) {
        const double qsi = pt[0], eta = pt[1];
        phi[0] = 0.25 * (1. - qsi) * (1. - eta);
        phi[1] = 0.25 * (1. + qsi) * (1. - eta);
        phi[2] = 0.25 * (1. + qsi) * (1. + eta);
        phi[3] = 0.25 * (1. - qsi) * (1. + eta);
        nn = 4;
    } else if (topo == ORC_TRI) {
        phi[0] = 1.0 - pt[0] - pt[1];
        phi[1] = pt[0];
        phi[2] = pt[1];
        nn = 3;
    } else {
        return -1;
    }
    for (int i = 0; i < 3; i++) {
        x[i] = 0.0;
        for (int j = 0; j < nn; j++) x[i] += phi[j] * coords[3 * j + i];
    }
    return 0;
}

/* Mesh/pzgeoel.cpp:1167-1356: 3-D branch :1296-1344, 2-D Gram-Schmidt branch :1228-1295.
 * Returns detjac (signed), jacinv[dim][dim]. */
static double jacobian_of(int dim, double gradx[3][3], double jacinv[3][3], double axes[3][3]) {
    double detjac = 0.0;
    if (dim == 1) { /* Mesh/pzgeoel.cpp:1185-1225 */
        double n1 = 0.0;
        for (int i = 0; i < 3; i++) n1 += gradx[i][0] * gradx[i][0];
        n1 = sqrt(n1);
        detjac = n1;
        if (fabs(detjac) < 1.e-12) detjac = 1.e-12;
        jacinv[0][0] = 1.0 / detjac;
        for (int i = 0; i < 3; i++) axes[0][i] = gradx[i][0] / n1;
        return detjac;
    }
    if (dim == 3) {
        double (*jac)[3] = gradx;
        detjac -= jac[0][2] * jac[1][1] * jac[2][0];
        detjac += jac[0][1] * jac[1][2] * jac[2][0];
        detjac += jac[0][2] * jac[1][0] * jac[2][1];
        detjac -= jac[0][0] * jac[1][2] * jac[2][1];
        detjac -= jac[0][1] * jac[1][0] * jac[2][2];
        detjac += jac[0][0] * jac[1][1] * jac[2][2];
        if (fabs(detjac) < 1.e-12) detjac = 1.e-12; /* IsZero -> ZeroTolerance(), Common/pzreal.h */
        jacinv[0][0] = (-jac[1][2] * jac[2][1] + jac[1][1] * jac[2][2]) / detjac;
        jacinv[0][1] = (jac[0][2] * jac[2][1] - jac[0][1] * jac[2][2]) / detjac;
        jacinv[0][2] = (-jac[0][2] * jac[1][1] + jac[0][1] * jac[1][2]) / detjac;
        jacinv[1][0] = (jac[1][2] * jac[2][0] - jac[1][0] * jac[2][2]) / detjac;
        jacinv[1][1] = (-jac[0][2] * jac[2][0] + jac[0][0] * jac[2][2]) / detjac;
        jacinv[1][2] = (jac[0][2] * jac[1][0] - jac[0][0] * jac[1][2]) / detjac;
        jacinv[2][0] = (-jac[1][1] * jac[2][0] + jac[1][0] * jac[2][1]) / detjac;
        jacinv[2][1] = (jac[0][1] * jac[2][0] - jac[0][0] * jac[2][1]) / detjac;
        jacinv[2][2] = (-jac[0][1] * jac[1][0] + jac[0][0] * jac[1][1]) / detjac;
        return detjac;
    }
    /* dim == 2 */
    double v1[3], v2[3], v1t[3], v2t[3];
    for (int i = 0; i < 3; i++) { v1[i] = gradx[i][0]; v2[i] = gradx[i][1]; }
    double n1 = 0.0, n2 = 0.0, dot = 0.0;
    for (int i = 0; i < 3; i++) { n1 += v1[i] * v1[i]; dot += v1[i] * v2[i]; }
    n1 = sqrt(n1);
    for (int i = 0; i < 3; i++) {
        v1t[i] = v1[i] / n1;
        v2t[i] = v2[i] - dot * v1t[i] / n1;
        n2 += v2t[i] * v2t[i];
    }
    n2 = sqrt(n2);
    const double j00 = n1, j01 = dot / n1, j10 = 0.0, j11 = n2;
    detjac = j00 * j11 - j10 * j01;
    jacinv[0][0] = +j11 / detjac;
    jacinv[1][1] = +j00 / detjac;
    jacinv[0][1] = -j01 / detjac;
    jacinv[1][0] = -j10 / detjac;
    if (fabs(detjac) < 1.e-12) detjac = 1.e-12;
    for (int i = 0; i < 3; i++) { /* Mesh/pzgeoel.cpp:1287-1291 */
        v2t[i] /= n2;
        axes[0][i] = v1t[i];
        axes[1][i] = v2t[i];
    }
    return detjac;
}

/* ------------------------------------------------------------------------------------------
 * Element descriptor + CalcStiff.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t topo, p, kind, bctype;
    double coords[24]; /* [node][3] */
    /* Poisson: mat[0]=fScale, mat[1]=constant source f (what the forcing std::function returns)
     * Elasticity3D: mat[0..2]=C1,C2,C3 ; mat[3..5]=fForce ; mat[6..8]=fPreStress
     * BC: mat[0]=big number ; mat[1..9]=val1 (row-major ns x ns) ; mat[10..12]=val2 ; mat[13]=fScale */
    double mat[16];
    int32_t nq, pad;
    const double *qpts; /* [nq][dim] */
    const double *qw;   /* [nq] */
    int64_t ids[8];     /* global corner-node indices (orientation of the sides, p >= 3) */
    const double *bcval2; /* optional [nq][3]: values of the material's std::function at every integration point (evaluated at data.x
                             by the caller).  Boundary kinds: val2 when the boundary condition carries a forcing function
                             (TPZMatPoisson.cpp:62-64 / TPZElasticity3D.cpp:637-662 / TPZElasticity2D.cpp:244-246), replacing
                             mat[10..12].  Domain kinds: the forcing function of the material - the source of TPZMatPoisson
                             (TPZMatPoisson.cpp:24-27, replaces mat[1]), the body force of TPZElasticity3D (:271-274, mat[3..5]) and of
                             TPZElasticity2D (:120-127, mat[3..4]).  NULL: the constants */
    double outward[3];    /* boundary faces: centre of the face minus centre of the neighbouring volume element, the vector
                             TPZInterpolationSpace::ComputeNormal orients data.normal with (Mesh/pzinterpolationspace.cpp:338-381);
                             used by TPZElasticity3D boundary type 4 only */
} orc_elem_t;

static int topo_dim(int topo) {
    return (topo == ORC_HEX || topo == ORC_TET || topo == ORC_PRISM || topo == ORC_PYR) ? 3 : (topo == ORC_LINE ? 1 : 2);
}

/* Material/Poisson/TPZMatPoisson.cpp:19-42 */
static void contribute_poisson(int fdim, int n, const double *phi, const double *dphix, double weight, const double *mat, double *ek, double *ef) {
    const double fScale = mat[0], force = mat[1];
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) {
            double s = 0;
            for (int x = 0; x < fdim; x++) s += dphix[x * n + i] * dphix[x * n + j];
            ek[j * n + i] += weight * fScale * s;
        }
        ef[i] += weight * fScale * phi[i] * force;
    }
}

/* Material/Elasticity/TPZElasticity3D.cpp:269-372 (CODE3 branch); ek column-major ndof x ndof */
static void contribute_elast(int n, const double *phi, const double *dphi, double weight, const double *mat, double *ek, double *ef) {
    const double C1 = mat[0], C2 = mat[1], C3 = mat[2];
    const double *locForce = mat + 3, *fPreStress = mat + 6;
    const int nd = 3 * n;
#define EK(i, j) ek[(size_t)(j) * nd + (i)]
    for (int jn = 0; jn < n; jn++) {
        double dphij[3];
        for (int kd = 0; kd < 3; kd++) {
            dphij[kd] = dphi[kd * n + jn];
            ef[jn * 3 + kd] += weight * (locForce[kd] * phi[jn] - fPreStress[kd] * dphi[kd * n + jn]);
        }
        for (int in = 0; in < n; in++) {
            double D[3][3];
            for (int ud = 0; ud < 3; ud++)
                for (int vd = 0; vd < 3; vd++) D[vd][ud] = dphi[vd * n + in] * dphij[ud];
            EK(in * 3 + 0, jn * 3 + 0) += weight * ((D[1][1] + D[2][2]) * C1 + D[0][0] * C3);
            EK(in * 3 + 1, jn * 3 + 0) += weight * (D[0][1] * C1 - D[1][0] * C2);
            EK(in * 3 + 2, jn * 3 + 0) += weight * (D[0][2] * C1 - D[2][0] * C2);
            EK(in * 3 + 0, jn * 3 + 1) += weight * (D[1][0] * C1 - D[0][1] * C2);
            EK(in * 3 + 1, jn * 3 + 1) += weight * ((D[0][0] + D[2][2]) * C1 + D[1][1] * C3);
            EK(in * 3 + 2, jn * 3 + 1) += weight * (D[1][2] * C1 - D[2][1] * C2);
            EK(in * 3 + 0, jn * 3 + 2) += weight * (D[2][0] * C1 - D[0][2] * C2);
            EK(in * 3 + 1, jn * 3 + 2) += weight * (D[2][1] * C1 - D[1][2] * C2);
            EK(in * 3 + 2, jn * 3 + 2) += weight * ((D[0][0] + D[1][1]) * C1 + D[2][2] * C3);
        }
    }
#undef EK
}

/* Material/Poisson/TPZMatPoisson.cpp:45-121 (types 0, 1 and 2 as computed there: the Robin branch adds the penalty load vector
 * and BigNumber * Val1(0,0) * dphix(0,in) * dphix(0,jn), then falls through to the "not implemented" message) */
static int contribute_poisson_bc(int n, const double *phi, const double *dphix, double weight, int type, const double *mat, double *ek, double *ef) {
    const double big = mat[0], v2 = mat[10], fScale = mat[13];
    if (type == 2) {
        const double v1 = mat[1];
        for (int in = 0; in < n; in++) {
            ef[in] += big * v2 * phi[in] * weight;
            for (int jn = 0; jn < n; jn++) ek[jn * n + in] += big * v1 * dphix[in] * dphix[jn] * weight;
        }
        return 0;
    }
    if (type == 0) {
        for (int in = 0; in < n; in++) {
            ef[in] += big * v2 * phi[in] * weight;
            for (int jn = 0; jn < n; jn++) ek[jn * n + in] += big * phi[in] * phi[jn] * weight;
        }
        return 0;
    }
    if (type == 1) {
        for (int in = 0; in < n; in++) ef[in] += v2 * fScale * phi[in] * weight;
        return 0;
    }
    return -1;
}

/* Material/Elasticity/TPZElasticity3D.cpp:616-773 (types 0,1,2), BIGNUMBER=1e12 (:630) */
static int contribute_elast_bc(int n, const double *phi, double weight, int type, const double *mat, const double axes[3][3],
                               const double *outward, double *ek, double *ef) {
    const double BIG = 1.e12;
    const double *val1 = mat + 1, *val2 = mat + 10;
    const int nd = 3 * n;
#define EK(i, j) ek[(size_t)(j) * nd + (i)]
    switch (type) {
        case 0:
            for (int in = 0; in < n; in++) {
                for (int k = 0; k < 3; k++) ef[3 * in + k] += BIG * val2[k] * phi[in] * weight;
                for (int jn = 0; jn < n; jn++)
                    for (int k = 0; k < 3; k++) EK(3 * in + k, 3 * jn + k) += BIG * phi[in] * phi[jn] * weight;
            }
            return 0;
        case 1:
            for (int in = 0; in < n; in++)
                for (int k = 0; k < 3; k++) ef[3 * in + k] += val2[k] * phi[in] * weight;
            return 0;
        case 2:
            for (int in = 0; in < n; in++) {
                for (int k = 0; k < 3; k++) ef[3 * in + k] += val2[k] * phi[in] * weight;
                for (int jn = 0; jn < n; jn++)
                    for (int idf = 0; idf < 3; idf++)
                        for (int jdf = 0; jdf < 3; jdf++)
                            EK(3 * in + idf, 3 * jn + jdf) += val1[3 * idf + jdf] * weight * phi[in] * phi[jn];
            }
            return 0;
        case 3: /* directional null Dirichlet, TPZElasticity3D.cpp:715-723 */
            for (int in = 0; in < n; in++)
                for (int jn = 0; jn < n; jn++)
                    for (int k = 0; k < 3; k++) EK(3 * in + k, 3 * jn + k) += BIG * phi[in] * phi[jn] * weight * val2[k];
            return 0;
        case 4: { /* stress-field Neumann, :724-737, with data.normal of Mesh/pzinterpolationspace.cpp:300-383:
                     VectorialProd(axes(0), axes(1), normal, unitary) turned towards `vec` */
            double nrm[3];
            nrm[0] = axes[0][1] * axes[1][2] - axes[0][2] * axes[1][1];
            nrm[1] = -axes[0][0] * axes[1][2] + axes[0][2] * axes[1][0];
            nrm[2] = axes[0][0] * axes[1][1] - axes[0][1] * axes[1][0];
            double size = 0.;
            for (int i = 0; i < 3; i++) size += nrm[i] * nrm[i];
            size = sqrt(size);
            for (int i = 0; i < 3; i++) nrm[i] /= size;
            double dot = 0.;
            for (int i = 0; i < 3; i++) dot += nrm[i] * outward[i];
            if (dot < 0.)
                for (int i = 0; i < 3; i++) nrm[i] *= -1.;
            double v2l[3];
            for (int i = 0; i < 3; i++) v2l[i] = -(val1[3 * i + 0] * nrm[0] + val1[3 * i + 1] * nrm[1] + val1[3 * i + 2] * nrm[2]);
            for (int in = 0; in < n; in++)
                for (int k = 0; k < 3; k++) ef[3 * in + k] += v2l[k] * phi[in] * weight;
            return 0;
        }
        case 5: case 6: case 7: case 8: { /* directional Dirichlet on x / y / z / x and z, :739-772 */
            const int on[3] = {type == 5 || type == 8, type == 6, type == 7 || type == 8};
            for (int in = 0; in < n; in++) {
                for (int k = 0; k < 3; k++)
                    if (on[k]) ef[3 * in + k] += BIG * val2[k] * phi[in] * weight;
                for (int jn = 0; jn < n; jn++)
                    for (int k = 0; k < 3; k++)
                        if (on[k]) EK(3 * in + k, 3 * jn + k) += BIG * phi[in] * phi[jn] * weight;
            }
            return 0;
        }
    }
#undef EK
    return -1;
}

/* Material/Elasticity/TPZElasticity2D.cpp:86-203.  mat[0]=E, mat[1]=nu, mat[2]=fPlaneStress, mat[3..4]=ff,
 * mat[5..7]=fPreStressXX, XY, YY.  dphi = dphix in the element's axes, axes[2][3]. */
static void contribute_elast2d(int n, const double *phi, const double *dphi, double axes[3][3], double weight, const double *mat,
                               double *ek, double *ef) {
    const double E = mat[0], nu = mat[1];
    const int planestress = mat[2] != 0.0;
    const double floc[2] = {mat[3], mat[4]};
    const double sxx = mat[5], sxy = mat[6], syy = mat[7];
    const double Eover1MinNu2 = E / (1 - nu * nu);
    const double Eover21PlusNu = E / (2. * (1 + nu));
    const double nu1 = 1. - nu;
    const double nu2 = (1. - 2. * nu) / 2.;
    const double F = E / ((1. + nu) * (1. - 2. * nu));
    const int nd = 2 * n;
#define EK(i, j) ek[(size_t)(j) * nd + (i)]
    for (int in = 0; in < n; in++) {
        const double du00 = dphi[0 * n + in] * axes[0][0] + dphi[1 * n + in] * axes[1][0];
        const double du10 = dphi[0 * n + in] * axes[0][1] + dphi[1 * n + in] * axes[1][1];
        ef[2 * in] += weight * (floc[0] * phi[in] - du00 * sxx - du10 * sxy);
        ef[2 * in + 1] += weight * (floc[1] * phi[in] - du00 * sxy - du10 * syy);
        for (int jn = 0; jn < n; jn++) {
            const double du01 = dphi[0 * n + jn] * axes[0][0] + dphi[1 * n + jn] * axes[1][0];
            const double du11 = dphi[0 * n + jn] * axes[0][1] + dphi[1 * n + jn] * axes[1][1];
            if (!planestress) {
                EK(2 * in, 2 * jn) += weight * (nu1 * du00 * du01 + nu2 * du10 * du11) * F;
                EK(2 * in, 2 * jn + 1) += weight * (nu * du00 * du11 + nu2 * du10 * du01) * F;
                EK(2 * in + 1, 2 * jn) += weight * (nu * du10 * du01 + nu2 * du00 * du11) * F;
                EK(2 * in + 1, 2 * jn + 1) += weight * (nu1 * du10 * du11 + nu2 * du00 * du01) * F;
            } else {
                EK(2 * in, 2 * jn) += weight * (Eover1MinNu2 * du00 * du01 + Eover21PlusNu * du10 * du11);
                EK(2 * in, 2 * jn + 1) += weight * (Eover1MinNu2 * nu * du00 * du11 + Eover21PlusNu * du10 * du01);
                EK(2 * in + 1, 2 * jn) += weight * (Eover1MinNu2 * nu * du10 * du01 + Eover21PlusNu * du00 * du11);
                EK(2 * in + 1, 2 * jn + 1) += weight * (Eover1MinNu2 * du10 * du11 + Eover21PlusNu * du00 * du01);
            }
        }
    }
#undef EK
}

/* Material/Elasticity/TPZElasticity2D.cpp:205-330: types 0 (Dirichlet), 1 (Neumann), 2 (mixed), 3 (directional null
 * Dirichlet).  mat[0] = TPZMaterial::fBigNumber, mat[1..9] = val1 (3x3 row-major, 2x2 used), mat[10..11] = val2 */
static int contribute_elast2d_bc(int n, const double *phi, double weight, int type, const double *mat, double *ek, double *ef) {
    const double BIG = mat[0];
    const double *val1 = mat + 1, *val2 = mat + 10;
    const int nd = 2 * n;
#define EK(i, j) ek[(size_t)(j) * nd + (i)]
    switch (type) {
        case 0:
            for (int in = 0; in < n; in++) {
                ef[2 * in] += BIG * val2[0] * phi[in] * weight;
                ef[2 * in + 1] += BIG * val2[1] * phi[in] * weight;
                for (int jn = 0; jn < n; jn++) {
                    EK(2 * in, 2 * jn) += BIG * phi[in] * phi[jn] * weight;
                    EK(2 * in + 1, 2 * jn + 1) += BIG * phi[in] * phi[jn] * weight;
                }
            }
            return 0;
        case 1:
            for (int in = 0; in < n; in++) {
                ef[2 * in] += val2[0] * phi[in] * weight;
                ef[2 * in + 1] += val2[1] * phi[in] * weight;
            }
            return 0;
        case 2:
            for (int in = 0; in < n; in++) {
                ef[2 * in] += val2[0] * phi[in] * weight;
                ef[2 * in + 1] += val2[1] * phi[in] * weight;
                for (int jn = 0; jn < n; jn++) {
                    EK(2 * in, 2 * jn) += val1[0] * phi[in] * phi[jn] * weight;
                    EK(2 * in + 1, 2 * jn) += val1[3] * phi[in] * phi[jn] * weight;
                    EK(2 * in + 1, 2 * jn + 1) += val1[4] * phi[in] * phi[jn] * weight;
                    EK(2 * in, 2 * jn + 1) += val1[1] * phi[in] * phi[jn] * weight;
                }
            }
            return 0;
        case 3:
            for (int in = 0; in < n; in++)
                for (int jn = 0; jn < n; jn++) {
                    EK(2 * in, 2 * jn) += BIG * phi[in] * phi[jn] * weight * val2[0];
                    EK(2 * in + 1, 2 * jn + 1) += BIG * phi[in] * phi[jn] * weight * val2[1];
                }
            return 0;
    }
#undef EK
    return -1;
}

/* Mesh/pzinterpolationspace.cpp:404-473 (quadrature loop), :266-297 (ComputeRequiredData),
 * Mesh/TPZCompElH1.cpp:140-149 (dphix = jacinv^T dphi via Matrix/pzfmatrix.cpp:608-622).
 * ek: column-major ndof x ndof (zeroed here), ef: ndof. returns ndof or <0 */
int orc_calcstiff(const orc_elem_t *e, double *ek, double *ef) {
    const int dim = topo_dim(e->topo);
    const int ns = (e->kind == ORC_ELAST3D || e->kind == ORC_ELAST3D_BC) ? 3 : ((e->kind == ORC_ELAST2D || e->kind == ORC_ELAST2D_BC) ? 2 : 1);
    static __thread double phi[ORC_MAXSHAPE], dphi[3 * ORC_MAXSHAPE], dphix[3 * ORC_MAXSHAPE];
    double pt0[3] = {0, 0, 0};
    const int n = orc_shape_ids(e->topo, e->p, e->ids, pt0, phi, dphi);
    if (n < 0) return -1;
    const int nd = n * ns;
    memset(ek, 0, sizeof(double) * nd * nd);
    memset(ef, 0, sizeof(double) * nd);
    for (int q = 0; q < e->nq; q++) {
        const double *pt = e->qpts + (size_t)q * dim;
        double weight = e->qw[q];
        double gradx[3][3], jacinv[3][3], axes[3][3];
        gradx_of(e->topo, e->coords, pt, gradx);
        double detjac = jacobian_of(dim, gradx, jacinv, axes);
        detjac = fabs(detjac);
        orc_shape_ids(e->topo, e->p, e->ids, pt, phi, dphi);
        for (int j = 0; j < n; j++)
            for (int c = 0; c < dim; c++) {
                double val = 0.;
                for (int k = 0; k < dim; k++) val += jacinv[k][c] * dphi[k * n + j];
                dphix[c * n + j] = 0. + 1. * val;
            }
        weight *= fabs(detjac);
        int rc = 0;
        double bmat[16];
        const double *bm = e->mat;
        if (e->bcval2) { /* data from a function: the values of this point */
            const double *v = e->bcval2 + (size_t)q * 3;
            memcpy(bmat, e->mat, sizeof(bmat));
            if (e->kind == ORC_POISSON) bmat[1] = v[0];
            else if (e->kind == ORC_ELAST3D) { bmat[3] = v[0]; bmat[4] = v[1]; bmat[5] = v[2]; }
            else if (e->kind == ORC_ELAST2D) { bmat[3] = v[0]; bmat[4] = v[1]; }
            else for (int k = 0; k < 3; k++) bmat[10 + k] = v[k];
            bm = bmat;
        }
        switch (e->kind) {
            case ORC_POISSON: contribute_poisson(dim, n, phi, dphix, weight, bm, ek, ef); break;
            case ORC_ELAST2D: contribute_elast2d(n, phi, dphix, axes, weight, bm, ek, ef); break;
            case ORC_ELAST2D_BC: rc = contribute_elast2d_bc(n, phi, weight, e->bctype, bm, ek, ef); break;
            case ORC_ELAST3D: contribute_elast(n, phi, dphix, weight, bm, ek, ef); break;
            case ORC_POISSON_BC: rc = contribute_poisson_bc(n, phi, dphix, weight, e->bctype, bm, ek, ef); break;
            case ORC_ELAST3D_BC: rc = contribute_elast_bc(n, phi, weight, e->bctype, bm, axes, e->outward, ek, ef); break;
            default: rc = -1;
        }
        if (rc) return -2;
    }
    return nd;
}

/* single-point Contribute of TPZElasticity3D, for the reference's known-answer test
 * UnitTest_PZ/TestMaterial/TestMaterial.cpp:18-40 (dphix = I3, weight = 8) */
void orc_elast_contribute_point(int n, const double *phi, const double *dphix, double weight, const double *mat, double *ek, double *ef) {
    contribute_elast(n, phi, dphix, weight, mat, ek, ef);
}

/* TPZElasticity3D::SetC, Material/Elasticity/TPZElasticity3D.h:183-188 */
void orc_elast_constants(double E, double nu, double *C) {
    C[0] = E / (2. + 2. * nu);
    C[1] = E * nu / (-1. + nu + 2. * nu * nu);
    C[2] = E * (nu - 1.) / (-1. + nu + 2. * nu * nu);
}

/* ------------------------------------------------------------------------------------------
 * CSR pattern.  Mesh/pzcmesh.cpp:1223-1267 (element graph = seqnums of each element's connects),
 * External/TPZRenumbering.cpp:30-110 (node->element graph, then node->node graph as an ascending
 * std::set with self erased), StrMatrix/TPZSSpStructMatrix.cpp:50-193 (symmetric, upper) and
 * StrMatrix/TPZSpStructMatrix.cpp:53-190 (full, rows sorted).  No equation filter.
 * Two-call protocol: ja == NULL -> only count; returns nnz.  ia has neq+1 entries.
 * ------------------------------------------------------------------------------------------ */
static int cmp_i64(const void *a, const void *b) {
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

int64_t orc_pattern(int symmetric, int64_t nel, const int64_t *elgraphindex, const int64_t *elgraph, int64_t nblock,
                    const int64_t *blockpos, const int64_t *blocksize, int64_t *ia, int64_t *ja) {
    /* NodeToElGraph */
    int64_t *n2e_idx = (int64_t *)calloc(nblock + 1, sizeof(int64_t));
    const int64_t last = elgraphindex[nel];
    for (int64_t k = 0; k < last; k++) n2e_idx[elgraph[k] + 1]++;
    for (int64_t b = 0; b < nblock; b++) n2e_idx[b + 1] += n2e_idx[b];
    int64_t *n2e = (int64_t *)malloc(sizeof(int64_t) * (last > 0 ? last : 1));
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (nblock + 1));
    memcpy(cursor, n2e_idx, sizeof(int64_t) * (nblock + 1));
    for (int64_t el = 0; el < nel; el++)
        for (int64_t k = elgraphindex[el]; k < elgraphindex[el + 1]; k++) n2e[cursor[elgraph[k]]++] = el;
    /* ConvertGraph + SetupMatrixData, block row by block row */
    int64_t cap = 4096, *buf = (int64_t *)malloc(sizeof(int64_t) * cap);
    int64_t pos = 0, ieq = 0;
    for (int64_t i = 0; i < nblock; i++) {
        int64_t cnt = 0;
        for (int64_t e = n2e_idx[i]; e < n2e_idx[i + 1]; e++) {
            const int64_t gel = n2e[e];
            for (int64_t k = elgraphindex[gel]; k < elgraphindex[gel + 1]; k++) {
                if (cnt == cap) { cap *= 2; buf = (int64_t *)realloc(buf, sizeof(int64_t) * cap); }
                buf[cnt++] = elgraph[k];
            }
        }
        qsort(buf, cnt, sizeof(int64_t), cmp_i64);
        int64_t m = 0;
        for (int64_t k = 0; k < cnt; k++)
            if (buf[k] != i && (m == 0 || buf[m - 1] != buf[k])) buf[m++] = buf[k];
        const int64_t iblsize = blocksize[i], iblpos = blockpos[i];
        if (iblsize == 0) continue; /* NumActive == 0 */
        for (int64_t ibleq = 0; ibleq < iblsize; ibleq++) {
            ia[ieq] = pos;
            if (symmetric) {
                for (int64_t j = 0; j < iblsize; j++) {
                    const int64_t jeq = iblpos + j;
                    if (jeq < ieq) continue;
                    if (ja) ja[pos] = jeq;
                    pos++;
                }
                for (int64_t k = 0; k < m; k++) {
                    const int64_t col = buf[k];
                    if (col < i) continue;
                    for (int64_t j = 0; j < blocksize[col]; j++) {
                        const int64_t jeq = blockpos[col] + j;
                        if (jeq < ieq) continue;
                        if (ja) ja[pos] = jeq;
                        pos++;
                    }
                }
            } else {
                const int64_t first = pos;
                for (int64_t j = 0; j < iblsize; j++) { if (ja) ja[pos] = iblpos + j; pos++; }
                for (int64_t k = 0; k < m; k++) {
                    const int64_t col = buf[k];
                    for (int64_t j = 0; j < blocksize[col]; j++) { if (ja) ja[pos] = blockpos[col] + j; pos++; }
                }
                if (ja) qsort(ja + first, pos - first, sizeof(int64_t), cmp_i64); /* std::stable_sort of distinct keys */
            }
            ieq++;
        }
    }
    ia[ieq] = pos;
    free(buf); free(cursor); free(n2e); free(n2e_idx);
    return pos;
}

/* ------------------------------------------------------------------------------------------
 * Scatter-add.  Matrix/pzsysmp.cpp:370-411 (sym: only jpos>=ipos, skip |v|<1e-12, sequential
 * hit k++ else linear search of the row), Matrix/pzysmp.cpp:178-218 (full),
 * Matrix/pzfmatrix.cpp:285-299 (AddFel).  Returns the number of entries NOT found (must be 0).
 * ------------------------------------------------------------------------------------------ */
int64_t orc_addkel(int symmetric, const int64_t *ia, const int64_t *ja, double *a, int nd, const double *ek, const int64_t *dest) {
    int64_t k = 0, missing = 0;
    for (int i = 0; i < nd; i++) {
        for (int j = 0; j < nd; j++) {
            const int64_t ipos = dest[i], jpos = dest[j];
            if (symmetric && jpos < ipos) continue;
            const double value = ek[(size_t)j * nd + i];
            if (!(fabs(value) < 1.e-12)) {
                int flag = 0;
                k++;
                if (k >= ia[ipos] && k < ia[ipos + 1] && ja[k] == jpos) {
                    a[k] += value;
                    flag = 1;
                } else {
                    for (k = ia[ipos]; k < ia[ipos + 1]; k++) {
                        if (ja[k] == jpos) { a[k] += value; flag = 1; break; }
                    }
                }
                if (!flag) missing++;
            }
        }
    }
    return missing;
}

void orc_addfel(double *rhs, int nd, const double *ef, const int64_t *dest) {
    for (int i = 0; i < nd; i++) rhs[dest[i]] += ef[i];
}

/* ------------------------------------------------------------------------------------------
 * Whole serial assembly (StrMatrix/pzstrmatrixor.cpp:104-365): elements in index order,
 * CalcStiff -> destination indices -> AddKel -> AddFel.
 * elems[nel]; dest_ptr[nel+1], dest[] concatenated (Mesh/pzelmat.cpp:37-70 done by the caller,
 * who owns the connect/block tables).  a and rhs must be zeroed by the caller.
 * ------------------------------------------------------------------------------------------ */
int64_t orc_assemble(int symmetric, int64_t nel, const orc_elem_t *elems, const int64_t *dest_ptr, const int64_t *dest,
                     const int64_t *ia, const int64_t *ja, double *a, double *rhs) {
    const size_t ndmax = 3 * ORC_MAXSHAPE;
    double *ek = (double *)malloc(sizeof(double) * ndmax * ndmax), *ef = (double *)malloc(sizeof(double) * ndmax);
    int64_t missing = 0;
    for (int64_t el = 0; el < nel; el++) {
        const int nd = orc_calcstiff(&elems[el], ek, ef);
        if (nd < 0) { missing = -1; break; }
        if (dest_ptr[el + 1] - dest_ptr[el] != nd) { missing = -2; break; }
        missing += orc_addkel(symmetric, ia, ja, a, nd, ek, dest + dest_ptr[el]);
        orc_addfel(rhs, nd, ef, dest + dest_ptr[el]);
    }
    free(ek); free(ef);
    return missing;
}

int orc_sizeof_elem(void) { return (int)sizeof(orc_elem_t); }
