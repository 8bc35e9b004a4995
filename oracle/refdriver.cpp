// TEST INFRASTRUCTURE (oracle) — not product code, never linked into the shipped library.
//
// Driver for the UNMODIFIED reference (NeoPZ, built by oracle/Makefile.ref into oracle/_ref/libpz.so).
// It only calls the reference's public API:
//   dump : builds a grid mesh, runs the reference's own assembly
//          (TPZLinearAnalysis::Assemble, Analysis/TPZLinearAnalysis.cpp:42-94, with
//          TPZSSpStructMatrix / TPZSpStructMatrix + TPZStructMatrixOR) and writes every
//          intermediate the parity tests need as .npy files (golden fixtures).
//   time : times the reference's threaded TPZStructMatrixOR assembly (second Assemble(),
//          pattern reused — Analysis/TPZLinearAnalysis.cpp:73-77) = the CPU baseline.
//
// Mesh recipe = SURVEY.md §8(d): TPZGeoMeshTools::CreateGeoMeshOnGrid on the unit cube,
// matid 1 volume, matid -1 on all six faces (Dirichlet type 0), optional smooth node
// perturbation x += amp*h*sin(2*pi*id/97 + axis) to make Jacobians non-constant.
#include "pzgmesh.h"
#include "pzcmesh.h"
#include "pzgnode.h"
#include "TPZGeoMeshTools.h"
#include "TPZLinearAnalysis.h"
#include "TPZSSpStructMatrix.h"
#include "TPZSpStructMatrix.h"
#include "pzskylstrmatrix.h"
#include "pzstepsolver.h"
#include "pzsysmp.h"
#include "pzysmp.h"
#include "pzintel.h"
#include "pzinterpolationspace.h"
#include "TPZElementMatrixT.h"
#include "TPZShapeH1.h"
#include "TPZShapeData.h"
#include "pzshapecube.h"
#include "pzshapetetra.h"
#include "pzshapequad.h"
#include "pzshapetriang.h"
#include "pzshapeprism.h"
#include "pzshapepiram.h"
#include "Poisson/TPZMatPoisson.h"
#include "Elasticity/TPZElasticity3D.h"
#include "Elasticity/TPZElasticity2D.h"
#include "pzshapelinear.h"
#include "TPZBndCond.h"
#include "TPZBndCondT.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using clk = std::chrono::steady_clock;

// ---------------------------------------------------------------- .npy writer
template <class T> struct NpyType;
template <> struct NpyType<double> { static const char *s() { return "<f8"; } };
template <> struct NpyType<int64_t> { static const char *s() { return "<i8"; } };
template <> struct NpyType<int32_t> { static const char *s() { return "<i4"; } };

template <class T>
static void save_npy(const std::string &path, const T *data, const std::vector<int64_t> &shape) {
    std::ostringstream h;
    h << "{'descr': '" << NpyType<T>::s() << "', 'fortran_order': False, 'shape': (";
    int64_t n = 1;
    for (size_t i = 0; i < shape.size(); i++) { h << shape[i] << (shape.size() == 1 || i + 1 < shape.size() ? "," : ""); n *= shape[i]; }
    h << "), }";
    std::string hs = h.str();
    size_t total = 10 + hs.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hs.append(pad, ' ');
    hs.push_back('\n');
    std::ofstream f(path, std::ios::binary);
    const char magic[] = {(char)0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    f.write(magic, 8);
    uint16_t hl = (uint16_t)hs.size();
    f.write((const char *)&hl, 2);
    f.write(hs.data(), hs.size());
    f.write((const char *)data, sizeof(T) * n);
}
template <class T>
static void save_vec(const std::string &dir, const std::string &name, const std::vector<T> &v, std::vector<int64_t> shape = {}) {
    if (shape.empty()) shape = {(int64_t)v.size()};
    save_npy<T>(dir + "/" + name + ".npy", v.data(), shape);
}

// ---------------------------------------------------------------- mesh recipe
struct Case {
    int n = 4, p = 1, phys = 0, tet = 0;  // tet: 0 hexahedra / quadrilaterals, 1 tetrahedra / triangles, 2 prisms (MMeshType::EPrismatic),
                                          // 3 hexahedra + pyramids (MMeshType::EHexaPyrMixed: every other cell split into six pyramids)
    double perturb = 0.0;
    int bctype = 0;  // type of the BC on matid -1 (0 Dirichlet, 1 Neumann on zmax only -> matid -2)
    int dim = 3;       // 2: plane mesh (TPZGenGrid2D): phys 0 = TPZMatPoisson(dim 2), phys 2 / 3 = TPZElasticity2D plane
                       // strain / plane stress; boundary = line elements, matid -2 on the top side when bctype >= 1
    int bcfunc = 0;    // 1: the Dirichlet data on matid -1 (and the data of matid -2 when bctype == 2) come from a function of x
                       // (TPZBndCondT::SetForcingFunctionBC); the functions are those of tests/golden_util.py
    int scramble = 0;  // != 0: node indices permuted by a seeded Fisher-Yates shuffle, so that the side
                       // orientations (transform ids, Shape/pzgenericshape.cpp:57-68) differ between elements
};
static std::vector<int64_t> g_node_perm;  // new index of every grid node (identity without scramble)

static TPZCompMesh *build_mesh(const Case &c) {
    TPZManVector<REAL, 3> minX(3, 0.), maxX(3, 1.);
    TPZManVector<int, 7> matids(c.dim == 3 ? 7 : 5, -1);
    matids[0] = 1;
    if (c.bctype >= 1) matids[c.dim == 3 ? 6 : 3] = -2;  // zmax face (3-D) / top side (2-D): Neumann (1) or BC type c.bctype (>= 2)
    TPZManVector<int, 3> ndiv(c.dim, c.n);
    TPZGeoMesh *gmesh = c.dim == 3
        ? TPZGeoMeshTools::CreateGeoMeshOnGrid(3, minX, maxX, matids, ndiv,
                                               c.tet == 1 ? MMeshType::ETetrahedral : (c.tet == 2 ? MMeshType::EPrismatic : (c.tet == 3 ? MMeshType::EHexaPyrMixed : MMeshType::EHexahedral)), true)
        : TPZGeoMeshTools::CreateGeoMeshOnGrid(2, minX, maxX, matids, ndiv, c.tet ? MMeshType::ETriangular : MMeshType::EQuadrilateral, true);
    if (c.perturb != 0.0) {
        const double h = 1.0 / c.n;
        const int64_t nn = gmesh->NNodes();
        for (int64_t i = 0; i < nn; i++) {
            TPZGeoNode &nd = gmesh->NodeVec()[i];
            for (int d = 0; d < c.dim; d++) {  // plane meshes stay in z = 0
                double x = nd.Coord(d);
                x += c.perturb * h * std::sin(2.0 * M_PI * (double)i / 97.0 + (double)d);
                nd.SetCoord(d, x);
            }
        }
    }
    g_node_perm.resize(gmesh->NNodes());
    for (int64_t i = 0; i < gmesh->NNodes(); i++) g_node_perm[i] = i;
    if (c.scramble) {
        const int64_t nn = gmesh->NNodes();
        uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)c.scramble + 12345u;
        for (int64_t i = nn - 1; i > 0; i--) {
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            const int64_t j = (int64_t)((st >> 33) % (uint64_t)(i + 1));
            std::swap(g_node_perm[i], g_node_perm[j]);
        }
        std::vector<TPZGeoNode> old(nn);
        for (int64_t i = 0; i < nn; i++) old[i] = gmesh->NodeVec()[i];
        for (int64_t i = 0; i < nn; i++) {
            TPZGeoNode &nd = gmesh->NodeVec()[g_node_perm[i]];
            nd = old[i];
            nd.SetNodeId((int)g_node_perm[i]);
        }
        for (int64_t e = 0; e < gmesh->NElements(); e++) {
            TPZGeoEl *gel = gmesh->Element(e);
            if (!gel) continue;
            for (int k = 0; k < gel->NNodes(); k++) gel->SetNodeIndex(k, g_node_perm[gel->NodeIndex(k)]);
        }
        gmesh->ResetConnectivities();
        gmesh->BuildConnectivity();
    }
    TPZCompMesh *cmesh = new TPZCompMesh(gmesh);
    cmesh->SetDimModel(c.dim);
    cmesh->SetDefaultOrder(c.p);
    if (c.phys >= 2) {
        // TPZElasticity2D: E = 1000, nu = 0.3, body force (0.5, -1), plane strain (phys 2) or plane stress (phys 3)
        // (the reference's six-argument constructor has an empty body, Material/Elasticity/TPZElasticity2D.cpp:49-53)
        auto *m = new TPZElasticity2D(1);
        m->SetElasticity(1000., 0.3);
        m->SetBodyForce(0.5, -1.0);
        if (c.phys == 3) m->SetPlaneStress(); else m->SetPlaneStrain();
        cmesh->InsertMaterialObject(m);
        TPZFNMatrix<4, STATE> v1(2, 2, 0.);
        TPZManVector<STATE, 2> v2(2, 0.);
        cmesh->InsertMaterialObject(m->CreateBC(m, -1, 0, v1, v2));
        if (c.bctype == 1) {
            TPZManVector<STATE, 2> v2n(2, 0.);
            v2n[0] = 0.25; v2n[1] = -0.5;
            cmesh->InsertMaterialObject(m->CreateBC(m, -2, 1, v1, v2n));
        } else if (c.bctype >= 2) {
            TPZFNMatrix<4, STATE> v1m(2, 2, 0.);
            v1m(0, 0) = 4.0; v1m(0, 1) = 0.5; v1m(1, 0) = 0.5; v1m(1, 1) = 3.0;
            TPZManVector<STATE, 2> v2m(2, 0.);
            v2m[0] = 0.3; v2m[1] = -0.2;
            cmesh->InsertMaterialObject(m->CreateBC(m, -2, c.bctype, v1m, v2m));
        }
    } else if (c.phys == 0) {
        auto *m = new TPZMatPoisson<STATE>(1, c.dim);
        m->SetForcingFunction([](const TPZVec<REAL> &x, TPZVec<STATE> &f) { f[0] = 1.0; }, 0);
        cmesh->InsertMaterialObject(m);
        TPZFNMatrix<1, STATE> v1(1, 1, 0.);
        TPZManVector<STATE, 1> v2(1, 0.);
        auto *bcd = m->CreateBC(m, -1, 0, v1, v2);
        if (c.bcfunc)
            bcd->SetForcingFunctionBC([](const TPZVec<REAL> &x, TPZVec<STATE> &u, TPZFMatrix<STATE> &du) { u[0] = 0.3 + x[0] * x[1] - 0.5 * x[2] * x[2]; });
        cmesh->InsertMaterialObject(bcd);
        if (c.bctype >= 1) {
            TPZManVector<STATE, 1> v2n(1, 0.75);
            // type 2: the Robin branch of TPZMatPoisson::ContributeBC (penalty load vector + BigNumber * Val1(0,0) * dphix0 dphix0)
            TPZFNMatrix<1, STATE> v1n(1, 1, 0.);
            if (c.bctype == 2) v1n(0, 0) = 2.5e-15;
            auto *bcn = m->CreateBC(m, -2, c.bctype == 2 ? 2 : 1, c.bctype == 2 ? v1n : v1, v2n);
            if (c.bcfunc)
                bcn->SetForcingFunctionBC([](const TPZVec<REAL> &x, TPZVec<STATE> &u, TPZFMatrix<STATE> &du) { u[0] = 0.75 + 2.0 * x[0] - x[1] * x[1]; });
            cmesh->InsertMaterialObject(bcn);
        }
    } else {
        TPZManVector<STATE, 3> force(3, 0.);
        force[2] = -1.;
        auto *m = new TPZElasticity3D(1, 1000., 0.3, force);
        cmesh->InsertMaterialObject(m);
        TPZFNMatrix<9, STATE> v1(3, 3, 0.);
        TPZManVector<STATE, 3> v2(3, 0.);
        auto elastfunc = [](const TPZVec<REAL> &x, TPZVec<STATE> &u, TPZFMatrix<STATE> &du) {
            u[0] = 0.01 * x[1];
            u[1] = -0.02 * x[0] * x[2];
            u[2] = 0.005 + 0.01 * x[2];
        };
        auto *bcd = m->CreateBC(m, -1, 0, v1, v2);
        if (c.bcfunc) bcd->SetForcingFunctionBC(elastfunc);
        cmesh->InsertMaterialObject(bcd);
        if (c.bctype == 1) {
            TPZManVector<STATE, 3> v2n(3, 0.);
            v2n[0] = 0.25; v2n[1] = -0.5; v2n[2] = 2.0;
            cmesh->InsertMaterialObject(m->CreateBC(m, -2, 1, v1, v2n));
        } else if (c.bctype >= 2) {
            // the other types of TPZElasticity3D::ContributeBC (mixed, directional (null) Dirichlet) on the zmax face
            TPZFNMatrix<9, STATE> v1m(3, 3, 0.);
            v1m(0, 0) = 4.0; v1m(0, 1) = 0.5; v1m(1, 0) = 0.5; v1m(1, 1) = 3.0; v1m(1, 2) = 0.25; v1m(2, 1) = 0.25; v1m(2, 2) = 5.0;
            TPZManVector<STATE, 3> v2m(3, 0.);
            v2m[0] = 0.3; v2m[1] = -0.2; v2m[2] = 0.7;
            auto *bcm = m->CreateBC(m, -2, c.bctype, v1m, v2m);
            if (c.bcfunc && c.bctype == 2) bcm->SetForcingFunctionBC(elastfunc);  // val2loc = val1 * function (TPZElasticity3D.cpp:646-654)
            cmesh->InsertMaterialObject(bcm);
        }
    }
    cmesh->SetAllCreateFunctionsContinuous();
    cmesh->AutoBuild();
    if (getenv("REFDRIVER_DEBUG")) std::cerr << "gmesh elements " << gmesh->NElements() << " materials " << cmesh->NMaterials() << " cmesh elements " << cmesh->NElements() << std::endl;
    cmesh->AdjustBoundaryElements();
    cmesh->CleanUpUnconnectedNodes();
    return cmesh;
}

static int eltype(TPZCompEl *cel) {
    switch (cel->Reference()->Type()) {
        case ECube: return 0;
        case ETetraedro: return 1;
        case EQuadrilateral: return 2;
        case ETriangle: return 3;
        case EOned: return 4;
        case EPrisma: return 5;
        case EPiramide: return 6;
        default: return -1;
    }
}

template <class TSHAPE>
static void dump_shape(const std::string &dir, const std::string &tag, TPZCompEl *cel) {
    auto *intel = dynamic_cast<TPZInterpolationSpace *>(cel);
    TPZGeoEl *gel = cel->Reference();
    const int nc = TSHAPE::NCornerNodes, ns = TSHAPE::NSides, dim = TSHAPE::Dimension;
    TPZManVector<int64_t, 8> ids(nc);
    TPZManVector<int, 27> orders(ns - nc);
    for (int i = 0; i < nc; i++) ids[i] = gel->NodeIndex(i);
    for (int i = nc; i < ns; i++) orders[i - nc] = cel->Connect(i).Order();
    TPZShapeData sd;
    TPZShapeH1<TSHAPE>::Initialize(ids, orders, sd);
    const TPZIntPoints &rule = intel->GetIntegrationRule();
    const int nq = rule.NPoints();
    const int nshape = sd.fPhi.Rows();
    std::vector<double> pts(nq * dim), w(nq), phi((size_t)nq * nshape), dphi((size_t)nq * dim * nshape);
    TPZManVector<REAL, 3> pt(dim);
    for (int q = 0; q < nq; q++) {
        REAL wq;
        rule.Point(q, pt, wq);
        w[q] = wq;
        for (int d = 0; d < dim; d++) pts[q * dim + d] = pt[d];
        TPZShapeH1<TSHAPE>::Shape(pt, sd);
        for (int i = 0; i < nshape; i++) {
            phi[(size_t)q * nshape + i] = sd.fPhi(i, 0);
            for (int d = 0; d < dim; d++) dphi[((size_t)q * dim + d) * nshape + i] = sd.fDPhi(d, i);
        }
    }
    save_vec(dir, "rule_" + tag + "_pts", pts, {nq, dim});
    save_vec(dir, "rule_" + tag + "_w", w);
    save_vec(dir, "shape_" + tag + "_phi", phi, {nq, nshape});
    save_vec(dir, "shape_" + tag + "_dphi", dphi, {nq, dim, nshape});
    std::vector<int64_t> idv(ids.begin(), ids.end());
    save_vec(dir, "shape_" + tag + "_ids", idv);
}

// shape functions of EVERY element of one topology at a few integration points (orientation-dependent for p >= 3)
template <class TSHAPE>
static void dump_shape_all(const std::string &dir, const std::string &tag, TPZCompMesh *cmesh, MElementType type) {
    const int nc = TSHAPE::NCornerNodes, ns = TSHAPE::NSides, dim = TSHAPE::Dimension;
    std::vector<double> phi, dphi;
    std::vector<int64_t> ids_all, elidx, qsel;
    int nshape = 0, nqs = 0;
    for (int64_t iel = 0; iel < cmesh->NElements(); iel++) {
        TPZCompEl *cel = cmesh->Element(iel);
        if (!cel || cel->Reference()->Type() != type) continue;
        auto *intel = dynamic_cast<TPZInterpolationSpace *>(cel);
        TPZGeoEl *gel = cel->Reference();
        TPZManVector<int64_t, 8> ids(nc);
        TPZManVector<int, 27> orders(ns - nc);
        for (int i = 0; i < nc; i++) ids[i] = gel->NodeIndex(i);
        for (int i = nc; i < ns; i++) orders[i - nc] = cel->Connect(i).Order();
        TPZShapeData sd;
        TPZShapeH1<TSHAPE>::Initialize(ids, orders, sd);
        const TPZIntPoints &rule = intel->GetIntegrationRule();
        const int nq = rule.NPoints();
        nshape = sd.fPhi.Rows();
        if (qsel.empty()) {
            qsel = {0, nq / 3, (2 * nq) / 3, nq - 1};
            nqs = 4;
        }
        TPZManVector<REAL, 3> pt(dim);
        for (int k = 0; k < nqs; k++) {
            REAL wq;
            rule.Point((int)qsel[k], pt, wq);
            TPZShapeH1<TSHAPE>::Shape(pt, sd);
            for (int i = 0; i < nshape; i++) phi.push_back(sd.fPhi(i, 0));
            for (int d = 0; d < dim; d++)
                for (int i = 0; i < nshape; i++) dphi.push_back(sd.fDPhi(d, i));
        }
        for (int i = 0; i < nc; i++) ids_all.push_back(ids[i]);
        elidx.push_back(iel);
    }
    if (elidx.empty()) return;
    const int64_t ne = (int64_t)elidx.size();
    save_vec(dir, "shapeall_" + tag + "_phi", phi, {ne, nqs, nshape});
    save_vec(dir, "shapeall_" + tag + "_dphi", dphi, {ne, nqs, dim, nshape});
    save_vec(dir, "shapeall_" + tag + "_ids", ids_all, {ne, nc});
    save_vec(dir, "shapeall_" + tag + "_el", elidx);
    save_vec(dir, "shapeall_" + tag + "_q", qsel);
}

static int cmd_dump(const std::string &dir, const Case &c, int with_elmats) {
    TPZCompMesh *cmesh = build_mesh(c);
    TPZGeoMesh *gmesh = cmesh->Reference();
    const int64_t nn = gmesh->NNodes();
    std::vector<double> nodes(nn * 3);
    for (int64_t i = 0; i < nn; i++)
        for (int d = 0; d < 3; d++) nodes[i * 3 + d] = gmesh->NodeVec()[i].Coord(d);
    save_vec(dir, "nodes", nodes, {nn, 3});
    save_vec(dir, "node_perm", g_node_perm);
    if (c.p >= 3) {
        dump_shape_all<pzshape::TPZShapeCube>(dir, "hex", cmesh, ECube);
        dump_shape_all<pzshape::TPZShapeTetra>(dir, "tet", cmesh, ETetraedro);
        dump_shape_all<pzshape::TPZShapeQuad>(dir, "quad", cmesh, EQuadrilateral);
        dump_shape_all<pzshape::TPZShapeTriang>(dir, "tri", cmesh, ETriangle);
        dump_shape_all<pzshape::TPZShapeLinear>(dir, "line", cmesh, EOned);
        dump_shape_all<pzshape::TPZShapePrism>(dir, "prism", cmesh, EPrisma);
        dump_shape_all<pzshape::TPZShapePiram>(dir, "pyr", cmesh, EPiramide);
    }

    TPZLinearAnalysis an(cmesh, false);
    const int64_t ncel = cmesh->NElements();
    std::vector<int32_t> etype(ncel), ematid(ncel), encon(ncel);
    std::vector<int64_t> enodes(ncel * 8, -1), econseq(ncel * 27, -1), econidx(ncel * 27, -1);
    std::vector<int32_t> econorder(ncel * 27, -1);
    std::vector<int64_t> dest_ptr(ncel + 1, 0), dest, ek_ptr(ncel + 1, 0);
    std::vector<double> ekv, efv;
    bool done[7] = {false, false, false, false, false, false, false};
    for (int64_t iel = 0; iel < ncel; iel++) {
        TPZCompEl *cel = cmesh->Element(iel);
        if (!cel) { etype[iel] = -1; dest_ptr[iel + 1] = dest.size(); ek_ptr[iel + 1] = ekv.size(); continue; }
        TPZGeoEl *gel = cel->Reference();
        etype[iel] = eltype(cel);
        ematid[iel] = gel->MaterialId();
        for (int i = 0; i < gel->NCornerNodes(); i++) enodes[iel * 8 + i] = gel->NodeIndex(i);
        const int nc = cel->NConnects();
        encon[iel] = nc;
        for (int i = 0; i < nc; i++) {
            TPZConnect &con = cel->Connect(i);
            econidx[iel * 27 + i] = cel->ConnectIndex(i);
            econseq[iel * 27 + i] = con.SequenceNumber();
            econorder[iel * 27 + i] = con.Order();
        }
        TPZElementMatrixT<STATE> ek(cmesh, TPZElementMatrix::EK), ef(cmesh, TPZElementMatrix::EF);
        cel->CalcStiff(ek, ef);
        ek.ComputeDestinationIndices();
        for (int64_t k = 0; k < ek.fDestinationIndex.size(); k++) dest.push_back(ek.fDestinationIndex[k]);
        dest_ptr[iel + 1] = dest.size();
        if (with_elmats) {
            const int64_t nd = ek.fMat.Rows();
            for (int64_t j = 0; j < nd; j++)
                for (int64_t i = 0; i < nd; i++) ekv.push_back(ek.fMat(i, j));  // column-major like TPZFMatrix
            for (int64_t i = 0; i < nd; i++) efv.push_back(ef.fMat(i, 0));
        }
        ek_ptr[iel + 1] = ekv.size();
        const int t = etype[iel];
        if (t >= 0 && !done[t]) {
            done[t] = true;
            if (t == 0) dump_shape<pzshape::TPZShapeCube>(dir, "hex", cel);
            if (t == 1) dump_shape<pzshape::TPZShapeTetra>(dir, "tet", cel);
            if (t == 2) dump_shape<pzshape::TPZShapeQuad>(dir, "quad", cel);
            if (t == 3) dump_shape<pzshape::TPZShapeTriang>(dir, "tri", cel);
            if (t == 4) dump_shape<pzshape::TPZShapeLinear>(dir, "line", cel);
            if (t == 5) dump_shape<pzshape::TPZShapePrism>(dir, "prism", cel);
            if (t == 6) dump_shape<pzshape::TPZShapePiram>(dir, "pyr", cel);
        }
    }
    save_vec(dir, "el_type", etype);
    save_vec(dir, "el_matid", ematid);
    save_vec(dir, "el_ncon", encon);
    save_vec(dir, "el_nodes", enodes, {ncel, 8});
    save_vec(dir, "el_conseq", econseq, {ncel, 27});
    save_vec(dir, "el_conidx", econidx, {ncel, 27});
    save_vec(dir, "el_conorder", econorder, {ncel, 27});
    save_vec(dir, "el_dest_ptr", dest_ptr);
    save_vec(dir, "el_dest", dest);
    if (with_elmats) {
        save_vec(dir, "ek_ptr", ek_ptr);
        save_vec(dir, "ek", ekv);
        save_vec(dir, "ef", efv);
    }
    const int64_t nb = cmesh->Block().NBlocks();
    std::vector<int64_t> bpos(nb), bsize(nb);
    for (int64_t b = 0; b < nb; b++) { bpos[b] = cmesh->Block().Position(b); bsize[b] = cmesh->Block().Size(b); }
    save_vec(dir, "block_pos", bpos);
    save_vec(dir, "block_size", bsize);

    // symmetric CSR through the reference's own Assemble()
    const int64_t neq = cmesh->NEquations();
    {
        TPZSSpStructMatrix<STATE> strmat(cmesh);
        strmat.SetNumThreads(0);
        an.SetStructuralMatrix(strmat);
        TPZStepSolver<STATE> step;
        step.SetDirect(ELDLt);
        an.SetSolver(step);
        an.Assemble();
        auto mtx = an.MatrixSolver<STATE>().Matrix();
        auto *sp = dynamic_cast<TPZSYsmpMatrix<STATE> *>(mtx.operator->());
        if (!sp) { std::cerr << "not a TPZSYsmpMatrix\n"; return 2; }
        std::vector<int64_t> ia(sp->IA().begin(), sp->IA().end()), ja(sp->JA().begin(), sp->JA().end());
        std::vector<double> a(sp->A().begin(), sp->A().end());
        save_vec(dir, "sym_ia", ia);
        save_vec(dir, "sym_ja", ja);
        save_vec(dir, "sym_a", a);
        TPZFMatrix<STATE> &rhs = an.Rhs();
        std::vector<double> r(neq);
        for (int64_t i = 0; i < neq; i++) r[i] = rhs(i, 0);
        save_vec(dir, "rhs", r);
    }
    // full CSR
    {
        TPZLinearAnalysis an2(cmesh, false);
        TPZSpStructMatrix<STATE> strmat(cmesh);
        strmat.SetNumThreads(0);
        an2.SetStructuralMatrix(strmat);
        TPZStepSolver<STATE> step;
        step.SetDirect(ELU);
        an2.SetSolver(step);
        an2.Assemble();
        auto mtx = an2.MatrixSolver<STATE>().Matrix();
        auto *sp = dynamic_cast<TPZFYsmpMatrix<STATE> *>(mtx.operator->());
        if (!sp) { std::cerr << "not a TPZFYsmpMatrix\n"; return 2; }
        TPZVec<int64_t> ia, ja;
        TPZVec<STATE> a;
        sp->GetData(ia, ja, a);  // copies (Matrix/pzysmp.h:284-288)
        const int64_t nnz = ia[neq];
        save_npy<int64_t>(dir + "/full_ia.npy", ia.begin(), {neq + 1});
        save_npy<int64_t>(dir + "/full_ja.npy", ja.begin(), {nnz});
        save_npy<double>(dir + "/full_a.npy", a.begin(), {nnz});
    }
    // reference solution by the reference's skyline LDLt (robust with penalty BCs)
    {
        TPZLinearAnalysis an3(cmesh, false);
        TPZSkylineStructMatrix<STATE> strmat(cmesh);
        strmat.SetNumThreads(0);
        an3.SetStructuralMatrix(strmat);
        TPZStepSolver<STATE> step;
        step.SetDirect(ELDLt);
        an3.SetSolver(step);
        an3.Run();
        TPZFMatrix<STATE> &sol = an3.Solution();
        std::vector<double> u(neq);
        for (int64_t i = 0; i < neq; i++) u[i] = sol(i, 0);
        save_vec(dir, "sol", u);
    }
    // material constants as the reference holds them
    std::ofstream meta(dir + "/meta.json");
    meta.precision(17);
    meta << "{\"n\": " << c.n << ", \"p\": " << c.p << ", \"phys\": " << c.phys << ", \"tet\": " << c.tet
         << ", \"perturb\": " << c.perturb << ", \"bctype\": " << c.bctype << ", \"dim\": " << c.dim << ", \"scramble\": " << c.scramble << ", \"bcfunc\": " << c.bcfunc << ", \"neq\": " << neq
         << ", \"ncel\": " << ncel << ", \"nnodes\": " << nn;
    TPZMaterial *mat = cmesh->FindMaterial(1);
    meta << ", \"bignumber\": " << mat->BigNumber();
    meta << "}\n";
    std::cout << "dumped " << dir << " neq " << neq << " ncel " << ncel << "\n";
    return 0;
}

// raw binary dump (bench.py's parity leg reads it with numpy.fromfile)
template <class T>
static void save_raw(const std::string &path, const T *data, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write((const char *)data, sizeof(T) * n);
}

static int cmd_time(const Case &c, int nthreads, int reps, const std::string &dumpdir = "") {
    auto t0 = clk::now();
    TPZCompMesh *cmesh = build_mesh(c);
    auto t1 = clk::now();
    TPZLinearAnalysis an(cmesh, false);
    TPZSSpStructMatrix<STATE> strmat(cmesh);
    strmat.SetNumThreads(nthreads);
    an.SetStructuralMatrix(strmat);
    TPZStepSolver<STATE> step;
    step.SetDirect(ELDLt);
    an.SetSolver(step);
    auto t2 = clk::now();
    an.Assemble();  // Create() + Assemble
    auto t3 = clk::now();
    double best = 1e300, sum = 0;
    for (int r = 0; r < reps; r++) {
        auto a = clk::now();
        an.Assemble();  // Zero() + Assemble only
        auto b = clk::now();
        double s = std::chrono::duration<double>(b - a).count();
        best = std::min(best, s);
        sum += s;
    }
    auto mtx = an.MatrixSolver<STATE>().Matrix();
    auto *sp = dynamic_cast<TPZSYsmpMatrix<STATE> *>(mtx.operator->());
    long double fro = 0;
    for (int64_t k = 0; k < sp->A().size(); k++) fro += (long double)sp->A()[k] * sp->A()[k];
    int64_t nvol = 0;  // volume elements = computational elements of the mesh dimension
    for (int64_t iel = 0; iel < cmesh->NElements(); iel++) {
        TPZCompEl *cel = cmesh->Element(iel);
        if (cel && cel->Reference() && cel->Reference()->Dimension() == c.dim) nvol++;
    }
    long double rn2 = 0;
    TPZFMatrix<STATE> &rhs = an.Rhs();
    for (int64_t k = 0; k < rhs.Rows(); k++) rn2 += (long double)rhs(k, 0) * rhs(k, 0);
    if (!dumpdir.empty()) {  // the assembled system of the timed mesh, for the parity check of the GPU arm on the SAME mesh
        save_raw(dumpdir + "/ia.bin", &sp->IA()[0], (size_t)sp->IA().size());
        save_raw(dumpdir + "/ja.bin", &sp->JA()[0], (size_t)sp->JA().size());
        save_raw(dumpdir + "/a.bin", &sp->A()[0], (size_t)sp->A().size());
        save_raw(dumpdir + "/rhs.bin", &rhs(0, 0), (size_t)rhs.Rows());
    }
    std::cout.precision(17);
    std::cout << "{\"n\": " << c.n << ", \"p\": " << c.p << ", \"phys\": " << c.phys << ", \"tet\": " << c.tet
              << ", \"perturb\": " << c.perturb << ", \"rhs2\": " << (double)rn2
              << ", \"threads\": " << nthreads << ", \"reps\": " << reps << ", \"vol_elements\": " << nvol
              << ", \"neq\": " << cmesh->NEquations() << ", \"nnz\": " << sp->JA().size()
              << ", \"mesh_s\": " << std::chrono::duration<double>(t1 - t0).count()
              << ", \"create_assemble_s\": " << std::chrono::duration<double>(t3 - t2).count()
              << ", \"assemble_s_best\": " << best << ", \"assemble_s_mean\": " << sum / reps
              << ", \"fro2\": " << (double)fro << "}" << std::endl;
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        std::cerr << "usage: refdriver dump <dir> n p phys tet perturb bctype with_elmats | time n p phys tet threads reps [perturb [dumpdir]]\n";
        return 1;
    }
    std::string cmd = argv[1];
    Case c;
    if (cmd == "dump" && argc >= 10) {
        std::string dir = argv[2];
        c.n = atoi(argv[3]); c.p = atoi(argv[4]); c.phys = atoi(argv[5]); c.tet = atoi(argv[6]);
        c.perturb = atof(argv[7]); c.bctype = atoi(argv[8]);
        if (argc >= 11) c.scramble = atoi(argv[10]);
        if (argc >= 12) c.dim = atoi(argv[11]);
        if (argc >= 13) c.bcfunc = atoi(argv[12]);
        return cmd_dump(dir, c, atoi(argv[9]));
    }
    if (cmd == "time" && argc >= 8) {
        c.n = atoi(argv[2]); c.p = atoi(argv[3]); c.phys = atoi(argv[4]); c.tet = atoi(argv[5]);
        if (argc >= 9) c.perturb = atof(argv[8]);
        return cmd_time(c, atoi(argv[6]), atoi(argv[7]), argc >= 10 ? argv[9] : "");
    }
    std::cerr << "bad arguments\n";
    return 1;
}
