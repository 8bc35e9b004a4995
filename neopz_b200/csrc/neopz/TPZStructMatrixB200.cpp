// TPZStructMatrixB200 — NeoPZ-side host code of the B200 assembly strategy (see the header).
//
// Everything here runs once per mesh (flatten) or once per Assemble() (node coordinates, material
// constants, one C-ABI call); the arithmetic of CalcStiff / Contribute / AddKel / AddFel happens in the
// CUDA library (neopz_b200/csrc/b200asm.cu).
#include "TPZStructMatrixB200.h"

#include <chrono>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "Elasticity/TPZElasticity2D.h"
#include "Elasticity/TPZElasticity3D.h"
#include "pzshapelinear.h"
#include "Poisson/TPZMatPoisson.h"
#include "TPZBndCondT.h"
#include "TPZCompElH1.h"
#include "TPZFMatrixRef.h"
#include "TPZMatLoadCases.h"
#include "TPZMaterial.h"
#include "TPZShapeData.h"
#include "TPZShapeH1.h"
#include "TPZStructMatrix.h"
#include "pzcmesh.h"
#include "pzfmatrix.h"
#include "pzgmesh.h"
#include "pzgnode.h"
#include "pzinterpolationspace.h"
#include "pzshapecube.h"
#include "pzshapepiram.h"
#include "pzshapeprism.h"
#include "pzshapequad.h"
#include "pzshapetetra.h"
#include "pzshapetriang.h"
#include "pzsysmp.h"
#include "pzysmp.h"

#include "../../../include/b200asm.h"

namespace {

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

[[noreturn]] void Fatal(const std::string &msg) {
    PZError << "TPZStructMatrixB200: " << msg << std::endl;
    DebugStop();
    std::abort();
}

// protected constants of TPZElasticity3D (Material/Elasticity/TPZElasticity3D.h:225-248)
struct ElastAccess : public TPZElasticity3D {
    using TPZElasticity3D::C1;
    using TPZElasticity3D::C2;
    using TPZElasticity3D::C3;
    using TPZElasticity3D::fForce;
    using TPZElasticity3D::fPreStress;
};

// one device group per (topology, material, side orders, side orientations): the shape functions of a side with more than one
// function depend on the global indices of its corner nodes (Shape/pzgenericshape.cpp:57-68); elements of one class share tables
// protected members of TPZElasticity2D (Material/Elasticity/TPZElasticity2D.h:196-232)
struct Elast2DAccess : public TPZElasticity2D {
    using TPZElasticity2D::ff;
    using TPZElasticity2D::fPlaneStress;
    using TPZElasticity2D::fPreStressXX;
    using TPZElasticity2D::fPreStressXY;
    using TPZElasticity2D::fPreStressYY;
};

// signature: the order of every side connect and, where a side carries more than one function, its transform id
// (TSHAPE::GetTransformId, the input of ComputeTransforms): equal signatures <=> equal TPZShapeH1 tables
struct GroupKey {
    int topology, matid;
    std::vector<int> signature;
    bool operator<(const GroupKey &o) const {
        return std::tie(topology, matid, signature) < std::tie(o.topology, o.matid, o.signature);
    }
};

struct HostGroup {
    b200asm_group meta{};
    TPZMaterial *material = nullptr;
    std::vector<int32_t> elnodes;
    std::vector<int64_t> dest;
    std::vector<double> qpts, qw, phi, dphi, force;
    std::vector<TPZCompEl *> elements;
    bool has_forcing = false;
};

int TopologyOf(MElementType t) {
    switch (t) {
        case ECube: return B200ASM_HEX;
        case ETetraedro: return B200ASM_TET;
        case EQuadrilateral: return B200ASM_QUAD;
        case ETriangle: return B200ASM_TRI;
        case EOned: return B200ASM_LINE;
        case EPrisma: return B200ASM_PRISM;
        case EPiramide: return B200ASM_PYRAMID;
        default: return -1;
    }
}

// orders of the side connects (the element may be p-refined: sides of different order) followed by the transform ids of the
// sides with more than one function (Shape/pzgenericshape.cpp:57-68)
template <class TSHAPE>
void SideSignature(TPZCompEl *cel, std::vector<int> &sig) {
    TPZGeoEl *gel = cel->Reference();
    const int nc = TSHAPE::NCornerNodes, ns = TSHAPE::NSides;
    TPZManVector<int64_t, 8> ids(nc);
    for (int i = 0; i < nc; i++) ids[i] = gel->NodeIndex(i);
    sig.clear();
    for (int side = nc; side < ns; side++) sig.push_back(cel->Connect(side).Order());
    for (int side = nc; side < ns; side++)
        sig.push_back(TSHAPE::NConnectShapeF(side, cel->Connect(side).Order()) > 1 ? TSHAPE::GetTransformId(side, ids) : 0);
}

template <class TSHAPE>
void ShapeTables(TPZCompEl *cel, HostGroup &g) {
    // integration rule and shape tables through the reference's own objects
    auto *intel = dynamic_cast<TPZInterpolationSpace *>(cel);
    TPZGeoEl *gel = cel->Reference();
    const int nc = TSHAPE::NCornerNodes, ns = TSHAPE::NSides, dim = TSHAPE::Dimension;
    TPZManVector<int64_t, 8> ids(nc);
    TPZManVector<int, 27> orders(ns - nc, 1);
    for (int i = 0; i < nc; i++) ids[i] = gel->NodeIndex(i);
    for (int side = nc; side < ns; side++) orders[side - nc] = cel->Connect(side).Order();
    TPZShapeData sd;
    TPZShapeH1<TSHAPE>::Initialize(ids, orders, sd);
    const TPZIntPoints &rule = intel->GetIntegrationRule();
    const int nq = rule.NPoints();
    const int nshape = sd.fPhi.Rows();
    g.qpts.resize((size_t)nq * dim);
    g.qw.resize(nq);
    g.phi.resize((size_t)nq * nshape);
    g.dphi.resize((size_t)nq * dim * nshape);
    TPZManVector<REAL, 3> pt(dim);
    for (int q = 0; q < nq; q++) {
        REAL w;
        rule.Point(q, pt, w);
        g.qw[q] = w;
        for (int d = 0; d < dim; d++) g.qpts[(size_t)q * dim + d] = pt[d];
        TPZShapeH1<TSHAPE>::Shape(pt, sd);
        for (int i = 0; i < nshape; i++) {
            g.phi[(size_t)q * nshape + i] = sd.fPhi(i, 0);
            for (int d = 0; d < dim; d++) g.dphi[((size_t)q * dim + d) * nshape + i] = sd.fDPhi(d, i);
        }
    }
    g.meta.nqp = nq;
    g.meta.nshape = nshape;
}

}  // namespace

struct TPZB200AssemblyCache {
    b200asm_ctx *ctx = nullptr;
    TPZCompMesh *mesh = nullptr;
    int64_t nelem = -1, neq = -1, nconnects = -1, nnz = -1;
    int64_t nactive = -1;  // equations of the assembled system: NEquations(), or NActiveEquations() of an active filter
    int symmetric = -1;
    bool pattern_set = false;
    std::vector<HostGroup> groups;
    double flatten_ms = 0, pattern_ms = 0, assemble_ms = 0;
    void *pinned = nullptr;  // value array page-locked by SetPinHostMatrix
    size_t pinned_bytes = 0;
    void Unpin() {
        if (pinned && ctx) b200asm_unpin_host(ctx, pinned);
        pinned = nullptr;
        pinned_bytes = 0;
    }
    ~TPZB200AssemblyCache() {
        Unpin();
        if (ctx) b200asm_destroy(ctx);
    }
};

namespace {

void Check(TPZB200AssemblyCache &c, int rc, const char *what) {
    if (rc < 0) Fatal(std::string(what) + " failed: " + b200asm_last_error(c.ctx));
}

// material constants of a group, re-read at every Assemble (they may change between assemblies)
void FillCoef(HostGroup &g) {
    double *coef = g.meta.coef;
    std::memset(coef, 0, sizeof(double) * 16);
    TPZMaterial *mat = g.material;
    if (auto *bc = dynamic_cast<TPZBndCondT<STATE> *>(mat)) {
        // (a forcing function replaces val2: FillForce tabulates the load-vector coefficient per integration point and the
        //  kernel ignores coef[9..11]; the matrix part below stays constant)
        TPZMaterial *vol = bc->Material();
        const int type = bc->Type();
        const TPZFMatrix<STATE> &v1 = bc->Val1();
        const TPZVec<STATE> &v2 = bc->Val2();
        if (auto *pois = dynamic_cast<TPZMatPoisson<STATE> *>(vol)) {
            const double big = pois->BigNumber();
            if (type == 0) {  // Material/Poisson/TPZMatPoisson.cpp:79-90
                coef[0] = big;
                coef[9] = big * v2[0];
            } else if (type == 1) {  // :93-100
                coef[9] = v2[0] * pois->ScaleFactor();
            } else {
                Fatal("TPZMatPoisson boundary condition type " + std::to_string(type) + " is not supported");
            }
        } else if (auto *e2 = dynamic_cast<TPZElasticity2D *>(vol)) {
            const double big = e2->BigNumber();  // Material/Elasticity/TPZElasticity2D.cpp:219
            if (type == 0) {  // :256-270
                for (int a = 0; a < 2; a++) {
                    coef[a * 3 + a] = big;
                    coef[9 + a] = big * v2[a];
                }
            } else if (type == 1) {  // :273-283
                for (int a = 0; a < 2; a++) coef[9 + a] = v2[a];
            } else if (type == 3) {  // :305-316
                for (int a = 0; a < 2; a++) coef[a * 3 + a] = big * v2[a];
            } else {
                Fatal("TPZElasticity2D boundary condition type " + std::to_string(type) + " is not supported");
            }
        } else if (dynamic_cast<TPZElasticity3D *>(vol)) {
            const double big = 1.e12;  // Material/Elasticity/TPZElasticity3D.cpp:630
            if (type == 0) {
                for (int a = 0; a < 3; a++) {
                    coef[a * 3 + a] = big;
                    coef[9 + a] = big * v2[a];
                }
            } else if (type == 1) {
                for (int a = 0; a < 3; a++) coef[9 + a] = v2[a];
            } else if (type == 2) {
                for (int a = 0; a < 3; a++) {
                    for (int b = 0; b < 3; b++) coef[a * 3 + b] = v1.GetVal(a, b);
                    coef[9 + a] = v2[a];
                }
            } else if (type == 3) {  // directional null Dirichlet, TPZElasticity3D.cpp:715-723
                for (int a = 0; a < 3; a++) coef[a * 3 + a] = big * v2[a];
            } else if (type >= 5 && type <= 8) {  // directional Dirichlet on x / y / z / x and z, :739-772
                const bool on[3] = {type == 5 || type == 8, type == 6, type == 7 || type == 8};
                for (int a = 0; a < 3; a++)
                    if (on[a]) {
                        coef[a * 3 + a] = big;
                        coef[9 + a] = big * v2[a];
                    }
            } else {
                Fatal("TPZElasticity3D boundary condition type " + std::to_string(type) + " is not supported (type 4 needs the face normal)");
            }
        } else {
            Fatal("boundary condition of an unsupported material");
        }
        return;
    }
    if (auto *pois = dynamic_cast<TPZMatPoisson<STATE> *>(mat)) {
        coef[0] = pois->ScaleFactor();
        coef[1] = 0.0;  // the source comes from the forcing-function table (or is absent)
        return;
    }
    if (auto *e2 = dynamic_cast<TPZElasticity2D *>(mat)) {
        // the three constants of Material/Elasticity/TPZElasticity2D.cpp:141-199 (plane strain / plane stress)
        auto *acc = static_cast<Elast2DAccess *>(e2);
        const double E = e2->E(), nu = e2->Nu();
        if (acc->fPlaneStress) {
            coef[0] = E / (1 - nu * nu);
            coef[1] = E / (2. * (1 + nu));
            coef[2] = E / (1 - nu * nu) * nu;
        } else {
            const double F = E / ((1. + nu) * (1. - 2. * nu));
            coef[0] = (1. - nu) * F;
            coef[1] = (1. - 2. * nu) / 2. * F;
            coef[2] = nu * F;
        }
        coef[3] = acc->ff[0];
        coef[4] = acc->ff[1];
        coef[5] = acc->fPreStressXX;
        coef[6] = acc->fPreStressXY;
        coef[7] = acc->fPreStressYY;
        return;
    }
    if (auto *el = dynamic_cast<TPZElasticity3D *>(mat)) {
        auto *acc = static_cast<ElastAccess *>(el);
        coef[0] = acc->C1;
        coef[1] = acc->C2;
        coef[2] = acc->C3;
        for (int k = 0; k < 3; k++) {
            coef[3 + k] = acc->fForce[k];
            coef[6 + k] = acc->fPreStress[k];
        }
        return;
    }
    Fatal("unsupported material (TPZMatPoisson<STATE>, TPZElasticity3D and their TPZBndCondT only)");
}

// forcing std::function evaluated on the host at every integration point (Material/TPZMatTypes.h:15-19)
void FillForce(HostGroup &g) {
    g.has_forcing = false;
    g.meta.force = nullptr;
    const int ns = g.meta.nstate;
    if (g.meta.kind == B200ASM_BC) {
        // boundary data given by a function (TPZBndCondT::ForcingFunctionBC): evaluated on the host at data.x of every
        // integration point, stored as the coefficient of phi_i * weight in ef (what the constant case keeps in coef[9..11])
        auto *bc = dynamic_cast<TPZBndCondT<STATE> *>(g.material);
        if (!bc || !bc->HasForcingFunctionBC()) return;
        TPZMaterial *vol = bc->Material();
        auto *pois = dynamic_cast<TPZMatPoisson<STATE> *>(vol);
        auto *e2 = dynamic_cast<TPZElasticity2D *>(vol);
        auto *e3 = dynamic_cast<TPZElasticity3D *>(vol);
        const int type = bc->Type();
        const bool ok = pois ? (type == 0 || type == 1) : (e2 ? (type == 0 || type == 1) : (e3 ? (type == 0 || type == 2 || (type >= 5 && type <= 8)) : false));
        if (!ok) Fatal("boundary condition type " + std::to_string(type) + " with a forcing function is not supported");
        const int fdim = g.meta.topology == B200ASM_LINE ? 1 : 2;
        const int nq = g.meta.nqp;
        g.force.assign((size_t)g.meta.nel * nq * ns, 0.0);
        TPZManVector<REAL, 3> qsi(fdim), x(3);
        for (int64_t e = 0; e < g.meta.nel; e++) {
            TPZGeoEl *gel = g.elements[e]->Reference();
            for (int q = 0; q < nq; q++) {
                for (int d = 0; d < fdim; d++) qsi[d] = g.qpts[(size_t)q * fdim + d];
                gel->X(qsi, x);
                TPZManVector<STATE, 3> v2(e3 ? 3 : ns, 0.);
                TPZFNMatrix<9, STATE> v1(bc->Val1());
                bc->ForcingFunctionBC()(x, v2, v1);
                double *out = &g.force[((size_t)e * nq + q) * ns];
                if (pois) {  // Material/Poisson/TPZMatPoisson.cpp:79-100
                    out[0] = type == 0 ? pois->BigNumber() * v2[0] : v2[0] * pois->ScaleFactor();
                } else if (e2) {  // Material/Elasticity/TPZElasticity2D.cpp:256-283
                    for (int a = 0; a < 2; a++) out[a] = type == 0 ? e2->BigNumber() * v2[a] : v2[a];
                } else {  // Material/Elasticity/TPZElasticity3D.cpp:637-697,739-772
                    const double big = 1.e12;
                    if (type == 0) {
                        for (int a = 0; a < 3; a++) out[a] = big * v2[a];
                    } else if (type == 2) {  // val2loc = Val1 * function value
                        const TPZFMatrix<STATE> &m = bc->Val1();
                        for (int a = 0; a < 3; a++) {
                            double t = 0.;
                            for (int b = 0; b < 3; b++) t += m.GetVal(a, b) * v2[b];
                            out[a] = t;
                        }
                    } else {
                        const bool on[3] = {type == 5 || type == 8, type == 6, type == 7 || type == 8};
                        for (int a = 0; a < 3; a++) out[a] = on[a] ? big * v2[a] : 0.;
                    }
                }
            }
        }
        g.has_forcing = true;
        g.meta.force = g.force.data();
        return;
    }
    const int dim = (g.meta.topology == B200ASM_HEX || g.meta.topology == B200ASM_TET || g.meta.topology == B200ASM_PRISM ||
                     g.meta.topology == B200ASM_PYRAMID) ? 3 : 2;
    if (dim != 3) {
        // plane domain elements: TPZMatPoisson(dim 2) source, TPZElasticity2D body force (TPZElasticity2D.cpp:120-127)
        auto *p2 = dynamic_cast<TPZMatPoisson<STATE> *>(g.material);
        auto *e2 = dynamic_cast<TPZElasticity2D *>(g.material);
        if (!((p2 && p2->HasForcingFunction()) || (e2 && e2->HasForcingFunction()))) return;
        const int nq = g.meta.nqp;
        g.force.assign((size_t)g.meta.nel * nq * ns, 0.0);
        TPZManVector<REAL, 3> qsi(2), x(3);
        for (int64_t e = 0; e < g.meta.nel; e++) {
            TPZGeoEl *gel = g.elements[e]->Reference();
            for (int q = 0; q < nq; q++) {
                for (int d = 0; d < 2; d++) qsi[d] = g.qpts[(size_t)q * 2 + d];
                gel->X(qsi, x);
                TPZManVector<STATE, 3> f(3, 0.);
                if (p2) p2->ForcingFunction()(x, f); else e2->ForcingFunction()(x, f);
                for (int k = 0; k < ns; k++) g.force[((size_t)e * nq + q) * ns + k] = f[k];
            }
        }
        g.has_forcing = true;
        g.meta.force = g.force.data();
        return;
    }
    auto *pois = dynamic_cast<TPZMatPoisson<STATE> *>(g.material);
    auto *el = dynamic_cast<TPZElasticity3D *>(g.material);
    const bool hasf = pois ? pois->HasForcingFunction() : (el ? el->HasForcingFunction() : false);
    if (!hasf) return;
    const int nq = g.meta.nqp;
    g.force.assign((size_t)g.meta.nel * nq * ns, 0.0);
    TPZManVector<REAL, 3> qsi(3), x(3);
    for (int64_t e = 0; e < g.meta.nel; e++) {
        TPZGeoEl *gel = g.elements[e]->Reference();
        for (int q = 0; q < nq; q++) {
            for (int d = 0; d < 3; d++) qsi[d] = g.qpts[(size_t)q * 3 + d];
            gel->X(qsi, x);
            TPZManVector<STATE, 3> f(ns, 0.);
            if (pois) {
                pois->ForcingFunction()(x, f);
            } else {
                auto *acc = static_cast<ElastAccess *>(el);
                for (int k = 0; k < 3; k++) f[k] = acc->fForce[k];  // locForce(fForce) then the callback
                el->ForcingFunction()(x, f);
            }
            for (int k = 0; k < ns; k++) g.force[((size_t)e * nq + q) * ns + k] = f[k];
        }
    }
    g.has_forcing = true;
    g.meta.force = g.force.data();
}

void Flatten(TPZB200AssemblyCache &c, TPZStructMatrix *strmat) {
    TPZCompMesh *cmesh = strmat->Mesh();
    c.groups.clear();
    // active equation filter (StrMatrix/TPZEquationFilter.h): destination indices become the condensed numbers, removed
    // equations get -1 — what TPZEquationFilter::Filter does per element in the reference (pzstrmatrixor.cpp:205-214)
    std::vector<int64_t> eqmap;
    const TPZEquationFilter &filter = strmat->EquationFilter();
    if (filter.IsActive()) {
        const int64_t neq = cmesh->NEquations();
        TPZVec<int64_t> orig(neq), dest(neq);
        for (int64_t i = 0; i < neq; i++) orig[i] = dest[i] = i;
        filter.Filter(orig, dest);
        eqmap.assign(neq, -1);
        for (int64_t k = 0; k < orig.size(); k++) eqmap[orig[k]] = dest[k];
    }
    std::map<GroupKey, size_t> index;
    const int64_t nel = cmesh->NElements();
    for (int64_t iel = 0; iel < nel; iel++) {
        TPZCompEl *cel = cmesh->Element(iel);
        if (!cel) continue;
        TPZMaterial *mat = cel->Material();
        if (!mat) continue;  // the reference skips elements without material too
        if (!strmat->ShouldCompute(mat->Id())) continue;  // TPZStructMatrix::SetMaterialIds filter
        TPZGeoEl *gel = cel->Reference();
        if (!gel) Fatal("computational element without geometric reference");
        const int topo = TopologyOf(gel->Type());
        const bool h1 = dynamic_cast<TPZCompElH1<pzshape::TPZShapeCube> *>(cel) || dynamic_cast<TPZCompElH1<pzshape::TPZShapeTetra> *>(cel) ||
                        dynamic_cast<TPZCompElH1<pzshape::TPZShapeQuad> *>(cel) || dynamic_cast<TPZCompElH1<pzshape::TPZShapeTriang> *>(cel) ||
                        dynamic_cast<TPZCompElH1<pzshape::TPZShapeLinear> *>(cel) || dynamic_cast<TPZCompElH1<pzshape::TPZShapePrism> *>(cel) ||
                        dynamic_cast<TPZCompElH1<pzshape::TPZShapePiram> *>(cel);
        if (topo < 0 || !h1) Fatal("element " + std::to_string(iel) + " is not an H1 hexahedron/tetrahedron/prism/pyramid/quadrilateral/triangle/line");
        if (!gel->IsLinearMapping()) Fatal("element " + std::to_string(iel) + " has a non-(multi)linear geometric map");
        // no constraints; the sides may carry different orders (p-refined neighbours)
        const int ncon = cel->NConnects();
        const int ncorner = gel->NCornerNodes();
        int porder = 1;
        for (int i = 0; i < ncon; i++) {
            TPZConnect &con = cel->Connect(i);
            if (con.HasDependency() || con.IsCondensed()) Fatal("hanging nodes / condensed connects are not supported");
            if (i >= ncorner) porder = std::max<int>(porder, con.Order());
        }
        if (porder > 9) Fatal("polynomial order " + std::to_string(porder) + " is not supported (<= 9)");
        std::vector<int> signature;
        switch (topo) {
            case B200ASM_HEX: SideSignature<pzshape::TPZShapeCube>(cel, signature); break;
            case B200ASM_TET: SideSignature<pzshape::TPZShapeTetra>(cel, signature); break;
            case B200ASM_QUAD: SideSignature<pzshape::TPZShapeQuad>(cel, signature); break;
            case B200ASM_LINE: SideSignature<pzshape::TPZShapeLinear>(cel, signature); break;
            case B200ASM_PRISM: SideSignature<pzshape::TPZShapePrism>(cel, signature); break;
            case B200ASM_PYRAMID: SideSignature<pzshape::TPZShapePiram>(cel, signature); break;
            default: SideSignature<pzshape::TPZShapeTriang>(cel, signature); break;
        }
        const GroupKey key{topo, mat->Id(), signature};
        auto it = index.find(key);
        if (it == index.end()) {
            it = index.emplace(key, c.groups.size()).first;
            c.groups.emplace_back();
            HostGroup &g = c.groups.back();
            g.material = mat;
            g.meta.topology = topo;
            g.meta.porder = porder;
            g.meta.nstate = mat->NStateVariables();
            const bool isbc = dynamic_cast<TPZBndCond *>(mat) != nullptr;
            if (isbc) g.meta.kind = B200ASM_BC;
            else if (dynamic_cast<TPZMatPoisson<STATE> *>(mat)) g.meta.kind = B200ASM_POISSON;
            else if (dynamic_cast<TPZElasticity3D *>(mat)) g.meta.kind = B200ASM_ELASTICITY3D;
            else if (dynamic_cast<TPZElasticity2D *>(mat)) g.meta.kind = B200ASM_ELASTICITY2D;
            else Fatal("unsupported material id " + std::to_string(mat->Id()));
            if (auto *lc = dynamic_cast<TPZMatLoadCasesBase *>(mat))
                if (lc->NumLoadCases() != 1) Fatal("more than one load case is not supported");
            switch (topo) {
                case B200ASM_HEX: ShapeTables<pzshape::TPZShapeCube>(cel, g); break;
                case B200ASM_TET: ShapeTables<pzshape::TPZShapeTetra>(cel, g); break;
                case B200ASM_QUAD: ShapeTables<pzshape::TPZShapeQuad>(cel, g); break;
                case B200ASM_LINE: ShapeTables<pzshape::TPZShapeLinear>(cel, g); break;
                case B200ASM_PRISM: ShapeTables<pzshape::TPZShapePrism>(cel, g); break;
                case B200ASM_PYRAMID: ShapeTables<pzshape::TPZShapePiram>(cel, g); break;
                default: ShapeTables<pzshape::TPZShapeTriang>(cel, g); break;
            }
        }
        HostGroup &g = c.groups[it->second];
        auto *intel = dynamic_cast<TPZInterpolationSpace *>(cel);
        if (intel->GetIntegrationRule().NPoints() != g.meta.nqp) Fatal("elements of one material/order use different integration rules");
        for (int i = 0; i < ncorner; i++) g.elnodes.push_back((int32_t)gel->NodeIndex(i));
        // TPZElementMatrix::ComputeDestinationIndices, Mesh/pzelmat.cpp:37-70
        int ndof = 0;
        for (int i = 0; i < ncon; i++) {
            TPZConnect &con = cel->Connect(i);
            const int64_t seq = con.SequenceNumber();
            const int64_t first = cmesh->Block().Position(seq);
            const int ndf = cmesh->Block().Size(seq);
            for (int idf = 0; idf < ndf; idf++) g.dest.push_back(eqmap.empty() ? first + idf : eqmap[first + idf]);
            ndof += ndf;
        }
        if (ndof != g.meta.nshape * g.meta.nstate) Fatal("element " + std::to_string(iel) + ": unexpected number of equations");
        g.elements.push_back(cel);
    }
    // node table
    TPZGeoMesh *gmesh = cmesh->Reference();
    const int64_t nnodes = gmesh->NNodes();
    std::vector<double> xyz((size_t)nnodes * 3);
    for (int64_t i = 0; i < nnodes; i++)
        for (int d = 0; d < 3; d++) xyz[(size_t)i * 3 + d] = gmesh->NodeVec()[i].Coord(d);
    Check(c, b200asm_clear_groups(c.ctx), "b200asm_clear_groups");
    Check(c, b200asm_set_nodes(c.ctx, nnodes, xyz.data()), "b200asm_set_nodes");
    for (HostGroup &g : c.groups) {
        g.meta.nel = (int64_t)g.elements.size();
        g.meta.elnodes = g.elnodes.data();
        g.meta.dest = g.dest.data();
        g.meta.qpts = g.qpts.data();
        g.meta.qwts = g.qw.data();
        g.meta.phi = g.phi.data();
        g.meta.dphi = g.dphi.data();
        FillCoef(g);
        FillForce(g);
        Check(c, b200asm_add_group(c.ctx, &g.meta), "b200asm_add_group");
    }
    c.mesh = cmesh;
    c.nelem = nel;
    c.neq = cmesh->NEquations();
    c.nactive = filter.IsActive() ? filter.NActiveEquations() : c.neq;
    c.nconnects = cmesh->NConnects();
    c.pattern_set = false;
}

}  // namespace

namespace {
// (re)flatten when the mesh changed, otherwise refresh node coordinates and material constants
void PrepareMesh(TPZB200AssemblyCache &c, TPZStructMatrix *strmat, int device) {
    TPZCompMesh *cmesh = strmat->Mesh();
    if (!c.ctx) {
        if (b200asm_create(&c.ctx, device) != 0) Fatal(std::string("b200asm_create: ") + b200asm_last_error(nullptr));
    }
    auto t0 = clk::now();
    c.flatten_ms = 0;
    const TPZEquationFilter &filter = strmat->EquationFilter();
    const int64_t nactive = filter.IsActive() ? filter.NActiveEquations() : cmesh->NEquations();
    if (c.mesh != cmesh || c.nelem != cmesh->NElements() || c.neq != cmesh->NEquations() || c.nconnects != cmesh->NConnects() ||
        c.nactive != nactive) {
        Flatten(c, strmat);
        c.flatten_ms = ms_since(t0);
        return;
    }
    TPZGeoMesh *gmesh = cmesh->Reference();
    const int64_t nnodes = gmesh->NNodes();
    std::vector<double> xyz((size_t)nnodes * 3);
    for (int64_t i = 0; i < nnodes; i++)
        for (int d = 0; d < 3; d++) xyz[(size_t)i * 3 + d] = gmesh->NodeVec()[i].Coord(d);
    Check(c, b200asm_set_nodes(c.ctx, nnodes, xyz.data()), "b200asm_set_nodes");
    int gi = 0;
    for (HostGroup &g : c.groups) {
        FillCoef(g);
        Check(c, b200asm_set_group_coef(c.ctx, gi++, g.meta.coef), "b200asm_set_group_coef");
    }
}
}  // namespace

template <class TVar>
TPZStructMatrixB200<TVar>::TPZStructMatrixB200() : fCache(std::make_shared<TPZB200AssemblyCache>()) {}

// copies (TPZStructMatrix::Clone through TPZAnalysis::SetStructuralMatrix) start with an empty cache
template <class TVar>
TPZStructMatrixB200<TVar>::TPZStructMatrixB200(const TPZStructMatrixB200 &copy)
    : TPZStrMatParInterface(copy), fDevice(copy.fDevice), fPinHost(copy.fPinHost), fCache(std::make_shared<TPZB200AssemblyCache>()) {}

template <class TVar>
TPZStructMatrixB200<TVar> &TPZStructMatrixB200<TVar>::operator=(const TPZStructMatrixB200 &copy) {
    TPZStrMatParInterface::operator=(copy);
    fDevice = copy.fDevice;
    fPinHost = copy.fPinHost;
    fCache = std::make_shared<TPZB200AssemblyCache>();
    return *this;
}

template <class TVar>
TPZStructMatrixB200<TVar>::~TPZStructMatrixB200() = default;

template <class TVar>
void TPZStructMatrixB200<TVar>::UnpinHostMatrix() {
    fCache->Unpin();
}

template <class TVar>
void TPZStructMatrixB200<TVar>::LastTimings(double &flatten_ms, double &pattern_ms, double &assemble_ms) const {
    flatten_ms = fCache->flatten_ms;
    pattern_ms = fCache->pattern_ms;
    assemble_ms = fCache->assemble_ms;
}

template <class TVar>
void TPZStructMatrixB200<TVar>::Assemble(TPZBaseMatrix &stiffness, TPZBaseMatrix &rhs, TPZAutoPointer<TPZGuiInterface> guiInterface) {
    auto *strmat = dynamic_cast<TPZStructMatrix *>(this);
    if (!strmat) Fatal("the strategy must be combined with a TPZStructMatrix");
    const TPZEquationFilter &filter = strmat->EquationFilter();
    auto *rhsmat = dynamic_cast<TPZFMatrix<STATE> *>(&rhs);
    auto *sym = dynamic_cast<TPZSYsmpMatrix<STATE> *>(&stiffness);
    auto *full = dynamic_cast<TPZFYsmpMatrix<STATE> *>(&stiffness);
    if (!rhsmat) Fatal("rhs is not a TPZFMatrix<STATE>");
    if (!sym && !full) Fatal("the stiffness matrix must be TPZSYsmpMatrix<STATE> or TPZFYsmpMatrix<STATE>");
    TPZCompMesh *cmesh = strmat->Mesh();
    TPZB200AssemblyCache &c = *fCache;
    PrepareMesh(c, strmat, fDevice);
    if (guiInterface && guiInterface->AmIKilled()) return;
    // pattern of the matrix Create() produced
    auto t0 = clk::now();
    c.pattern_ms = 0;
    const int symmetric = sym ? 1 : 0;
    double *values = nullptr;
    int64_t nnz = 0;
    if (sym) {
        nnz = sym->JA().size();
        values = &sym->A()[0];
    } else {
        // values in place through the public TPZMatrix::Storage() (Matrix/pzmatrix.cpp:61-67)
        TPZFMatrixRef<STATE> st = full->Storage();
        nnz = st.Rows();
        values = &st(0, 0);
    }
    if (!c.pattern_set || c.symmetric != symmetric || c.nnz != nnz) {
        if (sym) {
            Check(c, b200asm_set_pattern(c.ctx, sym->Rows(), &sym->IA()[0], &sym->JA()[0], 1), "b200asm_set_pattern");
        } else {
            TPZVec<int64_t> ia, ja;
            TPZVec<STATE> a;
            full->GetData(ia, ja, a);
            Check(c, b200asm_set_pattern(c.ctx, full->Rows(), &ia[0], &ja[0], 0), "b200asm_set_pattern");
        }
        c.pattern_set = true;
        c.symmetric = symmetric;
        c.nnz = nnz;
        c.pattern_ms = ms_since(t0);
    }
    if (guiInterface && guiInterface->AmIKilled()) return;
    // assemble; the reference ADDS into rhs (TPZFMatrix::AddFel), the matrix arrives zeroed
    t0 = clk::now();
    const int64_t neq = rhsmat->Rows();
    if (neq != c.neq || rhsmat->Cols() != 1) Fatal("rhs has the wrong size (one load case, NEquations rows)");
    if (stiffness.Rows() != c.nactive) Fatal("the matrix does not have NActiveEquations rows");
    TPZFMatrix<STATE> rhsloc;  // condensed numbering when the filter is active
    double *rhsptr = nullptr;
    if (ComputeRhs()) {
        rhsloc.Redim(c.nactive, 1);
        rhsptr = &rhsloc(0, 0);
    }
    if (fPinHost && nnz > 0 && (c.pinned != values || c.pinned_bytes != (size_t)nnz * sizeof(double))) {
        c.Unpin();
        Check(c, b200asm_pin_host(c.ctx, values, (size_t)nnz * sizeof(double)), "b200asm_pin_host");
        c.pinned = values;
        c.pinned_bytes = (size_t)nnz * sizeof(double);
    }
    Check(c, b200asm_assemble(c.ctx, values, rhsptr), "b200asm_assemble");
    if (rhsptr) {
        if (filter.IsActive()) {
            filter.Scatter(rhsloc, rhs);  // StrMatrix/pzstrmatrixor.cpp:47-62
        } else {
            double *dst = &(*rhsmat)(0, 0);
            for (int64_t i = 0; i < neq; i++) dst[i] += rhsloc(i, 0);  // the reference adds into rhs (TPZFMatrix::AddFel)
        }
    }
    c.assemble_ms = ms_since(t0);
}

// Right-hand side only (TPZLinearAnalysis::AssembleResidual, Analysis/TPZLinearAnalysis.cpp:96-120; the OR strategy's
// CalcResidual loop, StrMatrix/pzstrmatrixor.cpp:253-365).  For the supported (linear) materials CalcResidual evaluates
// the same ef as CalcStiff, so the device runs the assembly kernels without the Gram products and the matrix scatter.
template <class TVar>
void TPZStructMatrixB200<TVar>::Assemble(TPZBaseMatrix &rhs, TPZAutoPointer<TPZGuiInterface> guiInterface) {
    auto *strmat = dynamic_cast<TPZStructMatrix *>(this);
    if (!strmat) Fatal("the strategy must be combined with a TPZStructMatrix");
    const TPZEquationFilter &filter = strmat->EquationFilter();
    auto *rhsmat = dynamic_cast<TPZFMatrix<STATE> *>(&rhs);
    if (!rhsmat) Fatal("rhs is not a TPZFMatrix<STATE>");
    TPZB200AssemblyCache &c = *fCache;
    PrepareMesh(c, strmat, fDevice);
    if (guiInterface && guiInterface->AmIKilled()) return;
    auto t0 = clk::now();
    const int64_t neq = rhsmat->Rows();
    if (neq != c.neq || rhsmat->Cols() != 1) Fatal("rhs has the wrong size (one load case, NEquations rows)");
    TPZFMatrix<STATE> rhsloc(c.nactive, 1, 0.);
    Check(c, b200asm_assemble_rhs(c.ctx, &rhsloc(0, 0)), "b200asm_assemble_rhs");
    if (filter.IsActive()) {
        filter.Scatter(rhsloc, rhs);  // StrMatrix/pzstrmatrixor.cpp:79-99
    } else {
        double *dst = &(*rhsmat)(0, 0);
        for (int64_t i = 0; i < neq; i++) dst[i] += rhsloc(i, 0);  // the reference ADDS into rhs (TPZFMatrix::AddFel)
    }
    c.assemble_ms = ms_since(t0);
}

// CG on the device-resident matrix of the last Assemble(stiffness, rhs): the reference's algorithm
// (Solvers/LinearSolvers/cg.h:44-120) with TPZStepSolver::SetJacobi(1, 0., 0) or no preconditioner.
template <class TVar>
void TPZStructMatrixB200<TVar>::SolveCG(const TPZFMatrix<TVar> &F, TPZFMatrix<TVar> &result, int64_t &numiterations, REAL &tol,
                                        bool jacobi, int fromcurrent) {
    TPZB200AssemblyCache &c = *fCache;
    if (!c.ctx || !c.pattern_set) Fatal("SolveCG: Assemble(stiffness, rhs) must run first (the matrix lives on the device)");
    if constexpr (std::is_same<TVar, double>::value) {
        if (F.Rows() != c.nactive || F.Cols() != 1) Fatal("SolveCG: F has the wrong size");
        if (!fromcurrent || result.Rows() != c.nactive || result.Cols() != 1) result.Redim(c.nactive, 1);
        int64_t iters = 0;
        double resid = 0;
        Check(c, b200asm_cg_solve(c.ctx, jacobi ? 1 : 0, numiterations, tol, fromcurrent, &F.g(0, 0), &result(0, 0), &iters, &resid),
              "b200asm_cg_solve");
        numiterations = iters;
        tol = resid;
    } else {
        Fatal("SolveCG: STATE = double only");
    }
}

template <class TVar>
int TPZStructMatrixB200<TVar>::ClassId() const {
    return Hash("TPZStructMatrixB200") ^ TPZStrMatParInterface::ClassId() << 1 ^ ClassIdOrHash<TVar>() << 2;
}

template <class TVar>
void TPZStructMatrixB200<TVar>::Read(TPZStream &buf, void *context) {
    TPZStrMatParInterface::Read(buf, context);
    buf.Read(&fDevice);
}

template <class TVar>
void TPZStructMatrixB200<TVar>::Write(TPZStream &buf, int withclassid) const {
    TPZStrMatParInterface::Write(buf, withclassid);
    buf.Write(&fDevice);
}

template class TPZStructMatrixB200<STATE>;
template class TPZRestoreClass<TPZStructMatrixB200<STATE>>;

// The member functions of the struct matrices are defined in their .cpp and explicitly instantiated only
// for OR/OT/TBBFlow (StrMatrix/TPZSSpStructMatrix.cpp:217-223, TPZSpStructMatrix.cpp:214-220), so a new
// parallel layer instantiates them here.
#include "TPZSSpStructMatrix.cpp"
#include "TPZSpStructMatrix.cpp"
template class TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>;
template class TPZSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>;

// ---- Create() on the device ------------------------------------------------------------------------------------
template <class TVar>
void TPZStructMatrixB200<TVar>::CreatePatternOnDevice(bool symmetric, TPZStack<int64_t> &elgraph, TPZVec<int64_t> &elgraphindex,
                                                      TPZVec<int64_t> &ia, TPZVec<int64_t> &ja) {
    auto *strmat = dynamic_cast<TPZStructMatrix *>(this);
    if (!strmat) Fatal("the strategy must be combined with a TPZStructMatrix");
    if (strmat->EquationFilter().IsActive())
        Fatal("Create() on the device does not support an active equation filter: use TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>");
    TPZCompMesh *cmesh = strmat->Mesh();
    TPZB200AssemblyCache &c = *fCache;
    if (!c.ctx) {
        if (b200asm_create(&c.ctx, fDevice) != 0) Fatal(std::string("b200asm_create: ") + b200asm_last_error(nullptr));
    }
    auto t0 = clk::now();
    // block table: TPZBlock::Position / Size per sequence number of the independent connects
    // (External/TPZRenumbering.cpp:76-110 works on NIndependentConnects() blocks)
    const int64_t nblock = cmesh->NIndependentConnects();
    std::vector<int64_t> bpos(nblock), bsize(nblock);
    for (int64_t b = 0; b < nblock; b++) {
        bpos[b] = cmesh->Block().Position(b);
        bsize[b] = cmesh->Block().Size(b);
    }
    const int64_t nel = elgraphindex.size() - 1;
    int64_t neq = 0, nnz = 0;
    Check(c, b200asm_build_pattern_device(c.ctx, symmetric ? 1 : 0, nel, &elgraphindex[0], nel && elgraph.size() ? &elgraph[0] : nullptr,
                                          nblock, bpos.data(), bsize.data(), &neq, &nnz),
          "b200asm_build_pattern_device");
    if (neq != cmesh->NEquations()) Fatal("device pattern: unexpected number of equations");
    ia.resize(neq + 1);
    ja.resize(nnz);
    Check(c, b200asm_get_pattern(c.ctx, &ia[0], nnz ? &ja[0] : nullptr), "b200asm_get_pattern");
    c.pattern_set = true;  // the next Assemble() finds its pattern resident (same nnz / storage kind)
    c.symmetric = symmetric ? 1 : 0;
    c.nnz = nnz;
    c.pattern_ms = ms_since(t0);
}

template <class TVar>
TPZMatrix<TVar> *TPZSSpStructMatrixB200<TVar>::SetupMatrixData(TPZStack<int64_t> &elgraph, TPZVec<int64_t> &elgraphindex) {
    const int64_t neq = this->fEquationFilter.NActiveEquations();
    auto *mat = new TPZSYsmpMatrix<TVar>(neq, neq);
    // written in place through the public accessors (Matrix/pzsysmp.h:114-127): no second copy of the pattern
    this->CreatePatternOnDevice(true, elgraph, elgraphindex, mat->IA(), mat->JA());
    mat->A().Resize(mat->JA().size());
    mat->A().Fill(0.);
    mat->ComputeDiagonal();  // what SetData does after storing the arrays (Matrix/pzsysmp.h:233-240)
    return mat;
}

template <class TVar>
TPZMatrix<TVar> *TPZSpStructMatrixB200<TVar>::SetupMatrixData(TPZStack<int64_t> &elgraph, TPZVec<int64_t> &elgraphindex) {
    const int64_t neq = this->fEquationFilter.NActiveEquations();
    auto *mat = new TPZFYsmpMatrix<TVar>(neq, neq);
    TPZVec<int64_t> ia, ja;
    this->CreatePatternOnDevice(false, elgraph, elgraphindex, ia, ja);
    TPZVec<TVar> a(ja.size(), 0.);
    mat->SetData(ia, ja, a);
    return mat;
}

template class TPZSSpStructMatrixB200<STATE>;
template class TPZSpStructMatrixB200<STATE>;
