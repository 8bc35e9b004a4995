// TPZStructMatrixB200 — NeoPZ-side host code of the B200 assembly strategy (see the header).
//
// Everything here runs once per mesh (flatten) or once per Assemble() (node coordinates, material
// constants, one C-ABI call); the arithmetic of CalcStiff / Contribute / AddKel / AddFel happens in the
// CUDA library (neopz_b200/csrc/b200asm.cu).
#include "TPZStructMatrixB200.h"

#include <chrono>
#include <cstring>
#include <map>
#include <thread>
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
    // the engine: ONE context (one GPU) or a b200asm_multi (several GPUs of this process); same calls either way
    b200asm_ctx *ctx = nullptr;
    b200asm_multi *multi = nullptr;
    std::vector<int> devices;  // the devices the engine was created on
    TPZCompMesh *mesh = nullptr;
    int64_t nelem = -1, neq = -1, nconnects = -1, nnz = -1;
    int64_t nactive = -1;  // equations of the assembled system: NEquations(), or NActiveEquations() of an active filter
    uint64_t signature = 0;     // hash of what the flattened arrays were derived from (MeshSignature)
    uint64_t pattern_hash = 0;  // hash of the IA / JA the engine holds
    int symmetric = -1;
    bool pattern_set = false;
    bool drop_tiny = false;
    int nloadcases = 1;
    std::vector<HostGroup> groups;
    double flatten_ms = 0, pattern_ms = 0, assemble_ms = 0;
    void *pinned = nullptr;  // value array page-locked by SetPinHostMatrix
    size_t pinned_bytes = 0;
    std::vector<double> accum;  // previous values of the matrix (SetAccumulate)

    b200asm_ctx *AnyContext() {
        if (ctx) return ctx;
        b200asm_ctx *c0 = nullptr;
        if (multi) b200asm_multi_context(multi, 0, &c0);
        return c0;
    }
    const char *LastError() const { return multi ? b200asm_multi_last_error(multi) : b200asm_last_error(ctx); }
    int SetOption(const char *name, int64_t v) { return multi ? b200asm_multi_set_option(multi, name, v) : b200asm_set_option(ctx, name, v); }
    int SetNodes(int64_t n, const double *xyz) { return multi ? b200asm_multi_set_nodes(multi, n, xyz) : b200asm_set_nodes(ctx, n, xyz); }
    int AddGroup(const b200asm_group *g) { return multi ? b200asm_multi_add_group(multi, g) : b200asm_add_group(ctx, g); }
    int SetGroupCoef(int gi, const double *coef) { return multi ? b200asm_multi_set_group_coef(multi, gi, coef) : b200asm_set_group_coef(ctx, gi, coef); }
    int SetGroupForce(int gi, const double *f) { return multi ? b200asm_multi_set_group_force(multi, gi, f) : b200asm_set_group_force(ctx, gi, f); }
    int ClearGroups() { return multi ? b200asm_multi_clear_groups(multi) : b200asm_clear_groups(ctx); }
    int SetPattern(int64_t n, const int64_t *ia, const int64_t *ja, int sym) {
        return multi ? b200asm_multi_set_pattern(multi, n, ia, ja, sym) : b200asm_set_pattern(ctx, n, ia, ja, sym);
    }
    int Assemble(double *a, double *rhs) { return multi ? b200asm_multi_assemble(multi, a, rhs) : b200asm_assemble(ctx, a, rhs); }
    int AssembleRhs(double *rhs) { return multi ? b200asm_multi_assemble_rhs(multi, rhs) : b200asm_assemble_rhs(ctx, rhs); }

    void Unpin() {
        if (pinned && AnyContext()) b200asm_unpin_host(AnyContext(), pinned);
        pinned = nullptr;
        pinned_bytes = 0;
    }
    void DestroyEngine() {
        Unpin();
        if (ctx) b200asm_destroy(ctx);
        if (multi) b200asm_multi_destroy(multi);
        ctx = nullptr;
        multi = nullptr;
        devices.clear();
        mesh = nullptr;
        nelem = neq = nconnects = nnz = nactive = -1;
        signature = pattern_hash = 0;
        pattern_set = false;
        groups.clear();
    }
    ~TPZB200AssemblyCache() { DestroyEngine(); }
};

namespace {

void Check(TPZB200AssemblyCache &c, int rc, const char *what) {
    if (rc < 0) Fatal(std::string(what) + " failed: " + c.LastError());
}

// 64-bit mix (splitmix64 finaliser) folded over a stream of words
inline void HashWord(uint64_t &h, uint64_t v) {
    v += 0x9E3779B97F4A7C15ull + h;
    v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
    v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
    h = (h * 0x100000001B3ull) ^ (v ^ (v >> 31));
}

// host threads of the per-Assemble host work (TPZStructMatrixB200::SetHostThreads), set at the entry points of the strategy
thread_local int t_host_threads = 1;

// body(first, last) over contiguous ranges of [0, n), one range per host thread (small n: the calling thread alone)
template <class F>
void ParallelFor(int64_t n, int64_t grain, F body) {
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(t_host_threads, n / std::max<int64_t>(grain, 1)));
    if (nthreads == 1) {
        body((int64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back([=]() { body(n * t / nthreads, n * (t + 1) / nthreads); });
    for (std::thread &th : pool) th.join();
}

// What the flattened arrays depend on besides node coordinates and material constants (those are refreshed at every
// Assemble): the element list, the material object and id of every element, the ShouldCompute outcome, per connect its block
// position, block size and order, the geometric corner nodes, and the equation filter.  Connect renumbering (Permute), p
// changes, SetMaterialIds and replaced material objects all change it.
uint64_t MeshSignature(TPZStructMatrix *strmat) {
    TPZCompMesh *cmesh = strmat->Mesh();
    uint64_t h = 0x243F6A8885A308D3ull;
    const int64_t nel = cmesh->NElements();
    HashWord(h, (uint64_t)nel);
    HashWord(h, (uint64_t)cmesh->NEquations());
    const TPZEquationFilter &filter = strmat->EquationFilter();
    HashWord(h, filter.IsActive() ? (uint64_t)filter.NActiveEquations() + 1 : 0);
    // read-only walk over the elements: contiguous element ranges are hashed by separate threads (the walk is the largest
    // host cost of a repeated Assemble() on a cached mesh), the range hashes are folded in element order
    auto range_hash = [&](int64_t e0, int64_t e1) {
        uint64_t hr = 0x452821E638D01377ull;
        for (int64_t iel = e0; iel < e1; iel++) {
            TPZCompEl *cel = cmesh->Element(iel);
            if (!cel) { HashWord(hr, 1); continue; }
            TPZMaterial *mat = cel->Material();
            if (!mat) { HashWord(hr, 2); continue; }
            if (!strmat->ShouldCompute(mat->Id())) { HashWord(hr, 3); continue; }
            HashWord(hr, (uint64_t)(uintptr_t)mat);
            HashWord(hr, (uint64_t)(int64_t)mat->Id());
            const int ncon = cel->NConnects();
            for (int i = 0; i < ncon; i++) {
                TPZConnect &con = cel->Connect(i);
                const int64_t seq = con.SequenceNumber();
                HashWord(hr, (uint64_t)cmesh->Block().Position(seq) * 64u + (uint64_t)con.Order());
                HashWord(hr, (uint64_t)cmesh->Block().Size(seq));
            }
            TPZGeoEl *gel = cel->Reference();
            if (gel) {
                const int nc = gel->NCornerNodes();
                for (int i = 0; i < nc; i++) HashWord(hr, (uint64_t)gel->NodeIndex(i));
            }
        }
        return hr;
    };
    // fixed blocks of 4096 elements, whatever the thread count: the signature does not depend on SetHostThreads
    constexpr int64_t kBlock = 4096;
    const int64_t nblocks = (nel + kBlock - 1) / kBlock;
    std::vector<uint64_t> part((size_t)nblocks, 0);
    ParallelFor(nblocks, 1, [&](int64_t b0, int64_t b1) {
        for (int64_t b = b0; b < b1; b++) part[b] = range_hash(b * kBlock, std::min(nel, (b + 1) * kBlock));
    });
    for (int64_t b = 0; b < nblocks; b++) HashWord(h, part[b]);
    return h;
}

uint64_t PatternHash(int64_t neq, const int64_t *ia, const int64_t *ja) {
    uint64_t h = 0x13198A2E03707344ull;
    HashWord(h, (uint64_t)neq);
    for (int64_t i = 0; i <= neq; i++) HashWord(h, (uint64_t)ia[i]);
    const int64_t nnz = ia[neq];
    const int64_t stride = std::max<int64_t>(1, nnz / (1 << 16));  // the row pointers pin the shape; the columns are sampled (65 k probes)
    for (int64_t k = 0; k < nnz; k += stride) HashWord(h, (uint64_t)ja[k]);
    return h;
}

// material constants of a group, re-read at every Assemble (they may change between assemblies)
// load case lc (TPZMatLoadCases: only TPZMatPoisson evaluates more than one, Material/Poisson/TPZMatPoisson.cpp:23,56-70)
void FillCoef(HostGroup &g, int lc = 0) {
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
            // the value of load case lc: TPZMatLoadCasesBC::GetBCRhsVal (TPZMatPoisson.cpp:64-68; Val2() without a value vector)
            double v2l = v2[0];
            if (auto *lcbc = dynamic_cast<TPZMatLoadCasesBC<STATE> *>(bc)) v2l = lcbc->GetBCRhsVal(lc)[0];
            if (type == 0) {  // Material/Poisson/TPZMatPoisson.cpp:79-90
                coef[0] = big;
                coef[9] = big * v2l;
            } else if (type == 1) {  // :93-100
                coef[9] = v2l * pois->ScaleFactor();
            } else if (type == 2) {  // :104-118 as the reference computes it (then "not implemented" is printed): penalty load
                                     // vector and BigNumber * Val1(0,0) * dphix(0,i) * dphix(0,j) in the matrix
                coef[9] = big * v2l;
                coef[12] = big * v1.GetVal(0, 0);
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
            } else if (type == 4) {
                // stress-field Neumann, :724-737: ef += -(Val1 . normal) phi w; the normal of every integration point is
                // tabulated by FillForce (the kernel reads the load-vector coefficients from the table), no matrix part
            } else if (type >= 5 && type <= 8) {  // directional Dirichlet on x / y / z / x and z, :739-772
                const bool on[3] = {type == 5 || type == 8, type == 6, type == 7 || type == 8};
                for (int a = 0; a < 3; a++)
                    if (on[a]) {
                        coef[a * 3 + a] = big;
                        coef[9 + a] = big * v2[a];
                    }
            } else {
                Fatal("TPZElasticity3D boundary condition type " + std::to_string(type) + " is not supported");
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
// the outward unit normal of a boundary face at a point, as TPZInterpolationSpace::ComputeNormal builds it
// (Mesh/pzinterpolationspace.cpp:300-383): axes(0) x axes(1) of TPZGeoEl::Jacobian, normalised, turned so that it points away
// from the centre of the neighbouring volume element
void FaceNormal(TPZCompMesh *cmesh, TPZGeoEl *gel, TPZVec<REAL> &qsi, const TPZVec<REAL> &towards, TPZVec<REAL> &normal) {
    TPZFNMatrix<9, REAL> jac, axes, jacinv;
    REAL detjac;
    gel->Jacobian(qsi, jac, axes, detjac, jacinv);
    normal.Resize(3);
    normal[0] = axes(0, 1) * axes(1, 2) - axes(0, 2) * axes(1, 1);
    normal[1] = -axes(0, 0) * axes(1, 2) + axes(0, 2) * axes(1, 0);
    normal[2] = axes(0, 0) * axes(1, 1) - axes(0, 1) * axes(1, 0);
    REAL size = 0.;
    for (int i = 0; i < 3; i++) size += normal[i] * normal[i];
    size = sqrt(size);
    for (int i = 0; i < 3; i++) normal[i] /= size;
    REAL dot = 0.;
    for (int i = 0; i < 3; i++) dot += normal[i] * towards[i];
    if (dot < 0.)
        for (int i = 0; i < 3; i++) normal[i] *= -1.;
}

// vector from the centre of the neighbouring volume element to the centre of the face (ComputeNormal's `vec`, :316-345)
bool OutwardVector(TPZCompMesh *cmesh, TPZGeoEl *gel, TPZVec<REAL> &vec) {
    const int face = gel->NSides() - 1;
    TPZGeoElSide thisside(gel, face);
    TPZGeoElSide neigh = gel->Neighbour(face);
    while (neigh != thisside) {
        const int matid = neigh.Element()->MaterialId();
        if (cmesh->FindMaterial(matid) && neigh.Element()->Dimension() > gel->Dimension()) break;
        neigh = neigh.Neighbour();
    }
    TPZGeoEl *vol = neigh.Element();
    if (vol == gel) return false;
    TPZManVector<REAL, 3> c0(gel->Dimension(), 0.), c1(vol->Dimension(), 0.), x0(3, 0.), x1(3, 0.);
    gel->CenterPoint(face, c0);
    vol->CenterPoint(vol->NSides() - 1, c1);
    gel->X(c0, x0);
    vol->X(c1, x1);
    vec.Resize(3);
    for (int i = 0; i < 3; i++) vec[i] = -x1[i] + x0[i];
    return true;
}

void FillForce(HostGroup &g, int lc = 0) {
    g.has_forcing = false;
    g.meta.force = nullptr;
    const int ns = g.meta.nstate;
    if (g.meta.kind == B200ASM_BC) {
        if (auto *bc4 = dynamic_cast<TPZBndCondT<STATE> *>(g.material))
            if (bc4->Type() == 4 && dynamic_cast<TPZElasticity3D *>(bc4->Material())) {
                // TPZElasticity3D stress-field Neumann (TPZElasticity3D.cpp:724-737): val2loc = -(Val1 . normal) per point
                if (g.meta.topology != B200ASM_QUAD && g.meta.topology != B200ASM_TRI) Fatal("boundary condition type 4 on an element that is not a face");
                const int nq = g.meta.nqp;
                g.force.assign((size_t)g.meta.nel * nq * 3, 0.0);
                TPZManVector<REAL, 3> qsi(2), x(3), vec(3, 0.), normal(3, 0.);
                for (int64_t e = 0; e < g.meta.nel; e++) {
                    TPZGeoEl *gel = g.elements[e]->Reference();
                    const bool has_neighbour = OutwardVector(g.elements[e]->Mesh(), gel, vec);
                    for (int q = 0; q < nq; q++) {
                        for (int d = 0; d < 2; d++) qsi[d] = g.qpts[(size_t)q * 2 + d];
                        TPZFNMatrix<9, STATE> v1(bc4->Val1());
                        if (bc4->HasForcingFunctionBC()) {  // :637-645: the function may replace Val1 (and val2loc[0], unused here)
                            gel->X(qsi, x);
                            TPZManVector<STATE, 3> v2(3, 0.);
                            bc4->ForcingFunctionBC()(x, v2, v1);
                        }
                        if (has_neighbour) FaceNormal(g.elements[e]->Mesh(), gel, qsi, vec, normal); else normal.Fill(0.);
                        double *out = &g.force[((size_t)e * nq + q) * 3];
                        for (int a = 0; a < 3; a++) out[a] = -(v1(a, 0) * normal[0] + v1(a, 1) * normal[1] + v1(a, 2) * normal[2]);
                    }
                }
                g.has_forcing = true;
                g.meta.force = g.force.data();
                return;
            }
        // boundary data given by a function (TPZBndCondT::ForcingFunctionBC): evaluated on the host at data.x of every
        // integration point, stored as the coefficient of phi_i * weight in ef (what the constant case keeps in coef[9..11])
        auto *bc = dynamic_cast<TPZBndCondT<STATE> *>(g.material);
        if (!bc || !bc->HasForcingFunctionBC()) return;
        TPZMaterial *vol = bc->Material();
        auto *pois = dynamic_cast<TPZMatPoisson<STATE> *>(vol);
        auto *e2 = dynamic_cast<TPZElasticity2D *>(vol);
        auto *e3 = dynamic_cast<TPZElasticity3D *>(vol);
        const int type = bc->Type();
        const bool ok = pois ? (type == 0 || type == 1 || type == 2) : (e2 ? (type == 0 || type == 1) : (e3 ? (type == 0 || type == 2 || (type >= 5 && type <= 8)) : false));
        if (!ok) Fatal("boundary condition type " + std::to_string(type) + " with a forcing function is not supported");
        const int fdim = g.meta.topology == B200ASM_LINE ? 1 : 2;
        const int nq = g.meta.nqp;
        g.force.assign((size_t)g.meta.nel * nq * ns, 0.0);
        ParallelFor(g.meta.nel, 512, [&, fdim, nq, ns, type, lc](int64_t e_first, int64_t e_last) {
        TPZManVector<REAL, 3> qsi(fdim), x(3);
        for (int64_t e = e_first; e < e_last; e++) {
            TPZGeoEl *gel = g.elements[e]->Reference();
            for (int q = 0; q < nq; q++) {
                for (int d = 0; d < fdim; d++) qsi[d] = g.qpts[(size_t)q * fdim + d];
                gel->X(qsi, x);
                TPZManVector<STATE, 10> v2(e3 ? 3 : (pois ? std::max(1, pois->NumLoadCases()) : ns), 0.);
                TPZFNMatrix<9, STATE> v1(bc->Val1());
                bc->ForcingFunctionBC()(x, v2, v1);
                double *out = &g.force[((size_t)e * nq + q) * ns];
                if (pois) {  // Material/Poisson/TPZMatPoisson.cpp:79-118: v2[nvars * l + iv] of load case l
                    out[0] = type == 1 ? v2[lc] * pois->ScaleFactor() : pois->BigNumber() * v2[lc];
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
        });
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
        ParallelFor(g.meta.nel, 512, [&, nq, ns, lc](int64_t e_first, int64_t e_last) {
        TPZManVector<REAL, 3> qsi(2), x(3);
        for (int64_t e = e_first; e < e_last; e++) {
            TPZGeoEl *gel = g.elements[e]->Reference();
            for (int q = 0; q < nq; q++) {
                for (int d = 0; d < 2; d++) qsi[d] = g.qpts[(size_t)q * 2 + d];
                gel->X(qsi, x);
                TPZManVector<STATE, 10> f(p2 ? std::max(3, p2->NumLoadCases()) : 3, 0.);
                if (p2) p2->ForcingFunction()(x, f); else e2->ForcingFunction()(x, f);
                if (p2) g.force[((size_t)e * nq + q)] = f[lc];
                else for (int k = 0; k < ns; k++) g.force[((size_t)e * nq + q) * ns + k] = f[k];
            }
        }
        });
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
    // (the reference calls the function from every assembly thread at once, Material/Poisson/TPZMatPoisson.cpp:24-27 under
    //  StrMatrix/pzstrmatrixor.cpp:476-513; so does this table, range by range)
    ParallelFor(g.meta.nel, 256, [&, nq, ns, lc](int64_t e_first, int64_t e_last) {
    TPZManVector<REAL, 3> qsi(3), x(3);
    for (int64_t e = e_first; e < e_last; e++) {
        TPZGeoEl *gel = g.elements[e]->Reference();
        for (int q = 0; q < nq; q++) {
            for (int d = 0; d < 3; d++) qsi[d] = g.qpts[(size_t)q * 3 + d];
            gel->X(qsi, x);
            TPZManVector<STATE, 10> f(pois ? std::max(1, pois->NumLoadCases()) : ns, 0.);
            if (pois) {
                pois->ForcingFunction()(x, f);
                f[0] = f[lc];
            } else {
                auto *acc = static_cast<ElastAccess *>(el);
                for (int k = 0; k < 3; k++) f[k] = acc->fForce[k];  // locForce(fForce) then the callback
                el->ForcingFunction()(x, f);
            }
            for (int k = 0; k < ns; k++) g.force[((size_t)e * nq + q) * ns + k] = f[k];
        }
    }
    });
    g.has_forcing = true;
    g.meta.force = g.force.data();
}

void Flatten(TPZB200AssemblyCache &c, TPZStructMatrix *strmat) {
    TPZCompMesh *cmesh = strmat->Mesh();
    c.groups.clear();
    c.nloadcases = 1;
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
            if (auto *lc = dynamic_cast<TPZMatLoadCasesBase *>(mat)) {
                // several load cases: only TPZMatPoisson evaluates them (TPZElasticity3D / 2D write column 0 only)
                if (lc->NumLoadCases() > 1 && !isbc && !dynamic_cast<TPZMatPoisson<STATE> *>(mat))
                    Fatal("more than one load case is supported for TPZMatPoisson only");
                c.nloadcases = std::max(c.nloadcases, lc->NumLoadCases());
            }
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
    Check(c, c.ClearGroups(), "b200asm_clear_groups");
    Check(c, c.SetNodes(nnodes, xyz.data()), "b200asm_set_nodes");
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
        Check(c, c.AddGroup(&g.meta), "b200asm_add_group");
    }
    c.mesh = cmesh;
    c.nelem = nel;
    c.neq = cmesh->NEquations();
    c.nactive = filter.IsActive() ? filter.NActiveEquations() : c.neq;
    c.nconnects = cmesh->NConnects();
    // One context keeps a resident pattern across b200asm_clear_groups (only the scatter maps are rebuilt), e.g. the pattern
    // Create() built on the device; Assemble() compares its hash with the matrix it is handed.  Several GPUs: the row
    // partition follows the elements, so the pattern is set again.
    if (c.multi) c.pattern_set = false;
}

}  // namespace

namespace {
// the devices of the engine: an explicit list, else the first min(n, present) devices for SetNumThreads(n >= 2), else one
std::vector<int> WantedDevices(int device, const std::vector<int> &devices, int numthreads) {
    if (!devices.empty()) return devices;
    if (numthreads >= 2) {
        const int present = b200asm_device_count();
        const int n = std::min(numthreads, std::min(present, B200ASM_MAX_PEERS / 2));
        if (n >= 2) {
            std::vector<int> d(n);
            for (int k = 0; k < n; k++) d[k] = k;
            return d;
        }
    }
    return {device};
}

void EnsureEngine(TPZB200AssemblyCache &c, const std::vector<int> &want, bool drop_tiny) {
    if ((c.ctx || c.multi) && c.devices == want && c.drop_tiny == drop_tiny) return;
    c.DestroyEngine();
    if (want.size() == 1) {
        if (b200asm_create(&c.ctx, want[0]) != 0) Fatal(std::string("b200asm_create: ") + b200asm_last_error(nullptr));
    } else {
        if (b200asm_multi_create(&c.multi, (int)want.size(), want.data()) != 0)
            Fatal(std::string("b200asm_multi_create: ") + b200asm_multi_last_error(nullptr));
    }
    c.devices = want;
    c.drop_tiny = drop_tiny;
    if (drop_tiny) Check(c, c.SetOption("drop_tiny", 1), "option drop_tiny");
}

// (re)flatten when the mesh changed, otherwise refresh node coordinates, material constants and forcing tables
void PrepareMesh(TPZB200AssemblyCache &c, TPZStructMatrix *strmat, bool check_mesh, bool static_forcing) {
    TPZCompMesh *cmesh = strmat->Mesh();
    auto t0 = clk::now();
    c.flatten_ms = 0;
    const TPZEquationFilter &filter = strmat->EquationFilter();
    const int64_t nactive = filter.IsActive() ? filter.NActiveEquations() : cmesh->NEquations();
    bool stale = c.mesh != cmesh || c.nelem != cmesh->NElements() || c.neq != cmesh->NEquations() || c.nconnects != cmesh->NConnects() ||
                 c.nactive != nactive;
    uint64_t sig = c.signature;
    if (!stale && check_mesh) {
        sig = MeshSignature(strmat);
        stale = sig != c.signature;
    }
    if (stale) {
        Flatten(c, strmat);
        c.signature = check_mesh ? (sig != c.signature ? sig : MeshSignature(strmat)) : 0;
        if (check_mesh && c.signature == 0) c.signature = 1;
        c.flatten_ms = ms_since(t0);
        return;
    }
    TPZGeoMesh *gmesh = cmesh->Reference();
    const int64_t nnodes = gmesh->NNodes();
    std::vector<double> xyz((size_t)nnodes * 3);
    for (int64_t i = 0; i < nnodes; i++)
        for (int d = 0; d < 3; d++) xyz[(size_t)i * 3 + d] = gmesh->NodeVec()[i].Coord(d);
    Check(c, c.SetNodes(nnodes, xyz.data()), "b200asm_set_nodes");
    int gi = 0;
    for (HostGroup &g : c.groups) {
        FillCoef(g);
        Check(c, c.SetGroupCoef(gi, g.meta.coef), "b200asm_set_group_coef");
        // forcing functions and boundary functions are called again at every assembly, like the reference's Contribute does
        // (they may depend on time, and the nodes may have moved)
        if (g.has_forcing && !static_forcing) {
            FillForce(g);
            if (!g.has_forcing) Fatal("a forcing function was removed from a material: call Invalidate()");
            Check(c, c.SetGroupForce(gi, g.force.data()), "b200asm_set_group_force");
        }
        gi++;
    }
    c.flatten_ms = ms_since(t0);  // (cached mesh: signature walk + refresh of coordinates, constants and tables)
}

// load case lc >= 1: the groups whose data depend on the load case get the coefficients / tables of that case
void SelectLoadCase(TPZB200AssemblyCache &c, int lc) {
    int gi = 0;
    for (HostGroup &g : c.groups) {
        FillCoef(g, lc);
        Check(c, c.SetGroupCoef(gi, g.meta.coef), "b200asm_set_group_coef");
        if (g.has_forcing) {
            FillForce(g, lc);
            Check(c, c.SetGroupForce(gi, g.force.data()), "b200asm_set_group_force");
        }
        gi++;
    }
}
}  // namespace

template <class TVar>
TPZStructMatrixB200<TVar>::TPZStructMatrixB200() : fCache(std::make_shared<TPZB200AssemblyCache>()) {}

// copies (TPZStructMatrix::Clone through TPZAnalysis::SetStructuralMatrix) start with an empty cache
template <class TVar>
TPZStructMatrixB200<TVar>::TPZStructMatrixB200(const TPZStructMatrixB200 &copy)
    : TPZStrMatParInterface(copy), fDevice(copy.fDevice), fDevices(copy.fDevices), fPinHost(copy.fPinHost), fAccumulate(copy.fAccumulate),
      fDropTiny(copy.fDropTiny), fStaticForcing(copy.fStaticForcing), fCheckMesh(copy.fCheckMesh), fHostThreads(copy.fHostThreads),
      fCache(std::make_shared<TPZB200AssemblyCache>()) {}

template <class TVar>
TPZStructMatrixB200<TVar> &TPZStructMatrixB200<TVar>::operator=(const TPZStructMatrixB200 &copy) {
    TPZStrMatParInterface::operator=(copy);
    fDevice = copy.fDevice;
    fDevices = copy.fDevices;
    fPinHost = copy.fPinHost;
    fAccumulate = copy.fAccumulate;
    fDropTiny = copy.fDropTiny;
    fStaticForcing = copy.fStaticForcing;
    fCheckMesh = copy.fCheckMesh;
    fHostThreads = copy.fHostThreads;
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
void TPZStructMatrixB200<TVar>::Invalidate() {
    fCache->mesh = nullptr;
    fCache->signature = 0;
    fCache->pattern_set = false;
}

template <class TVar>
int TPZStructMatrixB200<TVar>::NumDevicesUsed() const {
    return (int)fCache->devices.size();
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
    TPZB200AssemblyCache &c = *fCache;
    t_host_threads = fHostThreads > 0 ? fHostThreads : std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));
    EnsureEngine(c, WantedDevices(fDevice, fDevices, this->fNumThreads), fDropTiny);
    PrepareMesh(c, strmat, fCheckMesh, fStaticForcing);
    if (guiInterface && guiInterface->AmIKilled()) return;
    // pattern of the matrix Create() produced
    auto t0 = clk::now();
    c.pattern_ms = 0;
    const int symmetric = sym ? 1 : 0;
    double *values = nullptr;
    int64_t nnz = 0;
    TPZVec<int64_t> ia_full, ja_full;
    const int64_t *ia = nullptr, *ja = nullptr;
    if (sym) {
        nnz = sym->JA().size();
        values = &sym->A()[0];
        ia = &sym->IA()[0];
        ja = nnz ? &sym->JA()[0] : nullptr;
    } else {
        // values in place through the public TPZMatrix::Storage() (Matrix/pzmatrix.cpp:61-67)
        TPZFMatrixRef<STATE> st = full->Storage();
        nnz = st.Rows();
        values = &st(0, 0);
    }
    // Is the pattern the engine holds the pattern of this matrix?  Symmetric storage exposes IA / JA: their hash decides.
    // TPZFYsmpMatrix only hands out copies (GetData, Matrix/pzysmp.h:284-288): fetched when the sizes disagree.
    bool upload = !c.pattern_set || c.symmetric != symmetric || c.nnz != nnz;
    uint64_t phash = c.pattern_hash;
    if (sym) {
        phash = PatternHash(stiffness.Rows(), ia, ja);
        upload = upload || phash != c.pattern_hash;
    } else if (upload) {
        TPZVec<STATE> a;
        full->GetData(ia_full, ja_full, a);
        ia = &ia_full[0];
        ja = nnz ? &ja_full[0] : nullptr;
        phash = PatternHash(stiffness.Rows(), ia, ja);
    }
    if (upload) {
        Check(c, c.SetPattern(stiffness.Rows(), ia, ja, symmetric), "b200asm_set_pattern");
        c.pattern_set = true;
        c.symmetric = symmetric;
        c.nnz = nnz;
        c.pattern_hash = phash;
    }
    c.pattern_ms = ms_since(t0);  // (pattern hash; + upload when the pattern changed)
    if (guiInterface && guiInterface->AmIKilled()) return;
    // assemble; the reference ADDS into rhs (TPZFMatrix::AddFel), the matrix arrives zeroed
    t0 = clk::now();
    const int64_t neq = rhsmat->Rows();
    const int ncols = (int)rhsmat->Cols();
    if (neq != c.neq) Fatal("rhs has the wrong size (NEquations rows)");
    if (ComputeRhs() && ncols < c.nloadcases)
        Fatal("rhs has " + std::to_string(ncols) + " columns but the materials define " + std::to_string(c.nloadcases) + " load cases");
    if (stiffness.Rows() != c.nactive) Fatal("the matrix does not have NActiveEquations rows");
    TPZFMatrix<STATE> rhsloc;  // condensed numbering when the filter is active; one column per load case
    double *rhsptr = nullptr;
    if (ComputeRhs()) {
        rhsloc.Redim(c.nactive, c.nloadcases);
        rhsptr = &rhsloc(0, 0);
    }
    if (fPinHost && nnz > 0 && (c.pinned != values || c.pinned_bytes != (size_t)nnz * sizeof(double))) {
        c.Unpin();
        Check(c, b200asm_pin_host(c.AnyContext(), values, (size_t)nnz * sizeof(double)), "b200asm_pin_host");
        c.pinned = values;
        c.pinned_bytes = (size_t)nnz * sizeof(double);
    }
    if (fAccumulate) c.accum.assign(values, values + nnz);  // TPZSYsmpMatrix::AddKel adds to what the matrix holds
    Check(c, c.Assemble(values, rhsptr), "b200asm_assemble");
    if (fAccumulate) {
        for (int64_t k = 0; k < nnz; k++) values[k] += c.accum[k];
        c.accum.clear();
        c.accum.shrink_to_fit();
    }
    if (rhsptr) {
        // the other load cases (TPZMatPoisson.cpp:23-41,56-100): ek does not depend on the load case, so columns 1.. are
        // load-vector-only assemblies with the data of that case
        for (int lc = 1; lc < c.nloadcases; lc++) {
            if (guiInterface && guiInterface->AmIKilled()) return;
            SelectLoadCase(c, lc);
            Check(c, c.AssembleRhs(&rhsloc(0, lc)), "b200asm_assemble_rhs");
        }
        if (c.nloadcases > 1) SelectLoadCase(c, 0);
        if (filter.IsActive()) {
            filter.Scatter(rhsloc, rhs);  // StrMatrix/pzstrmatrixor.cpp:47-62
        } else {
            for (int lc = 0; lc < c.nloadcases; lc++) {
                double *dst = &(*rhsmat)(0, lc);
                const double *src = &rhsloc(0, lc);
                for (int64_t i = 0; i < neq; i++) dst[i] += src[i];  // the reference adds into rhs (TPZFMatrix::AddFel)
            }
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
    t_host_threads = fHostThreads > 0 ? fHostThreads : std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));
    EnsureEngine(c, WantedDevices(fDevice, fDevices, this->fNumThreads), fDropTiny);
    PrepareMesh(c, strmat, fCheckMesh, fStaticForcing);
    if (guiInterface && guiInterface->AmIKilled()) return;
    auto t0 = clk::now();
    const int64_t neq = rhsmat->Rows();
    if (neq != c.neq) Fatal("rhs has the wrong size (NEquations rows)");
    if (rhsmat->Cols() < c.nloadcases) Fatal("rhs has fewer columns than the materials define load cases");
    if (c.multi && !c.pattern_set) Fatal("Assemble(rhs) on several GPUs needs the row partition of a previous Assemble(stiffness, rhs)");
    TPZFMatrix<STATE> rhsloc(c.nactive, c.nloadcases, 0.);
    for (int lc = 0; lc < c.nloadcases; lc++) {
        if (lc > 0) SelectLoadCase(c, lc);
        Check(c, c.AssembleRhs(&rhsloc(0, lc)), "b200asm_assemble_rhs");
    }
    if (c.nloadcases > 1) SelectLoadCase(c, 0);
    if (filter.IsActive()) {
        filter.Scatter(rhsloc, rhs);  // StrMatrix/pzstrmatrixor.cpp:79-99
    } else {
        for (int lc = 0; lc < c.nloadcases; lc++) {
            double *dst = &(*rhsmat)(0, lc);
            const double *src = &rhsloc(0, lc);
            for (int64_t i = 0; i < neq; i++) dst[i] += src[i];  // the reference ADDS into rhs (TPZFMatrix::AddFel)
        }
    }
    c.assemble_ms = ms_since(t0);
}

// CG on the device-resident matrix of the last Assemble(stiffness, rhs): the reference's algorithm
// (Solvers/LinearSolvers/cg.h:44-120) with TPZStepSolver::SetJacobi(1, 0., 0) or no preconditioner.
template <class TVar>
void TPZStructMatrixB200<TVar>::SolveCG(const TPZFMatrix<TVar> &F, TPZFMatrix<TVar> &result, int64_t &numiterations, REAL &tol,
                                        bool jacobi, int fromcurrent) {
    TPZB200AssemblyCache &c = *fCache;
    if ((!c.ctx && !c.multi) || !c.pattern_set) Fatal("SolveCG: Assemble(stiffness, rhs) must run first (the matrix lives on the device)");
    if constexpr (std::is_same<TVar, double>::value) {
        if (F.Rows() != c.nactive || F.Cols() != 1) Fatal("SolveCG: F has the wrong size");
        if (!fromcurrent || result.Rows() != c.nactive || result.Cols() != 1) result.Redim(c.nactive, 1);
        int64_t iters = 0;
        double resid = 0;
        // several GPUs: every GPU multiplies its row block, halos over NVLink (b200asm_multi_cg_solve)
        const int rc = c.multi ? b200asm_multi_cg_solve(c.multi, jacobi ? 1 : 0, numiterations, tol, fromcurrent, &F.g(0, 0), &result(0, 0), &iters, &resid)
                               : b200asm_cg_solve(c.ctx, jacobi ? 1 : 0, numiterations, tol, fromcurrent, &F.g(0, 0), &result(0, 0), &iters, &resid);
        if (rc < 0) Fatal(std::string("b200asm_cg_solve failed: ") + c.LastError());
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
    t_host_threads = fHostThreads > 0 ? fHostThreads : std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));
    EnsureEngine(c, WantedDevices(fDevice, fDevices, this->fNumThreads), fDropTiny);
    b200asm_ctx *pctx = c.AnyContext();  // (several GPUs: the first one builds the pattern, Assemble() then shards it)
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
    if (b200asm_build_pattern_device(pctx, symmetric ? 1 : 0, nel, &elgraphindex[0], nel && elgraph.size() ? &elgraph[0] : nullptr,
                                     nblock, bpos.data(), bsize.data(), &neq, &nnz) < 0)
        Fatal(std::string("b200asm_build_pattern_device failed: ") + b200asm_last_error(pctx));
    if (neq != cmesh->NEquations()) Fatal("device pattern: unexpected number of equations");
    ia.resize(neq + 1);
    ja.resize(nnz);
    if (b200asm_get_pattern(pctx, &ia[0], nnz ? &ja[0] : nullptr) < 0) Fatal(std::string("b200asm_get_pattern failed: ") + b200asm_last_error(pctx));
    // one GPU: the next Assemble() finds its pattern resident (same hash / nnz / storage kind) and only builds the scatter maps
    c.pattern_set = c.multi == nullptr;
    c.symmetric = symmetric ? 1 : 0;
    c.nnz = nnz;
    c.pattern_hash = PatternHash(neq, &ia[0], nnz ? &ja[0] : nullptr);
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
