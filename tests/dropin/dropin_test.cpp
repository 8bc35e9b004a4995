// Drop-in test of the B200 strategy inside the UNMODIFIED reference (test code; reads like
// UnitTest_PZ/TestStruct/StructMatrixUnitTest.cpp "Compare parallel and serial matrices", :215-246).
//
// Same mesh, same TPZLinearAnalysis user code, two struct matrices:
//   TPZSSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>    (reference CPU path, serial + threaded)
//   TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>  (this repo, CUDA)
// Checks: IA/JA identical (memcmp), ||A_gpu - A_ref||_F / ||A_ref||_F <= 1e-12 (all rows and rows without
// penalty entries), rhs likewise, and the reference's own CG (TPZStepSolver::SetCG, Jacobi preconditioner)
// on both systems gives solutions within 1e-10.  Prints one JSON line; exit code 0 iff all checks pass.
#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <thread>

#include "Elasticity/TPZElasticity2D.h"
#include "Elasticity/TPZElasticity3D.h"
#include "Poisson/TPZMatPoisson.h"
#include "TPZBndCondT.h"
#include "TPZGeoMeshTools.h"
#include "TPZLinearAnalysis.h"
#include "TPZSSpStructMatrix.h"
#include "TPZSpStructMatrix.h"
#include "TPZStructMatrixB200.h"
#include "pzcmesh.h"
#include "pzgmesh.h"
#include "pzgnode.h"
#include "pzintel.h"
#include "pzstepsolver.h"
#include "pzsysmp.h"
#include "pzysmp.h"

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

// environment knobs (kept out of the positional arguments the existing tests use)
static int EnvInt(const char *name, int def) { const char *v = getenv(name); return v ? atoi(v) : def; }
static double EnvDouble(const char *name, double def) { const char *v = getenv(name); return v ? atof(v) : def; }
static int g_gpus = 0;        // B200_GPUS: TPZStructMatrixB200::SetNumThreads(n) = GPUs that share the assembly
static int g_bctype = 1;      // B200_BCTYPE: type of the boundary condition on matid -2 (default 1, Neumann)
static int g_loadcases = 1;   // B200_LOADCASES: TPZMatPoisson with several load cases (rhs columns)
static int g_droptiny = 0;    // B200_DROPTINY: SetDropTinyEntries(true) (AddKel's IsZero drop, Matrix/pzsysmp.cpp:381)
static double g_scale = 1.0;  // B200_SCALE: TPZMatPoisson::SetScaleFactor (a tiny factor puts element entries below AddKel's 1e-12 drop)
static int g_skip_serial = 0; // B200_SKIP_SERIAL: the threaded OR run is the reference (large meshes: timing runs)
static int g_accumulate = 0;  // B200_ACCUMULATE: a second Assemble() straight into the non-zero matrix, both strategies

// dirichlet: value imposed on matid -1.  The CG comparison uses 0: with a non-zero value the right-hand side norm
// is ~1e16 (penalty), and a relative residual tolerance of 1e-15 no longer constrains the interior equations.
// phys 2 / 3: TPZElasticity2D (plane strain / plane stress) on a plane mesh of quadrilaterals (tet = 0) or triangles
// tet (3-D): 0 hexahedra, 1 tetrahedra, 2 prisms (EPrismatic), 3 hexahedra + pyramids (EHexaPyrMixed)
// prefine: every third domain element is p-refined by one order (TPZInterpolatedElement::PRefine): elements, faces and edges
// of different order meet, the boundary elements follow in AdjustBoundaryElements
// bcfunc: the Dirichlet data on matid -1 come from a function of x (TPZBndCondT::SetForcingFunctionBC)
static TPZCompMesh *BuildMesh(int n, int p, int phys, int tet, double perturb, double dirichlet, int prefine = 0, int bcfunc = 0) {
    const int dim = phys >= 2 ? 2 : 3;
    TPZManVector<REAL, 3> minX(3, 0.), maxX(3, 1.);
    TPZManVector<int, 7> matids(dim == 3 ? 7 : 5, -1);
    matids[0] = 1;
    matids[dim == 3 ? 6 : 3] = -2;  // zmax face / top side: Neumann
    TPZManVector<int, 3> ndiv(dim, n);
    TPZGeoMesh *gmesh = dim == 3
        ? TPZGeoMeshTools::CreateGeoMeshOnGrid(3, minX, maxX, matids, ndiv,
                                               tet == 1 ? MMeshType::ETetrahedral : (tet == 2 ? MMeshType::EPrismatic : (tet == 3 ? MMeshType::EHexaPyrMixed : MMeshType::EHexahedral)), true)
        : TPZGeoMeshTools::CreateGeoMeshOnGrid(2, minX, maxX, matids, ndiv, tet ? MMeshType::ETriangular : MMeshType::EQuadrilateral, true);
    if (perturb != 0.0) {
        const double h = 1.0 / n;
        for (int64_t i = 0; i < gmesh->NNodes(); i++)
            for (int d = 0; d < dim; d++) {
                TPZGeoNode &nd = gmesh->NodeVec()[i];
                nd.SetCoord(d, nd.Coord(d) + perturb * h * std::sin(2.0 * M_PI * (double)i / 97.0 + (double)d));
            }
    }
    TPZCompMesh *cmesh = new TPZCompMesh(gmesh);
    cmesh->SetDimModel(dim);
    cmesh->SetDefaultOrder(p);
    if (phys >= 2) {
        auto *m = new TPZElasticity2D(1);
        m->SetElasticity(1000., 0.3);
        m->SetBodyForce(0.5, -1.0);
        m->SetPreStress(0.2, -0.1, 0.05, 0.);
        if (phys == 3) m->SetPlaneStress(); else m->SetPlaneStrain();
        if (bcfunc >= 2)  // x-dependent body force: the host-evaluated forcing table of the plane kernel
            m->SetForcingFunction([](const TPZVec<REAL> &x, TPZVec<STATE> &f) { f[0] = 0.5 + x[0] * x[1]; f[1] = -1.0 + 0.3 * x[1]; f[2] = 0.; }, 2);
        cmesh->InsertMaterialObject(m);
        TPZFNMatrix<4, STATE> v1(2, 2, 0.);
        TPZManVector<STATE, 2> v2(2, 0.), v2n(2, 0.);
        v2[0] = dirichlet / 30.;
        v2n[0] = 0.25; v2n[1] = -0.5;
        auto *bcd = m->CreateBC(m, -1, 0, v1, v2);
        if (bcfunc)
            bcd->SetForcingFunctionBC([](const TPZVec<REAL> &x, TPZVec<STATE> &u, TPZFMatrix<STATE> &du) {
                u[0] = 0.01 * x[1] + 0.02 * x[0] * x[1];
                u[1] = -0.03 * x[0];
            });
        cmesh->InsertMaterialObject(bcd);
        cmesh->InsertMaterialObject(m->CreateBC(m, -2, 1, v1, v2n));
    } else if (phys == 0) {
        auto *m = new TPZMatPoisson<STATE>(1, 3);
        const int nlc = g_loadcases;
        if (g_scale != 1.0) m->SetScaleFactor(g_scale);
        m->SetNumLoadCases(nlc);  // (before CreateBC: the boundary conditions copy the count, TPZMatLoadCases.cpp:49-60)
        // x-dependent source: exercises the host-evaluated forcing table; one value per load case (TPZMatPoisson.cpp:23-27)
        m->SetForcingFunction([nlc](const TPZVec<REAL> &x, TPZVec<STATE> &f) {
            for (int l = 0; l < nlc; l++) f[l] = (1.0 + x[0] * x[1] - 0.5 * x[2]) * (1.0 + 0.5 * l) + 0.25 * l * x[2];
        }, 2);
        cmesh->InsertMaterialObject(m);
        TPZFNMatrix<1, STATE> v1(1, 1, 0.);
        TPZManVector<STATE, 1> v2(1, dirichlet), v2n(1, 0.75);
        auto *bcd = m->CreateBC(m, -1, 0, v1, v2);
        if (bcfunc)
            bcd->SetForcingFunctionBC([nlc](const TPZVec<REAL> &x, TPZVec<STATE> &u, TPZFMatrix<STATE> &du) {
                for (int l = 0; l < nlc; l++) u[l] = (0.3 + x[0] * x[1] - 0.5 * x[2] * x[2]) * (1.0 - 0.25 * l);
            });
        cmesh->InsertMaterialObject(bcd);
        TPZFNMatrix<1, STATE> v1n(1, 1, 0.);
        v1n(0, 0) = 2.5e-15;  // type 2 multiplies it by BigNumber (TPZMatPoisson.cpp:112)
        auto *bcn = m->CreateBC(m, -2, g_bctype, v1n, v2n);
        if (nlc > 1) {  // TPZMatLoadCasesBC::SetBCRhsValVec: one value per load case
            TPZVec<TPZVec<STATE>> vals(nlc);
            for (int l = 0; l < nlc; l++) vals[l] = TPZManVector<STATE, 1>(1, 0.75 - 0.5 * l);
            dynamic_cast<TPZMatLoadCasesBC<STATE> *>(bcn)->SetBCRhsValVec(vals);
        }
        cmesh->InsertMaterialObject(bcn);
    } else {
        TPZManVector<STATE, 3> force(3, 0.);
        force[2] = -1.;
        auto *m = new TPZElasticity3D(1, 1000., 0.3, force);
        cmesh->InsertMaterialObject(m);
        TPZFNMatrix<9, STATE> v1(3, 3, 0.);
        TPZManVector<STATE, 3> v2(3, 0.), v2n(3, 0.);
        v2[0] = dirichlet / 30.;
        v2n[0] = 0.25; v2n[1] = -0.5; v2n[2] = 2.0;
        auto *bcd = m->CreateBC(m, -1, 0, v1, v2);
        if (bcfunc)
            bcd->SetForcingFunctionBC([](const TPZVec<REAL> &x, TPZVec<STATE> &u, TPZFMatrix<STATE> &du) {
                u[0] = 0.01 * x[1];
                u[1] = -0.02 * x[0] * x[2];
                u[2] = 0.005 + 0.01 * x[2];
            });
        cmesh->InsertMaterialObject(bcd);
        if (g_bctype == 1) {
            cmesh->InsertMaterialObject(m->CreateBC(m, -2, 1, v1, v2n));
        } else {  // the other types of TPZElasticity3D::ContributeBC (:616-773); 4 = stress field times the face normal
            TPZFNMatrix<9, STATE> v1m(3, 3, 0.);
            v1m(0, 0) = 4.0; v1m(0, 1) = 0.5; v1m(0, 2) = -0.3; v1m(1, 0) = 0.5; v1m(1, 1) = 3.0; v1m(1, 2) = 0.25; v1m(2, 0) = -0.3; v1m(2, 1) = 0.25; v1m(2, 2) = 5.0;
            TPZManVector<STATE, 3> v2m(3, 0.);
            v2m[0] = 0.3; v2m[1] = -0.2; v2m[2] = 0.7;
            cmesh->InsertMaterialObject(m->CreateBC(m, -2, g_bctype, v1m, v2m));
        }
    }
    cmesh->SetAllCreateFunctionsContinuous();
    cmesh->AutoBuild();
    if (prefine) {
        int64_t count = 0;
        const int64_t nel0 = cmesh->NElements();
        for (int64_t iel = 0; iel < nel0; iel++) {
            TPZCompEl *cel = cmesh->Element(iel);
            if (!cel || !cel->Reference() || cel->Reference()->Dimension() != dim) continue;
            if (count++ % 3) continue;
            if (auto *intel = dynamic_cast<TPZInterpolatedElement *>(cel)) intel->PRefine(p + 1);
        }
        cmesh->ExpandSolution();
    }
    cmesh->AdjustBoundaryElements();
    cmesh->CleanUpUnconnectedNodes();
    return cmesh;
}

struct Csr {
    std::vector<int64_t> ia, ja;
    std::vector<double> a, rhs, sol;
    std::vector<double> a_twice;       // B200_ACCUMULATE: values after a second Assemble() into the NON-zeroed matrix
    int devices = 1;
    std::vector<double> residual_rhs;  // TPZLinearAnalysis::AssembleResidual() -> Assemble(rhs)
    std::vector<double> sol_device;    // B200 only: CG on the device-resident matrix (TPZB200CGSolver through an.Solve())
    int64_t device_cg_iters = 0;
    double prepare_ms = 0, pattern_ms = 0, device_ms = 0;  // B200 only: TPZStructMatrixB200::LastTimings of the second Assemble()
};

static int g_pin = 0;     // 1: TPZStructMatrixB200::SetPinHostMatrix(true) — the matrix values are page-locked for the download
static int g_filter = 0;  // 1: an equation filter keeps the upper three quarters of the equations (TPZEquationFilter::SetMinMaxEq)

template <class TStrMat>
static void Run(TPZCompMesh *cmesh, int nthreads, bool symmetric, bool solve, Csr &out, double &t_first, double &t_second) {
    TPZLinearAnalysis an(cmesh, false);
    TStrMat strmat(cmesh);
    strmat.SetNumThreads(nthreads);
    if (auto *b = dynamic_cast<TPZStructMatrixB200<STATE> *>(&strmat)) {
        if (g_gpus) b->SetNumThreads(g_gpus);
        if (g_droptiny) b->SetDropTinyEntries(true);
    }
    if (g_filter) strmat.EquationFilter().SetMinMaxEq(cmesh->NEquations() / 4, cmesh->NEquations());
    if (g_pin)
        if (auto *b = dynamic_cast<TPZStructMatrixB200<STATE> *>(&strmat)) b->SetPinHostMatrix(true);
    an.SetStructuralMatrix(strmat);
    TPZStepSolver<STATE> step;
    step.SetDirect(symmetric ? ELDLt : ELU);  // never decomposed: the CG below is run on the assembled matrix
    an.SetSolver(step);
    auto t0 = clk::now();
    if (g_loadcases > 1) {
        // TPZAnalysis::ComputeNumberofLoadCases (Analysis/TPZAnalysis.cpp:307-335) never finds more than one load case (its
        // everyMatHasLoadCase flag starts false and is never set), so TPZLinearAnalysis::Assemble sizes fRhs with one column and
        // the reference's own AddFel then aborts.  Several load cases are driven the way the struct matrix supports them: the
        // caller sizes the rhs (StrMatrix/TPZStrMatParInterface.cpp:13-14 keeps its columns).
        an.Rhs().Redim(cmesh->NEquations(), g_loadcases);
        auto *mat = an.StructMatrix()->CreateAssemble(an.Rhs(), nullptr);
        an.MatrixSolver<STATE>().SetMatrix(mat);
    } else {
        an.Assemble();
    }
    auto t1 = clk::now();
    if (g_loadcases > 1) {
        an.MatrixSolver<STATE>().Matrix()->Zero();
        an.Rhs().Zero();
        an.StructMatrix()->Assemble(*an.MatrixSolver<STATE>().Matrix().operator->(), an.Rhs(), nullptr);
    } else {
        an.Assemble();
    }
    auto t2 = clk::now();
    t_first = secs(t0, t1);
    t_second = secs(t1, t2);
    if (auto *b = dynamic_cast<TPZStructMatrixB200<STATE> *>(an.StructMatrix().operator->())) b->LastTimings(out.prepare_ms, out.pattern_ms, out.device_ms);
    auto mtx = an.MatrixSolver<STATE>().Matrix();
    const int64_t neq = cmesh->NEquations();
    if (symmetric) {
        auto *sp = dynamic_cast<TPZSYsmpMatrix<STATE> *>(mtx.operator->());
        out.ia.assign(sp->IA().begin(), sp->IA().end());
        out.ja.assign(sp->JA().begin(), sp->JA().end());
        out.a.assign(sp->A().begin(), sp->A().end());
    } else {
        auto *sp = dynamic_cast<TPZFYsmpMatrix<STATE> *>(mtx.operator->());
        TPZVec<int64_t> ia, ja;
        TPZVec<STATE> a;
        sp->GetData(ia, ja, a);
        out.ia.assign(ia.begin(), ia.end());
        out.ja.assign(ja.begin(), ja.end());
        out.a.assign(a.begin(), a.end());
    }
    TPZFMatrix<STATE> &rhsm = an.Rhs();
    const int ncols = (int)rhsm.Cols();  // one column per load case (Analysis/TPZLinearAnalysis.cpp:70-72)
    out.rhs.resize(neq * ncols);
    for (int l = 0; l < ncols; l++)
        for (int64_t i = 0; i < neq; i++) out.rhs[l * neq + i] = rhsm(i, l);
    auto *b200 = dynamic_cast<TPZStructMatrixB200<STATE> *>(an.StructMatrix().operator->());
    if (b200) out.devices = b200->NumDevicesUsed();
    if (g_accumulate && symmetric) {
        // TPZStrMatParInterface::Assemble straight into the matrix that already holds one assembly: AddKel accumulates
        if (b200) b200->SetAccumulate(true);
        TPZFMatrix<STATE> rhs2(neq, ncols, 0.);
        an.StructMatrix()->Assemble(*mtx.operator->(), rhs2, nullptr);
        auto *sp = dynamic_cast<TPZSYsmpMatrix<STATE> *>(mtx.operator->());
        out.a_twice.assign(sp->A().begin(), sp->A().end());
        if (b200) b200->SetAccumulate(false);
        an.Assemble();  // back to one assembly for the checks below
    }
    if (solve) {
        // the reference's own CG (Solvers/LinearSolvers/cg.h:44-120 through TPZMatrix::SolveCG) with its Jacobi
        // preconditioner, on the assembled system
        TPZStepSolver<STATE> pre(mtx);
        pre.SetJacobi(1, 0., 0);
        TPZStepSolver<STATE> cg(mtx);
        cg.SetCG(50000, pre, 1.e-15, 0);
        TPZFMatrix<STATE> sol(neq, 1, 0.), f(rhsm);
        cg.Solve(f, sol);
        out.sol.resize(neq);
        for (int64_t i = 0; i < neq; i++) out.sol[i] = sol(i, 0);
        // the same solve on the GPU, through the unmodified TPZLinearAnalysis::Solve(): the matrix the strategy left on
        // the device, the reference's CG algorithm with its Jacobi(1) preconditioner
        if (b200) {  // (several GPUs: the row-sharded CG, halos over NVLink)
            TPZB200CGSolver<STATE> dev(b200, 50000, 1.e-15, true, 0);
            dev.SetMatrix(mtx);
            an.SetSolver(dev);
            an.Solve();
            TPZFMatrix<STATE> &u = an.Solution();
            out.sol_device.resize(neq);
            for (int64_t i = 0; i < neq; i++) out.sol_device[i] = u(i, 0);
            auto *used = dynamic_cast<TPZB200CGSolver<STATE> *>(an.Solver());
            out.device_cg_iters = used ? used->NumIterations() : -1;
        }
    }
    // right-hand side only: TPZLinearAnalysis::AssembleResidual() -> TPZStrMatParInterface::Assemble(rhs)
    if (g_loadcases > 1) {
        an.Rhs().Redim(neq, g_loadcases);
        an.StructMatrix()->Assemble(an.Rhs(), nullptr);
    } else {
        an.AssembleResidual();
    }
    {
        TPZFMatrix<STATE> &res = an.Rhs();
        const int rc = (int)res.Cols();
        out.residual_rhs.resize(neq * rc);
        for (int l = 0; l < rc; l++)
            for (int64_t i = 0; i < neq; i++) out.residual_rhs[l * neq + i] = res(i, l);
    }
    // the value array belongs to the analysis' matrix: release the page lock before the analysis goes away
    if (auto *b = dynamic_cast<TPZStructMatrixB200<STATE> *>(an.StructMatrix().operator->())) b->UnpinHostMatrix();
}

static double RelF(const std::vector<double> &x, const std::vector<double> &ref) {
    long double num = 0, den = 0;
    for (size_t i = 0; i < x.size(); i++) {
        num += (long double)(x[i] - ref[i]) * (x[i] - ref[i]);
        den += (long double)ref[i] * ref[i];
    }
    return den > 0 ? (double)std::sqrt(num / den) : (double)std::sqrt(num);
}

int main(int argc, char **argv) {
    if (argc < 6) {
        std::cerr << "usage: dropin_test n p phys(0|1|2|3) tet(0|1|2|3) symmetric(0|1) [solve(0|1)] [cpu_threads] [device_create] [equation_filter] [pin_host] [prefine] [bcfunc]\n";
        return 2;
    }
    const int n = atoi(argv[1]), p = atoi(argv[2]), phys = atoi(argv[3]), tet = atoi(argv[4]), symmetric = atoi(argv[5]);
    const int solve = argc > 6 ? atoi(argv[6]) : 1;
    const int threads = argc > 7 ? atoi(argv[7]) : (int)std::thread::hardware_concurrency();
    // 1: TPZSSpStructMatrixB200 / TPZSpStructMatrixB200, whose Create() builds the CSR pattern on the GPU
    const int device_create = argc > 8 ? atoi(argv[8]) : 0;
    g_filter = argc > 9 ? atoi(argv[9]) : 0;
    g_pin = argc > 10 ? atoi(argv[10]) : 0;
    const int prefine = argc > 11 ? atoi(argv[11]) : 0;
    const int bcfunc = argc > 12 ? atoi(argv[12]) : 0;
    g_gpus = EnvInt("B200_GPUS", 0);
    g_bctype = EnvInt("B200_BCTYPE", 1);
    g_loadcases = EnvInt("B200_LOADCASES", 1);
    g_droptiny = EnvInt("B200_DROPTINY", 0);
    g_scale = EnvDouble("B200_SCALE", 1.0);
    g_skip_serial = EnvInt("B200_SKIP_SERIAL", 0);
    g_accumulate = EnvInt("B200_ACCUMULATE", 0);
    TPZCompMesh *cmesh = BuildMesh(n, p, phys, tet, 0.12, solve ? 0.0 : 0.3, prefine, bcfunc);
    Csr ref, refmt, gpu;
    double t1 = 0, t2 = 0, tm1, tm2, g1, g2;
    if (g_skip_serial) {  // timing runs on large meshes: the threaded OR assembly (bit-identical to the serial one) is the reference
        if (symmetric) {
            Run<TPZSSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>>(cmesh, threads, true, solve, ref, tm1, tm2);
            if (device_create) Run<TPZSSpStructMatrixB200<STATE>>(cmesh, 0, true, solve, gpu, g1, g2);
            else Run<TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>>(cmesh, 0, true, solve, gpu, g1, g2);
        } else {
            Run<TPZSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>>(cmesh, threads, false, false, ref, tm1, tm2);
            if (device_create) Run<TPZSpStructMatrixB200<STATE>>(cmesh, 0, false, false, gpu, g1, g2);
            else Run<TPZSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>>(cmesh, 0, false, false, gpu, g1, g2);
        }
        refmt.a = ref.a;
    } else if (symmetric) {
        Run<TPZSSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>>(cmesh, 0, true, solve, ref, t1, t2);
        Run<TPZSSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>>(cmesh, threads, true, false, refmt, tm1, tm2);
        if (device_create) Run<TPZSSpStructMatrixB200<STATE>>(cmesh, 0, true, solve, gpu, g1, g2);
        else Run<TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>>(cmesh, 0, true, solve, gpu, g1, g2);
    } else {
        Run<TPZSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>>(cmesh, 0, false, false, ref, t1, t2);
        Run<TPZSpStructMatrix<STATE, TPZStructMatrixOR<STATE>>>(cmesh, threads, false, false, refmt, tm1, tm2);
        if (device_create) Run<TPZSpStructMatrixB200<STATE>>(cmesh, 0, false, false, gpu, g1, g2);
        else Run<TPZSpStructMatrix<STATE, TPZStructMatrixB200<STATE>>>(cmesh, 0, false, false, gpu, g1, g2);
    }
    const bool same_ia = ref.ia.size() == gpu.ia.size() && !memcmp(ref.ia.data(), gpu.ia.data(), ref.ia.size() * 8);
    const bool same_ja = ref.ja.size() == gpu.ja.size() && !memcmp(ref.ja.data(), gpu.ja.data(), ref.ja.size() * 8);
    const double errA = RelF(gpu.a, ref.a), errR = RelF(gpu.rhs, ref.rhs);
    // rows without penalty entries
    const int64_t neq = (int64_t)ref.ia.size() - 1;
    std::vector<char> bigrow(neq, 0);
    for (int64_t r = 0; r < neq; r++)
        for (int64_t k = ref.ia[r]; k < ref.ia[r + 1]; k++)
            if (std::fabs(ref.a[k]) > 1e9) bigrow[r] = 1;
    long double num = 0, den = 0;
    double maxrel = 0;
    for (int64_t r = 0; r < neq; r++) {
        if (bigrow[r]) continue;
        double rowmax = 0;
        for (int64_t k = ref.ia[r]; k < ref.ia[r + 1]; k++) rowmax = std::max(rowmax, std::fabs(ref.a[k]));
        for (int64_t k = ref.ia[r]; k < ref.ia[r + 1]; k++) {
            const double d = gpu.a[k] - ref.a[k];
            num += (long double)d * d;
            den += (long double)ref.a[k] * ref.a[k];
            if (rowmax > 0) maxrel = std::max(maxrel, std::fabs(d) / rowmax);
        }
    }
    const double errInt = den > 0 ? (double)std::sqrt(num / den) : 0.0;
    const double errMT = RelF(refmt.a, ref.a);
    double errSol = 0;
    if (solve && symmetric) errSol = RelF(gpu.sol, ref.sol);
    double errSolDev = 0;
    if (solve && symmetric && !gpu.sol_device.empty()) errSolDev = RelF(gpu.sol_device, ref.sol);
    const double errTwice = (g_accumulate && symmetric) ? RelF(gpu.a_twice, ref.a_twice) : 0.0;
    const double errRes = RelF(gpu.residual_rhs, ref.residual_rhs), errResVsRhs = RelF(ref.residual_rhs, ref.rhs);
    int64_t nvol = 0;  // elements of the mesh dimension
    for (int64_t iel = 0; iel < cmesh->NElements(); iel++) {
        TPZCompEl *cel = cmesh->Element(iel);
        if (cel && cel->Reference() && cel->Reference()->Dimension() == cmesh->Dimension()) nvol++;
    }
    const bool ok = same_ia && same_ja && errA <= 1e-12 && errR <= 1e-12 && errInt <= 1e-12 && errSol <= 1e-10 && errSolDev <= 1e-10 && errRes <= 1e-12 && errTwice <= 1e-12 &&
                    (g_gpus < 2 || gpu.devices >= 2);
    std::cout.precision(6);
    std::cout << "{\"n\": " << n << ", \"p\": " << p << ", \"phys\": " << phys << ", \"tet\": " << tet << ", \"symmetric\": " << symmetric << ", \"device_create\": " << device_create << ", \"equation_filter\": " << g_filter << ", \"pin_host\": " << g_pin << ", \"prefine\": " << prefine << ", \"bcfunc\": " << bcfunc
              << ", \"gpus\": " << gpu.devices << ", \"bctype\": " << g_bctype << ", \"loadcases\": " << g_loadcases << ", \"droptiny\": " << g_droptiny
              << ", \"scale\": " << g_scale << ", \"relF_A_accumulated_twice\": " << errTwice << ", \"cpu_first_assemble_s\": " << t1
              << ", \"neq\": " << neq << ", \"nnz\": " << ref.ja.size() << ", \"vol_elements\": " << nvol
              << ", \"ia_identical\": " << same_ia << ", \"ja_identical\": " << same_ja << ", \"relF_A\": " << errA
              << ", \"relF_A_nonpenalty_rows\": " << errInt << ", \"max_entry_err_over_rowmax\": " << maxrel
              << ", \"relF_rhs\": " << errR << ", \"relF_cg_solution\": " << errSol << ", \"relF_device_cg_solution\": " << errSolDev
              << ", \"device_cg_iterations\": " << gpu.device_cg_iters << ", \"relF_residual_rhs\": " << errRes
              << ", \"ref_residual_rhs_vs_rhs\": " << errResVsRhs << ", \"relF_A_ref_threads_vs_serial\": " << errMT
              << ", \"cpu_serial_assemble_s\": " << t2 << ", \"cpu_threads\": " << threads << ", \"cpu_threaded_assemble_s\": " << tm2
              << ", \"gpu_first_assemble_s\": " << g1 << ", \"gpu_second_assemble_s\": " << g2
              << ", \"gpu_second_prepare_ms\": " << gpu.prepare_ms << ", \"gpu_second_pattern_ms\": " << gpu.pattern_ms
              << ", \"gpu_second_assemble_call_ms\": " << gpu.device_ms << ", \"ok\": " << ok << "}" << std::endl;
    return ok ? 0 : 1;
}
