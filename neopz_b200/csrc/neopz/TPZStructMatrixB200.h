/**
 * @file TPZStructMatrixB200.h
 * @brief Parallel layer for NeoPZ struct matrices that assembles on an NVIDIA B200 (sm_100a).
 *
 * A fourth TPZStrMatParInterface strategy next to TPZStructMatrixOR / OT / TBBFlow
 * (StrMatrix/TPZStrMatParInterface.h:34-92, pattern of StrMatrix/pzstrmatrixor.h:29-119).  Use it as
 *
 *     TPZSSpStructMatrix<STATE, TPZStructMatrixB200<STATE>> strmat(cmesh);   // or TPZSpStructMatrix
 *     an.SetStructuralMatrix(strmat);
 *     an.Assemble();                                                         // unchanged user code
 *
 * Assemble(stiffness, rhs) flattens the TPZCompMesh once (node coordinates, corner nodes, destination
 * indices exactly as Mesh/pzelmat.cpp:37-70, integration rules read through TPZIntPoints::Point, shape
 * tables through TPZShapeH1<TSHAPE>::Shape, material constants, forcing functions evaluated on the
 * host at the integration points) and drives the CUDA engine through the C ABI of include/b200asm.h.
 * The CSR pattern is the one the struct matrix's own Create() produced; values are written in place
 * into TPZSYsmpMatrix::A() / TPZFYsmpMatrix storage and the rhs into the TPZFMatrix.
 *
 * Supported (everything else is reported on PZError followed by DebugStop(), the reference's error convention,
 * pzstrmatrixor.cpp:107-114; there is no CPU fallback):
 *  - H1 TPZCompElH1 elements on hexahedra, tetrahedra, prisms, pyramids (volume) and quadrilaterals, triangles, lines (boundary
 *    faces, or the domain and boundary of a plane problem), any order up to 9 and sides of different order (p-refined meshes);
 *    specialised kernels for uniform p <= 4 on hexahedra / tetrahedra and p <= 2 on prisms / pyramids, the generic runtime-size
 *    kernel for the rest; (multi)linear geometric maps, no hanging nodes / condensed connects;
 *  - TPZMatPoisson<STATE> (dim 2, 3), TPZElasticity3D, TPZElasticity2D (plane strain / stress, prestress), with constant data or
 *    std::function forcing (tabulated by the host at every integration point, re-evaluated at every Assemble unless
 *    SetForcingIsStatic(true));
 *  - TPZBndCondT<STATE>: TPZMatPoisson types 0, 1, 2; TPZElasticity3D types 0-8; TPZElasticity2D types 0, 1, 3; constant data or
 *    TPZBndCondT::SetForcingFunctionBC;
 *  - several load cases of TPZMatPoisson (TPZMatLoadCases: the rhs has NumLoadCases() columns);
 *  - symmetric (TPZSYsmpMatrix) and full (TPZFYsmpMatrix) storage, material-id subsets (SetMaterialIds), an active
 *    TPZEquationFilter;
 *  - one GPU (SetDevice) or several (SetNumThreads(n >= 2) = the first n CUDA devices, or SetDevices): elements partitioned by
 *    their smallest destination equation, every GPU owns a row block of the CSR, interface rows travel over NVLink.
 *
 * Semantics that differ from TPZStructMatrixOR on purpose:
 *  - Assemble(stiffness, rhs) OVERWRITES the value array of `stiffness` with the assembled matrix (the reference's AddKel
 *    accumulates into whatever the matrix holds; TPZLinearAnalysis::Assemble zeroes it first, so the two agree there).  A
 *    caller that assembles several passes into one matrix turns SetAccumulate(true) on: the previous values are then added
 *    on the host after the download.  The rhs is always ADDED to, like TPZFMatrix::AddFel does.
 *  - TPZSYsmpMatrix::AddKel drops element entries with |value| < 1e-12 (Matrix/pzsysmp.cpp:381, IsZero); the GPU adds
 *    them unless SetDropTinyEntries(true) asks for the reference's behaviour.
 *  - the flattened mesh is cached between calls and revalidated with a signature of the mesh (element list, connect
 *    sequence numbers / orders / block positions, materials, material-id filter, equation filter) at every Assemble;
 *    SetCheckMesh(false) skips that walk (then call Invalidate() after changing the mesh).
 */
#ifndef TPZSTRUCTMATRIXB200_H
#define TPZSTRUCTMATRIXB200_H

#include <memory>
#include <vector>

#include "TPZMatrixSolver.h"
#include "TPZStrMatParInterface.h"
#include "pzfmatrix.h"
#include "pzreal.h"
#include "pzstack.h"
#include "pzvec.h"

class TPZBaseMatrix;
class TPZStructMatrix;
class TPZCompMesh;

struct TPZB200AssemblyCache;  // flattened mesh + device context, shared by copies (Clone())

template <class TVar>
class TPZStructMatrixB200 : public virtual TPZStrMatParInterface {
public:
    TPZStructMatrixB200();
    TPZStructMatrixB200(const TPZStructMatrixB200 &copy);
    TPZStructMatrixB200 &operator=(const TPZStructMatrixB200 &copy);
    virtual ~TPZStructMatrixB200();

    //! Assemble the global system of equations into a matrix that has already been created.
    void Assemble(TPZBaseMatrix &stiffness, TPZBaseMatrix &rhs, TPZAutoPointer<TPZGuiInterface> guiInterface) override;
    //! Assemble the global right hand side vector.
    void Assemble(TPZBaseMatrix &rhs, TPZAutoPointer<TPZGuiInterface> guiInterface) override;

    //! Conjugate gradients on the device-resident matrix of the last Assemble(stiffness, rhs) — the reference's CG
    //! (Solvers/LinearSolvers/cg.h:44-120) with TPZStepSolver::SetJacobi(1,0.,0) (jacobi) or no preconditioner.
    //! numiterations / tol: in = limits, out = iterations done / relative residual reached (like TPZMatrix::SolveCG).
    void SolveCG(const TPZFMatrix<TVar> &F, TPZFMatrix<TVar> &result, int64_t &numiterations, REAL &tol, bool jacobi = true,
                 int fromcurrent = 0);

    //! CUDA device used by this strategy (default 0) when it runs on one GPU.
    void SetDevice(int device) { fDevice = device; fDevices.clear(); }
    int Device() const { return fDevice; }
    //! Several GPUs of this process share the assembly (b200asm_multi_*): the GPU counterpart of SetNumThreads
    //! (StrMatrix/TPZStrMatParInterface.h:62-69).  Without an explicit list, SetNumThreads(n >= 2) selects the first
    //! min(n, devices present) CUDA devices; n = 0 / 1 (the default) keeps the single device of SetDevice.
    void SetDevices(const std::vector<int> &devices) { fDevices = devices; }
    const std::vector<int> &Devices() const { return fDevices; }
    //! GPUs the last Assemble() ran on.
    int NumDevicesUsed() const;
    //! Add the assembled matrix to the values `stiffness` already holds (TPZSYsmpMatrix::AddKel semantics) instead of overwriting.
    void SetAccumulate(bool accumulate) { fAccumulate = accumulate; }
    //! Drop element entries with |value| < 1e-12 like TPZSYsmpMatrix::AddKel does (Matrix/pzsysmp.cpp:381).  Symmetric storage only.
    void SetDropTinyEntries(bool drop) { fDropTiny = drop; }
    //! Forcing functions / boundary functions do not change between assemblies: tabulate them once per flatten only.
    void SetForcingIsStatic(bool is_static) { fStaticForcing = is_static; }
    //! Host threads for the per-Assemble host work (tabulating forcing / boundary functions at the integration points, the mesh
    //! signature walk).  Default 0 = min(16, hardware threads).  The callbacks are then called concurrently, exactly as
    //! TPZStructMatrixOR does with SetNumThreads(n > 0) (StrMatrix/pzstrmatrixor.cpp:476-513); SetHostThreads(1) for callbacks
    //! that are not thread-safe.
    void SetHostThreads(int n) { fHostThreads = n; }
    //! Revalidate the cached flattened mesh at every Assemble (default true).
    void SetCheckMesh(bool check) { fCheckMesh = check; }
    //! Forget the cached flattened mesh (the next Assemble flattens again).
    void Invalidate();
    //! Page-lock the value array of the matrix handed to Assemble() (cudaHostRegister, once per array): the download then
    //! runs at PCIe speed and overlaps the kernels.  Off by default because the array belongs to the caller's matrix:
    //! call UnpinHostMatrix() (or destroy / reassign this object) BEFORE that matrix is destroyed.
    void SetPinHostMatrix(bool pin) { fPinHost = pin; }
    void UnpinHostMatrix();
    //! Milliseconds the last Assemble() spent in flatten / pattern upload / device assembly + copies.
    void LastTimings(double &flatten_ms, double &pattern_ms, double &assemble_ms) const;

    int ClassId() const override;
    void Read(TPZStream &buf, void *context) override;
    void Write(TPZStream &buf, int withclassid) const override;

protected:
    //! Pattern of the struct matrix built on the device from the element graph; fills ia / ja (host copies, reference
    //! layout) and leaves the pattern resident so that the next Assemble() does not upload it again.
    void CreatePatternOnDevice(bool symmetric, TPZStack<int64_t> &elgraph, TPZVec<int64_t> &elgraphindex, TPZVec<int64_t> &ia,
                               TPZVec<int64_t> &ja);
    template <class T>
    friend class TPZSSpStructMatrixB200;
    template <class T>
    friend class TPZSpStructMatrixB200;
    int fDevice{0};
    std::vector<int> fDevices;
    bool fPinHost{false};
    bool fAccumulate{false};
    bool fDropTiny{false};
    bool fStaticForcing{false};
    bool fCheckMesh{true};
    int fHostThreads{0};
    std::shared_ptr<TPZB200AssemblyCache> fCache;
};

extern template class TPZStructMatrixB200<STATE>;

#include "TPZSSpStructMatrix.h"
#include "TPZSpStructMatrix.h"

/**
 * @brief TPZSSpStructMatrix / TPZSpStructMatrix with the B200 strategy whose Create() builds the CSR pattern ON THE GPU
 * (b200asm_build_pattern_device) instead of the serial std::set walk of TPZSSpStructMatrix::SetupMatrixData
 * (StrMatrix/TPZSSpStructMatrix.cpp:50-193; 1.7 s for 36 k equations, more than the assembly itself).  The pattern is
 * bit-identical; it stays on the device for the assembly and is copied once into the TPZSYsmpMatrix / TPZFYsmpMatrix
 * the caller receives.  Everything else is inherited: use it exactly like TPZSSpStructMatrix<STATE>.
 */
template <class TVar = STATE>
class TPZSSpStructMatrixB200 : public TPZSSpStructMatrix<TVar, TPZStructMatrixB200<TVar>> {
public:
    using TPZSSpStructMatrix<TVar, TPZStructMatrixB200<TVar>>::TPZSSpStructMatrix;
    TPZStructMatrix *Clone() override { return new TPZSSpStructMatrixB200<TVar>(*this); }

protected:
    TPZMatrix<TVar> *SetupMatrixData(TPZStack<int64_t> &elgraph, TPZVec<int64_t> &elgraphindex) override;
};
template <class TVar = STATE>
class TPZSpStructMatrixB200 : public TPZSpStructMatrix<TVar, TPZStructMatrixB200<TVar>> {
public:
    using TPZSpStructMatrix<TVar, TPZStructMatrixB200<TVar>>::TPZSpStructMatrix;
    TPZStructMatrix *Clone() override { return new TPZSpStructMatrixB200<TVar>(*this); }

protected:
    TPZMatrix<TVar> *SetupMatrixData(TPZStack<int64_t> &elgraph, TPZVec<int64_t> &elgraphindex) override;
};
extern template class TPZSSpStructMatrixB200<STATE>;
extern template class TPZSpStructMatrixB200<STATE>;

/**
 * @brief TPZMatrixSolver that solves with the matrix the B200 strategy left on the device (no factorisation, no copy of
 * the matrix): `TPZB200CGSolver<STATE> solver(strategy); an.SetSolver(solver); an.Solve();` — the drop-in for
 * `TPZStepSolver::SetCG(numiter, Jacobi, tol, fromcurrent)`.
 */
template <class TVar>
class TPZB200CGSolver : public TPZMatrixSolver<TVar> {
public:
    TPZB200CGSolver(TPZStructMatrixB200<TVar> *strategy, int64_t maxiter = 50000, REAL tol = 1.e-15, bool jacobi = true, int fromcurrent = 0)
        : fStrategy(strategy), fMaxIter(maxiter), fTol(tol), fJacobi(jacobi), fFromCurrent(fromcurrent) {}
    void Solve(const TPZFMatrix<TVar> &F, TPZFMatrix<TVar> &result, TPZFMatrix<TVar> *residual = 0) override {
        fNumIterations = fMaxIter;
        fResidual = fTol;
        fStrategy->SolveCG(F, result, fNumIterations, fResidual, fJacobi, fFromCurrent);
    }
    TPZSolver *Clone() const override { return new TPZB200CGSolver<TVar>(*this); }
    int64_t NumIterations() const { return fNumIterations; }
    REAL Residual() const { return fResidual; }

private:
    TPZStructMatrixB200<TVar> *fStrategy;
    int64_t fMaxIter, fNumIterations{0};
    REAL fTol, fResidual{0};
    bool fJacobi;
    int fFromCurrent;
};
#endif
