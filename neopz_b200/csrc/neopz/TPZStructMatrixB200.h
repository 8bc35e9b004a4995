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
 * Supported: H1 TPZCompElH1 elements of uniform order p in {1,2} on hexahedra / tetrahedra with
 * TPZMatPoisson<STATE> or TPZElasticity3D, boundary faces with TPZBndCondT<STATE> (types 0,1; 2 for
 * elasticity), no hanging nodes, inactive equation filter, one load case.  Anything else is reported on
 * PZError followed by DebugStop(), the reference's error convention (pzstrmatrixor.cpp:107-114).
 * There is no CPU fallback.
 */
#ifndef TPZSTRUCTMATRIXB200_H
#define TPZSTRUCTMATRIXB200_H

#include <memory>

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

    //! CUDA device used by this strategy (default 0).
    void SetDevice(int device) { fDevice = device; }
    int Device() const { return fDevice; }
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
    bool fPinHost{false};
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
