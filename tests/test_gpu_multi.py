"""Several GPUs behind the C ABI in ONE process (b200asm_multi_*, what TPZStructMatrixB200::SetNumThreads(n) drives): element
partition by the smallest destination equation, row-sharded CSR, staged interface rows pushed over NVLink while the interior is
assembled.  The caller sees the same global IA / JA / A / rhs as with one context; checked against the oracle on the undivided
mesh (1e-12) and against the one-GPU result, on the grid numbering (banded: every GPU pushes to its neighbour) and on a random
renumbering of the connects (every GPU pushes to every GPU above it).  Skipped on a one-GPU box."""
import numpy as np
import pytest
import torch

from neopz_b200 import gridmesh, strmatrix as sm
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, materials_for, relF

pytestmark = pytest.mark.gpu


def renumbered(mesh, seed):
    """The same mesh with the sequence numbers of its connects shuffled (what TPZCompMesh::Permute / a renumbering does)."""
    rng = np.random.default_rng(seed)
    ncon = len(mesh.block_size)
    perm = rng.permutation(ncon)           # old sequence number -> new
    size = np.empty_like(mesh.block_size)
    size[perm] = mesh.block_size
    pos = np.concatenate([[0], np.cumsum(size)[:-1]]).astype(np.int64)
    out = gridmesh.FlatMesh(porder=mesh.porder, nstate=mesh.nstate, nodes=mesh.nodes, block_pos=pos, block_size=size, neq=mesh.neq)
    for b in mesh.blocks:
        conn = perm[b.connects]
        out.blocks.append(gridmesh.ElementBlock(topology=b.topology, matid=b.matid, first=b.first, elnodes=b.elnodes, connects=conn,
                                                dest=gridmesh.destination_indices(b.topology, conn, pos, mesh.porder, mesh.nstate)))
    return out


@pytest.mark.parametrize("ndev", [2, 4])
@pytest.mark.parametrize("n,p,phys,tet,symmetric,shuffle", [(6, 2, 0, False, True, 0), (5, 2, 1, False, True, 0), (5, 2, 0, True, False, 0),
                                                            (6, 2, 0, False, True, 7), (4, 2, 1, True, True, 3), (4, 3, 0, False, False, 5),
                                                            (12, 2, 0, False, True, 0), (10, 1, 1, False, False, 11)])
def test_multi_device_against_oracle(ndev, n, p, phys, tet, symmetric, shuffle):
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=tet, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.1)
    if shuffle:
        mesh = renumbered(mesh, shuffle)
    mats = materials_for(phys, neumann=True)
    one = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    ia, ja, a1, rhs1 = one.CreateAssemble()
    one.ctx.close()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
    multi = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, devices=list(range(ndev)))
    multi.ctx.set_option("overlap_min_elements", 384)
    multi.ctx.set_option("exchange_timeout_ms", 20000)
    ia2, ja2 = multi.Create()
    assert np.array_equal(ia, ia2) and np.array_equal(ja, ja2)
    rb, nel, staged = multi.ctx.partition()
    assert rb[0] == 0 and rb[-1] == mesh.neq and np.all(np.diff(rb) >= 0) and nel.sum() == mesh.nelements
    assert nel.min() > 0, "every GPU gets elements"
    for _ in range(3):  # re-assembly reproduces (the step counters advance)
        a, rhs = multi.Assemble()
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        assert relF(a, a1) <= 1e-14 and relF(rhs, rhs1) <= 1e-14
    assert relF(multi.AssembleRhs(), rhs_ref) <= TOL
    # page-locked global arrays: every GPU downloads its slice while its kernels run
    a_pin = torch.empty(len(ja), dtype=torch.float64).pin_memory()
    r_pin = torch.empty(mesh.neq, dtype=torch.float64).pin_memory()
    multi.Assemble(a_pin.numpy(), r_pin.numpy())
    assert relF(a_pin.numpy(), a_ref) <= TOL and relF(r_pin.numpy(), rhs_ref) <= TOL
    multi.ctx.close()


@pytest.mark.parametrize("ndev", [2, 4])
@pytest.mark.parametrize("n,p,phys,tet,symmetric,shuffle", [(6, 2, 0, False, True, 0), (4, 2, 1, False, True, 0), (5, 2, 0, True, True, 3),
                                                            (6, 1, 0, False, False, 0), (5, 2, 0, False, False, 9)])
def test_sharded_cg_against_one_gpu_and_scipy(ndev, n, p, phys, tet, symmetric, shuffle):
    """N2 across GPUs (b200asm_multi_cg_solve): the reference's CG (Solvers/LinearSolvers/cg.h:44-120, Jacobi(1) preconditioner)
    on the row-sharded resident matrix - row-block products, halos of p and q over NVLink - against the one-GPU device CG and a
    direct solve of the assembled system (1e-10, the north star's bound for the CG solution)."""
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=tet, perturb=0.1)   # Dirichlet on all faces: definite
    if shuffle:
        mesh = renumbered(mesh, shuffle)
    mats = materials_for(phys)
    one = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    ia, ja, a1, rhs1 = one.CreateAssemble()
    x1, it1, res1 = one.SolveCG(max_iter=20000, tol=1e-14)
    one.ctx.close()
    multi = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, devices=list(range(ndev)))
    multi.Create()
    a, rhs = multi.Assemble()
    x, it, res = multi.SolveCG(max_iter=20000, tol=1e-14)
    assert res <= 1e-14 and it > 0
    U = sp.csr_matrix((a1, ja, ia), shape=(mesh.neq, mesh.neq))
    A = (U + sp.triu(U, 1).T) if symmetric else U
    xd = spla.spsolve(sp.csc_matrix(A), rhs1)
    assert relF(x, xd) <= 1e-10 and relF(x1, xd) <= 1e-10
    assert relF(x, x1) <= 1e-10
    # warm start from the solution: nothing left to do (the TRUE residual of a solution whose recurrence residual reached
    # 1e-14 sits around cond(A) * eps, so the restart is asked for 1e-10); a second right-hand side through the same matrix
    x2, it2, res2 = multi.SolveCG(max_iter=20000, tol=1e-10, x0=x)
    assert it2 <= 2 and res2 <= 1e-10
    f2 = np.linspace(0.5, 1.5, mesh.neq)
    x3, _it3, _res3 = multi.SolveCG(max_iter=20000, tol=1e-14, f=f2)
    assert relF(x3, spla.spsolve(sp.csc_matrix(A), f2)) <= 1e-10
    multi.ctx.close()
