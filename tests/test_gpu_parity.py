"""GPU parity tests proper: the CUDA path (through the C ABI) against
  (1) the committed fixtures produced by the unmodified reference (pattern bit-exact, values 1e-12),
  (2) the oracle on the same seeded/perturbed meshes at sizes it finishes in seconds,
  (3) size-independent properties at larger sizes (rigid-body modes, constants in the kernel of K,
      symmetric-vs-full storage consistency, reassembly idempotence).
Tolerance (BASELINE.json north_star): pattern bit-exact; ||A-Aref||_F/||Aref||_F <= 1e-12, rhs same."""
import numpy as np
import pytest

from neopz_b200 import capi, gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble

pytestmark = pytest.mark.gpu
TOL = 1e-12


def materials_for(phys, neumann=False):
    if phys == 0:
        m = sm.TPZMatPoisson(1, 3)
        m.SetForcingFunction(1.0)
        mats = {1: m, -1: m.CreateBC(-1, 0, [[0.0]], [0.0])}
        if neumann:
            mats[-2] = m.CreateBC(-2, 1, [[0.0]], [gu.NEUMANN_POISSON])
    else:
        m = sm.TPZElasticity3D(1, gu.E_MOD, gu.NU, gu.ELAST_FORCE)
        mats = {1: m, -1: m.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3))}
        if neumann:
            mats[-2] = m.CreateBC(-2, 1, np.zeros((3, 3)), gu.NEUMANN_ELAST)
    return mats


def fixture_setup(g):
    """(mesh, materials) of a committed reference fixture: oracle/refdriver.cpp's recipe."""
    m = g["meta"]
    bct = m["bctype"]
    if m.get("dim", 3) == 2:  # plane meshes: TPZMatPoisson(dim 2) / TPZElasticity2D, line elements on the boundary
        mesh = gridmesh.grid_mesh_2d(m["n"], m["p"], 2 if m["phys"] >= 2 else 1, triangles=bool(m["tet"]),
                                     bc_matids=(-1, -1, -2 if bct >= 1 else -1, -1), perturb=m["perturb"],
                                     node_perm=g["node_perm"] if m.get("scramble") else None)
        if m["phys"] >= 2:
            mat = sm.TPZElasticity2D(1, gu.E_MOD, gu.NU, *gu.E2D_FORCE, planestress=m["phys"] == 3)
            mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((2, 2)), np.zeros(2))}
            if bct == 1:
                mats[-2] = mat.CreateBC(-2, 1, np.zeros((2, 2)), gu.NEUMANN_ELAST2D)
            elif bct >= 2:
                mats[-2] = mat.CreateBC(-2, bct, gu.BC2D_VAL1, gu.BC2D_VAL2)
        else:
            mat = sm.TPZMatPoisson(1, 2)
            mat.SetForcingFunction(1.0)
            mats = {1: mat, -1: mat.CreateBC(-1, 0, [[0.0]], [0.0])}
            if bct >= 1:
                mats[-2] = mat.CreateBC(-2, 1, [[0.0]], [gu.NEUMANN_POISSON])
        return mesh, mats
    bc = (-1, -1, -1, -1, -1, -2 if bct >= 1 else -1)
    if m["tet"] == 3:  # hexahedra + pyramids (MMeshType::EHexaPyrMixed)
        mesh = gridmesh.hexpyr_mesh(m["n"], m["p"], 3 if m["phys"] == 1 else 1, bc_matids=bc, perturb=m["perturb"])
    else:
        mesh = gridmesh.grid_mesh(m["n"], m["p"], 3 if m["phys"] == 1 else 1, tetrahedra=m["tet"] == 1, prisms=m["tet"] == 2,
                                  bc_matids=bc, perturb=m["perturb"], node_perm=g["node_perm"] if m.get("scramble") else None)
    mats = materials_for(m["phys"], neumann=bct >= 1)
    if m["phys"] == 0 and bct == 2:  # TPZMatPoisson::ContributeBC type 2 with refdriver's Val1
        mats[-2] = mats[1].CreateBC(-2, 2, [[gu.POISSON_BC2_VAL1]], [gu.NEUMANN_POISSON])
    if m["phys"] == 1 and bct >= 2:  # the other TPZElasticity3D::ContributeBC types on the zmax face
        mats[-2] = mats[1].CreateBC(-2, bct, gu.BC_VAL1, gu.BC_VAL2)
    if m.get("bcfunc"):  # boundary data from the functions of oracle/refdriver.cpp (vectorised over the points)
        if m["phys"] == 0:
            mats[-1].SetForcingFunctionBC(lambda x: (0.3 + x[:, 0] * x[:, 1] - 0.5 * x[:, 2] * x[:, 2])[:, None])
            if -2 in mats:
                mats[-2].SetForcingFunctionBC(lambda x: (0.75 + 2.0 * x[:, 0] - x[:, 1] * x[:, 1])[:, None])
        else:
            f = lambda x: np.stack([0.01 * x[:, 1], -0.02 * x[:, 0] * x[:, 2], 0.005 + 0.01 * x[:, 2]], axis=1)  # noqa: E731
            mats[-1].SetForcingFunctionBC(f)
            if bct == 2:
                mats[-2].SetForcingFunctionBC(f)
    return mesh, mats


def relF(x, ref):
    return np.linalg.norm(x - ref) / np.linalg.norm(ref)


def interior_relF(ia, a, ref, big_rows):
    """Frobenius error restricted to rows without a penalty entry."""
    rows = np.repeat(np.arange(len(ia) - 1), np.diff(ia))
    keep = ~big_rows[rows]
    return np.linalg.norm((a - ref)[keep]) / max(np.linalg.norm(ref[keep]), 1e-300)


@pytest.mark.parametrize("name", gu.CORE_CASES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_against_reference_fixtures(name, symmetric):
    g = gu.load(name)
    m = g["meta"]
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    ia, ja, a, rhs = strmat.CreateAssemble()
    pre = "sym" if symmetric else "full"
    assert np.array_equal(ia, g[pre + "_ia"]) and np.array_equal(ja, g[pre + "_ja"])  # pattern: bit-exact
    assert relF(a, g[pre + "_a"]) <= TOL
    assert relF(rhs, g["rhs"]) <= TOL
    # stricter than the north star: rows without penalty entries, and the largest entry-wise error
    big = np.zeros(m["neq"], dtype=bool)
    rows = np.repeat(np.arange(m["neq"]), np.diff(ia))
    big[rows[np.abs(g[pre + "_a"]) > 1e9]] = True
    assert interior_relF(ia, a, g[pre + "_a"], big) <= TOL
    # the reference solution solves OUR system: residual of the reference's skyline-LDLt solution
    u = g["sol"]
    if symmetric:
        import scipy.sparse as sp
        U = sp.csr_matrix((a, ja, ia), shape=(m["neq"], m["neq"]))
        A = U + sp.triu(U, 1).T
    else:
        import scipy.sparse as sp
        A = sp.csr_matrix((a, ja, ia), shape=(m["neq"], m["neq"]))
    assert np.linalg.norm(A @ u - rhs) / np.linalg.norm(rhs) < 1e-9


@pytest.mark.parametrize("n,p,phys,tet", [(5, 1, 0, 0), (4, 2, 0, 0), (3, 2, 1, 0), (4, 1, 1, 0),
                                          (3, 2, 0, 1), (3, 2, 1, 1), (4, 1, 0, 1), (3, 1, 1, 1)])
def test_against_oracle(n, p, phys, tet):
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet),
                              bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL
        assert relF(rhs, rhs_ref) <= TOL
        # ragged last batch + re-assembly must reproduce (Zero()+Assemble path, TPZLinearAnalysis.cpp:73-77)
        a2, rhs2 = strmat.Assemble()
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL


@pytest.mark.parametrize("name", gu.CORE_CASES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_device_pattern_bit_exact(name, symmetric):
    """b200asm_build_pattern_device (CSR pattern built on the GPU) == the reference's Create(): memcmp of IA and JA;
    assembling into it gives the reference's values."""
    g = gu.load(name)
    m = g["meta"]
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    ia, ja = strmat.Create(on_device=True)
    pre = "sym" if symmetric else "full"
    assert np.array_equal(ia, g[pre + "_ia"]) and np.array_equal(ja, g[pre + "_ja"])
    a, rhs = strmat.Assemble()
    assert relF(a, g[pre + "_a"]) <= TOL and relF(rhs, g["rhs"]) <= TOL


@pytest.mark.parametrize("n,p,ns,tet", [(24, 2, 1, 0), (12, 2, 3, 1), (10, 3, 1, 0), (40, 1, 3, 0)])
def test_device_pattern_matches_host_builder_at_scale(n, p, ns, tet):
    mesh = gridmesh.grid_mesh(n, p, ns, tetrahedra=bool(tet))
    idx, graph = mesh.element_graph()
    for symmetric in (True, False):
        ia, ja = capi.build_pattern(symmetric, idx, graph, mesh.block_pos, mesh.block_size, 0)
        ctx = capi.Context(0)
        neq, nnz = ctx.build_pattern_device(symmetric, idx, graph, mesh.block_pos, mesh.block_size)
        assert (neq, nnz) == (mesh.neq, len(ja))
        ia_d, ja_d = ctx.get_pattern(neq, nnz)
        assert np.array_equal(ia, ia_d) and np.array_equal(ja, ja_d)
        ctx.close()


@pytest.mark.parametrize("name", ["hex_p2_poisson_n3", "hex_p2_elast_n2_pert", "tet_p2_poisson_n2_pert", "hex_p3_poisson_n3_pert",
                                  "hex_p1_poisson_n3_pert"])
@pytest.mark.parametrize("symmetric", [True, False])
def test_device_cg_reproduces_reference_solution(name, symmetric):
    """GPU CG (Jacobi-preconditioned, the reference's algorithm) on the device-resident assembled system against the
    reference's own direct (skyline LDLt) solution of the same mesh: 1e-10 (north_star)."""
    g = gu.load(name)
    m = g["meta"]
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    strmat.CreateAssemble()
    u, iters, resid = strmat.SolveCG(max_iter=20000, tol=1e-15)
    assert resid <= 1e-14 and 0 < iters < 20000, (iters, resid)
    assert relF(u, g["sol"]) <= 1e-10
    # restart from the solution: converged at iteration 0 (FromCurrent path)
    # (the true residual b - A u of the penalised system is a little above the recurrence residual the iteration stopped on)
    u2, it2, res2 = strmat.SolveCG(max_iter=10, tol=1e-9, x0=u)
    assert it2 == 0 and res2 <= 1e-9 and np.array_equal(u2, u)


@pytest.mark.parametrize("n,p,phys,tet,engine,symmetric", [(12, 2, 0, 0, 1, True), (10, 2, 1, 0, 1, True), (8, 2, 1, 1, 0, False),
                                                           (16, 1, 0, 0, 0, True), (9, 2, 0, 1, 1, False), (8, 3, 0, 0, 1, True)])
def test_overlapped_download_matches_plain_download(n, p, phys, tet, engine, symmetric):
    """b200asm_assemble(a_host, rhs_host) copies finished rows while later element chunks run (option "overlap"): the
    host result must equal the plain assemble-then-download of the same context, and the oracle."""
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.1)
    mats = materials_for(phys, neumann=True)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=engine, scatter="colored")
    strmat.ctx.set_option("overlap", 0)
    ia, ja, a_plain, rhs_plain = strmat.CreateAssemble()          # deterministic reference run (coloured, no overlap)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=engine)
    strmat.ctx.set_option("overlap_min_elements", 128)
    strmat.ctx.set_option("overlap_min_bytes", 0)
    strmat.SetPattern(ia, ja)
    import torch
    a = torch.full((len(ja),), float("nan"), dtype=torch.float64).pin_memory().numpy()  # the overlap needs page-locked memory
    rhs = torch.full((mesh.neq,), float("nan"), dtype=torch.float64).pin_memory().numpy()
    launches0 = strmat.ctx.counters()[0]
    strmat.Assemble(a, rhs)
    assert strmat.ctx.counters()[0] - launches0 > len(mesh.blocks)  # the volume group really ran in chunks
    assert relF(a, a_plain) <= TOL and relF(rhs, rhs_plain) <= TOL
    if mesh.nelements <= 3000:
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    a2, _ = strmat.Assemble()                                     # again: events / cached frontiers are reused
    assert relF(a2, a_plain) <= TOL


@pytest.mark.parametrize("n,p,phys,tet,engine", [(4, 2, 0, 0, 1), (3, 2, 1, 0, 1), (3, 2, 1, 1, 0), (3, 4, 0, 0, 1), (5, 1, 0, 0, 0)])
def test_rhs_only_assembly(n, p, phys, tet, engine):
    """TPZStrMatParInterface::Assemble(rhs): load vector only, equal to the rhs of the full assembly (and the oracle's);
    the CSR values on the device stay untouched; also works before any pattern exists (AssembleResidual first)."""
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    fresh = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, engine=engine)
    fresh._flatten()
    rhs0 = np.empty(mesh.neq)
    fresh.ctx.assemble_rhs(rhs0)                       # no pattern yet
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, engine=engine)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    rhs1 = strmat.AssembleRhs()
    assert relF(rhs1, rhs_ref) <= TOL and relF(rhs0, rhs_ref) <= TOL
    a_after = np.empty_like(a)
    strmat.ctx.download(a_after, None)
    assert np.array_equal(a_after, a)                  # matrix untouched by the rhs-only pass


@pytest.mark.parametrize("n,p,tet,engine", [(3, 2, 0, 1), (4, 1, 0, 1), (3, 2, 1, 0), (2, 3, 0, 0)])
def test_elasticity_prestress_load_vector(n, p, tet, engine):
    """ef(3j+k) += w (f_k phi_j - prestress_k dphix(k,j))  (TPZElasticity3D.cpp:278) with a non-zero prestress: the
    point-wise branch of the load vector in every kernel family."""
    mesh = gridmesh.grid_mesh(n, p, 3, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    m = sm.TPZElasticity3D(1, gu.E_MOD, gu.NU, (0.3, -0.2, -1.0), prestress=(1.5, -0.75, 2.25))
    mats = {1: m, -1: m.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3)), -2: m.CreateBC(-2, 1, np.zeros((3, 3)), gu.NEUMANN_ELAST)}
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, engine=engine)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


@pytest.mark.parametrize("bctype", [2, 3, 5, 6, 7, 8])
@pytest.mark.parametrize("tet,p", [(0, 2), (1, 2), (0, 3)])
def test_elasticity_bc_types(bctype, tet, p):
    """TPZElasticity3D::ContributeBC types 2 (mixed), 3 (directional null Dirichlet), 5-8 (directional Dirichlet)
    (Material/Elasticity/TPZElasticity3D.cpp:684-772) on one face, full Dirichlet elsewhere."""
    mesh = gridmesh.grid_mesh(3 if p < 3 else 2, p, 3, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    m = sm.TPZElasticity3D(1, gu.E_MOD, gu.NU, gu.ELAST_FORCE)
    val1 = np.array([[4.0, 0.5, 0.0], [0.5, 3.0, 0.25], [0.0, 0.25, 5.0]])
    mats = {1: m, -1: m.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3)), -2: m.CreateBC(-2, bctype, val1, [0.3, -0.2, 0.7])}
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,tri,planestress,scatter", [(7, 2, 0, 0, "atomic"), (6, 2, 1, 1, "atomic"), (9, 1, 0, 1, "colored"),
                                                          (8, 1, 1, 0, "atomic"), (5, 2, 0, 1, "colored")])
def test_plane_elasticity_against_oracle(n, p, tri, planestress, scatter):
    """TPZElasticity2D (Material/Elasticity/TPZElasticity2D.cpp:86-203) on plane meshes, with prestress, Dirichlet /
    Neumann / directional-null-Dirichlet sides (line elements), symmetric and full storage."""
    mesh = gridmesh.grid_mesh_2d(n, p, 2, triangles=bool(tri), bc_matids=(-1, -3, -2, -1), perturb=0.12)
    mat = sm.TPZElasticity2D(1, gu.E_MOD, gu.NU, 0.5, -1.0, planestress=bool(planestress))
    mat.SetPreStress(0.2, -0.1, 0.05)
    mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((2, 2)), [0.01, -0.02]), -2: mat.CreateBC(-2, 1, np.zeros((2, 2)), [0.25, -0.5]),
            -3: mat.CreateBC(-3, 3, np.zeros((2, 2)), [1.0, 0.0])}
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, scatter=scatter)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        a2, rhs2 = strmat.Assemble()
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,phys,planestress,shuffle", [(5, 3, 2, 0, 1), (4, 4, 2, 1, 1), (6, 4, 0, 0, 1), (5, 3, 0, 0, 0),
                                                          (12, 4, 2, 0, 0)])
def test_high_order_plane_against_oracle(n, p, phys, planestress, shuffle):
    """p = 3, 4 quadrilaterals of plane meshes with oriented line boundary elements (TPZShapeLinear of order p,
    Shape/pzshapelinear.cpp:306-312,360-363): one group per side-orientation class."""
    perm = np.random.default_rng(77 + n + p).permutation((n + 1) ** 2) if shuffle else None
    mesh = gridmesh.grid_mesh_2d(n, p, 2 if phys else 1, bc_matids=(-1, -3 if phys else -1, -2, -1), perturb=0.12, node_perm=perm)
    if phys:
        mat = sm.TPZElasticity2D(1, gu.E_MOD, gu.NU, 0.5, -1.0, planestress=bool(planestress))
        mat.SetPreStress(0.2, -0.1, 0.05)
        mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((2, 2)), [0.01, -0.02]),
                -2: mat.CreateBC(-2, 1, np.zeros((2, 2)), [0.25, -0.5]), -3: mat.CreateBC(-3, 3, np.zeros((2, 2)), [1.0, 0.0])}
    else:
        mat = sm.TPZMatPoisson(1, 2)
        mat.SetForcingFunction(1.5)
        mats = {1: mat, -1: mat.CreateBC(-1, 0, [[0.0]], [0.25]), -2: mat.CreateBC(-2, 1, [[0.0]], [0.75])}
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,tri", [(8, 2, 0), (7, 2, 1), (10, 1, 0), (9, 1, 1)])
def test_plane_poisson_against_oracle(n, p, tri):
    """TPZMatPoisson(dim 2) on plane meshes."""
    mesh = gridmesh.grid_mesh_2d(n, p, 1, triangles=bool(tri), bc_matids=(-1, -1, -2, -1), perturb=0.12)
    mat = sm.TPZMatPoisson(1, 2)
    mat.SetForcingFunction(1.5)
    mats = {1: mat, -1: mat.CreateBC(-1, 0, [[0.0]], [0.25]), -2: mat.CreateBC(-2, 1, [[0.0]], [0.75])}
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble(); a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,phys,tet,engine", [(4, 2, 0, 0, 1), (3, 2, 1, 0, 1), (3, 2, 1, 1, 1), (4, 1, 0, 1, 0), (3, 3, 0, 0, 1)])
def test_filtered_equations(n, p, phys, tet, engine):
    """Destination index -1 = equation removed by a TPZEquationFilter (TPZEquationFilter::Filter,
    StrMatrix/TPZEquationFilter.h:120-141): the condensed system of the upper three quarters of the equations equals
    the corresponding rows / columns of the full assembly."""
    import copy
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    idx, graph = mesh.element_graph()
    ia, ja = capi.build_pattern(True, idx, graph, mesh.block_pos, mesh.block_size, 0)
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    lo = mesh.neq // 4
    cond = copy.copy(mesh)
    cond.blocks = []
    for b in mesh.blocks:
        b2 = copy.copy(b)
        b2.dest = np.where(b.dest >= lo, b.dest - lo, -1)
        cond.blocks.append(b2)
    cond.neq = mesh.neq - lo
    # symmetric upper storage: the rows >= lo only hold columns >= lo, so the condensed pattern is a slice
    ia_c = ia[lo:] - ia[lo]
    ja_c = ja[ia[lo]:] - lo
    for scatter in ("atomic", "colored"):
        strmat = sm.TPZStructMatrixB200(cond, mats, symmetric=True, engine=engine, scatter=scatter)
        strmat.SetPattern(ia_c, ja_c)
        a, rhs = strmat.Assemble()
        assert relF(a, a_ref[ia[lo]:]) <= TOL and relF(rhs, rhs_ref[lo:]) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref[lo:]) <= TOL


def _shuffled(n, seed):
    """Random renumbering of the (n+1)^3 grid nodes: every element gets its own side orientations (p >= 3)."""
    return np.random.default_rng(seed).permutation((n + 1) ** 3)


@pytest.mark.parametrize("n,p,phys,engine,shuffle", [(3, 3, 0, 1, 1), (3, 4, 0, 1, 1), (2, 3, 1, 1, 1), (4, 3, 0, 0, 0),
                                                     (3, 4, 0, 0, 1), (5, 4, 0, 1, 0), (5, 3, 0, 1, 0)])
def test_high_order_hex_against_oracle(n, p, phys, engine, shuffle):
    """p = 3, 4 hexahedra (BASELINE config C4): orientation-dependent shape tables, one group per orientation class;
    DMMA superblock team kernels (engine 1) and register-tile kernels (engine 0); quadrilateral faces of order p."""
    perm = _shuffled(n, 1234 + n + p) if shuffle else None
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12, node_perm=perm)
    mats = materials_for(phys, neumann=True)
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=engine)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL
        assert relF(rhs, rhs_ref) <= TOL
        a2, rhs2 = strmat.Assemble()
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL


@pytest.mark.parametrize("p,phys,variant,scatter", [(2, 1, 0, "atomic"), (2, 0, 0, "atomic"), (1, 1, 0, "atomic"), (1, 0, 0, "colored"),
                                                    (2, 1, 0, "colored"), (2, 1, 7, "atomic"), (2, 0, 7, "atomic")])
def test_tetrahedra_kernels_against_oracle(p, phys, variant, scatter):
    """Straight-sided tetrahedra: the closed-form kernel (affine_simplex.cuh, default) and the DMMA Gram kernels it replaced
    (variant 7), with prestress, Neumann faces, symmetric and full storage, re-assembly and the rhs-only path."""
    mesh = gridmesh.grid_mesh(4, p, 3 if phys else 1, tetrahedra=True, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.15)
    mats = materials_for(phys, neumann=True)
    if phys:
        mats[1].fPreStress = [0.3, -0.2, 0.1]
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, variant=variant, scatter=scatter)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        a2, rhs2 = strmat.Assemble()
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


def _sheared(mesh):
    """An affine image of the grid: every hexahedron becomes a general parallelepiped."""
    A = np.array([[1.3, 0.2, -0.1], [0.15, 0.8, 0.25], [-0.2, 0.1, 1.1]])
    mesh.nodes[:] = mesh.nodes @ A.T + np.array([0.3, -0.2, 0.5])
    return mesh


@pytest.mark.parametrize("p,phys,scatter,shear", [(2, 0, "atomic", 1), (2, 1, "atomic", 1), (1, 0, "atomic", 1), (1, 1, "colored", 0),
                                                  (2, 1, "colored", 1), (2, 0, "atomic", 0)])
def test_parallelepiped_hexahedra_closed_form(p, phys, scatter, shear):
    """Hexahedral groups whose elements are all parallelepipeds take the closed-form kernel (affine_hex.cuh; decided on the
    device from the node coordinates): against the oracle, symmetric and full, with prestress, re-assembly, rhs only; and the
    decision follows the coordinates (b200asm_set_nodes: perturbed -> Gram/DMMA kernel -> back)."""
    n = 4
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, bc_matids=(-1, -1, -1, -1, -1, -2))
    if shear:
        _sheared(mesh)
    mats = materials_for(phys, neumann=True)
    if phys:
        mats[1].fPreStress = [0.3, -0.2, 0.1]
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, scatter=scatter)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        a2, rhs2 = strmat.Assemble()
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL
        # the same context with the closed form switched off must agree
        strmat.ctx.set_option("affine", 0)
        a3, rhs3 = strmat.Assemble()
        assert relF(a3, a_ref) <= TOL and relF(rhs3, rhs_ref) <= TOL
        strmat.ctx.set_option("affine", 1)
        # move the nodes: no longer parallelepipeds -> the general kernel (and its scatter-map layout) takes over
        moved = gridmesh.grid_mesh(n, p, 3 if phys else 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.1)
        strmat.ctx.set_nodes(moved.nodes)
        a4, rhs4 = strmat.Assemble()
        a4_ref, rhs4_ref = oracle_assemble(moved, mats, symmetric, ia, ja)
        assert relF(a4, a4_ref) <= TOL and relF(rhs4, rhs4_ref) <= TOL
        strmat.ctx.set_nodes(mesh.nodes)
        a5, rhs5 = strmat.Assemble()
        assert relF(a5, a_ref) <= TOL and relF(rhs5, rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,phys,tet,perturb", [(17, 1, 0, 0, 0.1), (10, 2, 1, 1, 0.1), (17, 1, 1, 0, 0.0), (16, 2, 0, 0, 0.1)])
def test_locality_order_matches_oracle(n, p, phys, tet, perturb):
    """Groups of >= 4096 elements are stored along a Morton curve (option "locality"): same matrix as the oracle's mesh-order
    assembly and as the context that keeps the mesh order; forcing tables and the overlapped download follow the permutation."""
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=perturb)
    assert len(mesh.blocks[0].elnodes) >= 4096
    mats = materials_for(phys, neumann=True)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    strmat.ctx.set_option("overlap_min_elements", 1024)
    strmat.ctx.set_option("overlap_min_bytes", 0)
    ia, ja, a, rhs = strmat.CreateAssemble()
    plain = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    plain.ctx.set_option("locality", 0)
    plain.SetPattern(ia, ja)
    a0, rhs0 = plain.Assemble()
    assert relF(a, a0) <= TOL and relF(rhs, rhs0) <= TOL
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    import torch
    ah = torch.full((len(ja),), float("nan"), dtype=torch.float64).pin_memory().numpy()
    rh = torch.full((mesh.neq,), float("nan"), dtype=torch.float64).pin_memory().numpy()
    strmat.Assemble(ah, rh)   # overlapped download, chunks of the permuted group
    assert relF(ah, a_ref) <= TOL and relF(rh, rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,phys,tet", [(5, 2, 0, 0), (4, 2, 0, 1)])
def test_engines_agree(n, p, phys, tet):
    """Register-tile DFMA kernels (engine 0) and DMMA panel kernels (engine 1) against the oracle and each other."""
    mesh = gridmesh.grid_mesh(n, p, 1, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    res = {}
    for engine in (0, 1):
        for symmetric in (True, False):
            strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=engine)
            ia, ja, a, rhs = strmat.CreateAssemble()
            a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
            assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
            res[(engine, symmetric)] = a
    assert relF(res[(0, True)], res[(1, True)]) <= TOL


@pytest.mark.parametrize("n,p,phys,tet,engine", [(5, 2, 0, 0, 1), (4, 2, 1, 0, 1), (4, 2, 1, 1, 0), (6, 1, 0, 0, 0), (4, 2, 0, 1, 1)])
def test_colored_scatter_is_deterministic(n, p, phys, tet, engine):
    """Conflict-free scatter by element colouring (no atomics): parity with the oracle and bit-reproducible results
    (the reference is bit-reproducible too); the atomic mode agrees to rounding."""
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    for symmetric in (True, False):
        runs = []
        for _rep in range(2):
            strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=engine, scatter="colored")
            ia, ja, a, rhs = strmat.CreateAssemble()
            a2, rhs2 = strmat.Assemble()
            assert np.array_equal(a, a2) and np.array_equal(rhs, rhs2)
            runs.append((a, rhs))
        assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(runs[0][0], a_ref) <= TOL and relF(runs[0][1], rhs_ref) <= TOL


def _full_from_sym(ia, ja, a, neq):
    import scipy.sparse as sp
    U = sp.csr_matrix((a, ja, ia), shape=(neq, neq))
    return U + sp.triu(U, 1).T


@pytest.mark.parametrize("tet", [0, 1])
def test_properties_at_scale(tet):
    """Sizes beyond the oracle: without boundary faces K annihilates constants (Poisson) and rigid-body
    modes (elasticity); symmetric and full storage agree; rhs sums to the total load."""
    import scipy.sparse as sp
    n = 20 if not tet else 12
    for phys in (0, 1):
        ns = 3 if phys else 1
        nodes, blocks = gridmesh.grid_elements(n, tetrahedra=bool(tet), perturb=0.1)
        blocks = [b for b in blocks if b[1] == 1]  # volume elements only -> singular K with known kernel
        mesh = gridmesh.flatten(nodes, blocks, 2, ns)
        mats = materials_for(phys)
        s_sym = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
        ia, ja, a, rhs = s_sym.CreateAssemble()
        K = _full_from_sym(ia, ja, a, mesh.neq)
        s_full = sm.TPZStructMatrixB200(mesh, mats, symmetric=False)
        iaf, jaf, af, rhsf = s_full.CreateAssemble()
        Kf = sp.csr_matrix((af, jaf, iaf), shape=(mesh.neq, mesh.neq))
        scale = np.abs(a).max()
        assert abs(K - Kf).max() <= 1e-12 * scale
        assert relF(rhsf, rhs) <= 1e-13
        # kernel vectors expressed in the hierarchical basis: vertex functions reproduce linears,
        # edge/face/interior coefficients are zero
        nvert = len(nodes)
        vert_eq = np.zeros((nvert, ns), dtype=np.int64)
        b = mesh.blocks[0]
        ncorner = b.elnodes.shape[1]
        for c in range(ncorner):
            for s in range(ns):
                vert_eq[b.elnodes[:, c], s] = b.dest[:, c * ns + s]
        modes = []
        if phys == 0:
            u = np.zeros(mesh.neq)
            u[vert_eq[:, 0]] = 1.0
            modes.append(u)
        else:
            for s in range(3):  # translations
                u = np.zeros(mesh.neq)
                u[vert_eq[:, s]] = 1.0
                modes.append(u)
            for (i, j) in ((0, 1), (1, 2), (2, 0)):  # rotations: u_i = x_j, u_j = -x_i
                u = np.zeros(mesh.neq)
                u[vert_eq[:, i]] = nodes[:, j]
                u[vert_eq[:, j]] = -nodes[:, i]
                modes.append(u)
        if tet or True:
            # geometry is multilinear: linear fields are reproduced by the vertex functions on tets and,
            # for hexes, by vertex functions too (trilinear map of trilinear functions)
            for u in modes:
                r = K @ u
                assert np.abs(r).max() <= 1e-9 * scale * max(1.0, np.abs(u).max()), np.abs(r).max()
        # total load: sum of rhs over vertex+edge... for Poisson f=1: sum_i rhs_i * (coefficients of u=1) = volume
        if phys == 0:
            vol = modes[0] @ rhs
            assert abs(vol - _mesh_volume(nodes, blocks[0])) < 1e-10


def _mesh_volume(nodes, block):
    topo, _m, el = block
    X = nodes[el]
    if X.shape[1] == 4:
        d = X[:, 1:] - X[:, :1]
        return np.abs(np.linalg.det(d)).sum() / 6.0
    # hexahedron: split into 5 tets is not exact for warped hexes -> integrate |J| with 2x2x2 Gauss (exact for trilinear)
    from neopz_b200 import capi
    qp, qw = capi.tensor_rule(capi.HEX, 4)
    _phi, dN = capi.shape_tables(capi.HEX, 1, qp)  # [q][3][8]
    J = np.einsum("qda,eak->eqkd", dN, X)  # [el][q][k][d]
    return (np.abs(np.linalg.det(J)) * qw[None, :]).sum()
