"""GPU parity of the prism / pyramid configurations: the CUDA path through the C ABI against
  (1) the fixtures of the unmodified reference on MMeshType::EPrismatic and EHexaPyrMixed grids (pattern bit-exact,
      values 1e-12, both storages, host and device pattern builders),
  (2) the oracle on fresh meshes (prism grids of other sizes; the fixture's hexahedra + pyramids with moved nodes),
  (3) properties: constants / rigid-body modes in the kernel of K, symmetric vs full storage.
Same tolerance as tests/test_gpu_parity.py."""
import numpy as np
import pytest

from neopz_b200 import capi, gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, fixture_setup, interior_relF, materials_for, relF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gu.WEDGE_CASES)
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
    big = np.zeros(m["neq"], dtype=bool)
    rows = np.repeat(np.arange(m["neq"]), np.diff(ia))
    big[rows[np.abs(g[pre + "_a"]) > 1e9]] = True
    assert interior_relF(ia, a, g[pre + "_a"], big) <= TOL
    import scipy.sparse as sp
    U = sp.csr_matrix((a, ja, ia), shape=(m["neq"], m["neq"]))
    A = U + sp.triu(U, 1).T if symmetric else U
    assert np.linalg.norm(A @ g["sol"] - rhs) / np.linalg.norm(rhs) < 1e-9  # the reference's LDLt solution solves our system


@pytest.mark.parametrize("name", gu.WEDGE_CASES)
def test_device_pattern_and_colored_scatter(name):
    """Pattern built on the GPU == the reference's Create(); the deterministic coloured scatter gives the same values."""
    g = gu.load(name)
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, scatter="colored")
    ia, ja = strmat.Create(on_device=True)
    assert np.array_equal(ia, g["sym_ia"]) and np.array_equal(ja, g["sym_ja"])
    a, rhs = strmat.Assemble()
    assert relF(a, g["sym_a"]) <= TOL and relF(rhs, g["rhs"]) <= TOL
    a2, rhs2 = strmat.Assemble()
    assert np.array_equal(a, a2) and np.array_equal(rhs, rhs2)  # bit-reproducible


@pytest.mark.parametrize("n,p,phys", [(5, 1, 0), (4, 2, 0), (3, 2, 1), (4, 1, 1)])
def test_prisms_against_oracle(n, p, phys):
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, prisms=True, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        a2, rhs2 = strmat.Assemble()  # ragged last batch + re-assembly
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
        r3 = strmat.AssembleRhs()     # TPZStrMatParInterface::Assemble(rhs)
        assert relF(r3, rhs_ref) <= TOL


@pytest.mark.parametrize("name", ["hexpyr_p2_poisson_n2_pert", "hexpyr_p1_elast_n3", "hexpyr_p2_elast_n2_pert"])
def test_pyramids_moved_nodes_against_oracle(name):
    """The fixture's hexahedra + pyramids with differently perturbed nodes (b200asm_set_nodes): new geometry, same pattern."""
    g = gu.load(name)
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    ia, ja = strmat.Create()
    rng = np.random.default_rng(5)
    h = 1.0 / g["meta"]["n"]
    mesh.nodes = mesh.nodes + 0.08 * h * rng.uniform(-1.0, 1.0, mesh.nodes.shape)
    strmat.ctx.set_nodes(mesh.nodes)
    a, rhs = strmat.Assemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL


@pytest.mark.parametrize("prisms", [True, False])
def test_kernel_of_K(prisms):
    """Without boundary terms the stiffness matrix annihilates constants (Poisson) and rigid-body modes (elasticity):
    holds at any size, here prisms 6^3 x 2 and the hexahedra + pyramids fixture mesh."""
    import scipy.sparse as sp
    for phys in (0, 1):
        ns = 3 if phys else 1
        if prisms:
            mesh = gridmesh.grid_mesh(6, 2, ns, prisms=True, perturb=0.1)
        else:
            g = gu.load("hexpyr_p2_poisson_n2_pert")
            mesh = gridmesh.mesh_from_elements(g["nodes"], g["el_type"], g["el_matid"], g["el_nodes"], 2, ns)
        mats = materials_for(phys)
        vol = [b for b in mesh.blocks if gridmesh.DIM[b.topology] == 3]
        sub = gridmesh.FlatMesh(porder=mesh.porder, nstate=ns, nodes=mesh.nodes, blocks=vol, block_pos=mesh.block_pos,
                                block_size=mesh.block_size, neq=mesh.neq)
        strmat = sm.TPZStructMatrixB200(sub, {1: mats[1]}, symmetric=False)
        ia, ja, a, _rhs = strmat.CreateAssemble()
        A = sp.csr_matrix((a, ja, ia), shape=(mesh.neq, mesh.neq))
        # corner (vertex) functions interpolate; the higher functions vanish at the vertices: a constant / linear field is
        # represented by its nodal values on the vertex equations and zero elsewhere
        u = np.zeros((mesh.neq, 1 if phys == 0 else 6))
        for b in vol:
            nc = b.elnodes.shape[1]
            for k in range(nc):
                eq = mesh.block_pos[b.connects[:, k]]
                x = mesh.nodes[b.elnodes[:, k]]
                if phys == 0:
                    u[eq, 0] = 1.0
                else:
                    for d in range(3):
                        u[eq + d, d] = 1.0
                    u[eq + 1, 3], u[eq + 0, 3] = x[:, 0], -x[:, 1]
                    u[eq + 2, 4], u[eq + 1, 4] = x[:, 1], -x[:, 2]
                    u[eq + 0, 5], u[eq + 2, 5] = x[:, 2], -x[:, 0]
        assert np.abs(A @ u).max() <= 1e-10 * np.abs(a).max()


# n, p, phys, tet (2 prisms / 3 hexahedra + pyramids), symmetric, solve
DROPIN_CASES = [(4, 2, 0, 2, 1, 1), (3, 2, 1, 2, 1, 1), (4, 1, 1, 2, 0, 0), (4, 2, 0, 3, 1, 1), (2, 2, 1, 3, 1, 1), (4, 1, 0, 3, 0, 0)]


@pytest.mark.parametrize("case", DROPIN_CASES)
def test_dropin_strategy_matches_reference(case):
    """The unmodified TPZLinearAnalysis on EPrismatic / EHexaPyrMixed grids: TPZStructMatrixB200 vs TPZStructMatrixOR
    (tests/dropin/dropin_test.cpp; device-side Create() on every other case)."""
    import json
    import os
    import subprocess
    from tests.test_gpu_dropin import BIN
    if not os.path.exists(BIN):
        pytest.skip("tests/_bin/dropin_test not built (needs /root/reference at build time)")
    device_create = DROPIN_CASES.index(case) % 2
    out = subprocess.run([BIN] + [str(x) for x in case] + ["4", str(device_create)], capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[-1])
    assert r["ia_identical"] == 1 and r["ja_identical"] == 1
    assert r["relF_A"] <= 1e-12 and r["relF_A_nonpenalty_rows"] <= 1e-12 and r["relF_rhs"] <= 1e-12
    assert r["relF_cg_solution"] <= 1e-10 and r["relF_device_cg_solution"] <= 1e-10
    assert r["relF_residual_rhs"] <= 1e-12
    assert out.returncode == 0
