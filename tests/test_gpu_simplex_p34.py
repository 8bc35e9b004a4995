"""GPU parity of tetrahedra / triangles of order 3 and 4 (oriented edges and triangular sides, tetrahedron interior
function): the CUDA path through the C ABI against the fixtures of the unmodified reference (scrambled node numbering:
up to 21 orientation classes among 40 tetrahedra), against the oracle on fresh meshes, and the drop-in strategy inside
the unmodified TPZLinearAnalysis.  Same tolerance as tests/test_gpu_parity.py."""
import json
import os
import subprocess

import numpy as np
import pytest

from neopz_b200 import gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, fixture_setup, interior_relF, materials_for, relF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gu.SIMPLEX34_CASES)
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
    assert np.linalg.norm(A @ g["sol"] - rhs) / np.linalg.norm(rhs) < 1e-9


@pytest.mark.parametrize("name", gu.SIMPLEX34_CASES)
def test_device_pattern_and_colored_scatter(name):
    g = gu.load(name)
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, scatter="colored")
    ia, ja = strmat.Create(on_device=True)
    assert np.array_equal(ia, g["sym_ia"]) and np.array_equal(ja, g["sym_ja"])
    a, rhs = strmat.Assemble()
    assert relF(a, g["sym_a"]) <= TOL and relF(rhs, g["rhs"]) <= TOL
    a2, rhs2 = strmat.Assemble()
    assert np.array_equal(a, a2) and np.array_equal(rhs, rhs2)  # bit-reproducible


@pytest.mark.parametrize("n,p,phys,seed,variant", [(3, 3, 0, 0, 0), (2, 4, 0, 3, 0), (2, 3, 1, 4, 0), (2, 4, 1, 0, 0), (4, 3, 0, 9, 0),
                                                    (2, 3, 1, 4, 21), (2, 4, 0, 3, 21)])
def test_tetrahedra_against_oracle(n, p, phys, seed, variant):
    """Fresh meshes: perturbed nodes, node numbering shuffled with `seed` (0: the grid numbering), Neumann face.  variant 0: the
    closed-form kernel (default of tetrahedra p = 3, 4 since round 2), 21: the register-tile kernel it replaced."""
    nn = (n + 1) ** 3
    perm = np.random.default_rng(seed).permutation(nn) if seed else None
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=True, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12,
                              node_perm=perm)
    mats = materials_for(phys, neumann=True)
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, variant=variant)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        a2, rhs2 = strmat.Assemble()
        assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


@pytest.mark.parametrize("n,p,planestress,seed", [(4, 3, False, 2), (3, 4, True, 0), (5, 4, False, 6)])
def test_plane_triangles_against_oracle(n, p, planestress, seed):
    nn = (n + 1) ** 2
    perm = np.random.default_rng(seed).permutation(nn) if seed else None
    mesh = gridmesh.grid_mesh_2d(n, p, 2, triangles=True, bc_matids=(-1, -1, -2, -1), perturb=0.1, node_perm=perm)
    mat = sm.TPZElasticity2D(1, gu.E_MOD, gu.NU, *gu.E2D_FORCE, planestress=planestress)
    mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((2, 2)), np.zeros(2)), -2: mat.CreateBC(-2, 1, np.zeros((2, 2)), gu.NEUMANN_ELAST2D)}
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL


# n, p, phys, tet, symmetric, solve
DROPIN_CASES = [(3, 3, 0, 1, 1, 1), (2, 4, 0, 1, 1, 1), (2, 3, 1, 1, 1, 0), (2, 4, 1, 1, 0, 0), (5, 3, 2, 1, 1, 1), (4, 4, 3, 1, 1, 1)]


@pytest.mark.parametrize("case", DROPIN_CASES)
def test_dropin_strategy_matches_reference(case):
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
