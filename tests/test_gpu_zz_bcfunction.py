"""Boundary data given by a function (TPZBndCondT::SetForcingFunctionBC — the usual way Dirichlet data of a manufactured
solution enter a NeoPZ program): the host tabulates the function at the integration points of every boundary element, the
boundary kernel reads the table.  Against the oracle (same points) through the Python mirror, and against the unmodified
reference through the drop-in test (which also checks the points: data.x of the reference vs TPZGeoEl::X in the strategy)."""
import json
import os
import subprocess

import numpy as np
import pytest

from neopz_b200 import gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, materials_for, relF

pytestmark = pytest.mark.gpu


def g1(x):
    return (0.3 + x[:, 0] * x[:, 1] - 0.5 * x[:, 2] ** 2)[:, None]


def g3(x):
    return np.stack([0.01 * x[:, 1], -0.02 * x[:, 0] * x[:, 2], 0.005 + 0.01 * x[:, 2]], axis=1)


@pytest.mark.parametrize("n,p,phys,tet,prisms", [(4, 2, 0, 0, False), (3, 2, 1, 0, False), (3, 2, 1, 1, False), (3, 3, 0, 0, False),
                                                  (3, 1, 0, 0, True), (2, 4, 0, 1, False)])
def test_boundary_functions_against_oracle(n, p, phys, tet, prisms):
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), prisms=prisms, bc_matids=(-1, -1, -1, -1, -1, -2),
                              perturb=0.12)
    mats = materials_for(phys, neumann=True)
    mats[-1].SetForcingFunctionBC(g3 if phys else g1)           # Dirichlet data from a function
    if phys:
        mats[-2] = mats[1].CreateBC(-2, 2, gu.BC_VAL1, gu.BC_VAL2)  # mixed: val1 * function in the load vector
        mats[-2].SetForcingFunctionBC(g3)
    else:
        mats[-2].SetForcingFunctionBC(lambda x: 2.0 * g1(x))     # Neumann data from a function
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


def test_boundary_function_plane():
    mesh = gridmesh.grid_mesh_2d(5, 2, 2, bc_matids=(-1, -1, -2, -1), perturb=0.1)
    mat = sm.TPZElasticity2D(1, gu.E_MOD, gu.NU, *gu.E2D_FORCE)
    mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((2, 2)), np.zeros(2)), -2: mat.CreateBC(-2, 1, np.zeros((2, 2)), np.zeros(2))}
    mats[-1].SetForcingFunctionBC(lambda x: np.stack([0.01 * x[:, 1] + 0.02 * x[:, 0] * x[:, 1], -0.03 * x[:, 0]], axis=1))
    mats[-2].SetForcingFunctionBC(lambda x: np.stack([0.25 + x[:, 0], -0.5 * x[:, 0] ** 2], axis=1))
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL


# n, p, phys, tet, symmetric: Dirichlet data of matid -1 from a function of x (tests/dropin/dropin_test.cpp, bcfunc = 1)
# (bcfunc = 2 on plane meshes: the body force of TPZElasticity2D is a function of x too)
DROPIN_CASES = [(4, 2, 0, 0, 1), (3, 2, 1, 0, 1), (3, 2, 1, 1, 0), (6, 2, 2, 0, 1), (3, 3, 0, 0, 1), (3, 2, 0, 2, 1), (5, 1, 3, 1, 0),
                (6, 2, 2, 0, 1, 2), (4, 3, 3, 1, 0, 2)]


@pytest.mark.parametrize("case", DROPIN_CASES)
def test_dropin_strategy_matches_reference(case):
    from tests.test_gpu_dropin import BIN
    if not os.path.exists(BIN):
        pytest.skip("tests/_bin/dropin_test not built (needs /root/reference at build time)")
    args = [str(x) for x in case[:5]] + ["0", "4", "0", "0", "0", "0", str(case[5] if len(case) > 5 else 1)]
    out = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[-1])
    assert r["bcfunc"] >= 1
    assert r["ia_identical"] == 1 and r["ja_identical"] == 1
    assert r["relF_A"] <= 1e-12 and r["relF_A_nonpenalty_rows"] <= 1e-12 and r["relF_rhs"] <= 1e-12
    assert r["relF_residual_rhs"] <= 1e-12
    assert out.returncode == 0


def f1(x):
    return 1.0 + x[:, 0] * x[:, 1] - 0.5 * x[:, 2]


def f3(x):
    return np.stack([0.2 * x[:, 1], -0.1 * x[:, 0] * x[:, 2], -1.0 + 0.3 * x[:, 2] ** 2], axis=1)


# every kernel family reads the host-evaluated forcing table its own way: (n, p, phys, tet, prisms, perturb, engine)
FORCING_CASES = [(5, 2, 0, 0, False, 0.12, 1),   # DMMA one-warp kernel
                 (3, 4, 0, 0, False, 0.12, 1),   # DMMA team kernel, Poisson
                 (4, 2, 1, 0, False, 0.12, 1),   # DMMA team kernel, elasticity
                 (4, 1, 1, 0, False, 0.12, 1),
                 (4, 2, 1, 0, False, 0.0, 1),    # closed-form kernel for parallelepipeds
                 (4, 2, 0, 0, False, 0.0, 1),
                 (3, 2, 1, 1, False, 0.12, 1),   # closed-form kernel for tetrahedra
                 (3, 1, 0, 1, False, 0.12, 1),
                 (2, 3, 1, 1, False, 0.12, 1),   # register-tile kernel (tetrahedra of order 3)
                 (3, 2, 1, 0, True, 0.12, 1),    # register-tile kernel (prisms)
                 (4, 2, 1, 0, False, 0.12, 0),   # register-tile kernel (engine 0)
                 (3, 2, 1, 0, False, 0.12, 2),   # generic runtime-size kernel
                 (3, 3, 0, 0, False, 0.12, 2)]


@pytest.mark.parametrize("n,p,phys,tet,prisms,perturb,engine", FORCING_CASES)
def test_domain_forcing_functions_against_oracle(n, p, phys, tet, prisms, perturb, engine):
    """Source of TPZMatPoisson / body force of TPZElasticity3D as functions of x (std::function in the reference, evaluated by
    the host at the integration points): every kernel family against the oracle fed with the same point values."""
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=bool(tet), prisms=prisms, bc_matids=(-1, -1, -1, -1, -1, -2),
                              perturb=perturb)
    mats = materials_for(phys, neumann=True)
    mats[1].SetForcingFunction(f3 if phys else f1)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, engine=engine)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


def test_plane_body_force_function_against_oracle():
    mesh = gridmesh.grid_mesh_2d(5, 3, 2, bc_matids=(-1, -1, -2, -1), perturb=0.1)
    mat = sm.TPZElasticity2D(1, gu.E_MOD, gu.NU, *gu.E2D_FORCE)
    mat.SetForcingFunction(lambda x: np.stack([0.5 + x[:, 0] * x[:, 1], -1.0 + 0.3 * x[:, 1]], axis=1))
    mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((2, 2)), np.zeros(2)), -2: mat.CreateBC(-2, 1, np.zeros((2, 2)), gu.NEUMANN_ELAST2D)}
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
