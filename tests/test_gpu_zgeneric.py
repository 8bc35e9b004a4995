"""The generic runtime-size volume kernel (assemble_volume_generic_kernel: the kernel of every configuration without a
specialised one - higher orders, elements whose sides carry different orders) forced onto configurations that the
fixtures of the unmodified reference and the oracle cover (option engine = 2), and the drop-in strategy on meshes only
it can run: hexahedra of order 5, prisms / pyramids of order 3, p-refined (non-uniform order) meshes."""
import json
import os
import subprocess

import numpy as np
import pytest

from neopz_b200 import gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, fixture_setup, materials_for, relF

pytestmark = pytest.mark.gpu

CASES = ["hex_p1_poisson_n3_pert", "hex_p2_elast_n2_pert", "hex_p3_poisson_n2_pert_scr", "hex_p3_elast_n2_pert_scr",
         "hex_p4_poisson_n2_pert_scr", "tet_p2_elast_n2_pert", "tet_p4_poisson_n2_pert_scr", "tet_p3_elast_n2_pert_scr",
         "prism_p2_elast_n2_pert", "hexpyr_p2_poisson_n2_pert", "hexpyr_p2_elast_n2_pert", "hex_p2_elast_n2_bc3"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_generic_kernel_against_reference_fixtures(name, symmetric):
    g = gu.load(name)
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=2)
    ia, ja, a, rhs = strmat.CreateAssemble()
    pre = "sym" if symmetric else "full"
    assert np.array_equal(ia, g[pre + "_ia"]) and np.array_equal(ja, g[pre + "_ja"])
    assert relF(a, g[pre + "_a"]) <= TOL and relF(rhs, g["rhs"]) <= TOL
    assert relF(strmat.AssembleRhs(), g["rhs"]) <= TOL


@pytest.mark.parametrize("scatter", ["atomic", "colored"])
def test_generic_kernel_against_oracle_ragged(scatter):
    """More elements than CTAs of the persistent grid, both scatter modes, re-assembly."""
    mesh = gridmesh.grid_mesh(11, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)   # 1331 hexahedra > 148 x 8
    mats = materials_for(0, neumann=True)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, engine=2, scatter=scatter)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    a2, rhs2 = strmat.Assemble()
    assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL


# n, p, phys, tet, symmetric, solve [, prefine]: configurations only the generic kernel runs
DROPIN_CASES = [(2, 5, 0, 0, 1, 1, 0), (2, 3, 0, 2, 1, 1, 0), (2, 3, 1, 3, 1, 0, 0), (2, 3, 1, 2, 0, 0, 0),
                (3, 2, 0, 0, 1, 1, 1), (3, 2, 1, 0, 1, 1, 1), (2, 2, 0, 1, 1, 1, 1), (3, 1, 1, 2, 0, 0, 1), (3, 2, 0, 3, 1, 1, 1)]


@pytest.mark.parametrize("case", DROPIN_CASES)
def test_dropin_strategy_matches_reference(case):
    """prefine = 1: every third volume element is p-refined by one order (TPZInterpolatedElement::PRefine), so elements,
    faces and edges of different order meet (tests/dropin/dropin_test.cpp)."""
    from tests.test_gpu_dropin import BIN
    if not os.path.exists(BIN):
        pytest.skip("tests/_bin/dropin_test not built (needs /root/reference at build time)")
    device_create = DROPIN_CASES.index(case) % 2
    args = [str(x) for x in case[:6]] + ["4", str(device_create), "0", "0", str(case[6])]
    out = subprocess.run([BIN] + args, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[-1])
    assert r["ia_identical"] == 1 and r["ja_identical"] == 1
    assert r["relF_A"] <= 1e-12 and r["relF_A_nonpenalty_rows"] <= 1e-12 and r["relF_rhs"] <= 1e-12
    assert r["relF_cg_solution"] <= 1e-10 and r["relF_device_cg_solution"] <= 1e-10
    assert r["relF_residual_rhs"] <= 1e-12
    assert out.returncode == 0
