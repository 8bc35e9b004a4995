"""GPU tests written at the very end of round 1 (oracle against the device on generated hexahedron + pyramid meshes); green on
B200 since the first pass of round 2 (profiles/r02_late_tests.log).  The file name keeps their place late in the suite."""
import numpy as np
import pytest

from neopz_b200 import gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, materials_for, relF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,phys", [(4, 2, 0), (5, 1, 1), (3, 2, 1)])
def test_hexpyr_generated_meshes_against_oracle(n, p, phys):
    """Hexahedra + pyramids of other sizes (gridmesh.hexpyr_mesh replays the reference's generator, slot reuse included)."""
    mesh = gridmesh.hexpyr_mesh(n, p, 3 if phys else 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = materials_for(phys, neumann=True)
    for symmetric in (True, False):
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = strmat.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL


@pytest.mark.parametrize("name", gu.BCFUNC_CASES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_against_reference_fixtures(name, symmetric):
    """Fixtures of the unmodified reference with boundary data from functions (Dirichlet + Neumann for Poisson, Dirichlet + mixed
    for Elasticity3D): pattern bit-exact, values and load vector within 1e-12."""
    from tests.test_gpu_parity import fixture_setup
    g = gu.load(name)
    mesh, mats = fixture_setup(g)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    ia, ja, a, rhs = strmat.CreateAssemble()
    pre = "sym" if symmetric else "full"
    assert np.array_equal(ia, g[pre + "_ia"]) and np.array_equal(ja, g[pre + "_ja"])
    assert relF(a, g[pre + "_a"]) <= TOL and relF(rhs, g["rhs"]) <= TOL
