"""The sum-factorisation kernels for hexahedra of order 2 (neopz_b200/csrc/sumfact_hex.cuh; variant 0 / 20 = the default: one warp
per element, no block-wide barrier; 13 = one CTA of 64 threads per element, one barrier per (e,f)) and the one-warp DMMA Gram kernel
they replaced as the default (variant 16) against the oracle: both storages, coloured scatter, load vector only, forcing table.
(The same checks as tools/sumfact_check.py, whose B200 output is profiles/r01_sumfact_check.jsonl.)"""
import numpy as np
import pytest

from neopz_b200 import gridmesh, strmatrix as sm
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, relF

pytestmark = pytest.mark.gpu


def _mats(forcing=None):
    m = sm.TPZMatPoisson(1, 3)
    m.SetScaleFactor(1.7)
    m.SetForcingFunction(forcing if forcing else 1.0)
    return {1: m, -1: m.CreateBC(-1, 0, [[0.0]], [0.0]), -2: m.CreateBC(-2, 1, [[0.0]], [0.75])}


@pytest.mark.parametrize("variant", [0, 13, 16, 20])
@pytest.mark.parametrize("n,symmetric,scatter,forcing", [(5, True, "atomic", False), (4, False, "atomic", False), (5, True, "colored", False),
                                                         (7, True, "atomic", True)])
def test_sumfact_variants_against_oracle(variant, n, symmetric, scatter, forcing):
    mesh = gridmesh.grid_mesh(n, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    mats = _mats((lambda x: 1.0 + x[:, 0] * x[:, 1] - 0.5 * x[:, 2]) if forcing else None)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, variant=variant, scatter=scatter)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    a2, rhs2 = strmat.Assemble()
    assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
    assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL


def _elast_mats(prestress=False, forcing=None):
    m = sm.TPZElasticity3D(1, 1000.0, 0.3, (0.3, -0.2, -1.0), prestress=(1.5, -0.5, 0.25) if prestress else (0.0, 0.0, 0.0))
    if forcing:
        m.SetForcingFunction(forcing)
    return {1: m, -1: m.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3)), -2: m.CreateBC(-2, 1, np.zeros((3, 3)), np.array([0.1, 0.2, -0.3]))}


@pytest.mark.parametrize("variant", [0, 30, 31, 34])
@pytest.mark.parametrize("n,symmetric,scatter,prestress,forcing", [(4, True, "atomic", False, False), (3, False, "atomic", False, False),
                                                                   (4, True, "colored", False, False), (3, True, "atomic", True, False),
                                                                   (4, True, "atomic", False, True)])
def test_hex_p2_elasticity_variants_against_oracle(variant, n, symmetric, scatter, prestress, forcing):
    """Hexahedra p2 TPZElasticity3D (gram_mma_team.cuh): 0 / 31 = a pair of warps per element on a whole-element panel (the default),
    30 = one warp per element, 34 = the team of ten warps (the default of round 1); both storages, coloured scatter, prestress,
    forcing table, re-assembly and the load-vector-only pass."""
    mesh = gridmesh.grid_mesh(n, 2, 3, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
    force = (lambda x: np.stack([1.0 + x[:, 0], x[:, 1] * x[:, 2], -0.5 + x[:, 2]], axis=1)) if forcing else None
    mats = _elast_mats(prestress, force)
    strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, variant=variant, scatter=scatter)
    ia, ja, a, rhs = strmat.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    a2, rhs2 = strmat.Assemble()
    assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
    assert relF(strmat.AssembleRhs(), rhs_ref) <= TOL
    strmat.ctx.close()
