"""Product host code (neopz_b200/csrc/host_tables.cpp through the C ABI) against the reference fixtures:
quadrature rules bit-exact, closed-form shape tables within a few ulp of TPZShapeH1<TSHAPE>::Shape."""
import numpy as np
import pytest

from neopz_b200 import capi, strmatrix
from tests import golden_util as gu

TOPO = {"hex": capi.HEX, "tet": capi.TET, "quad": capi.QUAD, "tri": capi.TRI, "line": capi.LINE, "prism": capi.PRISM,
        "pyr": capi.PYRAMID}


@pytest.mark.parametrize("name", gu.ALL_CASES)
def test_rules_and_shape_tables(name):
    g = gu.load(name)
    p = g["meta"]["p"]
    for tag, topo in TOPO.items():
        if f"rule_{tag}_w" not in g:
            continue
        key = capi.orientation_keys(topo, g[f"shape_{tag}_ids"][None, :])[0]
        qpts, qw, phi, dphi = strmatrix.element_tables(topo, p, key)
        assert np.array_equal(qpts, g[f"rule_{tag}_pts"])  # integration points: bit-exact
        assert np.array_equal(qw, g[f"rule_{tag}_w"])
        assert phi.shape == g[f"shape_{tag}_phi"].shape
        assert np.abs(phi - g[f"shape_{tag}_phi"]).max() < 4e-16
        # (a few ulp of the largest gradient entry: 5.1 for tetrahedra of order 4)
        assert np.abs(dphi - g[f"shape_{tag}_dphi"]).max() < 1e-15 * max(1.0, np.abs(g[f"shape_{tag}_dphi"]).max())
        if f"shapeall_{tag}_phi" in g:  # p >= 3: every element of the fixture, each with its own orientation class
            keys = capi.orientation_keys(topo, g[f"shapeall_{tag}_ids"])
            pts = qpts[g[f"shapeall_{tag}_q"]]
            for e, k in enumerate(keys):
                phi, dphi = capi.shape_tables(topo, p, pts, k)
                assert np.abs(phi - g[f"shapeall_{tag}_phi"][e]).max() < 4e-16
                assert np.abs(dphi - g[f"shapeall_{tag}_dphi"][e]).max() < 2e-15


def test_gauss_legendre_orders():
    """1-D rules for every order the tensor rules can ask for, against the oracle's restatement of
    Integral/tpzgaussrule.cpp (bit-exact in double)."""
    import ctypes as C
    from oracle import oracle as orc
    fn = orc.lib().orc_gauss1d_ld
    fn.argtypes = [C.c_int, C.POINTER(C.c_longdouble), C.POINTER(C.c_longdouble)]
    for order in range(0, 21):
        loc = np.zeros(64)
        w = np.zeros(64)
        n = capi.lib().b200asm_gauss_legendre(order, capi.dptr(loc), capi.dptr(w))
        l2 = (C.c_longdouble * 64)()
        w2 = (C.c_longdouble * 64)()
        n2 = fn(order, l2, w2)
        assert n == n2
        # identical except the centre point of odd rules (a root at 0 found as +-1e-39 by either Newton)
        assert np.abs(loc[:n] - np.array([float(l2[i]) for i in range(n)])).max() < 1e-30
        assert np.array_equal(w[:n], np.array([float(w2[i]) for i in range(n)]))
