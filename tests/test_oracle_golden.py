"""Pins the oracle (oracle/oracle.c) against the reference: fixtures produced by running the
unmodified NeoPZ (tests/golden/make_golden.py) and the reference's own known-answer vector."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import golden_util as gu


@pytest.mark.parametrize("name", gu.ALL_CASES)
def test_rules_and_shapes(name):
    g = gu.load(name)
    p = g["meta"]["p"]
    for topo, tag in gu.TAGS.items():
        if f"rule_{tag}_w" not in g:
            continue
        if topo in (orc.HEX, orc.QUAD):
            pts, w = orc.rule(topo, 2 * p)  # Mesh/pzelctemp.cpp:35-47: order 2p
            assert np.array_equal(pts, g[f"rule_{tag}_pts"])  # bit-exact
            assert np.array_equal(w, g[f"rule_{tag}_w"])
        for q, pt in enumerate(g[f"rule_{tag}_pts"]):
            phi, dphi = orc.shape(topo, p, pt, g[f"shape_{tag}_ids"])
            assert np.array_equal(phi, g[f"shape_{tag}_phi"][q])  # same arithmetic order -> bit-exact
            assert np.array_equal(dphi, g[f"shape_{tag}_dphi"][q])
        if f"shapeall_{tag}_phi" in g:  # p >= 3: every element (its own side orientations) at a few points
            for e, ids in enumerate(g[f"shapeall_{tag}_ids"]):
                for k, q in enumerate(g[f"shapeall_{tag}_q"]):
                    phi, dphi = orc.shape(topo, p, g[f"rule_{tag}_pts"][q], ids)
                    assert np.array_equal(phi, g[f"shapeall_{tag}_phi"][e, k])
                    assert np.array_equal(dphi, g[f"shapeall_{tag}_dphi"][e, k])


@pytest.mark.parametrize("name", gu.ALL_CASES)
def test_element_matrices(name):
    g = gu.load(name)
    if "ek" not in g:
        pytest.skip("fixture without element matrices")
    elems, _keep = gu.oracle_elements(g)
    worst = 0.0
    for e, arr in enumerate(elems):
        nd = int(g["el_dest_ptr"][e + 1] - g["el_dest_ptr"][e])
        ek, ef = orc.calcstiff(arr[0], nd)
        ref = g["ek"][g["ek_ptr"][e]:g["ek_ptr"][e + 1]].reshape(nd, nd).T  # stored column-major
        scale = max(np.abs(ref).max(), 1e-300)  # Neumann faces have ek == 0
        worst = max(worst, np.abs(ek - ref).max() / scale)
        ef_ref = g["ef"][g["el_dest_ptr"][e]:g["el_dest_ptr"][e + 1]]
        assert np.abs(ef - ef_ref).max() <= 5e-15 * max(np.abs(ef_ref).max(), 1e-300)
    assert worst == 0.0, worst  # the restatement is bit-identical to the reference's CalcStiff


@pytest.mark.parametrize("name", gu.ALL_CASES)
def test_pattern_bit_exact(name):
    g = gu.load(name)
    idx, graph = gu.elgraph(g)
    ia, ja = orc.pattern(True, idx, graph, g["block_pos"], g["block_size"])
    assert np.array_equal(ia, g["sym_ia"]) and np.array_equal(ja, g["sym_ja"])
    ia, ja = orc.pattern(False, idx, graph, g["block_pos"], g["block_size"])
    assert np.array_equal(ia, g["full_ia"]) and np.array_equal(ja, g["full_ja"])


@pytest.mark.parametrize("name", gu.ALL_CASES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_assembly(name, symmetric):
    g = gu.load(name)
    elems, _keep = gu.oracle_elements(g)
    pre = "sym" if symmetric else "full"
    a, rhs = orc.assemble(symmetric, elems, g["el_dest_ptr"], g["el_dest"], g[pre + "_ia"], g[pre + "_ja"],
                          g["meta"]["neq"])
    ref = g[pre + "_a"]
    # same summation order as the reference's serial loop: agreement to a few ulps of each entry's row scale
    assert np.array_equal(a, ref)  # bit-identical to the reference's serial assembly
    assert np.array_equal(rhs, g["rhs"])
    nz = np.abs(ref) > 1e-9 * np.abs(ref).max()
    assert (np.abs(a - ref)[nz] / np.abs(ref)[nz]).max() < 1e-9
    assert np.linalg.norm(rhs - g["rhs"]) / np.linalg.norm(g["rhs"]) < 1e-15


def test_elasticity3d_known_answer():
    """UnitTest_PZ/TestMaterial/TestMaterial.cpp:18-40,68-86 + CubeStiffMatrix.txt: E=1000, nu=0.2,
    dphix = I3, weight 8 -> 9x9 matrix, margin 0.01, symmetric to 1e-8."""
    right = np.array([
        [8888.89, 0, 0, 0, 2222.22, 0, 0, 0, 2222.22],
        [0, 3333.33, 0, 3333.33, 0, 0, 0, 0, 0],
        [0, 0, 3333.33, 0, 0, 0, 3333.33, 0, 0],
        [0, 3333.33, 0, 3333.33, 0, 0, 0, 0, 0],
        [2222.22, 0, 0, 0, 8888.89, 0, 0, 0, 2222.22],
        [0, 0, 0, 0, 0, 3333.33, 0, 3333.33, 0],
        [0, 0, 3333.33, 0, 0, 0, 3333.33, 0, 0],
        [0, 0, 0, 0, 0, 3333.33, 0, 3333.33, 0],
        [2222.22, 0, 0, 0, 2222.22, 0, 0, 0, 8888.89]])
    mat = np.zeros(16)
    mat[0:3] = orc.elast_constants(1000.0, 0.2)
    phi = np.zeros(3)
    dphix = np.eye(3).reshape(-1).copy()  # dphix[d*n + i]
    ek = np.zeros(81)
    ef = np.zeros(9)
    orc.lib().orc_elast_contribute_point(3, orc._dp(phi), orc._dp(dphix), 8.0, orc._dp(mat), orc._dp(ek), orc._dp(ef))
    ek = ek.reshape(9, 9).T
    assert np.abs(ek - ek.T).max() < 1e-8
    assert np.abs(ek - right).max() < 0.01
