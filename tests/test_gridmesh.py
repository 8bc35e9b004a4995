"""neopz_b200.gridmesh must reproduce the reference's mesh, connect numbering, block table and
destination indices (fixtures dumped from the unmodified NeoPZ)."""
import numpy as np
import pytest

from neopz_b200 import capi, gridmesh
from tests import golden_util as gu


def _mesh_for(g):
    m = g["meta"]
    perm = g["node_perm"] if m.get("scramble") else None
    if m.get("dim", 3) == 2:
        bc = (-1, -1, -2 if m["bctype"] >= 1 else -1, -1)
        return gridmesh.grid_mesh_2d(m["n"], m["p"], 2 if m["phys"] >= 2 else 1, triangles=bool(m["tet"]), bc_matids=bc,
                                     perturb=m["perturb"], node_perm=perm)
    bc = (-1, -1, -1, -1, -1, -2 if m["bctype"] >= 1 else -1)
    if m["tet"] == 3:
        # hexahedra + pyramids (MMeshType::EHexaPyrMixed): the reference creates the pyramids by refinement and reuses the
        # slots of the deleted hexahedra, so element types interleave (gridmesh.hexpyr_elements replays that allocation)
        return gridmesh.hexpyr_mesh(m["n"], m["p"], 3 if m["phys"] == 1 else 1, bc_matids=bc, perturb=m["perturb"])
    return gridmesh.grid_mesh(m["n"], m["p"], 3 if m["phys"] == 1 else 1, tetrahedra=m["tet"] == 1, prisms=m["tet"] == 2,
                              bc_matids=bc, perturb=m["perturb"], node_perm=perm)


@pytest.mark.parametrize("name", gu.ALL_CASES)
def test_grid_mesh_matches_reference(name):
    g = gu.load(name)
    mesh = _mesh_for(g)
    assert np.array_equal(mesh.nodes, g["nodes"])  # bit-exact coordinates (same expression as TPZGenGrid3D)
    assert mesh.neq == g["meta"]["neq"]
    assert np.array_equal(mesh.block_pos, g["block_pos"][: len(mesh.block_pos)])
    assert np.array_equal(mesh.block_size, g["block_size"][: len(mesh.block_size)])
    e = 0
    for b in mesh.blocks:
        nel, nc = b.elnodes.shape
        if b.index is None:
            assert b.first == e
            idx = np.arange(e, e + nel)
        else:
            idx = b.index
        assert np.all(g["el_type"][idx] == b.topology)
        assert np.all(g["el_matid"][idx] == b.matid)
        assert np.array_equal(g["el_nodes"][idx, :nc], b.elnodes)
        ns = b.connects.shape[1]
        assert np.array_equal(g["el_conseq"][idx, :ns], b.connects)
        for k, el in enumerate(idx):
            lo, hi = g["el_dest_ptr"][el], g["el_dest_ptr"][el + 1]
            assert np.array_equal(g["el_dest"][lo:hi], b.dest[k])
        e += nel
    assert e == g["meta"]["ncel"]


@pytest.mark.parametrize("name", gu.ALL_CASES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_host_pattern_builder_bit_exact(name, symmetric):
    """b200asm_build_pattern == TPZSSpStructMatrix::Create / TPZSpStructMatrix::Create of the reference."""
    g = gu.load(name)
    mesh = _mesh_for(g)
    idx, graph = mesh.element_graph()
    ia, ja = capi.build_pattern(symmetric, idx, graph, mesh.block_pos, mesh.block_size, nthreads=3)
    pre = "sym" if symmetric else "full"
    assert np.array_equal(ia, g[pre + "_ia"])
    assert np.array_equal(ja, g[pre + "_ja"])
