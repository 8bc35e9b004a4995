"""Helpers shared by the tests: load the committed reference fixtures (tests/golden/*.npz, written by
tests/golden/make_golden.py from the unmodified reference) and turn them into oracle inputs."""
import json
import os

import numpy as np

from oracle import oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
# prisms / hexahedra + pyramids: their GPU tests live in tests/test_gpu_prism_pyramid.py
WEDGE_CASES = [c for c in ALL_CASES if c.startswith("prism") or c.startswith("hexpyr")]
# tetrahedra / triangles of order 3, 4: tests/test_gpu_simplex_p34.py
SIMPLEX34_CASES = [c for c in ALL_CASES if c.startswith(("tet_p3", "tet_p4", "tri_p3", "tri_p4"))]
# boundary data from functions: tests/test_gpu_zz_bcfunction.py
BCFUNC_CASES = [c for c in ALL_CASES if c.endswith("_bcfunc")]
WEDGE_CASES = [c for c in WEDGE_CASES if c not in BCFUNC_CASES]
CORE_CASES = [c for c in ALL_CASES if c not in WEDGE_CASES and c not in SIMPLEX34_CASES and c not in BCFUNC_CASES]

# material data of oracle/refdriver.cpp's recipe
E_MOD, NU = 1000.0, 0.3
ELAST_FORCE = (0.0, 0.0, -1.0)
NEUMANN_POISSON = 0.75
POISSON_BC2_VAL1 = 2.5e-15   # Val1(0,0) of the type-2 condition of oracle/refdriver.cpp
NEUMANN_ELAST = (0.25, -0.5, 2.0)
BC_VAL1 = np.array([[4.0, 0.5, 0.0], [0.5, 3.0, 0.25], [0.0, 0.25, 5.0]])
BC_VAL2 = (0.3, -0.2, 0.7)
TAGS = {orc.HEX: "hex", orc.TET: "tet", orc.QUAD: "quad", orc.TRI: "tri", orc.LINE: "line", orc.PRISM: "prism", orc.PYR: "pyr"}
# TPZElasticity2D of oracle/refdriver.cpp (phys 2 = plane strain, 3 = plane stress)
E2D_FORCE = (0.5, -1.0)
NEUMANN_ELAST2D = (0.25, -0.5)
BC2D_VAL1 = np.array([[4.0, 0.5], [0.5, 3.0]])
BC2D_VAL2 = (0.3, -0.2)


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(bytes(g.pop("meta_json")).decode())
    return g


def material_vector(g, topo, matid):
    """(kind, bctype, mat[16]) in the oracle's convention for an element of this topology/material id."""
    phys = g["meta"]["phys"]
    big = float(g["meta"]["bignumber"])
    mat = np.zeros(16)
    if phys >= 2:  # TPZElasticity2D
        if matid == 1:
            mat[0], mat[1], mat[2] = E_MOD, NU, 1.0 if phys == 3 else 0.0
            mat[3:5] = E2D_FORCE
            return orc.ELAST2D, 0, mat
        bctype = 0 if matid == -1 else max(1, g["meta"]["bctype"])
        mat[0] = big
        if bctype == 1:
            mat[10:12] = NEUMANN_ELAST2D
        elif bctype >= 2:
            v1 = np.zeros((3, 3))
            v1[:2, :2] = BC2D_VAL1
            mat[1:10] = v1.reshape(-1)
            mat[10:12] = BC2D_VAL2
        return orc.ELAST2D_BC, bctype, mat
    if matid == 1:
        if phys == 0:
            mat[0], mat[1] = 1.0, 1.0
            return orc.POISSON, 0, mat
        mat[0:3] = orc.elast_constants(E_MOD, NU)
        mat[3:6] = ELAST_FORCE
        return orc.ELAST3D, 0, mat
    bctype = 0 if matid == -1 else max(1, g["meta"]["bctype"])
    mat[0] = big
    mat[13] = 1.0
    if phys == 0:
        bctype = min(bctype, 2)
        mat[10] = 0.0 if bctype == 0 else NEUMANN_POISSON
        if bctype == 2:
            mat[1] = POISSON_BC2_VAL1
        return orc.POISSON_BC, bctype, mat
    if bctype == 1:
        mat[10:13] = NEUMANN_ELAST
    elif bctype >= 2:  # refdriver's data for the other TPZElasticity3D::ContributeBC types
        mat[1:10] = BC_VAL1.reshape(-1)
        mat[10:13] = BC_VAL2
    return orc.ELAST3D_BC, bctype, mat


# the boundary functions of oracle/refdriver.cpp (bcfunc = 1), in the same arithmetic order
def bc_function_poisson_dirichlet(x):
    return 0.3 + x[0] * x[1] - 0.5 * x[2] * x[2]


def bc_function_poisson_neumann(x):
    return 0.75 + 2.0 * x[0] - x[1] * x[1]


def bc_function_elast(x):
    return (0.01 * x[1], -0.02 * x[0] * x[2], 0.005 + 0.01 * x[2])


def bc_point_values(g, topo, matid, bctype, coords, qpts):
    """[1][nq][3] val2 at the integration points of one boundary element of a bcfunc fixture (None: constant data)."""
    m = g["meta"]
    if not m.get("bcfunc") or matid == 1:
        return None
    if m["phys"] == 1 and matid == -2 and bctype != 2:
        return None
    out = np.zeros((1, len(qpts), 3))
    for q, pt in enumerate(qpts):
        x = orc.point_x(topo, coords, pt)   # data.x as the reference computes it
        if m["phys"] == 0:
            out[0, q, 0] = bc_function_poisson_dirichlet(x) if matid == -1 else bc_function_poisson_neumann(x)
        else:
            u = bc_function_elast(x)
            if matid == -2:  # mixed condition: val2loc[i] = 0 + sum_j val1(i,j) * u[j]   (TPZElasticity3D.cpp:646-654)
                for i in range(3):
                    acc = 0.0
                    for j in range(3):
                        acc += BC_VAL1[i, j] * u[j]
                    out[0, q, i] = acc
            else:
                out[0, q, :] = u
    return out


def fixture_outward(g, e):
    """Centre of boundary face e minus centre of the volume element that owns it (the vector ComputeNormal orients with)."""
    topo = int(g["el_type"][e])
    fn = set(int(v) for v in g["el_nodes"][e, : orc.TOPO_NNODE[topo]])
    fc = g["nodes"][sorted(fn)].mean(axis=0)
    for v in range(len(g["el_type"])):
        vt = int(g["el_type"][v])
        if orc.TOPO_DIM[vt] != 3:
            continue
        vn = [int(x) for x in g["el_nodes"][v, : orc.TOPO_NNODE[vt]]]
        if fn.issubset(vn):
            return fc - g["nodes"][vn].mean(axis=0)
    return np.zeros(3)


def oracle_elements(g):
    """One ctypes Elem per computational element, in element order; returns (list_of_arrays, keepalive)."""
    p = g["meta"]["p"]
    ncel = len(g["el_type"])
    arrays, keep = [], []
    for e in range(ncel):
        topo = int(g["el_type"][e])
        nn = orc.TOPO_NNODE[topo]
        nodes = g["el_nodes"][e, :nn]
        coords = g["nodes"][nodes][None, :, :]
        kind, bctype, mat = material_vector(g, topo, int(g["el_matid"][e]))
        tag = TAGS[topo]
        bcv = bc_point_values(g, topo, int(g["el_matid"][e]), bctype, coords[0], g[f"rule_{tag}_pts"])
        outward = fixture_outward(g, e)[None, :] if (kind == orc.ELAST3D_BC and bctype == 4) else None
        arr, k = orc.make_elems(topo, p, kind, bctype, coords, mat, g[f"rule_{tag}_pts"], g[f"rule_{tag}_w"],
                                ids=nodes[None, :], bcval2=bcv, outward=outward)
        arrays.append(arr)
        keep.append(k)
    return arrays, keep


def elgraph(g):
    """Element graph of Mesh/pzcmesh.cpp:1223-1267: seqnums of every element's connects."""
    ncel = len(g["el_type"])
    idx = [0]
    graph = []
    for e in range(ncel):
        nc = int(g["el_ncon"][e])
        graph.extend(g["el_conseq"][e, :nc].tolist())
        idx.append(len(graph))
    return np.array(idx, dtype=np.int64), np.array(graph, dtype=np.int64)
