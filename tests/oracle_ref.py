"""Run the oracle (oracle/oracle.c, bit-identical to the reference on the golden fixtures) on a
neopz_b200.gridmesh.FlatMesh with neopz_b200.strmatrix materials.  Checker only."""
import os

import numpy as np

from oracle import oracle as orc

_SIMPLEX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neopz_b200", "data",
                        "simplex_rules.npz")


def _rule(topo, p):
    if topo in (orc.HEX, orc.QUAD, orc.LINE):
        return orc.rule(topo, 2 * p)
    z = np.load(_SIMPLEX)  # tables of the reference (tests/golden/make_simplex_rules.py)
    if topo == orc.PRISM:  # TPZIntPrism3D::Point (Integral/pzquad.cpp:423-436): triangle point fastest, w = w_line * w_tri
        lpts, lw = orc.rule(orc.LINE, 2 * p)
        tpts, tw = z[f"tri_order{2 * p}_pts"], z[f"tri_order{2 * p}_w"]
        pts = np.array([[t[0], t[1], zl[0]] for zl in lpts for t in tpts])
        return pts, np.array([wl * wt for wl in lw for wt in tw])
    tag = {orc.TET: "tet", orc.TRI: "tri", orc.PYR: "pyr"}[topo]
    return z[f"{tag}_order{2 * p}_pts"], z[f"{tag}_order{2 * p}_w"]


def _mat_vector(mat):
    """oracle convention (oracle.c orc_elem_t.mat) from a strmatrix material object."""
    from neopz_b200 import strmatrix as sm
    m = np.zeros(16)
    if isinstance(mat, sm.TPZMatPoisson):
        m[0], m[1] = mat.fScale, mat.force
        return orc.POISSON, 0, m
    if isinstance(mat, sm.TPZElasticity3D):
        m[0:3] = orc.elast_constants(mat.fE, mat.fPoisson)
        m[3:6] = mat.fForce
        m[6:9] = mat.fPreStress
        return orc.ELAST3D, 0, m
    if isinstance(mat, sm.TPZElasticity2D):
        m[0], m[1], m[2] = mat.fE, mat.fPoisson, 1.0 if mat.fPlaneStress else 0.0
        m[3:5] = mat.ff
        m[5:8] = mat.fPreStress
        return orc.ELAST2D, 0, m
    base = mat.material
    m[1:10] = mat.val1.reshape(-1)
    m[10:13] = mat.val2
    if isinstance(base, sm.TPZElasticity2D):
        m[0] = base.fBigNumber
        return orc.ELAST2D_BC, mat.type, m
    if isinstance(base, sm.TPZMatPoisson):
        m[0] = base.fBigNumber
        m[13] = base.fScale
        return orc.POISSON_BC, mat.type, m
    m[0] = 1.e12
    return orc.ELAST3D_BC, mat.type, m


def oracle_assemble(mesh, materials, symmetric, ia, ja):
    """Serial reference-order assembly of the whole mesh; returns (a, rhs)."""
    elems, keep, dest_parts = [], [], []
    for b in sorted(mesh.blocks, key=lambda b: b.first):
        kind, bctype, mvec = _mat_vector(materials[b.matid])
        qpts, qw = _rule(b.topology, mesh.porder)
        coords = mesh.nodes[b.elnodes]
        bcval2 = None
        mat = materials[b.matid]
        if getattr(mat, "forcing", None) is not None:
            # boundary data from a function: val2 at every integration point (data.x from the product's geometry functions:
            # the checker and the device share the points, tests/dropin checks them against the reference)
            from neopz_b200 import strmatrix as sm
            x = sm.points_x(b.topology, qpts, coords)
            v2 = np.asarray(mat.forcing(x.reshape(-1, 3)), dtype=np.float64).reshape(len(coords), len(qw), -1)
            if kind in (orc.POISSON, orc.ELAST3D, orc.ELAST2D):
                pass  # domain material: the values are the source / body force themselves
            elif bctype == 2 and kind == orc.ELAST3D_BC:  # TPZElasticity3D.cpp:646-654: val2loc = val1 * fn
                w = np.zeros((len(coords), len(qw), 3))
                for i in range(3):
                    for j in range(3):
                        w[..., i] += mat.val1[i, j] * v2[..., j]
                v2 = w
            bcval2 = np.zeros((len(coords), len(qw), 3))
            bcval2[..., : v2.shape[-1]] = v2
        outward = None
        if kind == orc.ELAST3D_BC and bctype == 4:  # the vector data.normal is oriented with (ComputeNormal)
            from neopz_b200 import gridmesh
            outward = gridmesh.face_outward(mesh, b)
        arr, k = orc.make_elems(b.topology, mesh.porder, kind, bctype, coords, mvec, qpts, qw, ids=b.elnodes, bcval2=bcval2,
                                outward=outward)
        elems.append(arr)
        keep.append(k)
        dest_parts.append(b.dest)
    ptr = [0]
    for d in dest_parts:
        nel, nd = d.shape
        ptr.extend((ptr[-1] + nd * np.arange(1, nel + 1)).tolist())
    dest = np.concatenate([d.reshape(-1) for d in dest_parts])
    return orc.assemble(symmetric, elems, np.array(ptr, dtype=np.int64), dest, ia, ja, mesh.neq)
