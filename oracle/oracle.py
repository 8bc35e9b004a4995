"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle.so (the C restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Nothing under neopz_b200/ does (tests/test_no_oracle_in_product.py enforces it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

HEX, TET, QUAD, TRI, LINE, PRISM, PYR = 0, 1, 2, 3, 4, 5, 6
POISSON, ELAST3D, POISSON_BC, ELAST3D_BC, ELAST2D, ELAST2D_BC = 0, 1, 2, 3, 4, 5
TOPO_DIM = {HEX: 3, TET: 3, QUAD: 2, TRI: 2, LINE: 1, PRISM: 3, PYR: 3}
TOPO_NNODE = {HEX: 8, TET: 4, QUAD: 4, TRI: 3, LINE: 2, PRISM: 6, PYR: 5}


def build():
    """Compile the C oracle (gcc only; a second or two)."""
    src = os.path.join(_HERE, "oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


class Elem(C.Structure):
    _fields_ = [("topo", C.c_int32), ("p", C.c_int32), ("kind", C.c_int32), ("bctype", C.c_int32),
                ("coords", C.c_double * 24), ("mat", C.c_double * 16),
                ("nq", C.c_int32), ("pad", C.c_int32),
                ("qpts", C.POINTER(C.c_double)), ("qw", C.POINTER(C.c_double)),
                ("ids", C.c_int64 * 8),
                ("bcval2", C.POINTER(C.c_double)), ("outward", C.c_double * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int64)
        _lib.orc_rule_hex.argtypes = [C.c_int, dp, dp]
        _lib.orc_rule_quad.argtypes = [C.c_int, dp, dp]
        _lib.orc_rule_line.argtypes = [C.c_int, dp, dp]
        _lib.orc_shape.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        _lib.orc_shape_ids.argtypes = [C.c_int, C.c_int, ip, dp, dp, dp]
        _lib.orc_calcstiff.argtypes = [C.POINTER(Elem), dp, dp]
        _lib.orc_point_x.argtypes = [C.c_int, dp, dp, dp]
        _lib.orc_elast_contribute_point.argtypes = [C.c_int, dp, dp, C.c_double, dp, dp, dp]
        _lib.orc_elast_constants.argtypes = [C.c_double, C.c_double, dp]
        _lib.orc_pattern.argtypes = [C.c_int, C.c_int64, ip, ip, C.c_int64, ip, ip, ip, ip]
        _lib.orc_pattern.restype = C.c_int64
        _lib.orc_addkel.argtypes = [C.c_int, ip, ip, dp, C.c_int, dp, ip]
        _lib.orc_addkel.restype = C.c_int64
        _lib.orc_assemble.argtypes = [C.c_int, C.c_int64, C.POINTER(Elem), ip, ip, ip, ip, dp, dp]
        _lib.orc_assemble.restype = C.c_int64
        assert _lib.orc_sizeof_elem() == C.sizeof(Elem)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def rule(topo, order):
    """Tensor Gauss-Legendre rule of the reference for hex / quad (order = 2p)."""
    dim = TOPO_DIM[topo]
    pts = np.zeros((4096, dim))
    w = np.zeros(4096)
    fn = {HEX: lib().orc_rule_hex, QUAD: lib().orc_rule_quad, LINE: lib().orc_rule_line}[topo]
    n = fn(order, _dp(pts), _dp(w))
    return pts[:n].copy(), w[:n].copy()


def shape(topo, p, pt, ids=None):
    """phi[n], dphi[dim][n] at one master-element point; ids = global corner-node indices (needed for p >= 3)."""
    dim = TOPO_DIM[topo]
    phi = np.zeros(343)
    dphi = np.zeros(3 * 343)
    pt = np.ascontiguousarray(pt, dtype=np.float64)
    if ids is None:
        n = lib().orc_shape(topo, p, _dp(pt), _dp(phi), _dp(dphi))
    else:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n = lib().orc_shape_ids(topo, p, _ip(ids), _dp(pt), _dp(phi), _dp(dphi))
    assert n > 0
    return phi[:n].copy(), dphi[: dim * n].reshape(dim, n).copy()


def point_x(topo, coords, pt):
    """data.x of a boundary face (quadrilateral / triangle) at a master-element point, in the reference's arithmetic."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    pt = np.ascontiguousarray(pt, dtype=np.float64)
    x = np.zeros(3)
    assert lib().orc_point_x(topo, _dp(coords), _dp(pt), _dp(x)) == 0
    return x


def elast_constants(E, nu):
    c = np.zeros(3)
    lib().orc_elast_constants(E, nu, _dp(c))
    return c


def make_elems(topo, p, kind, bctype, coords, mat, qpts, qw, ids=None, bcval2=None, outward=None):
    """coords: (nel, nnode, 3); ids: (nel, nnode) global corner-node indices (orientation, p >= 3); bcval2: (nel, nq, 3) val2 of
    a boundary condition with a forcing function at every integration point.  Returns (ctypes array of Elem, keepalive)."""
    nel = coords.shape[0]
    nn = TOPO_NNODE[topo]
    qpts = np.ascontiguousarray(qpts, dtype=np.float64)
    qw = np.ascontiguousarray(qw, dtype=np.float64)
    arr = (Elem * nel)()
    if bcval2 is not None:
        bcval2 = np.ascontiguousarray(bcval2, dtype=np.float64)
        assert bcval2.shape == (nel, len(qw), 3)
    for e in range(nel):
        el = arr[e]
        el.topo, el.p, el.kind, el.bctype = topo, p, kind, bctype
        if bcval2 is not None:
            el.bcval2 = _dp(bcval2[e])
        if outward is not None:  # (nel, 3): face centre minus centre of the neighbouring volume element (boundary type 4)
            el.outward[:] = [float(v) for v in outward[e]]
        flat = np.zeros(24)
        flat[: nn * 3] = coords[e].reshape(-1)
        el.coords[:] = flat.tolist()
        m = np.zeros(16)
        m[: len(mat)] = mat
        el.mat[:] = m.tolist()
        if ids is not None:
            el.ids[:nn] = [int(v) for v in ids[e][:nn]]
        el.nq = len(qw)
        el.qpts = _dp(qpts)
        el.qw = _dp(qw)
    return arr, (qpts, qw, bcval2)


def calcstiff(elem, ndof):
    ek = np.zeros(ndof * ndof)
    ef = np.zeros(ndof)
    nd = lib().orc_calcstiff(C.byref(elem), _dp(ek), _dp(ef))
    assert nd == ndof, (nd, ndof)
    return ek.reshape(ndof, ndof).T.copy(), ef  # ek[i, j]


def pattern(symmetric, elgraphindex, elgraph, blockpos, blocksize):
    elgraphindex = np.ascontiguousarray(elgraphindex, dtype=np.int64)
    elgraph = np.ascontiguousarray(elgraph, dtype=np.int64)
    blockpos = np.ascontiguousarray(blockpos, dtype=np.int64)
    blocksize = np.ascontiguousarray(blocksize, dtype=np.int64)
    neq = int((blocksize).sum())
    ia = np.zeros(neq + 1, dtype=np.int64)
    nel = len(elgraphindex) - 1
    nnz = lib().orc_pattern(int(symmetric), nel, _ip(elgraphindex), _ip(elgraph), len(blockpos),
                            _ip(blockpos), _ip(blocksize), _ip(ia), None)
    ja = np.zeros(nnz, dtype=np.int64)
    nnz2 = lib().orc_pattern(int(symmetric), nel, _ip(elgraphindex), _ip(elgraph), len(blockpos),
                             _ip(blockpos), _ip(blocksize), _ip(ia), _ip(ja))
    assert nnz == nnz2
    return ia, ja


def assemble(symmetric, elems_list, dest_ptr, dest, ia, ja, neq):
    """elems_list: list of (ctypes Elem array) in ELEMENT ORDER (concatenated here)."""
    total = sum(len(a) for a in elems_list)
    allel = (Elem * total)()
    k = 0
    for a in elems_list:
        C.memmove(C.byref(allel, k * C.sizeof(Elem)), a, len(a) * C.sizeof(Elem))
        k += len(a)
    a_val = np.zeros(len(ja))
    rhs = np.zeros(neq)
    dest_ptr = np.ascontiguousarray(dest_ptr, dtype=np.int64)
    dest = np.ascontiguousarray(dest, dtype=np.int64)
    missing = lib().orc_assemble(int(symmetric), total, allel, _ip(dest_ptr), _ip(dest), _ip(ia), _ip(ja),
                                 _dp(a_val), _dp(rhs))
    assert missing == 0, missing
    return a_val, rhs
