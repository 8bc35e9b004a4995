"""Host logic of the multi-GPU layer (neopz_b200/csrc/multi.cpp) on the CPU: element partition by the smallest destination
equation, local numberings / patterns, staging segments and push maps.  tests/stub/multi_stub.cpp compiles multi.cpp against a
fake context that assembles synthetic integer-valued element matrices, applies the recorded push maps and compares every GPU's
owned block with the directly assembled global system (exact equality).  Meshes: hexahedra / tetrahedra with boundary faces,
grid numbering (banded) and randomly renumbered connects (all-to-all exchange), symmetric and full storage, 2..8 devices, an
equation filter (-1 destinations)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from neopz_b200 import capi, gridmesh, strmatrix as sm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_bin", "libmulti_stub.so")


@pytest.fixture(scope="module")
def stub():
    src = os.path.join(ROOT, "tests", "stub", "multi_stub.cpp")
    deps = [src, os.path.join(ROOT, "neopz_b200", "csrc", "multi.cpp"), os.path.join(ROOT, "include", "b200asm.h")]
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", src, "-o", SO])
    L = C.CDLL(SO)
    vp, ip64 = C.c_void_p, C.POINTER(C.c_int64)
    L.b200asm_multi_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
    L.b200asm_multi_destroy.argtypes = [vp]
    L.b200asm_multi_destroy.restype = None
    L.b200asm_multi_last_error.argtypes = [vp]
    L.b200asm_multi_last_error.restype = C.c_char_p
    L.b200asm_multi_add_group.argtypes = [vp, C.POINTER(capi.Group)]
    L.b200asm_multi_set_pattern.argtypes = [vp, C.c_int64, ip64, ip64, C.c_int]
    L.b200asm_multi_partition.argtypes = [vp, ip64, ip64, ip64]
    L.stub_check.argtypes = [vp, C.c_int64, ip64, ip64, C.c_int, ip64]
    return L


def renumbered(mesh, seed):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(mesh.block_size))
    size = np.empty_like(mesh.block_size)
    size[perm] = mesh.block_size
    pos = np.concatenate([[0], np.cumsum(size)[:-1]]).astype(np.int64)
    out = gridmesh.FlatMesh(porder=mesh.porder, nstate=mesh.nstate, nodes=mesh.nodes, block_pos=pos, block_size=size, neq=mesh.neq)
    for b in mesh.blocks:
        conn = perm[b.connects]
        out.blocks.append(gridmesh.ElementBlock(topology=b.topology, matid=b.matid, first=b.first, elnodes=b.elnodes, connects=conn,
                                                dest=gridmesh.destination_indices(b.topology, conn, pos, mesh.porder, mesh.nstate)))
    return out


@pytest.mark.parametrize("ndev", [2, 3, 8])
@pytest.mark.parametrize("n,p,ns,tet,symmetric,shuffle,filtered", [(5, 2, 1, False, True, 0, False), (4, 2, 3, False, True, 0, False),
                                                                   (4, 2, 1, True, False, 0, False), (5, 2, 1, False, True, 7, False),
                                                                   (3, 2, 3, True, True, 3, False), (4, 3, 1, False, False, 5, False),
                                                                   (6, 1, 3, False, False, 11, True), (8, 2, 1, False, True, 0, True)])
def test_partition_and_push_maps(stub, ndev, n, p, ns, tet, symmetric, shuffle, filtered):
    mesh = gridmesh.grid_mesh(n, p, ns, tetrahedra=tet, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.1)
    if shuffle:
        mesh = renumbered(mesh, shuffle)
    idx, graph = mesh.element_graph()
    ia, ja = capi.build_pattern(symmetric, idx, graph, mesh.block_pos, mesh.block_size, 2)
    h = C.c_void_p()
    assert stub.b200asm_multi_create(C.byref(h), ndev, None) == 0
    rng = np.random.default_rng(5)
    drop = rng.random(mesh.neq) < 0.05 if filtered else np.zeros(mesh.neq, dtype=bool)   # equations an active filter removes
    for b in mesh.blocks:
        qpts, qw, phi, dphi = sm.element_tables(b.topology, mesh.porder)
        dest = np.where(drop[b.dest], -1, b.dest)
        g, keep = capi.make_group(b.topology, mesh.porder, 0, ns, b.elnodes, dest, qpts, qw, phi, dphi, [1.0])
        assert stub.b200asm_multi_add_group(h, C.byref(g)) >= 0
    rc = stub.b200asm_multi_set_pattern(h, mesh.neq, capi.i64ptr(ia), capi.i64ptr(ja), int(symmetric))
    assert rc == 0, stub.b200asm_multi_last_error(h).decode()
    rb, nel, st = np.zeros(ndev + 1, np.int64), np.zeros(ndev, np.int64), np.zeros(ndev, np.int64)
    assert stub.b200asm_multi_partition(h, capi.i64ptr(rb), capi.i64ptr(nel), capi.i64ptr(st)) == 0
    assert rb[0] == 0 and rb[-1] == mesh.neq and np.all(np.diff(rb) >= 0) and nel.sum() == mesh.nelements
    if not shuffle and ndev <= 3:   # banded numbering: balanced parts
        assert nel.min() > 0.5 * nel.max(), nel
    detail = np.zeros(8, dtype=np.int64)
    rc = stub.stub_check(h, mesh.neq, capi.i64ptr(ia), capi.i64ptr(ja), int(symmetric), capi.i64ptr(detail))
    assert rc == 0, (rc, detail.tolist())
    assert detail[4] >= ndev - 1 and detail[5] > 0     # links exist and entries travel
    stub.b200asm_multi_destroy(h)
