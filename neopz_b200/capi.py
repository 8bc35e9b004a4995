"""ctypes binding of libb200asm.so (C ABI declared in include/b200asm.h).

The library is built in-tree by __graft_entry__.build() / neopz_b200/build.py.  There is no Python
or CPU fallback: if the shared library is missing or no CUDA device is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200asm.so")

HEX, TET, QUAD, TRI, LINE, PRISM, PYRAMID = 0, 1, 2, 3, 4, 5, 6
POISSON, ELASTICITY3D, BC, ELASTICITY2D = 0, 1, 2, 3
ENODEVICE = -2


class B200AsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200asm error {code}: {msg}")
        self.code = code


class Group(C.Structure):
    _fields_ = [("topology", C.c_int32), ("porder", C.c_int32), ("kind", C.c_int32), ("nstate", C.c_int32),
                ("nel", C.c_int64),
                ("elnodes", C.POINTER(C.c_int32)), ("dest", C.POINTER(C.c_int64)),
                ("nqp", C.c_int32), ("nshape", C.c_int32),
                ("qpts", C.POINTER(C.c_double)), ("qwts", C.POINTER(C.c_double)),
                ("phi", C.POINTER(C.c_double)), ("dphi", C.POINTER(C.c_double)),
                ("coef", C.c_double * 16),
                ("force", C.POINTER(C.c_double))]


class IpcMem(C.Structure):
    """b200asm_ipc_mem: a cudaIpcMemHandle_t and the offset of the array inside the allocation it names."""
    _fields_ = [("handle", C.c_ubyte * 64), ("offset", C.c_int64)]


_lib = None

# every symbol include/b200asm.h declares (tests/test_capi_symbols.py checks the header against this)
SYMBOLS = [
    "b200asm_create", "b200asm_destroy", "b200asm_last_error", "b200asm_set_stream", "b200asm_set_option",
    "b200asm_set_nodes", "b200asm_add_group", "b200asm_set_group_coef", "b200asm_clear_groups",
    "b200asm_set_pattern", "b200asm_assemble", "b200asm_assemble_async", "b200asm_synchronize",
    "b200asm_download", "b200asm_device_pointers", "b200asm_counters", "b200asm_scatter_add", "b200asm_group_time_ms",
    "b200asm_gauss_legendre", "b200asm_tensor_rule", "b200asm_shape_tables", "b200asm_build_pattern",
    "b200asm_nshape", "b200asm_orientation_keys", "b200asm_shape_tables_oriented",
    "b200asm_build_pattern_device", "b200asm_get_pattern", "b200asm_cg_solve", "b200asm_cg_solution_device",
    "b200asm_assemble_rhs", "b200asm_get_ja_range", "b200asm_pin_host", "b200asm_unpin_host", "b200asm_prism_rule",
    "b200asm_exchange_export", "b200asm_exchange_add_peer", "b200asm_exchange_set_map", "b200asm_exchange_clear",
    "b200asm_set_group_force",
    "b200asm_multi_create", "b200asm_multi_destroy", "b200asm_multi_last_error", "b200asm_multi_num_devices", "b200asm_multi_context",
    "b200asm_multi_set_option", "b200asm_multi_set_nodes", "b200asm_multi_add_group", "b200asm_multi_set_group_coef",
    "b200asm_multi_set_group_force", "b200asm_multi_clear_groups", "b200asm_multi_set_pattern", "b200asm_multi_partition",
    "b200asm_multi_assemble", "b200asm_multi_assemble_rhs", "b200asm_multi_assemble_async", "b200asm_multi_synchronize",
    "b200asm_multi_counters", "b200asm_device_count", "b200asm_multi_cg_solve", "b200asm_group_kernel",
]


def lib():
    """Load the native library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). neopz_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip64, ip32 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    L.b200asm_device_count.argtypes = []
    L.b200asm_create.argtypes = [C.POINTER(vp), C.c_int]
    L.b200asm_destroy.argtypes = [vp]
    L.b200asm_destroy.restype = None
    L.b200asm_last_error.argtypes = [vp]
    L.b200asm_last_error.restype = C.c_char_p
    L.b200asm_set_stream.argtypes = [vp, vp]
    L.b200asm_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.b200asm_set_nodes.argtypes = [vp, C.c_int64, dp]
    L.b200asm_add_group.argtypes = [vp, C.POINTER(Group)]
    L.b200asm_set_group_coef.argtypes = [vp, C.c_int, dp]
    L.b200asm_clear_groups.argtypes = [vp]
    L.b200asm_set_pattern.argtypes = [vp, C.c_int64, ip64, ip64, C.c_int]
    L.b200asm_assemble.argtypes = [vp, dp, dp]
    L.b200asm_assemble_async.argtypes = [vp]
    L.b200asm_assemble_rhs.argtypes = [vp, dp]
    L.b200asm_pin_host.argtypes = [vp, C.c_void_p, C.c_size_t]
    L.b200asm_unpin_host.argtypes = [vp, C.c_void_p]
    L.b200asm_synchronize.argtypes = [vp]
    L.b200asm_download.argtypes = [vp, dp, dp]
    L.b200asm_device_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.b200asm_counters.argtypes = [vp, ip64, ip64, ip64]
    L.b200asm_group_time_ms.argtypes = [vp, C.c_int, dp]
    L.b200asm_group_kernel.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.b200asm_scatter_add.argtypes = [vp, C.c_int, vp, vp, C.c_int64]
    L.b200asm_gauss_legendre.argtypes = [C.c_int, dp, dp]
    L.b200asm_tensor_rule.argtypes = [C.c_int, C.c_int, dp, dp]
    L.b200asm_shape_tables.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp]
    L.b200asm_prism_rule.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
    L.b200asm_build_pattern_device.argtypes = [vp, C.c_int, C.c_int64, ip64, ip64, C.c_int64, ip64, ip64, ip64, ip64]
    L.b200asm_get_pattern.argtypes = [vp, ip64, ip64]
    L.b200asm_get_ja_range.argtypes = [vp, C.c_int64, C.c_int64, ip64]
    L.b200asm_cg_solve.argtypes = [vp, C.c_int, C.c_int64, C.c_double, C.c_int, dp, dp, ip64, dp]
    L.b200asm_cg_solution_device.argtypes = [vp, C.POINTER(vp)]
    L.b200asm_nshape.argtypes = [C.c_int, C.c_int]
    L.b200asm_orientation_keys.argtypes = [C.c_int, C.c_int64, ip32, ip64]
    L.b200asm_shape_tables_oriented.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, dp, dp, dp]
    L.b200asm_build_pattern.argtypes = [C.c_int, C.c_int64, ip64, ip64, C.c_int64, ip64, ip64, ip64, ip64, C.c_int]
    L.b200asm_build_pattern.restype = C.c_int64
    L.b200asm_exchange_export.argtypes = [vp, C.POINTER(IpcMem)]
    L.b200asm_exchange_add_peer.argtypes = [vp, C.c_int, C.c_int, C.POINTER(IpcMem), vp, C.c_int64]
    L.b200asm_exchange_set_map.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, ip32, C.c_int64, ip32, ip32]
    L.b200asm_exchange_clear.argtypes = [vp]
    L.b200asm_set_group_force.argtypes = [vp, C.c_int, dp]
    L.b200asm_multi_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
    L.b200asm_multi_destroy.argtypes = [vp]
    L.b200asm_multi_destroy.restype = None
    L.b200asm_multi_last_error.argtypes = [vp]
    L.b200asm_multi_last_error.restype = C.c_char_p
    L.b200asm_multi_num_devices.argtypes = [vp]
    L.b200asm_multi_context.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.b200asm_multi_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.b200asm_multi_set_nodes.argtypes = [vp, C.c_int64, dp]
    L.b200asm_multi_add_group.argtypes = [vp, C.POINTER(Group)]
    L.b200asm_multi_set_group_coef.argtypes = [vp, C.c_int, dp]
    L.b200asm_multi_set_group_force.argtypes = [vp, C.c_int, dp]
    L.b200asm_multi_clear_groups.argtypes = [vp]
    L.b200asm_multi_set_pattern.argtypes = [vp, C.c_int64, ip64, ip64, C.c_int]
    L.b200asm_multi_partition.argtypes = [vp, ip64, ip64, ip64]
    L.b200asm_multi_assemble.argtypes = [vp, dp, dp]
    L.b200asm_multi_assemble_rhs.argtypes = [vp, dp]
    L.b200asm_multi_assemble_async.argtypes = [vp]
    L.b200asm_multi_synchronize.argtypes = [vp]
    L.b200asm_multi_counters.argtypes = [vp, ip64, ip64, ip64]
    L.b200asm_multi_cg_solve.argtypes = [vp, C.c_int, C.c_int64, C.c_double, C.c_int, dp, dp, ip64, dp]
    _lib = L
    return L


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def i64ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64)) if a is not None else None


def i32ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


# ---- host-side helpers (no device) -------------------------------------------------------------
def tensor_rule(topology, order):
    dim = {HEX: 3, QUAD: 2, LINE: 1}[topology]
    pts = np.zeros((4096, dim))
    w = np.zeros(len(pts))
    n = lib().b200asm_tensor_rule(topology, order, dptr(pts), dptr(w))
    if n < 0:
        raise B200AsmError(n, "tensor_rule")
    return pts[:n].copy(), w[:n].copy()


def prism_rule(order, tripts, triw):
    """TPZIntPrism3D: line rule of `order` x the given triangle rule (triangle point fastest)."""
    tripts = np.ascontiguousarray(tripts, dtype=np.float64)
    triw = np.ascontiguousarray(triw, dtype=np.float64)
    pts = np.zeros((64 * len(triw), 3))
    w = np.zeros(len(pts))
    n = lib().b200asm_prism_rule(order, len(triw), dptr(tripts), dptr(triw), dptr(pts), dptr(w))
    if n < 0:
        raise B200AsmError(n, "prism_rule")
    return pts[:n].copy(), w[:n].copy()


def nshape(topology, porder):
    n = lib().b200asm_nshape(topology, porder)
    if n < 0:
        raise B200AsmError(n, f"nshape: unsupported topology {topology} / order {porder}")
    return n


def orientation_keys(topology, elnodes):
    """Side-orientation class of every element (from the global corner-node indices); all zero for simplices."""
    elnodes = np.ascontiguousarray(elnodes, dtype=np.int32)
    keys = np.zeros(len(elnodes), dtype=np.int64)
    rc = lib().b200asm_orientation_keys(topology, len(elnodes), i32ptr(elnodes), i64ptr(keys))
    if rc < 0:
        raise B200AsmError(rc, "orientation_keys")
    return keys


def shape_tables(topology, porder, qpts, key=0):
    """phi[nq][n], dphi[nq][dim][n] of the orientation class `key` (irrelevant for p <= 2)."""
    qpts = np.ascontiguousarray(qpts, dtype=np.float64)
    nq, dim = qpts.shape
    n = nshape(topology, porder)
    phi = np.zeros((nq, n))
    dphi = np.zeros((nq, dim, n))
    rc = lib().b200asm_shape_tables_oriented(topology, porder, int(key), nq, dptr(qpts), dptr(phi), dptr(dphi))
    if rc != n:
        raise B200AsmError(rc, "shape_tables")
    return phi, dphi


def build_pattern(symmetric, elgraphindex, elgraph, blockpos, blocksize, nthreads=0):
    elgraphindex = np.ascontiguousarray(elgraphindex, dtype=np.int64)
    elgraph = np.ascontiguousarray(elgraph, dtype=np.int64)
    blockpos = np.ascontiguousarray(blockpos, dtype=np.int64)
    blocksize = np.ascontiguousarray(blocksize, dtype=np.int64)
    neq = int(blocksize.sum())
    ia = np.zeros(neq + 1, dtype=np.int64)
    nel = len(elgraphindex) - 1
    args = (int(bool(symmetric)), nel, i64ptr(elgraphindex), i64ptr(elgraph), len(blockpos), i64ptr(blockpos),
            i64ptr(blocksize), i64ptr(ia))
    nnz = lib().b200asm_build_pattern(*args, None, nthreads)
    if nnz < 0:
        raise B200AsmError(nnz, "build_pattern")
    ja = np.empty(nnz, dtype=np.int64)
    nnz2 = lib().b200asm_build_pattern(*args, i64ptr(ja), nthreads)
    assert nnz2 == nnz
    return ia, ja


def make_group(topology, porder, kind, nstate, elnodes, dest, qpts, qwts, phi, dphi, coef, force=None):
    """b200asm_group over numpy arrays; returns (struct, arrays to keep alive during the call)."""
    elnodes = np.ascontiguousarray(elnodes, dtype=np.int32)
    dest = np.ascontiguousarray(dest, dtype=np.int64)
    qpts = np.ascontiguousarray(qpts, dtype=np.float64)
    qwts = np.ascontiguousarray(qwts, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    dphi = np.ascontiguousarray(dphi, dtype=np.float64)
    g = Group()
    g.topology, g.porder, g.kind, g.nstate = topology, porder, kind, nstate
    g.nel = elnodes.shape[0]
    g.elnodes, g.dest = i32ptr(elnodes), i64ptr(dest)
    g.nqp, g.nshape = len(qwts), phi.shape[1]
    g.qpts, g.qwts, g.phi, g.dphi = dptr(qpts), dptr(qwts), dptr(phi), dptr(dphi)
    c = np.zeros(16)
    c[: len(coef)] = coef
    g.coef[:] = c.tolist()
    keep = [elnodes, dest, qpts, qwts, phi, dphi]
    if force is not None:
        force = np.ascontiguousarray(force, dtype=np.float64)
        g.force = dptr(force)
        keep.append(force)
    return g, keep


# ---- device context ------------------------------------------------------------------------------
class Context:
    """Owns one b200asm_ctx (one GPU)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().b200asm_create(C.byref(self._h), device)
        if rc != 0:
            raise B200AsmError(rc, lib().b200asm_last_error(None).decode())
        self._keep = []
        self._neq = 0

    def _check(self, rc):
        if rc < 0:
            raise B200AsmError(rc, lib().b200asm_last_error(self._h).decode())
        return rc

    def close(self):
        if self._h:
            lib().b200asm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        self._check(lib().b200asm_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def set_option(self, name, value):
        self._check(lib().b200asm_set_option(self._h, name.encode(), int(value)))

    def set_nodes(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._check(lib().b200asm_set_nodes(self._h, xyz.shape[0], dptr(xyz)))

    def add_group(self, topology, porder, kind, nstate, elnodes, dest, qpts, qwts, phi, dphi, coef, force=None):
        g, keep = make_group(topology, porder, kind, nstate, elnodes, dest, qpts, qwts, phi, dphi, coef, force)
        return self._check(lib().b200asm_add_group(self._h, C.byref(g)))

    def set_group_force(self, group, force):
        force = np.ascontiguousarray(force, dtype=np.float64)
        self._check(lib().b200asm_set_group_force(self._h, group, dptr(force)))

    def set_group_coef(self, group, coef):
        c = np.zeros(16)
        c[: len(coef)] = coef
        self._check(lib().b200asm_set_group_coef(self._h, group, dptr(c)))

    def clear_groups(self):
        self._check(lib().b200asm_clear_groups(self._h))

    def set_pattern(self, ia, ja, symmetric):
        ia = np.ascontiguousarray(ia, dtype=np.int64)
        ja = np.ascontiguousarray(ja, dtype=np.int64)
        self._check(lib().b200asm_set_pattern(self._h, len(ia) - 1, i64ptr(ia), i64ptr(ja), int(bool(symmetric))))
        self._neq = len(ia) - 1

    def build_pattern_device(self, symmetric, elgraphindex, elgraph, blockpos, blocksize):
        """CSR pattern of the reference built on the device from the element graph; becomes the current pattern.
        Returns (neq, nnz)."""
        elgraphindex = np.ascontiguousarray(elgraphindex, dtype=np.int64)
        elgraph = np.ascontiguousarray(elgraph, dtype=np.int64)
        blockpos = np.ascontiguousarray(blockpos, dtype=np.int64)
        blocksize = np.ascontiguousarray(blocksize, dtype=np.int64)
        neq, nnz = C.c_int64(), C.c_int64()
        self._check(lib().b200asm_build_pattern_device(self._h, int(bool(symmetric)), len(elgraphindex) - 1, i64ptr(elgraphindex),
                                                       i64ptr(elgraph), len(blockpos), i64ptr(blockpos), i64ptr(blocksize),
                                                       C.byref(neq), C.byref(nnz)))
        self._neq = neq.value
        return neq.value, nnz.value

    def get_pattern(self, neq, nnz, want_ja=True):
        ia = np.empty(neq + 1, dtype=np.int64)
        ja = np.empty(nnz, dtype=np.int64) if want_ja else None
        self._check(lib().b200asm_get_pattern(self._h, i64ptr(ia), i64ptr(ja)))
        return ia, ja

    def cg_solve(self, precond=1, max_iter=50000, tol=1e-15, x0=None, f=None, download=True):
        """CG on the device-resident matrix (reference algorithm: Solvers/LinearSolvers/cg.h:44-120).
        Returns (x or None, iterations, relative residual)."""
        n = self.neq()
        x = None
        if x0 is not None:
            x = np.ascontiguousarray(x0, dtype=np.float64).copy()
        elif download:
            x = np.zeros(n)
        if f is not None:
            f = np.ascontiguousarray(f, dtype=np.float64)
        it, res = C.c_int64(), C.c_double()
        self._check(lib().b200asm_cg_solve(self._h, int(precond), int(max_iter), float(tol), int(x0 is not None), dptr(f), dptr(x),
                                           C.byref(it), C.byref(res)))
        return x, it.value, res.value

    def neq(self):
        return self._neq

    def assemble(self, a_host=None, rhs_host=None):
        self._check(lib().b200asm_assemble(self._h, dptr(a_host), dptr(rhs_host)))

    def get_ja_range(self, first, count):
        ja = np.empty(int(count), dtype=np.int64)
        self._check(lib().b200asm_get_ja_range(self._h, int(first), int(count), i64ptr(ja)))
        return ja

    def assemble_rhs(self, rhs_host=None):
        self._check(lib().b200asm_assemble_rhs(self._h, dptr(rhs_host)))

    def assemble_async(self):
        self._check(lib().b200asm_assemble_async(self._h))

    def synchronize(self):
        self._check(lib().b200asm_synchronize(self._h))

    def download(self, a_host=None, rhs_host=None):
        self._check(lib().b200asm_download(self._h, dptr(a_host), dptr(rhs_host)))

    def scatter_add(self, target, positions_dev_ptr, values_dev_ptr, n):
        """dst[positions[k]] += values[k] on the device (target 0: CSR values, 1: rhs); raw device pointers."""
        self._check(lib().b200asm_scatter_add(self._h, target, C.c_void_p(positions_dev_ptr), C.c_void_p(values_dev_ptr), n))

    # ---- interface exchange of the row-sharded assembly (include/b200asm.h: b200asm_exchange_*) ----
    def exchange_export(self):
        """bytes (3 x 72) naming this context's CSR values, load vector and flag block for a peer in another process."""
        mem = (IpcMem * 3)()
        self._check(lib().b200asm_exchange_export(self._h, mem))
        return bytes(mem)

    def exchange_add_peer(self, push, slot_there, mem_bytes=None, peer=None, incoming_min_row=-1):
        mem = None
        if mem_bytes is not None:
            mem = (IpcMem * 3).from_buffer_copy(mem_bytes)
        return self._check(lib().b200asm_exchange_add_peer(self._h, int(bool(push)), int(slot_there), mem,
                                                           peer._h if peer is not None else None, int(incoming_min_row)))

    def exchange_set_map(self, link, a_src0, a_dst, rhs_src, rhs_dst):
        a_dst = np.ascontiguousarray(a_dst, dtype=np.int32)
        rhs_src = np.ascontiguousarray(rhs_src, dtype=np.int32)
        rhs_dst = np.ascontiguousarray(rhs_dst, dtype=np.int32)
        self._check(lib().b200asm_exchange_set_map(self._h, int(link), len(a_dst), int(a_src0), i32ptr(a_dst), len(rhs_src),
                                                   i32ptr(rhs_src), i32ptr(rhs_dst)))

    def exchange_clear(self):
        self._check(lib().b200asm_exchange_clear(self._h))

    def device_pointers(self):
        a, r = C.c_void_p(), C.c_void_p()
        self._check(lib().b200asm_device_pointers(self._h, C.byref(a), C.byref(r)))
        return a.value, r.value

    def group_time_ms(self, group):
        """CUDA-event duration of the kernel launches of `group` in the last assembly (option "timing" = 1)."""
        ms = C.c_double()
        self._check(lib().b200asm_group_time_ms(self._h, group, C.byref(ms)))
        return ms.value

    def counters(self):
        k, h, d = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(lib().b200asm_counters(self._h, C.byref(k), C.byref(h), C.byref(d)))
        return k.value, h.value, d.value

    def group_kernel(self, group):
        """Kernel family of the matrix part of a group in the last assembly (b200asm_group_kernel)."""
        buf = C.create_string_buffer(32)
        self._check(lib().b200asm_group_kernel(self._h, int(group), buf, 32))
        return buf.value.decode()


class _Borrowed(Context):
    """A context owned by a MultiContext (never destroyed from here)."""

    def __init__(self, handle):
        self._h = handle
        self._keep = []
        self._neq = 0

    def close(self):
        self._h = C.c_void_p()


class MultiContext:
    """Owns one b200asm_multi: the same calls as Context on several GPUs of this process (include/b200asm.h)."""

    def __init__(self, devices):
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = lib().b200asm_multi_create(C.byref(self._h), len(devices), devs)
        if rc != 0:
            raise B200AsmError(rc, lib().b200asm_multi_last_error(None).decode())
        self.ndev = len(devices)
        self._neq = 0

    def _check(self, rc):
        if rc < 0:
            raise B200AsmError(rc, lib().b200asm_multi_last_error(self._h).decode())
        return rc

    def close(self):
        if self._h:
            lib().b200asm_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def context(self, k):
        h = C.c_void_p()
        self._check(lib().b200asm_multi_context(self._h, k, C.byref(h)))
        return _Borrowed(h)

    def set_option(self, name, value):
        self._check(lib().b200asm_multi_set_option(self._h, name.encode(), int(value)))

    def set_nodes(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._check(lib().b200asm_multi_set_nodes(self._h, xyz.shape[0], dptr(xyz)))

    def add_group(self, topology, porder, kind, nstate, elnodes, dest, qpts, qwts, phi, dphi, coef, force=None):
        g, keep = make_group(topology, porder, kind, nstate, elnodes, dest, qpts, qwts, phi, dphi, coef, force)
        return self._check(lib().b200asm_multi_add_group(self._h, C.byref(g)))

    def set_group_coef(self, group, coef):
        c = np.zeros(16)
        c[: len(coef)] = coef
        self._check(lib().b200asm_multi_set_group_coef(self._h, group, dptr(c)))

    def set_group_force(self, group, force):
        force = np.ascontiguousarray(force, dtype=np.float64)
        self._check(lib().b200asm_multi_set_group_force(self._h, group, dptr(force)))

    def clear_groups(self):
        self._check(lib().b200asm_multi_clear_groups(self._h))

    def set_pattern(self, ia, ja, symmetric):
        ia = np.ascontiguousarray(ia, dtype=np.int64)
        ja = np.ascontiguousarray(ja, dtype=np.int64)
        self._check(lib().b200asm_multi_set_pattern(self._h, len(ia) - 1, i64ptr(ia), i64ptr(ja), int(bool(symmetric))))
        self._neq = len(ia) - 1

    def partition(self):
        """(row_begin[ndev + 1], elements[ndev], staged CSR entries[ndev]) of the current partition."""
        rb = np.zeros(self.ndev + 1, dtype=np.int64)
        el = np.zeros(self.ndev, dtype=np.int64)
        st = np.zeros(self.ndev, dtype=np.int64)
        self._check(lib().b200asm_multi_partition(self._h, i64ptr(rb), i64ptr(el), i64ptr(st)))
        return rb, el, st

    def neq(self):
        return self._neq

    def assemble(self, a_host=None, rhs_host=None):
        self._check(lib().b200asm_multi_assemble(self._h, dptr(a_host), dptr(rhs_host)))

    def assemble_rhs(self, rhs_host=None):
        self._check(lib().b200asm_multi_assemble_rhs(self._h, dptr(rhs_host)))

    def assemble_async(self):
        self._check(lib().b200asm_multi_assemble_async(self._h))

    def synchronize(self):
        self._check(lib().b200asm_multi_synchronize(self._h))

    def counters(self):
        k, h, d = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(lib().b200asm_multi_counters(self._h, C.byref(k), C.byref(h), C.byref(d)))
        return k.value, h.value, d.value

    def cg_solve(self, precond=1, max_iter=50000, tol=1e-15, x0=None, f=None, download=True):
        """CG on the row-sharded device-resident matrix (b200asm_multi_cg_solve); global vectors in and out.
        Returns (x or None, iterations, relative residual)."""
        n = self.neq()
        x = None
        if x0 is not None:
            x = np.ascontiguousarray(x0, dtype=np.float64).copy()
        elif download:
            x = np.zeros(n)
        if f is not None:
            f = np.ascontiguousarray(f, dtype=np.float64)
        it, res = C.c_int64(), C.c_double()
        self._check(lib().b200asm_multi_cg_solve(self._h, int(precond), int(max_iter), float(tol), int(x0 is not None), dptr(f), dptr(x),
                                                 C.byref(it), C.byref(res)))
        return x, it.value, res.value
