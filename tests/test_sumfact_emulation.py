"""CPU emulation of the sum-factorisation kernel for hexahedra of order 2 (neopz_b200/csrc/sumfact_hex.cuh): the kernel's phase
functions run with thread loops (tools/sumfact_emu.cpp) and must reproduce the element matrices of the unmodified reference
(fixtures) - index maps, factor tables and the coverage of the upper triangle are thereby checked without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from neopz_b200 import capi, strmatrix
from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "bin", "libsumfact_emu.so")


def _lib():
    src = os.path.join(ROOT, "tools", "sumfact_emu.cpp")
    hdr = os.path.join(ROOT, "neopz_b200", "csrc", "sumfact_hex.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", src, "-o", SO])
    lib = C.CDLL(SO)
    dp = C.POINTER(C.c_double)
    lib.sf_emulate.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp]
    lib.sf_emulate_mode.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp, C.c_int]
    return lib


@pytest.mark.parametrize("mode", [0])
@pytest.mark.parametrize("name", ["hex_p2_poisson_n2_pert"])
def test_emulated_kernel_reproduces_reference_element_matrices(name, mode):
    g = gu.load(name)
    lib = _lib()
    qpts, qw, phi, dphi = strmatrix.element_tables(capi.HEX, 2)
    assert np.array_equal(qpts, g["rule_hex_pts"])
    _geo, dng = capi.shape_tables(capi.HEX, 1, qpts)          # [27][3][8], the table the device kernels read
    x1d = np.ascontiguousarray(qpts[:3, 0])
    # the rule is the tensor rule the kernel assumes: q = q1 + 3 (q2 + 3 q3)
    for q in range(27):
        assert np.array_equal(qpts[q], [x1d[q % 3], x1d[(q // 3) % 3], x1d[q // 9]])
    nhex = 0
    for e in range(len(g["el_type"])):
        if g["el_type"][e] != capi.HEX:
            continue
        X = np.ascontiguousarray(g["nodes"][g["el_nodes"][e, :8]])
        K = np.zeros((27, 27))
        wd = np.zeros(27)
        n = lib.sf_emulate_mode(capi.dptr(X), capi.dptr(np.ascontiguousarray(dng)), capi.dptr(qw), capi.dptr(x1d), 1.0, capi.dptr(K),
                                capi.dptr(wd), mode)
        assert n == 378, n       # every entry of the upper triangle exactly once
        ref = g["ek"][g["ek_ptr"][e]:g["ek_ptr"][e + 1]].reshape(27, 27).T
        assert np.linalg.norm(K - ref) <= 1e-14 * np.linalg.norm(ref)
        assert np.abs(K - ref).max() <= 1e-14 * np.abs(ref).max()
        nhex += 1
    assert nhex == 8
