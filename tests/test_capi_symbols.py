"""The C-ABI library loads on a machine without a GPU, exports every symbol include/b200asm.h declares,
and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from neopz_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200asm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200asm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    L = capi.lib()
    declared = _declared_symbols()
    assert declared, "no declarations found"
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/b200asm.h but not exported"
    assert sorted(capi.SYMBOLS) == declared


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.B200AsmError) as ei:
        capi.Context(0)
    assert ei.value.code == capi.ENODEVICE


def test_product_never_imports_oracle():
    """Nothing under neopz_b200/ may import, link or call oracle/ (the oracle is a checker)."""
    bad = []
    for dp, _dn, files in os.walk(os.path.join(ROOT, "neopz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"(^|\s)(from|import)\s+oracle\b|liboracle|oracle/oracle|oracle\.c\b|orc_[a-z]+\(", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_enumerations_match_the_python_binding():
    """Topology / kind codes of include/b200asm.h == the constants of neopz_b200/capi.py (and of the oracle's, which shares the
    numbering of the element types)."""
    text = open(os.path.join(ROOT, "include", "b200asm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(B200ASM_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    for name in ("HEX", "TET", "QUAD", "TRI", "LINE", "PRISM", "PYRAMID", "POISSON", "ELASTICITY3D", "BC", "ELASTICITY2D"):
        assert enums["B200ASM_" + name] == getattr(capi, name), name
    assert enums["B200ASM_ENODEVICE"] == capi.ENODEVICE
    from oracle import oracle as orc
    assert (orc.HEX, orc.TET, orc.QUAD, orc.TRI, orc.LINE, orc.PRISM, orc.PYR) == (capi.HEX, capi.TET, capi.QUAD, capi.TRI, capi.LINE,
                                                                                  capi.PRISM, capi.PYRAMID)
