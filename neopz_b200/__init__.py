"""neopz_b200 — B200-native global assembly for NeoPZ (TPZStructMatrix strategy + C ABI + host mirror).

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/b200asm.h, and the NeoPZ
strategy class), capi.py (ctypes binding), gridmesh.py (flattened structured meshes numbered like
NeoPZ numbers them) and strmatrix.py (host-side mirror of the TPZStructMatrix interface).
"""
from . import capi  # noqa: F401
