#!/usr/bin/env python
"""Barrier-free sum-factorisation variants (option variant = 10, 11, 12): parity against the oracle, then timing next to
variant 8 (barrier form) on a 64^3 perturbed grid.  JSON lines on stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from neopz_b200 import gridmesh, strmatrix as sm  # noqa: E402
from tests.oracle_ref import oracle_assemble  # noqa: E402  (checker only)
from tools.sumfact_check import mats, relF  # noqa: E402

for variant in (10, 11, 12):
    worst = 0.0
    for n, sym, forcing in ((6, True, lambda x: 1.0 + x[:, 0] * x[:, 1] - 0.5 * x[:, 2]), (4, False, None)):
        mesh = gridmesh.grid_mesh(n, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
        mm = mats(forcing)
        s = sm.TPZStructMatrixB200(mesh, mm, symmetric=sym, variant=variant)
        ia, ja, a, rhs = s.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mm, sym, ia, ja)
        worst = max(worst, relF(a, a_ref), relF(rhs, rhs_ref), relF(s.AssembleRhs(), rhs_ref))
        s.ctx.close()
    print(json.dumps({"variant": variant, "parity_worst_relF": worst, "ok": worst <= 1e-12}), flush=True)
mesh = gridmesh.grid_mesh(64, 2, 1, perturb=0.1)
nvol = len(mesh.blocks[0].elnodes)
for variant in (8, 10, 11, 12):
    s = sm.TPZStructMatrixB200(mesh, mats(), symmetric=True, variant=variant)
    s.Create(on_device=True, download=False)
    for _ in range(3):
        s.ctx.assemble_async()
    s.ctx.synchronize()
    s.ctx.set_option("timing", 1)
    ms = []
    for _ in range(5):
        s.ctx.assemble_async()
        ms.append(s.ctx.group_time_ms(s.group_of_block[0]))
    t = float(np.mean(ms[1:]))
    print(json.dumps({"variant": variant, "volume_kernel_ms": t, "elements_per_s": nvol / (t * 1e-3)}), flush=True)
    s.ctx.close()
