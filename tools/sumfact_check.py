#!/usr/bin/env python
"""One-shot check of the sum-factorisation kernels (option variant, neopz_b200/csrc/sumfact_hex.cuh) on a GPU
(usage: sumfact_check.py [grid] [variant ...]):
parity against the oracle on small perturbed meshes (both storages, coloured scatter, load vector only, forcing table),
then the CUDA-event time of the volume group next to the default DMMA kernel.  JSON lines on stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from neopz_b200 import gridmesh, strmatrix as sm  # noqa: E402
from tests.oracle_ref import oracle_assemble  # noqa: E402  (checker only)


def mats(forcing=None):
    m = sm.TPZMatPoisson(1, 3)
    m.SetScaleFactor(1.7)
    m.SetForcingFunction(forcing if forcing else 1.0)
    return {1: m, -1: m.CreateBC(-1, 0, [[0.0]], [0.0]), -2: m.CreateBC(-2, 1, [[0.0]], [0.75])}


def relF(x, ref):
    return float(np.linalg.norm(x - ref) / np.linalg.norm(ref))


def main():
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    variants = [int(v) for v in sys.argv[2:]] or [8, 13]
    for variant in variants:
        worst = 0.0
        for n, sym, scatter, forcing in ((5, True, "atomic", None), (4, False, "atomic", None), (5, True, "colored", None),
                                          (7, True, "atomic", lambda x: 1.0 + x[:, 0] * x[:, 1] - 0.5 * x[:, 2])):
            mesh = gridmesh.grid_mesh(n, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
            mm = mats(forcing)
            s = sm.TPZStructMatrixB200(mesh, mm, symmetric=sym, variant=variant, scatter=scatter)
            ia, ja, a, rhs = s.CreateAssemble()
            a_ref, rhs_ref = oracle_assemble(mesh, mm, sym, ia, ja)
            a2, rhs2 = s.Assemble()
            r3 = s.AssembleRhs()
            worst = max(worst, relF(a, a_ref), relF(rhs, rhs_ref), relF(a2, a_ref), relF(rhs2, rhs_ref), relF(r3, rhs_ref))
            s.ctx.close()
        print(json.dumps({"variant": variant, "parity_worst_relF": worst, "ok": worst <= 1e-12}), flush=True)
    mesh = gridmesh.grid_mesh(n_time, 2, 1, perturb=0.1)
    nvol = len(mesh.blocks[0].elnodes)
    for variant in [0, 16] + variants:
        mm = mats()
        s = sm.TPZStructMatrixB200(mesh, mm, symmetric=True, variant=variant)
        s.Create(on_device=True, download=False)
        for _ in range(3):
            s.ctx.assemble_async()
        s.ctx.synchronize()
        s.ctx.set_option("timing", 1)
        ms = []
        for _ in range(6):
            s.ctx.assemble_async()
            ms.append(s.ctx.group_time_ms(s.group_of_block[0]))
        t = float(np.mean(ms[1:]))
        print(json.dumps({"variant": variant, "grid": n_time, "volume_elements": nvol, "volume_kernel_ms": t, "elements_per_s": nvol / (t * 1e-3)}), flush=True)
        s.ctx.close()


if __name__ == "__main__":
    main()
