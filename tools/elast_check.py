#!/usr/bin/env python
"""One-shot check of the hexahedra p2 TPZElasticity3D kernels (option variant: 0 / 31 = default, a pair of warps per element; 30 = one warp per
element; 34 = the team of ten warps, gram_mma_team.cuh) on a GPU: parity against the oracle on small perturbed meshes (both storages, coloured scatter,
load vector only, prestress, forcing table), then the CUDA-event time of the volume group.  JSON lines on stdout.

    python tools/elast_check.py [grid] [variant ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from neopz_b200 import gridmesh, strmatrix as sm  # noqa: E402
from tests.oracle_ref import oracle_assemble  # noqa: E402  (checker only)


def mats(prestress=False, forcing=None):
    m = sm.TPZElasticity3D(1, 1000.0, 0.3, (0.3, -0.2, -1.0), prestress=(1.5, -0.5, 0.25) if prestress else (0.0, 0.0, 0.0))
    if forcing:
        m.SetForcingFunction(forcing)
    return {1: m, -1: m.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3)), -2: m.CreateBC(-2, 1, np.zeros((3, 3)), np.array([0.1, 0.2, -0.3]))}


def relF(x, ref):
    return float(np.linalg.norm(x - ref) / np.linalg.norm(ref))


def main():
    n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    variants = [int(v) for v in sys.argv[2:]] or [30, 31, 34]
    force = lambda x: np.stack([1.0 + x[:, 0], x[:, 1] * x[:, 2], -0.5 + x[:, 2]], axis=1)  # noqa: E731
    for variant in variants:
        worst = 0.0
        for n, sym, scatter, pre, forcing in ((4, True, "atomic", False, None), (3, False, "atomic", False, None), (4, True, "colored", False, None),
                                               (3, True, "atomic", True, None), (4, True, "atomic", False, force)):
            mesh = gridmesh.grid_mesh(n, 2, 3, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
            mm = mats(pre, forcing)
            s = sm.TPZStructMatrixB200(mesh, mm, symmetric=sym, variant=variant, scatter=scatter)
            ia, ja, a, rhs = s.CreateAssemble()
            a_ref, rhs_ref = oracle_assemble(mesh, mm, sym, ia, ja)
            a2, rhs2 = s.Assemble()
            r3 = s.AssembleRhs()
            worst = max(worst, relF(a, a_ref), relF(rhs, rhs_ref), relF(a2, a_ref), relF(rhs2, rhs_ref), relF(r3, rhs_ref))
            s.ctx.close()
        print(json.dumps({"variant": variant, "parity_worst_relF": worst, "ok": worst <= 1e-12}), flush=True)
    mesh = gridmesh.grid_mesh(n_time, 2, 3, perturb=0.1)
    nvol = len(mesh.blocks[0].elnodes)
    for variant in [0] + variants:
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
        print(json.dumps({"variant": variant, "grid": n_time, "volume_elements": nvol, "volume_kernel_ms": t, "elements_per_s": nvol / (t * 1e-3),
                          "kernel": s.ctx.group_kernel(s.group_of_block[0])}), flush=True)
        s.ctx.close()


if __name__ == "__main__":
    main()
