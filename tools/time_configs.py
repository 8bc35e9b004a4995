#!/usr/bin/env python
"""Device-resident assembly throughput of the configurations outside bench.py's headline (prisms, tetrahedra of order 3-4,
the generic runtime-size kernel): CUDA-event duration of the volume group's launches (b200asm_group_time_ms), mean of a few
assemblies after warm-up.  One JSON line per configuration.

    python tools/time_configs.py [n]           # n = grid divisions (default 24)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from neopz_b200 import gridmesh, strmatrix as sm  # noqa: E402

CONFIGS = [  # name, p, phys, tetrahedra, prisms, engine[, variant]
    ("prism p1 poisson", 1, 0, False, True, 1), ("prism p2 poisson", 2, 0, False, True, 1), ("prism p2 elasticity", 2, 1, False, True, 1),
    ("tet p3 poisson", 3, 0, True, False, 1), ("tet p3 elasticity", 3, 1, True, False, 1),
    ("tet p4 poisson", 4, 0, True, False, 1), ("tet p4 elasticity", 4, 1, True, False, 1),
    ("hex p2 poisson, generic kernel", 2, 0, False, False, 2), ("hex p2 elasticity, generic kernel", 2, 1, False, False, 2),
    ("hex p2 poisson, register tiles", 2, 0, False, False, 0),
    # closed-form kernel with the table in shared memory (tuning variant 20; not run on a GPU in round 1)
    ("tet p3 poisson, closed form", 3, 0, True, False, 1, 20), ("tet p3 elasticity, closed form", 3, 1, True, False, 1, 20),
    ("tet p4 poisson, closed form", 4, 0, True, False, 1, 20), ("tet p4 elasticity, closed form", 4, 1, True, False, 1, 20),
]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    only20 = len(sys.argv) > 2 and sys.argv[2] == "variant20"
    for cfg in CONFIGS:
        name, p, phys, tet, prisms, engine = cfg[:6]
        variant = cfg[6] if len(cfg) > 6 else 0
        if variant == 20 and not only20 and os.environ.get("B200ASM_TIME_UNVERIFIED") != "1":
            continue   # (python tools/time_configs.py 24 variant20 measures them, after tests/ has checked their parity)
        if only20 and variant != 20:
            continue
        nn = max(6, n // 2) if (p >= 4 or (p >= 3 and phys)) else n
        mesh = gridmesh.grid_mesh(nn, p, 3 if phys else 1, tetrahedra=tet, prisms=prisms, perturb=0.1)
        if phys == 0:
            mat = sm.TPZMatPoisson(1, 3)
            mat.SetForcingFunction(1.0)
            mats = {1: mat, -1: mat.CreateBC(-1, 0, [[0.0]], [0.0])}
        else:
            mat = sm.TPZElasticity3D(1, 1000.0, 0.3, (0.0, 0.0, -1.0))
            mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3))}
        strmat = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, engine=engine, variant=variant)
        strmat.Create(on_device=True, download=False)
        ctx = strmat.ctx
        for _ in range(3):
            ctx.assemble_async()
        ctx.synchronize()
        ctx.set_option("timing", 1)
        ms = []
        for _ in range(5):
            ctx.assemble_async()
            ms.append(sum(ctx.group_time_ms(g) for g in strmat.groups_of_block[0]))
        nvol = len(mesh.blocks[0].elnodes)
        t = float(np.mean(ms))
        print(json.dumps({"config": name, "grid": nn, "volume_elements": nvol, "dof": mesh.neq, "orientation_groups": len(strmat.groups_of_block[0]),
                          "volume_kernel_ms": t, "elements_per_s": nvol / (t * 1e-3)}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
