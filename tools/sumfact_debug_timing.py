import json, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from neopz_b200 import gridmesh, strmatrix as sm
import tools.sumfact_check as sc
mesh = gridmesh.grid_mesh(64, 2, 1, perturb=0.1)
nvol = len(mesh.blocks[0].elnodes)
for variant, debug in ((0, 1), (8, 1), (9, 1), (8, 0)):
    s = sm.TPZStructMatrixB200(mesh, sc.mats(), symmetric=True, variant=variant)
    s.Create(on_device=True, download=False)
    if debug:
        s.ctx.set_option("debug", 1)
    for _ in range(3):
        s.ctx.assemble_async()
    s.ctx.synchronize()
    s.ctx.set_option("timing", 1)
    ms = []
    for _ in range(6):
        s.ctx.assemble_async()
        ms.append(s.ctx.group_time_ms(s.group_of_block[0]))
    t = float(np.mean(ms[1:]))
    print(json.dumps({"variant": variant, "scatter_dropped": bool(debug), "volume_kernel_ms": t, "elements_per_s": nvol / (t * 1e-3)}), flush=True)
    s.ctx.close()
