#!/bin/bash
# First GPU passes of round 2 (everything written after round 1's GPU budget was spent, plus the profile the sum-factorisation
# kernel still lacks).  Run from the repo root:
#   gpurun --timeout 900 -- 'bash tools/round2_first_runs.sh'
set -x
mkdir -p gpurun_out
# 1. the 17 GPU tests that only had their CPU halves verified, then the whole suite
python -m pytest tests/test_gpu_zzz_sumfact.py tests/test_gpu_zzzz_added_late.py -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r02_late_tests.log
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r02_gpu_suite.log
# 2. sum-factorisation variants: 8/9 barrier form, 11 barrier-free, 13/14 barrier form + prefetch, 15 barrier-free + prefetch
python - <<'PY' 2>&1 | tee gpurun_out/r02_sumfact_variants.jsonl
import json, sys
sys.path.insert(0, ".")
import numpy as np
from neopz_b200 import gridmesh, strmatrix as sm
from tests.oracle_ref import oracle_assemble
from tools.sumfact_check import mats, relF
small = gridmesh.grid_mesh(6, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12)
mesh = gridmesh.grid_mesh(96, 2, 1, perturb=0.1)
nvol = len(mesh.blocks[0].elnodes)
for variant in (0, 8, 9, 11, 13, 14, 15):
    mm = mats(lambda x: 1.0 + x[:, 0] * x[:, 1])
    s = sm.TPZStructMatrixB200(small, mm, symmetric=True, variant=variant)
    ia, ja, a, rhs = s.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(small, mm, True, ia, ja)
    err = max(relF(a, a_ref), relF(rhs, rhs_ref))
    s.ctx.close()
    s = sm.TPZStructMatrixB200(mesh, mats(), symmetric=True, variant=variant)
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
    print(json.dumps({"variant": variant, "parity_relF": err, "grid": 96, "volume_kernel_ms": t, "elements_per_s": nvol / (t * 1e-3)}), flush=True)
    s.ctx.close()
PY
# 3. ncu: where do the 1500 cycles per element of the sum-factorisation kernel go (variant 8), next to the DMMA kernel
bash tools/ncu_capture.sh r02_ncu_sumfact_v8 assemble_sumfact --grid 64 --variant 8
bash tools/ncu_capture.sh r02_ncu_mma assemble_gram_mma --grid 64
# 3b. closed-form kernel for tetrahedra of order 3, 4 (variant 20): parity against the oracle, then timing
python - <<'PY' 2>&1 | tee gpurun_out/r02_tet_closed_form_parity.jsonl
import json, sys
sys.path.insert(0, ".")
import numpy as np
from neopz_b200 import gridmesh, strmatrix as sm
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import materials_for, relF
for n, p, phys in ((3, 3, 0), (2, 4, 0), (2, 3, 1), (2, 4, 1)):
    perm = np.random.default_rng(3).permutation((n + 1) ** 3)
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=True, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12, node_perm=perm)
    mats = materials_for(phys, neumann=True)
    for sym in (True, False):
        s = sm.TPZStructMatrixB200(mesh, mats, symmetric=sym, variant=20)
        ia, ja, a, rhs = s.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, sym, ia, ja)
        print(json.dumps({"n": n, "p": p, "phys": phys, "symmetric": sym, "relF_A": relF(a, a_ref), "relF_rhs": relF(rhs, rhs_ref)}), flush=True)
        s.ctx.close()
PY
python tools/time_configs.py 24 variant20 2>&1 | tee gpurun_out/r02_time_tet_closed_form.jsonl
# 4. the new configurations at sizes that fill the GPU
# (time_configs 48 dropped: GPU budget)

