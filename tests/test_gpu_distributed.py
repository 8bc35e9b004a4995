"""Row-sharded assembly on 2 GPUs over NCCL (skipped on a single-GPU box): CUDA kernels + device-side exchange
(send of the ghost-row CSR prefix, b200asm_scatter_add on the receiver) against the single-mesh oracle assembly."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.test_distributed_gloo import _free_port

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, nxy, nzl, p, phys, tet, symmetric, q, pattern="host"):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from neopz_b200 import capi, distributed, gridmesh
        from tests.oracle_ref import oracle_assemble
        from tests.test_gpu_parity import materials_for
        ns = 3 if phys else 1
        bc = (-1, -1, -1, -1, -1, -2)
        mats = materials_for(phys, neumann=True)
        slab = distributed.slab_mesh(nxy, nzl * world, rank, world, p, ns, tetrahedra=bool(tet), bc_matids=bc, perturb=0.1)
        sh = distributed.ShardedStructMatrix(slab, mats, symmetric=symmetric, device=rank, pattern=pattern)
        sh.strmat.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        sh.Create()
        for _ in range(2):  # re-assembly must reproduce
            a, rhs = sh.Assemble()
        ia_o, ja_o, a_o, rhs_o = sh.own_rows(a, rhs)
        gm = gridmesh.grid_mesh((nxy, nxy, nzl * world), p, ns, tetrahedra=bool(tet), bc_matids=bc, perturb=0.1)
        idx, graph = gm.element_graph()
        ia, ja = capi.build_pattern(symmetric, idx, graph, gm.block_pos, gm.block_size, 2)
        a_ref, rhs_ref = oracle_assemble(gm, mats, symmetric, ia, ja)
        r0, r1 = slab.row0, slab.row0 + slab.nown
        lo, hi = ia[r0], ia[r1]
        assert np.array_equal(ia_o, ia[r0:r1 + 1] - lo) and np.array_equal(ja_o, ja[lo:hi])
        err_a = np.linalg.norm(a_o - a_ref[lo:hi]) / np.linalg.norm(a_ref[lo:hi])
        err_r = np.linalg.norm(rhs_o - rhs_ref[r0:r1]) / np.linalg.norm(rhs_ref[r0:r1])
        assert err_a <= 1e-12 and err_r <= 1e-12, (err_a, err_r)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("pattern", ["host", "device"])
@pytest.mark.parametrize("nxy,nzl,p,phys,tet,symmetric", [(6, 3, 2, 0, 0, True), (4, 2, 2, 1, 0, True), (4, 2, 2, 0, 1, False), (3, 2, 4, 0, 0, True)])
def test_sharded_cuda_nccl(nxy, nzl, p, phys, tet, symmetric, pattern):
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nxy, nzl, p, phys, tet, symmetric, q, pattern)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for rank, status in sorted(res):
        assert status == "ok", f"rank {rank}: {status}"
