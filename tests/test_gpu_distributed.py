"""Row-sharded assembly on 2 (and 4) GPUs, one process per GPU (skipped on a single-GPU box): CUDA kernels + device-side
exchange against the single-mesh oracle assembly.  Both transports: "p2p" = the C ABI's own exchange (interface elements first,
push kernel into the owner's memory mapped with CUDA IPC, step counters; csrc/exchange.cuh) and "nccl" = send of the ghost-row
CSR prefix + b200asm_scatter_add on the receiver.  Also the end-to-end call b200asm_assemble with page-locked host buffers
(download overlapped with the kernels and the exchange)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.test_distributed_gloo import _free_port

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, nxy, nzl, p, phys, tet, symmetric, q, pattern="host", exchange="p2p"):
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
        sh = distributed.ShardedStructMatrix(slab, mats, symmetric=symmetric, device=rank, pattern=pattern, exchange=exchange)
        if exchange == "p2p":
            sh.strmat.ctx.set_option("overlap_min_elements", 384)  # (element chunks and early copies on these small meshes)
            sh.strmat.ctx.set_option("overlap_min_bytes", 0)
            sh.strmat.ctx.set_option("exchange_timeout_ms", 20000)
        sh.Create()
        for _ in range(3):  # re-assembly must reproduce
            a, rhs = sh.Assemble()
        ia_o, ja_o, a_o, rhs_o = sh.own_rows(a, rhs)
        if exchange == "p2p":  # the end-to-end call: pinned host buffers, overlapped download, exchange inside
            a_pin = torch.empty(sh.nnz, dtype=torch.float64).pin_memory()
            r_pin = torch.empty(slab.mesh.neq, dtype=torch.float64).pin_memory()
            for _ in range(2):
                sh.strmat.ctx.assemble(a_pin.numpy(), r_pin.numpy())
            _i, _j, a_e, rhs_e = sh.own_rows(a_pin.numpy(), r_pin.numpy())
            assert np.linalg.norm(a_e - a_o) <= 1e-13 * np.linalg.norm(a_o) and np.linalg.norm(rhs_e - rhs_o) <= 1e-13 * np.linalg.norm(rhs_o)
            rhs_only = np.empty(slab.mesh.neq)
            sh.strmat.ctx.assemble_rhs(rhs_only)  # the load-vector-only path exchanges its ghost rows too
            sh.strmat.ctx.synchronize()
            r0l = slab.nghost
            assert np.linalg.norm(rhs_only[r0l:r0l + slab.nown] - rhs_o) <= 1e-13 * np.linalg.norm(rhs_o)
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
        sh.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()[-1500:]))


@pytest.mark.parametrize("pattern,exchange,world", [("host", "p2p", 2), ("device", "p2p", 2), ("device", "nccl", 2), ("device", "p2p", 4)])
@pytest.mark.parametrize("nxy,nzl,p,phys,tet,symmetric", [(6, 3, 2, 0, 0, True), (4, 2, 2, 1, 0, True), (4, 2, 2, 0, 1, False), (3, 2, 4, 0, 0, True),
                                                          (12, 6, 2, 0, 0, True)])
def test_sharded_cuda(nxy, nzl, p, phys, tet, symmetric, pattern, exchange, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nxy, nzl, p, phys, tet, symmetric, q, pattern, exchange)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for rank, status in sorted(res):
        assert status == "ok", f"rank {rank}: {status}"
