"""Row-sharded assembly (neopz_b200/distributed.py) on CPU: world_size 2 and 3 over gloo.  The host logic under
test is the product's (slab partition, global numbering without a global mesh, local patterns, receive maps,
the exchange); the local element arithmetic is injected from the oracle so that no GPU is needed.  Every
rank's owned rows must equal the corresponding rows of the single-mesh (global) reference assembly:
pattern bit-exact, values to 1e-13."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nxy, nzl, p, phys, tet, symmetric, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from neopz_b200 import capi, distributed, gridmesh
        from tests.oracle_ref import oracle_assemble
        from tests.test_gpu_parity import materials_for
        ns = 3 if phys else 1
        bc = (-1, -1, -1, -1, -1, -2)
        mats = materials_for(phys, neumann=True)
        slab = distributed.slab_mesh(nxy, nzl * world, rank, world, p, ns, tetrahedra=bool(tet), bc_matids=bc, perturb=0.1)
        sh = distributed.ShardedStructMatrix(slab, mats, symmetric=symmetric, local_assembler=oracle_assemble)
        sh.Create()
        a, rhs = sh.Assemble()
        ia_o, ja_o, a_o, rhs_o = sh.own_rows(a, rhs)
        # global reference on the undivided mesh (every rank builds it here only to check itself)
        gm = gridmesh.grid_mesh((nxy, nxy, nzl * world), p, ns, tetrahedra=bool(tet), bc_matids=bc, perturb=0.1)
        idx, graph = gm.element_graph()
        ia, ja = capi.build_pattern(symmetric, idx, graph, gm.block_pos, gm.block_size, 2)
        a_ref, rhs_ref = oracle_assemble(gm, mats, symmetric, ia, ja)
        assert slab.neq_global == gm.neq
        r0, r1 = slab.row0, slab.row0 + slab.nown
        lo, hi = ia[r0], ia[r1]
        assert np.array_equal(ia_o, ia[r0:r1 + 1] - lo), "row pointers of the owned rows"
        assert np.array_equal(ja_o, ja[lo:hi]), "column indices of the owned rows"
        scale = np.abs(a_ref[lo:hi]).max()
        err_a = np.abs(a_o - a_ref[lo:hi]).max() / scale
        err_r = np.abs(rhs_o - rhs_ref[r0:r1]).max() / max(np.abs(rhs_ref).max(), 1e-300)
        assert err_a < 1e-13 and err_r < 1e-13, (err_a, err_r)
        owned = np.array([r0, r1], dtype=np.int64)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", owned.tolist()))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()[-1500:], None))


@pytest.mark.parametrize("world,nxy,nzl,p,phys,tet,symmetric", [
    (2, 3, 2, 2, 0, 0, True), (2, 2, 2, 2, 1, 0, True), (3, 2, 1, 2, 0, 0, True), (2, 3, 2, 1, 0, 0, True),
    (2, 2, 2, 2, 0, 1, True), (2, 2, 2, 2, 0, 0, False), (2, 2, 1, 2, 1, 1, False)])
def test_sharded_assembly_matches_global(world, nxy, nzl, p, phys, tet, symmetric):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nxy, nzl, p, phys, tet, symmetric, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    res.sort()
    for rank, status, _ in res:
        assert status == "ok", f"rank {rank}: {status}"
    # the owned row blocks tile [0, neq) without gaps
    for (_, _, a), (_, _, b) in zip(res[:-1], res[1:]):
        assert a[1] == b[0]
    assert res[0][2][0] == 0
