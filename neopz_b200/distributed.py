"""Row-sharded assembly across GPUs: z-slab element partition, owner-computes rows,
interface-row contributions exchanged once per assembly.

Partition.  The global grid (nx, ny, nz_total) is cut into `world` slabs of element layers.  Equation numbers
follow the reference's first-touch numbering of the GLOBAL mesh (gridmesh.flatten), so the equations first
touched by the elements of a slab form one contiguous block of rows: rank r owns rows [row0, row0+nown).  The
connects on the plane between slab r-1 and slab r were first touched by slab r-1, hence belong to rank r-1:
the elements of slab r contribute to those rows ("ghost rows"), and — symmetric or full storage alike —
contributions only ever travel downwards (r -> r-1).

Local system of rank r (extended numbering, monotone in the global one):
    [ ghost rows (interface plane below, partial: own elements only) | own rows (complete pattern) | upper columns ]
The complete pattern of the own rows needs the first element layer of slab r+1 (its connects are the "upper
columns"); no rank ever builds the global mesh: a rank flattens its slab plus one halo layer on each side
and derives the global equation numbers from the (affine in the layer count) size of the mesh below it.

Exchange.  Ghost rows come first in the local CSR, so the staged contributions ARE the prefix A[0:nnz_ghost] (and
rhs[0:nghost]) — nothing is packed; their positions in the owner's CSR are agreed at setup, only doubles travel at
assembly time.  Two transports:
  * exchange="p2p" (default on GPUs): the C ABI's own exchange (include/b200asm.h, b200asm_exchange_*; csrc/exchange.cuh).
    Every rank maps the arrays of the rank below with CUDA IPC; the elements that touch ghost rows are launched first
    and a push kernel on a side stream adds the staged values straight into the owner's memory over NVLink while the
    interior elements are still being assembled; ranks order themselves with step counters in device memory.
    torch.distributed only carries the setup (sizes, index arrays, IPC handles).
  * exchange="nccl": torch.distributed send/recv of the prefix after the kernels + b200asm_scatter_add on the receiver
    (round 1's path; also what the CPU tests run over gloo with an injected local assembler).
"""
from dataclasses import dataclass

import numpy as np

from . import capi, gridmesh
from .gridmesh import DIM, SIDES, FlatMesh, ElementBlock, side_nshape


@dataclass
class SlabMesh:
    rank: int
    world: int
    mesh: FlatMesh              # own elements, extended local numbering
    graph_index: np.ndarray     # element graph (own + upper-halo volume elements) for the pattern builder
    graph: np.ndarray
    nghost: int                 # ghost equations (rows owned by rank-1)
    nown: int
    nupper: int
    row0: int                   # global number of the first own equation
    neq_global: int
    ext2glob: np.ndarray        # extended local equation -> global equation


def _neq_cumulative(nx, ny, porder, nstate, tetrahedra):
    """(a, b) with: equations first touched by the first L >= 1 layers = a + b*L (0 for L = 0)."""
    sizes = []
    for L in (1, 2):
        nodes, blocks = gridmesh.grid_elements((nx, ny, L), tetrahedra=tetrahedra)
        vol = [b for b in blocks if DIM[b[0]] == 3]
        sizes.append(gridmesh.flatten(nodes, vol, porder, nstate).neq)
    b = sizes[1] - sizes[0]
    return sizes[0] - b, b


def slab_mesh(nxy, nz_total, rank, world, porder, nstate, tetrahedra=False, bc_matids=(-1,) * 6, perturb=0.0):
    nx = ny = nxy
    if nz_total % world:
        raise ValueError("nz_total must be a multiple of the number of ranks")
    nzl = nz_total // world
    z0, z1 = rank * nzl, (rank + 1) * nzl
    lo, hi = (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)
    L0, L1 = z0 - lo, z1 + hi
    nodes, blocks = gridmesh.grid_elements((nx, ny, nz_total), tetrahedra=tetrahedra, bc_matids=bc_matids,
                                           perturb=perturb, z_layers=(L0, L1), with_layers=True)
    full = gridmesh.flatten(nodes, [(t, m, e) for t, m, e, _l in blocks], porder, nstate)
    ncon = len(full.block_size)
    size = full.block_size

    vol = full.blocks[0]
    vlayer = blocks[0][3]
    assert DIM[vol.topology] == 3

    def touched(mask):
        t = np.zeros(ncon, dtype=bool)
        t[vol.connects[mask].reshape(-1)] = True
        return t
    t_own = touched((vlayer >= z0) & (vlayer < z1))
    t_low = touched(vlayer < z0)
    t_up = touched(vlayer >= z1)
    ghost = t_low & t_own
    own = t_own & ~t_low
    upper = t_up & ~t_own

    a, b = _neq_cumulative(nx, ny, porder, nstate, tetrahedra)

    def cum(L):
        return 0 if L == 0 else a + b * L
    neq_global = cum(nz_total)
    row0 = cum(z0)

    def excl_cumsum(mask):
        s = np.where(mask, size, 0)
        return np.cumsum(s) - s
    glob = np.full(ncon, -1, dtype=np.int64)   # global position of the first equation of every kept connect
    glob[own] = row0 + excl_cumsum(own)[own]
    if hi:
        glob[upper] = cum(z1) + excl_cumsum(upper)[upper]
    if lo:
        # connects first touched (globally) by the lower halo layer: all of its connects except, when a layer
        # exists below it, those on its bottom plane
        first_by_halo = t_low.copy()
        if L0 > 0:
            plane = (nx + 1) * (ny + 1)
            sides = SIDES[vol.topology]
            halo_el = np.nonzero(vlayer < z0)[0]
            en = vol.elnodes[halo_el].astype(np.int64)
            for s_idx, loc in enumerate(sides):
                on_bottom = np.all(en[:, loc] < plane, axis=1)
                first_by_halo[vol.connects[halo_el[on_bottom], s_idx]] = False
        glob[ghost] = (cum(z0 - 1) + excl_cumsum(first_by_halo))[ghost]

    # extended local numbering: ghost | own | upper, each ascending
    order = np.concatenate([np.nonzero(ghost)[0], np.nonzero(own)[0], np.nonzero(upper)[0]])
    newid = np.full(ncon, -1, dtype=np.int64)
    newid[order] = np.arange(len(order))
    new_size = size[order]
    new_pos = np.concatenate([[0], np.cumsum(new_size)[:-1]]).astype(np.int64)
    nghost = int(size[ghost].sum())
    nown = int(size[own].sum())
    nupper = int(size[upper].sum())
    ext2glob = np.empty(nghost + nown + nupper, dtype=np.int64)
    starts = new_pos
    reps = new_size
    base = np.repeat(glob[order], reps)
    within = np.arange(len(base)) - np.repeat(starts, reps)
    ext2glob[:] = base + within
    assert np.all(np.diff(ext2glob) > 0), "extended numbering must be monotone in the global one"

    mesh = FlatMesh(porder=porder, nstate=nstate, nodes=full.nodes, block_pos=new_pos, block_size=new_size,
                    neq=nghost + nown + nupper)
    first = 0
    graph_parts = []
    for blk, (_t, _m, _e, layer) in zip(full.blocks, blocks):
        keep = (layer >= z0) & (layer < z1)
        conn = newid[blk.connects[keep]]
        if len(conn) == 0:
            continue
        assert conn.min() >= 0
        dest = gridmesh.destination_indices(blk.topology, conn, new_pos, porder, nstate)
        if len(conn):
            mesh.blocks.append(ElementBlock(topology=blk.topology, matid=blk.matid, first=first,
                                            elnodes=np.ascontiguousarray(blk.elnodes[keep]), connects=conn,
                                            dest=np.ascontiguousarray(dest)))
            first += len(conn)
        graph_parts.append(conn)
    if hi:
        graph_parts.append(newid[vol.connects[vlayer >= z1]])  # pattern-only elements
    idx = [0]
    flat = []
    for part in graph_parts:
        nel, ns = part.shape
        flat.append(part.reshape(-1))
        idx.append(idx[-1][-1] + ns * np.arange(1, nel + 1) if isinstance(idx[-1], np.ndarray) else ns * np.arange(1, nel + 1))
    graph_index = np.concatenate([[0]] + [np.asarray(x) for x in idx[1:]]).astype(np.int64) if len(idx) > 1 else np.zeros(1, np.int64)
    graph = np.concatenate(flat).astype(np.int64)
    return SlabMesh(rank=rank, world=world, mesh=mesh, graph_index=graph_index, graph=graph, nghost=nghost,
                    nown=nown, nupper=nupper, row0=row0, neq_global=neq_global, ext2glob=ext2glob)


def build_recv_maps(slab, ia, ja, sender_ext2glob_ghost, sender_ia_ghost, sender_ja_ghost_glob):
    """Positions, in this rank's local CSR / rhs, of the ghost-row entries the rank above sends.

    sender_ext2glob_ghost: global numbers of the sender's ghost equations; sender_ia_ghost[nghost+1];
    sender_ja_ghost_glob: global column numbers of the sender's ghost-row entries (sender CSR order).
    ja: the local column indices, or a callable ja(first, count) that fetches a range of them (the pattern may live
    on the device only; just the receiving rows are needed)."""
    nghost_s = len(sender_ext2glob_ghost)
    rows_glob = np.repeat(sender_ext2glob_ghost, np.diff(sender_ia_ghost))
    lo, hi = slab.row0, slab.row0 + slab.nown
    if not (np.all(rows_glob >= lo) and np.all(rows_glob < hi)):
        raise RuntimeError("received ghost rows that this rank does not own")
    rows_ext = slab.nghost + (rows_glob - lo)
    cols = sender_ja_ghost_glob
    upper_glob = slab.ext2glob[slab.nghost + slab.nown:]
    cols_ext = np.where((cols >= lo) & (cols < hi), slab.nghost + (cols - lo), -1)
    need = cols_ext < 0
    if need.any():
        k = np.searchsorted(upper_glob, cols[need])
        if np.any(k >= len(upper_glob)) or np.any(upper_glob[np.minimum(k, len(upper_glob) - 1)] != cols[need]):
            raise RuntimeError("received a column this rank does not know")
        cols_ext[need] = slab.nghost + slab.nown + k
    # look the (row, col) pairs up in the local CSR rows involved
    urows = np.unique(rows_ext)
    seg_len = ia[urows + 1] - ia[urows]
    seg_pos = np.repeat(ia[urows], seg_len) + (np.arange(seg_len.sum()) - np.repeat(np.cumsum(seg_len) - seg_len, seg_len))
    K = np.int64(len(slab.ext2glob) + 1)
    if callable(ja):
        lo_pos, hi_pos = int(ia[urows.min()]), int(ia[urows.max() + 1])
        ja_seg = ja(lo_pos, hi_pos - lo_pos)[seg_pos - lo_pos]
    else:
        ja_seg = ja[seg_pos]
    keys_local = np.repeat(urows, seg_len) * K + ja_seg
    keys = rows_ext * K + cols_ext
    k = np.searchsorted(keys_local, keys)
    if np.any(k >= len(keys_local)) or np.any(keys_local[np.minimum(k, len(keys_local) - 1)] != keys):
        raise RuntimeError("a received entry has no position in the local CSR pattern")
    a_map = seg_pos[k].astype(np.int32)
    rhs_map = (slab.nghost + (sender_ext2glob_ghost - lo)).astype(np.int32)
    assert len(rhs_map) == nghost_s
    return a_map, rhs_map


class ShardedStructMatrix:
    """One rank's part of a row-sharded TPZSSpStructMatrix/TPZSpStructMatrix assembly on z-slabs.

    backend "cuda": strmatrix.TPZStructMatrixB200 on this rank's GPU, exchange on device buffers (NCCL).
    A `local_assembler(mesh, materials, symmetric, ia, ja) -> (a, rhs)` may be injected instead (the CPU tests
    inject the oracle and run the exchange over gloo)."""

    def __init__(self, slab: SlabMesh, materials, symmetric=True, device=0, local_assembler=None, nthreads=0, engine=None,
                 scatter=None, pattern="host", variant=None, exchange=None):
        """pattern: "host" (threaded host builder, IA/JA kept on the host) or "device" (b200asm_build_pattern_device: the
        column indices stay on the GPU, only the interface rows are ever copied back).
        exchange: "p2p" (default with the CUDA backend) or "nccl", see the module docstring.  With "nccl" the context runs on
        torch's current stream (the collectives order themselves against that stream only)."""
        import torch.distributed as dist
        self.dist = dist
        self.slab = slab
        self.materials = materials
        self.symmetric = symmetric
        self.local_assembler = local_assembler
        self.strmat = None
        self.exchange = exchange or ("p2p" if local_assembler is None else "nccl")
        if local_assembler is None:
            import torch
            from .strmatrix import TPZStructMatrixB200
            self.strmat = TPZStructMatrixB200(slab.mesh, materials, symmetric=symmetric, device=device, nthreads=nthreads,
                                              engine=engine, scatter=scatter, variant=variant)
            if self.exchange == "p2p":
                # the elements that touch ghost rows (local rows [0, nghost)) are stored and launched first
                self.strmat.ctx.set_option("staging_lo", 0)
                self.strmat.ctx.set_option("staging_hi", slab.nghost)
            else:
                self.strmat.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        self.pattern = pattern if local_assembler is None else "host"
        self.nthreads = nthreads
        self.ia = self.ja = None

    # ---- setup: local pattern + exchange maps -----------------------------------------------------
    def Create(self):
        import torch
        s = self.slab
        if self.pattern == "device":
            st = self.strmat
            st._flatten()
            neq, self.nnz = st.ctx.build_pattern_device(self.symmetric, s.graph_index, s.graph, s.mesh.block_pos, s.mesh.block_size)
            assert neq == s.mesh.neq
            self.ia, _ = st.ctx.get_pattern(neq, self.nnz, want_ja=False)
            st.ia, st.ja, st.nnz = self.ia, None, self.nnz
            self.ja = None
            ja_fetch = st.ctx.get_ja_range
        else:
            self.ia, self.ja = capi.build_pattern(self.symmetric, s.graph_index, s.graph, s.mesh.block_pos, s.mesh.block_size,
                                                  self.nthreads)
            self.nnz = len(self.ja)
            if self.strmat is not None:
                self.strmat.SetPattern(self.ia, self.ja)
            ja_fetch = None
        self.nnz_ghost = int(self.ia[s.nghost])
        dist = self.dist
        cuda = self.strmat is not None and dist.get_backend() == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        # every rank tells the rank below which ghost rows it will send (sizes first, then the index arrays)
        up, down = s.rank + 1, s.rank - 1
        self.a_map = self.rhs_map = None
        sizes = torch.zeros(2, dtype=torch.int64, device=dev)
        reqs = []
        if down >= 0:
            reqs.append(dist.isend(torch.tensor([s.nghost, self.nnz_ghost], dtype=torch.int64, device=dev), down))
        if up < s.world:
            reqs.append(dist.irecv(sizes, up))
        for r in reqs:
            r.wait()
        reqs = []
        keep = []
        if down >= 0:
            ja_ghost = ja_fetch(0, self.nnz_ghost) if ja_fetch else self.ja[:self.nnz_ghost]
            for arr in (s.ext2glob[:s.nghost], self.ia[:s.nghost + 1], s.ext2glob[ja_ghost]):
                t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(dev)
                keep.append(t)
                reqs.append(dist.isend(t, down))
        if up < s.world:
            ng, nz = int(sizes[0]), int(sizes[1])
            bufs = [torch.empty(ng, dtype=torch.int64, device=dev), torch.empty(ng + 1, dtype=torch.int64, device=dev),
                    torch.empty(nz, dtype=torch.int64, device=dev)]
            for t in bufs:
                reqs.append(dist.irecv(t, up))
        for r in reqs:
            r.wait()
        if up < s.world:
            g_eq, g_ia, g_ja = [t.cpu().numpy() for t in bufs]
            self.a_map, self.rhs_map = build_recv_maps(s, self.ia, ja_fetch if ja_fetch else self.ja, g_eq, g_ia, g_ja)
            self.recv_nnz, self.recv_neq = len(self.a_map), len(self.rhs_map)
            if self.strmat is not None and self.exchange == "nccl":
                self.a_map_dev = torch.from_numpy(self.a_map).to(dev)
                self.rhs_map_dev = torch.from_numpy(self.rhs_map).to(dev)
                self.recv_a = torch.empty(self.recv_nnz, dtype=torch.float64, device=dev)
                self.recv_rhs = torch.empty(self.recv_neq, dtype=torch.float64, device=dev)
        if self.strmat is not None and self.exchange == "p2p":
            self._attach_peers(dev)
        elif self.strmat is not None:
            a_ptr, r_ptr = self.strmat.ctx.device_pointers()
            self.a_view = _device_view(a_ptr, self.nnz)
            self.rhs_view = _device_view(r_ptr, s.mesh.neq)
        return self.ia, self.ja

    def _attach_peers(self, dev):
        """p2p exchange: the owner of the ghost rows (rank - 1) tells this rank where its staged entries go, every rank maps
        the arrays of its neighbours (CUDA IPC handles travel through torch.distributed) and declares the links.
        Link order of a rank: [push link to rank - 1], [incoming link from rank + 1]."""
        import torch
        dist = self.dist
        s = self.slab
        ctx = self.strmat.ctx
        up, down = s.rank + 1, s.rank - 1
        handles = [None] * s.world
        dist.all_gather_object(handles, ctx.exchange_export())
        # position maps travel back to the sender
        reqs, keep = [], []
        if up < s.world:
            for arr in (self.a_map, self.rhs_map):
                t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32)).to(dev)
                keep.append(t)
                reqs.append(dist.isend(t, up))
        if down >= 0:
            a_dst = torch.empty(self.nnz_ghost, dtype=torch.int32, device=dev)
            rhs_dst = torch.empty(s.nghost, dtype=torch.int32, device=dev)
            reqs += [dist.irecv(a_dst, down), dist.irecv(rhs_dst, down)]
        for r in reqs:
            r.wait()
        if down >= 0:
            link = ctx.exchange_add_peer(True, slot_there=(1 if down > 0 else 0), mem_bytes=handles[down])
            ctx.exchange_set_map(link, 0, a_dst.cpu().numpy(), np.arange(s.nghost, dtype=np.int32), rhs_dst.cpu().numpy())
        if up < s.world:
            ctx.exchange_add_peer(False, slot_there=0, mem_bytes=handles[up],
                                  incoming_min_row=int(self.rhs_map.min()) if len(self.rhs_map) else -1)
        dist.barrier()

    def close(self):
        """Collective: no rank frees the arrays its neighbours still push into."""
        if self.strmat is not None:
            self.strmat.ctx.synchronize()
            self.dist.barrier()
            self.strmat.ctx.close()

    # ---- assembly --------------------------------------------------------------------------------
    def AssembleDevice(self):
        """Device-resident: local kernels, then the interface exchange (async on the current stream)."""
        dist = self.dist
        s = self.slab
        self.strmat.ctx.assemble_async()   # (p2p: the exchange is part of the call)
        if self.exchange == "p2p":
            return
        ops = []
        if s.rank > 0 and s.nghost:
            ops.append(dist.P2POp(dist.isend, self.a_view[:self.nnz_ghost], s.rank - 1))
            ops.append(dist.P2POp(dist.isend, self.rhs_view[:s.nghost], s.rank - 1))
        if self.a_map is not None:
            ops.append(dist.P2POp(dist.irecv, self.recv_a, s.rank + 1))
            ops.append(dist.P2POp(dist.irecv, self.recv_rhs, s.rank + 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        if self.a_map is not None:
            self.strmat.ctx.scatter_add(0, self.a_map_dev.data_ptr(), self.recv_a.data_ptr(), self.recv_nnz)
            self.strmat.ctx.scatter_add(1, self.rhs_map_dev.data_ptr(), self.recv_rhs.data_ptr(), self.recv_neq)

    def Assemble(self):
        """Returns this rank's (a, rhs) in local extended numbering; own rows are complete after the exchange."""
        import torch
        s = self.slab
        if self.strmat is not None:
            self.AssembleDevice()
            a = np.empty(self.nnz)
            rhs = np.empty(s.mesh.neq)
            self.strmat.ctx.download(a, rhs)
            return a, rhs
        a, rhs = self.local_assembler(s.mesh, self.materials, self.symmetric, self.ia, self.ja)
        dist = self.dist
        reqs = []
        if s.rank > 0 and s.nghost:
            reqs.append(dist.isend(torch.from_numpy(a[:self.nnz_ghost].copy()), s.rank - 1))
            reqs.append(dist.isend(torch.from_numpy(rhs[:s.nghost].copy()), s.rank - 1))
        if self.a_map is not None:
            ra = torch.empty(self.recv_nnz, dtype=torch.float64)
            rr = torch.empty(self.recv_neq, dtype=torch.float64)
            reqs.append(dist.irecv(ra, s.rank + 1))
            reqs.append(dist.irecv(rr, s.rank + 1))
        for r in reqs:
            r.wait()
        if self.a_map is not None:
            np.add.at(a, self.a_map, ra.numpy())
            np.add.at(rhs, self.rhs_map, rr.numpy())
        return a, rhs

    def own_rows(self, a, rhs):
        """(global ia offset-free ia, global ja, a, rhs) of the rows this rank owns."""
        s = self.slab
        r0, r1 = s.nghost, s.nghost + s.nown
        lo, hi = self.ia[r0], self.ia[r1]
        ja = self.ja[lo:hi] if self.ja is not None else self.strmat.ctx.get_ja_range(lo, hi - lo)
        return self.ia[r0:r1 + 1] - lo, s.ext2glob[ja], a[lo:hi], rhs[r0:r1]


class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _device_view(ptr, n):
    import torch
    return torch.as_tensor(_DevArr(ptr, n), device="cuda")
