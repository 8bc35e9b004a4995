"""Flattened H1 meshes on structured grids, numbered exactly as NeoPZ numbers them.

Host-side mirror (numpy) of what the reference produces for the benchmark meshes:
  * TPZGeoMeshTools::CreateGeoMeshOnGrid (Mesh/TPZGeoMeshTools.cpp:105-219) with TPZGenGrid3D
    (Pre/TPZGenGrid3D.cpp:51-180 volume elements, :245-400 boundary faces): node id
    iz*(nx+1)*(ny+1)+iy*(nx+1)+ix, hexahedra in (iz,iy,ix) order, 5 tetrahedra per cell with the
    parity rotation, then the z-, y- and x-faces;
  * TPZCompMesh::AutoBuild with SetAllCreateFunctionsContinuous: computational elements are created
    in geometric-element order (Mesh/pzcreateapproxspace.cpp:252-325) and every element asks, side
    by side (side order of Topology/tpzcube.cpp etc.), for the connect of that side, creating it when
    no neighbour has one yet.  Connect index == sequence number (no renumbering,
    TPZLinearAnalysis(cmesh,false)), block positions are the running sum of the block sizes
    (Matrix/pzblock.h) — so first-touch order over (element, side) reproduces the numbering;
  * TPZElementMatrix::ComputeDestinationIndices (Mesh/pzelmat.cpp:37-70).

tests/test_gridmesh.py compares every array against fixtures dumped from the reference.
The result is what the NeoPZ-side flattener (csrc/neopz/TPZStructMatrixB200.cpp) extracts from a
TPZCompMesh, in the layout the C ABI (include/b200asm.h) takes.
"""
from dataclasses import dataclass, field

import numpy as np

from . import capi

# local node ids of every side, in side order
HEX_SIDES = ([[i] for i in range(8)] +
             [[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 5], [2, 6], [3, 7], [4, 5], [5, 6], [6, 7], [7, 4]] +
             [[0, 1, 2, 3], [0, 1, 5, 4], [1, 2, 6, 5], [3, 2, 6, 7], [0, 3, 7, 4], [4, 5, 6, 7]] +
             [[0, 1, 2, 3, 4, 5, 6, 7]])
TET_SIDES = ([[i] for i in range(4)] + [[0, 1], [1, 2], [2, 0], [0, 3], [1, 3], [2, 3]] +
             [[0, 1, 2], [0, 1, 3], [1, 2, 3], [0, 2, 3]] + [[0, 1, 2, 3]])
QUAD_SIDES = [[i] for i in range(4)] + [[0, 1], [1, 2], [2, 3], [3, 0]] + [[0, 1, 2, 3]]
TRI_SIDES = [[i] for i in range(3)] + [[0, 1], [1, 2], [2, 0]] + [[0, 1, 2]]
LINE_SIDES = [[0], [1], [0, 1]]
# Topology/tpzprism.h:275-278, Topology/tpzpyramid.cpp:24-25 (FaceConnectLocId) / :898-925
PRISM_SIDES = ([[i] for i in range(6)] + [[0, 1], [1, 2], [2, 0], [0, 3], [1, 4], [2, 5], [3, 4], [4, 5], [5, 3]] +
               [[0, 1, 2], [0, 1, 4, 3], [1, 2, 5, 4], [0, 2, 5, 3], [3, 4, 5]] + [[0, 1, 2, 3, 4, 5]])
PYRAMID_SIDES = ([[i] for i in range(5)] + [[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 4], [2, 4], [3, 4]] +
                 [[0, 1, 2, 3], [0, 1, 4], [1, 2, 4], [3, 2, 4], [0, 3, 4]] + [[0, 1, 2, 3, 4]])
SIDES = {capi.HEX: HEX_SIDES, capi.TET: TET_SIDES, capi.QUAD: QUAD_SIDES, capi.TRI: TRI_SIDES, capi.LINE: LINE_SIDES,
         capi.PRISM: PRISM_SIDES, capi.PYRAMID: PYRAMID_SIDES}
NCORNER = {capi.HEX: 8, capi.TET: 4, capi.QUAD: 4, capi.TRI: 3, capi.LINE: 2, capi.PRISM: 6, capi.PYRAMID: 5}
DIM = {capi.HEX: 3, capi.TET: 3, capi.QUAD: 2, capi.TRI: 2, capi.LINE: 1, capi.PRISM: 3, capi.PYRAMID: 3}


def side_nshape(topology, side_nodes, p):
    """TSHAPE::NConnectShapeF(side, p) (Shape/pzshapecube.cpp:573-584, pzshapequad.cpp, pzshapetetra.cpp:447-466):
    vertex 1, edge p-1, quadrilateral (p-1)^2, hexahedron interior (p-1)^3, triangle (p-2)(p-1)/2, tetrahedron interior
    sum_{i<p-2} i(i+1)/2; prisms / pyramids are supported for p <= 2."""
    k = len(side_nodes)
    if k == 1:
        return 1
    if k == 2:
        return p - 1
    if topology in (capi.HEX, capi.QUAD):
        return (p - 1) ** 2 if k == 4 else (p - 1) ** 3
    if topology in (capi.TET, capi.TRI):
        # Shape/pzshapetetra.cpp:451-468, pzshapetriang.cpp:452-468: triangular side (p-2)(p-1)/2, tetrahedron interior sum i(i+1)/2
        return (p - 2) * (p - 1) // 2 if k == 3 else sum(i * (i + 1) // 2 for i in range(1, p - 2))
    if p > 2:
        raise ValueError("prisms / pyramids: uniform order p <= 2 only")
    if k == 4:
        return (p - 1) ** 2   # quadrilateral faces (Shape/pzshapeprism.cpp:761-775, pzshapepiram.cpp:679-696)
    return 0


def destination_indices(topology, connects, block_pos, porder, nstate):
    """TPZElementMatrix::ComputeDestinationIndices (Mesh/pzelmat.cpp:37-70): connects in side order, all
    equations of a connect consecutively (shape-major, state fastest)."""
    cols = []
    for s, loc in enumerate(SIDES[topology]):
        n = side_nshape(topology, loc, porder) * nstate
        if n:
            cols.append(block_pos[connects[:, s]][:, None] + np.arange(n, dtype=np.int64)[None, :])
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


@dataclass
class ElementBlock:
    """Consecutive computational elements of one topology and material id."""
    topology: int
    matid: int
    first: int              # index of the first element in the computational mesh
    elnodes: np.ndarray     # [nel][ncorner] int32
    connects: np.ndarray    # [nel][nsides] int64 sequence numbers (TPZCompEl::ConnectIndex -> SequenceNumber)
    dest: np.ndarray = None  # [nel][ndof] int64
    index: np.ndarray = None  # computational-element index of every element when the block is NOT a consecutive run
                              # (meshes whose element types interleave, e.g. hexahedra + pyramids); None: first + arange


@dataclass
class FlatMesh:
    """What the flattener hands to the C ABI (one entry of `blocks` becomes one b200asm_group)."""
    porder: int
    nstate: int
    nodes: np.ndarray                   # [nnodes][3]
    blocks: list = field(default_factory=list)
    block_pos: np.ndarray = None        # TPZBlock::Position per sequence number
    block_size: np.ndarray = None       # TPZBlock::Size
    neq: int = 0

    @property
    def nelements(self):
        return sum(len(b.elnodes) for b in self.blocks)

    def element_graph(self):
        """TPZCompMesh::ComputeElGraph (Mesh/pzcmesh.cpp:1223-1267): seqnums of all connects, element order."""
        idx = [0]
        parts = []
        for b in sorted(self.blocks, key=lambda b: b.first):
            nel, ns = b.connects.shape
            parts.append(b.connects.reshape(-1))
            idx.append(idx[-1] + nel * ns)
        graph = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        index = np.concatenate([np.arange(idx[k], idx[k + 1], b.connects.shape[1], dtype=np.int64)
                                for k, b in enumerate(sorted(self.blocks, key=lambda b: b.first))] +
                               [np.array([idx[-1]], dtype=np.int64)])
        return index, graph


def face_outward(mesh, block):
    """For every boundary face of `block`: centre of the face minus centre of the volume element it bounds - the vector
    TPZInterpolationSpace::ComputeNormal (Mesh/pzinterpolationspace.cpp:338-381) turns data.normal towards.  The centres are
    the images of the master-element centres (TPZGeoEl::CenterPoint + X), i.e. the corner averages of (multi)linear maps.
    Only its direction relative to the face matters.  [nel][3]; zero where no volume element owns the face."""
    fn = np.sort(np.asarray(block.elnodes, dtype=np.int64), axis=1)[:, :3]   # the three smallest nodes identify a face
    nn = np.int64(len(mesh.nodes) + 1)
    if float(nn) ** 3 >= 9.2e18:
        raise ValueError("face_outward: too many nodes for the packed face key")
    key = (fn[:, 0] * nn + fn[:, 1]) * nn + fn[:, 2]
    out = np.zeros((len(fn), 3))
    found = np.zeros(len(fn), dtype=bool)
    fc = mesh.nodes[block.elnodes].mean(axis=1)
    for vb in mesh.blocks:
        if DIM[vb.topology] != 3:
            continue
        en = np.asarray(vb.elnodes, dtype=np.int64)
        vc = mesh.nodes[vb.elnodes].mean(axis=1)
        for loc in SIDES[vb.topology]:
            if len(loc) not in (3, 4) or len(loc) == NCORNER[vb.topology]:
                continue
            sn = np.sort(en[:, loc], axis=1)[:, :3]
            vkey = (sn[:, 0] * nn + sn[:, 1]) * nn + sn[:, 2]
            order = np.argsort(vkey, kind="stable")
            pos = np.searchsorted(vkey[order], key)
            pos = np.minimum(pos, len(vkey) - 1)
            hit = (vkey[order][pos] == key) & ~found
            out[hit] = fc[hit] - vc[order[pos[hit]]]
            found |= hit
    return out


def _first_touch_ids(keys):
    """Rank of first occurrence: keys[k] -> index in order of first appearance (greedy connect creation)."""
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq), dtype=np.int64)
    return rank[inv], len(uniq)


def flatten(nodes, element_blocks, porder, nstate):
    """Number connects / equations of a conforming H1 mesh the way NeoPZ's AutoBuild does.

    element_blocks: list of (topology, matid, elnodes[nel][ncorner]) in computational-element order, or
    (topology, matid, elnodes, index) with the computational-element index of every element when elements of different
    blocks interleave (connects are created in element-index order whatever the block layout).
    """
    nnodes = len(nodes)
    # entity keys per (element, side) as (class, ab, c) triples, in element order
    triples = []
    shapes = []
    ncell = 0
    indexed = any(len(b) > 3 and b[3] is not None for b in element_blocks)
    touch = []  # creation position (element index, side) of every (element, side) entry
    for blk in element_blocks:
        topo, _matid, elnodes = blk[:3]
        sides = SIDES[topo]
        en = np.asarray(elnodes, dtype=np.int64)
        nel = en.shape[0]
        t = np.empty((nel, len(sides), 3), dtype=np.int64)
        for s, loc in enumerate(sides):
            k = len(loc)
            if k == 1:
                t[:, s, 0] = 0
                t[:, s, 1] = en[:, loc[0]]
                t[:, s, 2] = -1
            else:
                srt = np.sort(en[:, loc], axis=1)
                iscell = (k == NCORNER[topo] and DIM[topo] == 3)
                t[:, s, 0] = 1 if k == 2 else (3 if iscell else 2)
                if iscell:  # the volume side is never shared: keyed by the element itself
                    t[:, s, 1] = ncell + np.arange(nel, dtype=np.int64)
                    t[:, s, 2] = -1
                else:       # edges: both nodes; faces: the three smallest nodes (distinct faces share at most two)
                    t[:, s, 1] = srt[:, 0] * nnodes + srt[:, 1]
                    t[:, s, 2] = srt[:, 2] if k > 2 else -1
        if DIM[topo] == 3:
            ncell += nel
        triples.append(t.reshape(-1, 3))
        shapes.append((nel, len(sides)))
        if indexed:
            if len(blk) < 4 or blk[3] is None:
                raise ValueError("flatten: either every block carries element indices or none does")
            idx = np.asarray(blk[3], dtype=np.int64)
            touch.append((idx[:, None] * 32 + np.arange(len(sides), dtype=np.int64)[None, :]).reshape(-1))
    allt = np.concatenate(triples, axis=0)
    # two-level packing: (class, ab) -> dense id, then (id, c) -> one int64
    lvl1 = allt[:, 0] * (np.int64(nnodes) * nnodes) + allt[:, 1]
    _, inv1 = np.unique(lvl1, return_inverse=True)
    key = inv1.astype(np.int64) * (nnodes + 1) + (allt[:, 2] + 1)
    if indexed:  # first touch in element-index order, not in block order
        by_creation = np.argsort(np.concatenate(touch), kind="stable")
        ranked, nconnects = _first_touch_ids(key[by_creation])
        conn_flat = np.empty_like(ranked)
        conn_flat[by_creation] = ranked
    else:
        conn_flat, nconnects = _first_touch_ids(key)

    # block sizes: nshape(side) * nstate of the side that created the connect (same for all sharers)
    size = np.zeros(nconnects, dtype=np.int64)
    off = 0
    for blk, (nel, ns) in zip(element_blocks, shapes):
        topo = blk[0]
        nsh = np.array([side_nshape(topo, loc, porder) for loc in SIDES[topo]], dtype=np.int64) * nstate
        size[conn_flat[off:off + nel * ns]] = np.tile(nsh, nel)
        off += nel * ns
    pos = np.concatenate([[0], np.cumsum(size)[:-1]]).astype(np.int64)
    neq = int(size.sum())

    mesh = FlatMesh(porder=porder, nstate=nstate, nodes=np.ascontiguousarray(nodes, dtype=np.float64),
                    block_pos=pos, block_size=size, neq=neq)
    off = 0
    first = 0
    for blk, (nel, ns) in zip(element_blocks, shapes):
        topo, matid, elnodes = blk[:3]
        conn = conn_flat[off:off + nel * ns].reshape(nel, ns)
        off += nel * ns
        dest = destination_indices(topo, conn, pos, porder, nstate)
        index = np.asarray(blk[3], dtype=np.int64) if indexed else None
        mesh.blocks.append(ElementBlock(topology=topo, matid=matid, first=int(index.min()) if indexed and nel else first,
                                        elnodes=np.ascontiguousarray(elnodes, dtype=np.int32),
                                        connects=conn, dest=np.ascontiguousarray(dest), index=index))
        first += nel
    return mesh


def grid_nodes(n, min_x=(0., 0., 0.), max_x=(1., 1., 1.), perturb=0.0, z_layers=None):
    """Node coordinates of the grid (or of the node planes iz0..iz1 of a slab): Pre/TPZGenGrid3D.cpp:72-74, plus the
    deterministic smooth perturbation of oracle/refdriver.cpp that makes the Jacobians non-constant."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    iz0, iz1 = (0, nz) if z_layers is None else z_layers
    min_x = np.asarray(min_x, dtype=np.float64)
    max_x = np.asarray(max_x, dtype=np.float64)
    sy = nx + 1
    sz = (nx + 1) * (ny + 1)
    # node id = iz*(nx+1)*(ny+1) + iy*(nx+1) + ix
    K, J, I = np.meshgrid(np.arange(iz0, iz1 + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    nodes = np.empty((len(I), 3))
    # fMinX + ((fMaxX-fMinX) * i)/nel   (Pre/TPZGenGrid3D.cpp:72-74)
    nodes[:, 0] = min_x[0] + ((max_x[0] - min_x[0]) * I) / nx
    nodes[:, 1] = min_x[1] + ((max_x[1] - min_x[1]) * J) / ny
    nodes[:, 2] = min_x[2] + ((max_x[2] - min_x[2]) * K) / nz
    if perturb != 0.0:
        h = 1.0 / nx
        ids = (K * sz + J * sy + I).astype(np.float64)  # global node ids
        for d in range(3):
            nodes[:, d] += perturb * h * np.sin(2.0 * np.pi * ids / 97.0 + float(d))
    return nodes


def grid_elements(n, tetrahedra=False, bc_matids=(-1, -1, -1, -1, -1, -1), vol_matid=1,
                  min_x=(0., 0., 0.), max_x=(1., 1., 1.), perturb=0.0, z_layers=None, with_layers=False, prisms=False):
    """Nodes and element blocks of CreateGeoMeshOnGrid(3, minX, maxX, matids, {nx,ny,nz}, type, createBoundEls=true).

    n: divisions per direction (int or 3-tuple).  bc_matids = matids[1..6] of the reference call, i.e. the
    arguments of BuildBoundaryElements(Zmin, Xmin, Ymin, Xmax, Ymax, Zmax) (Pre/TPZGenGrid3D.cpp:230-232).
    z_layers=(iz0, iz1): only the element layers iz0 <= iz < iz1 of the global grid (slab of a sharded mesh);
    node indices are then local to the slab (global id - iz0*(nx+1)*(ny+1)), coordinates and the element /
    face parities are those of the global grid.  with_layers: blocks carry a 4th entry, the global layer of
    every element.  prisms: MMeshType::EPrismatic, two prisms (0,1,2,4,5,6), (0,2,3,4,6,7) per cell (Pre/TPZGenGrid3D.cpp:163-179),
    two triangles (0,1,2), (0,3,2) per z-face and one quadrilateral per x- / y-face (:297-307, :347-354, :385-392).
    """
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    iz0, iz1 = (0, nz) if z_layers is None else z_layers
    nzl = iz1 - iz0
    if prisms and tetrahedra:
        raise ValueError("grid_elements: tetrahedra and prisms are exclusive")
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    nodes = grid_nodes(n, min_x=min_x, max_x=max_x, perturb=perturb, z_layers=z_layers)

    ez, ey, ex = np.meshgrid(np.arange(nzl), np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey, ez = ex.reshape(-1), ey.reshape(-1), ez.reshape(-1)
    first = ez * sz + ey * sy + ex
    cube = np.stack([first, first + 1, first + 1 + sy, first + sy,
                     first + sz, first + 1 + sz, first + 1 + sy + sz, first + sy + sz], axis=1)
    blocks = []  # [topology, matid, elnodes, layer]
    if prisms:
        pr = np.stack([cube[:, [0, 1, 2, 4, 5, 6]], cube[:, [0, 2, 3, 4, 6, 7]]], axis=1).reshape(-1, 6)
        blocks.append([capi.PRISM, vol_matid, pr, np.repeat(ez + iz0, 2)])
    elif not tetrahedra:
        blocks.append([capi.HEX, vol_matid, cube, ez + iz0])
    else:
        perm = (ex + ey + ez + iz0) % 2
        rows = np.arange(len(cube))

        def cn(k, top=False):  # cubenode[(k+permut)%4 (+4)]
            return cube[rows, (k + perm) % 4 + (4 if top else 0)]
        t0 = np.stack([cn(0), cn(1), cn(3), cn(0, True)], axis=1)
        t1 = np.stack([cn(1, True), cn(0, True), cn(2, True), cn(1)], axis=1)
        t2 = np.stack([cn(2), cn(3), cn(1), cn(2, True)], axis=1)
        t3 = np.stack([cn(3, True), cn(2, True), cn(0, True), cn(3)], axis=1)
        t4 = np.stack([cn(0, True), cn(1), cn(3), cn(2, True)], axis=1)
        tets = np.stack([t0, t1, t2, t3, t4], axis=1).reshape(-1, 4)
        blocks.append([capi.TET, vol_matid, tets, np.repeat(ez + iz0, 5)])

    m_zmin, m_xmin, m_ymin, m_xmax, m_ymax, m_zmax = bc_matids
    per_face = 2 if tetrahedra else 1

    def emit(matid, faces, layer):
        face_topo = capi.TRI if faces.shape[1] == 3 else capi.QUAD
        layer = np.repeat(layer, len(faces) // len(layer))
        if blocks[-1][0] == face_topo and blocks[-1][1] == matid:
            blocks[-1][2] = np.concatenate([blocks[-1][2], faces], axis=0)
            blocks[-1][3] = np.concatenate([blocks[-1][3], layer], axis=0)
        else:
            blocks.append([face_topo, matid, faces, layer])

    # top/bottom: iZ in {0, nz}; loops iY, iX
    for izf, matid, present in ((0, m_zmin, iz0 == 0), (nz, m_zmax, iz1 == nz)):
        if not present:
            continue
        fy, fx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        fx, fy = fx.reshape(-1), fy.reshape(-1)
        f = (izf - iz0) * sz + fy * sy + fx
        if prisms:
            faces = np.stack([np.stack([f, f + 1, f + sy + 1], axis=1), np.stack([f, f + sy, f + sy + 1], axis=1)],
                             axis=1).reshape(-1, 3)
        elif not tetrahedra:
            faces = np.stack([f, f + 1, f + sy + 1, f + sy], axis=1)
        else:
            odd = ((fx + fy + izf) % 2) == 1
            ta = np.stack([f, f + 1, f + sy + np.where(odd, 1, 0)], axis=1)
            tb = np.stack([f + np.where(odd, 0, 1), f + sy, f + sy + 1], axis=1)
            faces = np.stack([ta, tb], axis=1).reshape(-1, 3)
        emit(matid, faces, np.full(len(f), 0 if izf == 0 else nz - 1))
    # left/right: for iZ: for iY in {0, ny}: for iX
    for izl in range(nzl):
        for iyf, matid in ((0, m_ymin), (ny, m_ymax)):
            fx = np.arange(nx)
            f = izl * sz + iyf * sy + fx
            cnt = fx + iyf + izl + iz0
            if not tetrahedra:
                faces = np.stack([f, f + 1, f + sz + 1, f + sz], axis=1)
            else:
                odd = (cnt % 2) == 1
                ta = np.stack([f, f + 1, f + sz + np.where(odd, 1, 0)], axis=1)
                tb = np.stack([f + np.where(odd, 0, 1), f + sz + 1, f + sz], axis=1)
                faces = np.stack([ta, tb], axis=1).reshape(-1, 3)
            emit(matid, faces, np.full(len(f), izl + iz0))
    # front/back: for iZ: for iY: for iX in {0, nx}
    fz, fy = np.meshgrid(np.arange(nzl), np.arange(ny), indexing="ij")
    fz, fy = fz.reshape(-1), fy.reshape(-1)
    pieces = []
    for ixf, matid in ((0, m_xmin), (nx, m_xmax)):
        f = fz * sz + fy * sy + ixf
        cnt = ixf + fy + fz + iz0
        if not tetrahedra:
            faces = np.stack([f, f + sy, f + sz + sy, f + sz], axis=1)[:, None, :]
        else:
            odd = (cnt % 2) == 1
            ta = np.stack([f, f + sy, f + sz + np.where(odd, sy, 0)], axis=1)
            tb = np.stack([f + np.where(odd, 0, sy), f + sz + sy, f + sz], axis=1)
            faces = np.stack([ta, tb], axis=1)  # [nface][2][3]
        pieces.append((matid, faces))
    # interleaved: for each (iZ,iY): the xmin face(s) then the xmax face(s)
    (m0, f0), (m1, f1) = pieces
    if m0 == m1:
        inter = np.stack([f0, f1], axis=1)  # [k][2][per_face][nc]
        emit(m0, inter.reshape(-1, f0.shape[-1]), np.repeat(fz + iz0, 2))
    else:
        for r in range(f0.shape[0]):
            emit(m0, f0[r].reshape(-1, f0.shape[-1]), np.array([fz[r] + iz0]))
            emit(m1, f1[r].reshape(-1, f1.shape[-1]), np.array([fz[r] + iz0]))
    if with_layers:
        return nodes, [tuple(b) for b in blocks]
    return nodes, [tuple(b[:3]) for b in blocks]


def grid_elements_2d(n, triangles=False, bc_matids=(-1, -1, -1, -1), dom_matid=1, min_x=(0., 0.), max_x=(1., 1.), perturb=0.0):
    """Nodes and element blocks of CreateGeoMeshOnGrid(2, minX, maxX, matids, {nx,ny}, type, createBoundEls=true)
    (Mesh/TPZGeoMeshTools.cpp:186-199 with Pre/TPZGenGrid2D.cpp): node id iy*(nx+1)+ix at x0 + ix*((x1-x0)/nx) (:545-606),
    quadrilaterals row by row (:625-636), two triangles (0,1,2), (0,2,3) per cell (:486-494), then the line elements of
    the sides 4 (bottom), 5 (right), 6 (top), 7 (left) in that order (SetBC, :706-765).  bc_matids = matids[1..4]."""
    nx, ny = (n, n) if np.isscalar(n) else n
    sy = nx + 1
    J, I = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    I, J = I.reshape(-1), J.reshape(-1)
    nodes = np.zeros((len(I), 3))
    dx, dy = (max_x[0] - min_x[0]) / nx, (max_x[1] - min_x[1]) / ny
    nodes[:, 0] = min_x[0] + I * dx
    nodes[:, 1] = min_x[1] + J * dy
    if perturb != 0.0:
        h = 1.0 / nx
        ids = (J * sy + I).astype(np.float64)
        for d in range(2):  # plane meshes stay in z = 0 (oracle/refdriver.cpp)
            nodes[:, d] += perturb * h * np.sin(2.0 * np.pi * ids / 97.0 + float(d))
    ey, ex = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey = ex.reshape(-1), ey.reshape(-1)
    f = ey * sy + ex
    cell = np.stack([f, f + 1, f + 1 + sy, f + sy], axis=1)  # [ncell][4]
    blocks = []
    if not triangles:
        blocks.append([capi.QUAD, dom_matid, cell])
    else:
        tri = np.stack([cell[:, [0, 1, 2]], cell[:, [0, 2, 3]]], axis=1).reshape(-1, 3)
        blocks.append([capi.TRI, dom_matid, tri])

    def emit(matid, lines):
        if blocks[-1][0] == capi.LINE and blocks[-1][1] == matid:
            blocks[-1][2] = np.concatenate([blocks[-1][2], lines], axis=0)
        else:
            blocks.append([capi.LINE, matid, lines])
    row0, rowl = cell[:nx], cell[nx * (ny - 1):]
    col0, coll = cell[0::nx], cell[nx - 1::nx]
    emit(bc_matids[0], row0[:, [0, 1]])   # side 4 of the bottom row
    emit(bc_matids[1], coll[:, [1, 2]])   # side 5 of the last column
    emit(bc_matids[2], rowl[:, [2, 3]])   # side 6 of the top row
    emit(bc_matids[3], col0[:, [3, 0]])   # side 7 of the first column
    return nodes, [tuple(b) for b in blocks]


def _renumber_nodes(nodes, blocks, node_perm):
    """node_perm[i] = new index of grid node i (a renumbered mesh: same geometry, different side orientations)."""
    if node_perm is None:
        return nodes, blocks
    node_perm = np.asarray(node_perm, dtype=np.int64)
    renum = np.empty_like(nodes)
    renum[node_perm] = nodes
    return renum, [(t, m, node_perm[np.asarray(e, dtype=np.int64)]) for t, m, e in blocks]


def grid_mesh_2d(n, porder, nstate, triangles=False, bc_matids=(-1,) * 4, perturb=0.0, node_perm=None):
    nodes, blocks = grid_elements_2d(n, triangles=triangles, bc_matids=bc_matids, perturb=perturb)
    nodes, blocks = _renumber_nodes(nodes, blocks, node_perm)
    return flatten(nodes, blocks, porder, nstate)


def mesh_from_elements(nodes, el_type, el_matid, el_nodes, porder, nstate):
    """FlatMesh of an arbitrary conforming mesh given element by element in computational-element order (topology code,
    material id, corner nodes padded with -1): elements are gathered into one block per (topology, material id) in order of
    first appearance, the connects are numbered in element order (AutoBuild)."""
    el_type = np.asarray(el_type)
    el_matid = np.asarray(el_matid)
    el_nodes = np.asarray(el_nodes, dtype=np.int64)
    blocks, seen = [], {}
    for e in range(len(el_type)):
        k = (int(el_type[e]), int(el_matid[e]))
        if k not in seen:
            seen[k] = len(blocks)
            blocks.append([k[0], k[1], []])
        blocks[seen[k]][2].append(e)
    out = []
    for topo, matid, idx in blocks:
        idx = np.array(idx, dtype=np.int64)
        out.append((topo, matid, el_nodes[idx, :NCORNER[topo]], idx))
    return flatten(np.asarray(nodes, dtype=np.float64), out, porder, nstate)


def hexpyr_elements(n, bc_matids=(-1, -1, -1, -1, -1, -1), vol_matid=1, perturb=0.0):
    """Nodes and elements (in computational-element order) of CreateGeoMeshOnGrid(..., MMeshType::EHexaPyrMixed, createBoundEls=true):
    every other cell ((iel + iel/nx + iel/(nx ny)) even) is split into six pyramids around a new centre node by the refinement
    pattern of Pre/TPZGenGrid3D.cpp:181-228; the sons and, later, the boundary quadrilaterals take the element slots the deleted
    hexahedra freed (TPZAdmChunkVector::AllocateNewElement pops the free stack first), so the element types interleave.
    Returns (nodes, el_type, el_matid, el_nodes[nel][8] padded with -1) ready for mesh_from_elements."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    sy, sz = nx + 1, (nx + 1) * (ny + 1)
    grid = grid_nodes(n, perturb=0.0)
    ncell = nx * ny * nz
    cells = []
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                f = iz * sz + iy * sy + ix
                cells.append([f, f + 1, f + 1 + sy, f + sy, f + sz, f + 1 + sz, f + 1 + sy + sz, f + sy + sz])
    slots = {i: (capi.HEX, vol_matid, c) for i, c in enumerate(cells)}
    free, nxt = [], ncell
    centres = []
    sons = ((0, 1, 2, 3), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0), (7, 6, 5, 4))

    def allocate():
        nonlocal nxt
        if free:
            return free.pop()
        nxt += 1
        return nxt - 1
    for iel in range(ncell):
        if (iel + iel // nx + iel // (nx * ny)) % 2 == 1:
            continue
        c = cells[iel]
        centre = len(grid) + len(centres)
        x = np.zeros(3)
        for a in range(8):   # X(0,0,0) of the trilinear map: corner functions 1/8, accumulated in node order
            x = x + 0.125 * grid[c[a]]
        centres.append(x)
        for s in sons:
            slots[allocate()] = (capi.PYRAMID, vol_matid, [c[s[0]], c[s[1]], c[s[2]], c[s[3]], centre])
        del slots[iel]
        free.append(iel)
    m_zmin, m_xmin, m_ymin, m_xmax, m_ymax, m_zmax = bc_matids
    for izf, matid in ((0, m_zmin), (nz, m_zmax)):
        for iy in range(ny):
            for ix in range(nx):
                f = izf * sz + iy * sy + ix
                slots[allocate()] = (capi.QUAD, matid, [f, f + 1, f + sy + 1, f + sy])
    for iz in range(nz):
        for iyf, matid in ((0, m_ymin), (ny, m_ymax)):
            for ix in range(nx):
                f = iz * sz + iyf * sy + ix
                slots[allocate()] = (capi.QUAD, matid, [f, f + 1, f + sz + 1, f + sz])
    for iz in range(nz):
        for iy in range(ny):
            for ixf, matid in ((0, m_xmin), (nx, m_xmax)):
                f = iz * sz + iy * sy + ixf
                slots[allocate()] = (capi.QUAD, matid, [f, f + sy, f + sz + sy, f + sz])
    nodes = np.concatenate([grid, np.array(centres).reshape(-1, 3)], axis=0)
    if perturb != 0.0:  # the deterministic perturbation of oracle/refdriver.cpp, by node index (centre nodes included)
        h = 1.0 / nx
        ids = np.arange(len(nodes), dtype=np.float64)
        for d in range(3):
            nodes[:, d] += perturb * h * np.sin(2.0 * np.pi * ids / 97.0 + float(d))
    order = sorted(slots)
    el_type = np.array([slots[i][0] for i in order], dtype=np.int32)
    el_matid = np.array([slots[i][1] for i in order], dtype=np.int32)
    el_nodes = np.full((len(order), 8), -1, dtype=np.int64)
    for k, i in enumerate(order):
        el_nodes[k, : len(slots[i][2])] = slots[i][2]
    return nodes, el_type, el_matid, el_nodes


def hexpyr_mesh(n, porder, nstate, bc_matids=(-1,) * 6, perturb=0.0):
    """Flattened mesh of hexahedra and pyramids side by side (MMeshType::EHexaPyrMixed), numbered as the reference numbers it."""
    nodes, el_type, el_matid, el_nodes = hexpyr_elements(n, bc_matids=bc_matids, perturb=perturb)
    return mesh_from_elements(nodes, el_type, el_matid, el_nodes, porder, nstate)


def grid_mesh(n, porder, nstate, tetrahedra=False, bc_matids=(-1,) * 6, perturb=0.0, node_perm=None, prisms=False):
    """node_perm[i] = new index of grid node i (a renumbered mesh: same geometry, different side orientations)."""
    nodes, blocks = grid_elements(n, tetrahedra=tetrahedra, bc_matids=bc_matids, perturb=perturb, prisms=prisms)
    nodes, blocks = _renumber_nodes(nodes, blocks, node_perm)
    return flatten(nodes, blocks, porder, nstate)
