"""Round-2 additions on the GPU:
  * TPZMatPoisson::ContributeBC type 2 and TPZElasticity3D::ContributeBC type 4 (the last two boundary types of SURVEY R20):
    reference fixtures are part of golden_util.CORE_CASES (tests/test_gpu_parity.py runs them); here the Python mirror against the
    oracle on fresh meshes and the C++ strategy inside the unmodified reference;
  * several load cases of TPZMatPoisson (rhs with NumLoadCases() columns);
  * SetDropTinyEntries (TPZSYsmpMatrix::AddKel's IsZero drop, Matrix/pzsysmp.cpp:381) on micro-scale geometry;
  * SetAccumulate (AddKel adds to what the matrix holds);
  * several GPUs behind the strategy: SetNumThreads(n) (StrMatrix/TPZStrMatParInterface.h:62-69) -> b200asm_multi_*;
  * the mesh signature: an Assemble after a connect renumbering / a changed material-id filter re-flattens;
  * parity at BASELINE sizes: C1's literal 32^3 p1 and the survey's probe sizes through the drop-in binary (threaded OR as the
    reference: bit-identical to serial, pzstrmatrixor.cpp:714-717)."""
import json
import os
import subprocess

import numpy as np
import pytest

from neopz_b200 import gridmesh, strmatrix as sm
from tests import golden_util as gu
from tests.oracle_ref import oracle_assemble
from tests.test_gpu_parity import TOL, materials_for, relF

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_bin", "dropin_test")


def dropin(case, env=None, extra=("4", "0"), timeout=900):
    if not os.path.exists(BIN):
        pytest.skip("tests/_bin/dropin_test not built (needs /root/reference at build time)")
    e = dict(os.environ)
    e.update({k: str(v) for k, v in (env or {}).items()})
    out = subprocess.run([BIN] + [str(x) for x in case] + list(extra), capture_output=True, text=True, timeout=timeout, env=e)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(lines[-1]), out.returncode


def check(r, rc):
    assert r["ia_identical"] == 1 and r["ja_identical"] == 1
    assert r["relF_A"] <= 1e-12 and r["relF_A_nonpenalty_rows"] <= 1e-12 and r["relF_rhs"] <= 1e-12
    assert r["relF_residual_rhs"] <= 1e-12
    assert rc == 0


@pytest.mark.parametrize("tet", [False, True])
@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("symmetric", [True, False])
def test_poisson_bc_type2_against_oracle(tet, p, symmetric):
    n = 4 if p < 3 else 3
    perm = np.random.default_rng(5).permutation((n + 1) ** 3) if p >= 3 else None   # p >= 3: oriented sides
    mesh = gridmesh.grid_mesh(n, p, 1, tetrahedra=tet, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12, node_perm=perm)
    mats = materials_for(0)
    mats[-2] = mats[1].CreateBC(-2, 2, [[3.0e-15]], [0.4])
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    ia, ja, a, rhs = st.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    # rows of the type-2 face carry no penalty mass term: compare them alone as well
    rows = np.repeat(np.arange(mesh.neq), np.diff(ia))
    big = np.zeros(mesh.neq, dtype=bool)
    big[rows[np.abs(a_ref) > 1e9]] = True
    keep = ~big[rows]
    assert np.linalg.norm((a - a_ref)[keep]) <= TOL * np.linalg.norm(a_ref[keep])
    st.ctx.close()


@pytest.mark.parametrize("tet,p", [(False, 1), (False, 2), (True, 2), (True, 1), (False, 3)])
def test_elasticity_bc_type4_against_oracle(tet, p):
    mesh = gridmesh.grid_mesh(3, p, 3, tetrahedra=tet, bc_matids=(-1, -4, -1, -4, -1, -4), perturb=0.12)
    mats = materials_for(1)
    v1 = np.array([[4.0, 0.5, -0.3], [0.5, 3.0, 0.25], [-0.3, 0.25, 5.0]])
    mats[-4] = mats[1].CreateBC(-4, 4, v1, np.zeros(3))
    for symmetric in (True, False):
        st = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        ia, ja, a, rhs = st.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        # the traction of a constant stress field over the three faces is in the load vector: not zero
        assert np.linalg.norm(rhs_ref) > 0
        st.ctx.close()


def test_face_normals_point_outwards():
    mesh = gridmesh.grid_mesh(3, 1, 3, bc_matids=(-1, -2, -3, -4, -5, -6), perturb=0.0)
    want = {-1: (0, 0, -1), -2: (-1, 0, 0), -3: (0, -1, 0), -4: (1, 0, 0), -5: (0, 1, 0), -6: (0, 0, 1)}
    qpts, _qw, _phi, _dphi = sm.element_tables(2, 1)
    for b in mesh.blocks:
        if b.matid == 1:
            continue
        n = sm.face_normals(b.topology, qpts, mesh.nodes[b.elnodes], gridmesh.face_outward(mesh, b))
        assert np.allclose(n, np.array(want[b.matid], dtype=float)[None, None, :], atol=1e-14)


@pytest.mark.parametrize("case,bctype", [((4, 2, 0, 0, 1, 0), 2), ((3, 2, 0, 1, 0, 0), 2), ((3, 3, 0, 0, 1, 0), 2),
                                         ((4, 2, 1, 0, 1, 0), 4), ((3, 2, 1, 1, 0, 0), 4), ((4, 1, 1, 0, 1, 0), 4),
                                         ((3, 2, 1, 0, 1, 0), 3), ((3, 2, 1, 0, 1, 0), 7)])
def test_dropin_boundary_types(case, bctype):
    r, rc = dropin(case, {"B200_BCTYPE": bctype})
    assert r["bctype"] == bctype
    check(r, rc)


@pytest.mark.parametrize("case", [(4, 2, 0, 0, 1, 0), (4, 2, 0, 1, 0, 0), (3, 3, 0, 0, 1, 0), (5, 1, 0, 0, 1, 0)])
@pytest.mark.parametrize("bcfunc", [0, 1])
def test_dropin_load_cases(case, bcfunc):
    """TPZMatPoisson with three load cases: fRhs has three columns (Analysis/TPZLinearAnalysis.cpp:70-72), the forcing function
    and the boundary values differ per case (TPZMatPoisson.cpp:23-41,56-100); relF_rhs covers all columns."""
    r, rc = dropin(case, {"B200_LOADCASES": 3}, extra=("4", "0", "0", "0", "0", str(bcfunc)))
    assert r["loadcases"] == 3
    check(r, rc)


def test_dropin_drop_tiny_entries():
    """A small TPZMatPoisson scale factor (1e-10) puts 13 % of the element entries below 1e-12; TPZSYsmpMatrix::AddKel drops them
    (pzsysmp.cpp:381; 2 % of the Frobenius norm of the element matrices, counted with the oracle).  With SetDropTinyEntries(true)
    the GPU result equals the reference's on the rows without penalty entries; without it the difference is visible (the
    documented deviation of the default).  (Micro-scale GEOMETRY does not get there in the reference: its Jacobian
    determinant is clamped at 1e-12 first, Mesh/pzgeoel.cpp:1316-1326.)"""
    for case in ((4, 2, 0, 0, 1, 0), (4, 2, 0, 0, 0, 0)):
        r, rc = dropin(case, {"B200_SCALE": 1e-10, "B200_DROPTINY": 1})
        assert r["droptiny"] == 1 and r["ia_identical"] == 1 and r["ja_identical"] == 1
        assert r["relF_A_nonpenalty_rows"] <= 1e-12 and r["relF_rhs"] <= 1e-12 and r["relF_A"] <= 1e-12
        r2, _rc2 = dropin(case, {"B200_SCALE": 1e-10, "B200_DROPTINY": 0})
        assert r2["relF_A_nonpenalty_rows"] > 1e-6, "the scaled material does not exercise the drop"


@pytest.mark.parametrize("case", [(4, 2, 0, 0, 1, 0), (3, 2, 1, 0, 1, 0)])
def test_dropin_accumulate(case):
    r, rc = dropin(case, {"B200_ACCUMULATE": 1})
    assert r["relF_A_accumulated_twice"] <= 1e-12
    check(r, rc)


@pytest.mark.parametrize("gpus", [2, 4])
@pytest.mark.parametrize("case,extra", [((6, 2, 0, 0, 1, 1), ("4", "0")), ((4, 2, 1, 0, 1, 0), ("4", "0")), ((4, 2, 0, 1, 0, 0), ("4", "0")),
                                        ((4, 3, 0, 0, 1, 0), ("4", "0")), ((5, 2, 0, 0, 1, 0), ("4", "1")),
                                        ((8, 2, 2, 0, 1, 0), ("4", "0")), ((5, 2, 0, 0, 1, 0), ("4", "0", "1")),
                                        ((5, 2, 1, 0, 1, 0), ("4", "0", "0", "1"))])
def test_dropin_several_gpus(gpus, case, extra):
    """TPZStructMatrixB200::SetNumThreads(n): n GPUs share one TPZLinearAnalysis::Assemble() of a TPZCompMesh (element partition by
    the smallest destination equation, row-sharded CSR, interface rows over NVLink); IA / JA memcmp and 1e-12 against OR.
    extra: device-side Create(), an equation filter, a page-locked host matrix."""
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    r, rc = dropin(case, {"B200_GPUS": gpus}, extra=extra)
    assert r["gpus"] == gpus
    check(r, rc)
    if case[5]:
        assert r["relF_cg_solution"] <= 1e-10


# parity at BASELINE sizes (VERDICT round 1, weak point 1): C1 literally, and SURVEY section 6's probe sizes
@pytest.mark.parametrize("case", [(32, 1, 0, 0, 1, 0), (24, 2, 0, 0, 1, 0), (16, 2, 1, 0, 1, 0), (16, 2, 1, 1, 1, 0), (8, 4, 0, 0, 1, 0)])
def test_dropin_at_baseline_sizes(case):
    r, rc = dropin(case, {"B200_SKIP_SERIAL": 1}, extra=("0", "1"), timeout=1800)   # cpu_threads 0 = hardware_concurrency; device Create()
    check(r, rc)


def test_mesh_signature_catches_a_material_filter_change():
    """the cache key of the strategy.  The Python mirror has no TPZCompMesh; the equivalent event - groups
    re-added on a resident pattern with a different element subset - must rebuild the scatter maps and give the subset's matrix."""
    mesh = gridmesh.grid_mesh(4, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.1)
    mats = materials_for(0, neumann=True)
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    ia, ja, a, rhs = st.CreateAssemble()
    # volume elements only (TPZStructMatrix::SetMaterialIds({1})): same pattern, boundary groups gone
    sub = gridmesh.FlatMesh(porder=mesh.porder, nstate=mesh.nstate, nodes=mesh.nodes, block_pos=mesh.block_pos,
                            block_size=mesh.block_size, neq=mesh.neq)
    sub.blocks = [b for b in mesh.blocks if b.matid == 1]
    st2 = sm.TPZStructMatrixB200(sub, mats, symmetric=True)
    st2.SetPattern(ia, ja)
    a2, rhs2 = st2.Assemble()
    a_ref, rhs_ref = oracle_assemble(sub, mats, True, ia, ja)
    assert relF(a2, a_ref) <= TOL and relF(rhs2, rhs_ref) <= TOL
    assert np.abs(a).max() > 1e10 and np.abs(a2).max() < 1e10
    st.ctx.close()
    st2.ctx.close()


# ---- owner-computes ("gather") assembly of the closed-form groups (csrc/gather_rows.cuh) ------------------------------------
@pytest.mark.parametrize("tet,p,phys", [(True, 1, 0), (True, 2, 0), (True, 1, 1), (True, 2, 1), (False, 1, 0), (False, 2, 0), (False, 1, 1), (False, 2, 1)])
@pytest.mark.parametrize("symmetric", [True, False])
def test_gather_equals_scatter_and_oracle(tet, p, phys, symmetric):
    """Straight-sided tetrahedra (perturbed nodes) and parallelepiped hexahedra (sheared lattice): every CSR row written once by
    the warp that owns its node, against the scatter kernels (option gather = 0) and the oracle; two assemblies bit-identical."""
    n = 5
    mesh = gridmesh.grid_mesh(n, p, 3 if phys else 1, tetrahedra=tet, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.12 if tet else 0.0)
    if not tet:  # affine image of the grid: still parallelepipeds, no axis-aligned symmetry left
        shear = np.array([[1.0, 0.2, -0.1], [0.05, 0.9, 0.15], [0.1, -0.2, 1.1]])
        mesh.nodes[:] = mesh.nodes @ shear.T
    mats = materials_for(phys, neumann=True)
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    st.ctx.set_option("gather", 1)
    ia, ja, a, rhs = st.CreateAssemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
    assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
    assert "gather_rows" in [st.ctx.group_kernel(g) for gids in st.groups_of_block for g in gids], "the gather kernel ran"
    a2, rhs2 = st.Assemble()
    # fixed summation order in the rows only the gather kernel writes; the rows of the boundary nodes also receive the
    # (atomic) penalty terms of the boundary elements first, so those agree to rounding only
    assert relF(a, a2) <= 1e-15
    same = a == a2
    assert same.mean() > 0.5
    st.ctx.close()
    sc = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
    sc.ctx.set_option("gather", 0)
    sc.SetPattern(ia, ja)
    a3, rhs3 = sc.Assemble()
    assert relF(a, a3) <= 1e-14 and relF(rhs, rhs3) <= 1e-14
    sc.ctx.close()


def test_gather_rows_shared_with_other_groups():
    """Hexahedra + pyramids on the unperturbed grid: the hexahedra are parallelepipeds and gather, the pyramids scatter into the
    rows they share (zeroed first, added to); two materials on the same rows likewise."""
    mesh = gridmesh.hexpyr_mesh(4, 2, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.0)
    mats = materials_for(0, neumann=True)
    for symmetric in (True, False):
        st = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric)
        st.ctx.set_option("gather", 1)
        ia, ja, a, rhs = st.CreateAssemble()
        a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
        assert relF(a, a_ref) <= TOL and relF(rhs, rhs_ref) <= TOL
        a2, _ = st.Assemble()
        assert relF(a2, a_ref) <= TOL
        st.ctx.close()


def test_gather_follows_the_geometry():
    """b200asm_set_nodes: parallelepipeds -> gather; perturbed -> quadrature kernels + scatter maps; back again."""
    n, p = 6, 2
    mesh = gridmesh.grid_mesh(n, p, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.0)
    mats = materials_for(0, neumann=True)
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    st.ctx.set_option("gather", 1)
    ia, ja, a0, rhs0 = st.CreateAssemble()
    ref0 = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a0, ref0[0]) <= TOL
    moved = gridmesh.grid_mesh(n, p, 1, bc_matids=(-1, -1, -1, -1, -1, -2), perturb=0.1)
    st.ctx.set_nodes(moved.nodes)
    a1, rhs1 = st.Assemble()
    ref1 = oracle_assemble(moved, mats, True, ia, ja)
    assert relF(a1, ref1[0]) <= TOL and relF(rhs1, ref1[1]) <= TOL
    st.ctx.set_nodes(mesh.nodes)
    a2, rhs2 = st.Assemble()
    assert relF(a2, a0) <= 1e-15 and relF(rhs2, rhs0) <= 1e-14
    st.ctx.close()


def test_gather_overlapped_download_in_chunks():
    """Page-locked host matrix + small chunks: node-block ranges are downloaded while later ranges are assembled."""
    import torch
    mesh = gridmesh.grid_mesh(10, 2, 3, tetrahedra=True, perturb=0.1)
    mats = materials_for(1)
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=True)
    st.ctx.set_option("gather", 1)
    st.ctx.set_option("overlap_min_elements", 256)
    st.ctx.set_option("overlap_min_bytes", 4096)
    ia, ja = st.Create(on_device=True)
    a_pin = torch.empty(len(ja), dtype=torch.float64).pin_memory()
    r_pin = torch.empty(mesh.neq, dtype=torch.float64).pin_memory()
    st.Assemble(a_pin.numpy(), r_pin.numpy())
    a_ref, rhs_ref = oracle_assemble(mesh, mats, True, ia, ja)
    assert relF(a_pin.numpy(), a_ref) <= TOL and relF(r_pin.numpy(), rhs_ref) <= TOL
    st.ctx.set_option("overlap", 0)
    a2, rhs2 = st.Assemble()
    assert relF(a2, a_pin.numpy()) <= 1e-15
    st.ctx.close()
