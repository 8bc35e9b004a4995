#!/usr/bin/env python
"""Small assemblies of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Families: sum factorisation (hexahedra p2 Poisson), one-warp DMMA (variant 16, hexahedra p1), warp-team DMMA (elasticity, p3 / p4 Poisson),
closed-form scatter (gather off) and gather kernels (tetrahedra, parallelepiped hexahedra), register tiles (engine 0, prisms),
generic runtime-size kernel (engine 2), plane kernel, boundary kernels; atomic and coloured scatter; device pattern builder; CG.
Every result is checked against the oracle (1e-12), so a sanitizer-clean run is also a parity run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from neopz_b200 import gridmesh, strmatrix as sm  # noqa: E402
from tests.oracle_ref import oracle_assemble  # noqa: E402
from tests.test_gpu_parity import materials_for, relF  # noqa: E402


def run(name, mesh, mats, symmetric=True, on_device=True, cg=False, **opts):
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=symmetric, engine=opts.pop("engine", None), scatter=opts.pop("scatter", None),
                                variant=opts.pop("variant", None))
    for k, v in opts.items():
        st.ctx.set_option(k, v)
    ia, ja = st.Create(on_device=on_device)
    a, rhs = st.Assemble()
    a_ref, rhs_ref = oracle_assemble(mesh, mats, symmetric, ia, ja)
    ea, er = relF(a, a_ref), relF(rhs, rhs_ref)
    if cg:
        _x, it, res = st.SolveCG(max_iter=200, tol=1e-8)
    st.ctx.close()
    print(f"{name}: relF(A) {ea:.2e} relF(rhs) {er:.2e}", flush=True)
    assert ea <= 1e-12 and er <= 1e-12, name


def main():
    bc = (-1, -1, -1, -1, -1, -2)
    pois, elas = materials_for(0, neumann=True), materials_for(1, neumann=True)
    hex2 = gridmesh.grid_mesh(3, 2, 1, bc_matids=bc, perturb=0.1)
    run("sumfact hex p2 poisson, atomic", hex2, pois, cg=True)
    run("sumfact hex p2 poisson, coloured", hex2, pois, scatter="colored")
    run("sumfact hex p2 poisson, one warp per element (variant 20), full storage", hex2, pois, symmetric=False, variant=20)
    run("one-warp DMMA hex p2 poisson (variant 16), full storage", hex2, pois, symmetric=False, variant=16)
    run("one-warp DMMA hex p1 poisson", gridmesh.grid_mesh(4, 1, 1, bc_matids=bc, perturb=0.1), pois)
    run("warp-pair DMMA hex p2 elasticity", gridmesh.grid_mesh(3, 2, 3, bc_matids=bc, perturb=0.1), elas)
    run("warp-pair DMMA hex p2 elasticity, coloured, full", gridmesh.grid_mesh(2, 2, 3, bc_matids=bc, perturb=0.1), elas, symmetric=False, scatter="colored")
    run("one-warp DMMA hex p2 elasticity (variant 30)", gridmesh.grid_mesh(2, 2, 3, bc_matids=bc, perturb=0.1), elas, variant=30)
    run("team DMMA hex p2 elasticity (variant 34)", gridmesh.grid_mesh(2, 2, 3, bc_matids=bc, perturb=0.1), elas, variant=34)
    run("team DMMA hex p1 elasticity", gridmesh.grid_mesh(3, 1, 3, bc_matids=bc, perturb=0.1), elas)
    run("team DMMA hex p3 poisson", gridmesh.grid_mesh(2, 3, 1, bc_matids=bc, perturb=0.1), pois)
    run("team DMMA hex p4 poisson", gridmesh.grid_mesh(2, 4, 1, bc_matids=bc, perturb=0.1), pois)
    tet2 = gridmesh.grid_mesh(3, 2, 3, tetrahedra=True, bc_matids=bc, perturb=0.1)
    run("gather tet p2 elasticity", tet2, elas, gather=1)
    run("gather tet p2 elasticity, full storage", tet2, elas, symmetric=False, gather=1)
    run("closed-form scatter tet p2 elasticity", tet2, elas)
    run("closed-form scatter tet p2 elasticity, coloured", tet2, elas, scatter="colored")
    uni = gridmesh.grid_mesh(4, 2, 1, bc_matids=bc, perturb=0.0)
    run("gather hex p2 poisson (parallelepipeds)", uni, pois, gather=1)
    run("closed-form scatter hex p2 poisson", uni, pois)
    run("gather hex p2 elasticity (parallelepipeds)", gridmesh.grid_mesh(3, 2, 3, bc_matids=bc, perturb=0.0), elas, gather=1)
    run("closed form tet p3 poisson", gridmesh.grid_mesh(2, 3, 1, tetrahedra=True, bc_matids=bc, perturb=0.1), pois)
    run("register tiles (engine 0) hex p2 poisson", hex2, pois, engine=0)
    run("register tiles prisms p2 elasticity", gridmesh.grid_mesh(2, 2, 3, prisms=True, bc_matids=bc, perturb=0.1), elas)
    run("generic kernel (engine 2) hex p2 elasticity, drop_tiny off", gridmesh.grid_mesh(2, 2, 3, bc_matids=bc, perturb=0.1), elas, engine=2)
    run("hexahedra + pyramids p2 poisson", gridmesh.hexpyr_mesh(2, 2, 1, bc_matids=bc, perturb=0.1), pois)
    print("sanitize_run: all families assembled")


if __name__ == "__main__":
    main()
