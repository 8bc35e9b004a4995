"""Regenerates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/refdriver, built
by oracle/Makefile.ref from /root/reference).  Only runs where the reference exists (this container);
the .npz files it writes are committed so the GPU box needs neither the reference nor this script.

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "refdriver")

# name: (n, p, phys, tet, perturb, bctype, with_elmats[, scramble seed[, dim[, bcfunc]]])   bcfunc 1: boundary data from functions of x
#   tet: 0 hex / quad, 1 tet / tri, 2 prism, 3 hex + pyramid
# dim 2: plane meshes (TPZGenGrid2D); phys 0 = TPZMatPoisson(dim 2), 2 / 3 = TPZElasticity2D plane strain / plane stress
# scramble != 0: node indices shuffled so that the side orientations differ from element to element (p >= 3)
CASES = {
    "hex_p1_poisson_n3": (3, 1, 0, 0, 0.0, 0, 1),
    "hex_p1_poisson_n3_pert": (3, 1, 0, 0, 0.15, 1, 1),
    "hex_p2_poisson_n2_pert": (2, 2, 0, 0, 0.15, 1, 1),
    "hex_p2_poisson_n3": (3, 2, 0, 0, 0.0, 0, 0),
    "hex_p1_elast_n2_pert": (2, 1, 1, 0, 0.15, 1, 1),
    "hex_p2_elast_n2_pert": (2, 2, 1, 0, 0.15, 1, 1),
    "tet_p1_poisson_n2_pert": (2, 1, 0, 1, 0.15, 0, 1),
    "tet_p2_poisson_n2_pert": (2, 2, 0, 1, 0.15, 1, 1),
    "tet_p2_elast_n2_pert": (2, 2, 1, 1, 0.15, 1, 1),
    "hex_p3_poisson_n2_pert_scr": (2, 3, 0, 0, 0.15, 1, 1, 7),
    "hex_p4_poisson_n2_pert_scr": (2, 4, 0, 0, 0.15, 1, 1, 11),
    "hex_p3_elast_n2_pert_scr": (2, 3, 1, 0, 0.15, 1, 1, 5),
    "hex_p4_poisson_n2": (2, 4, 0, 0, 0.0, 0, 0, 0),
    "hex_p3_poisson_n3_pert": (3, 3, 0, 0, 0.15, 1, 0, 0),
    # TPZElasticity3D::ContributeBC types 2 (mixed), 3 (directional null Dirichlet), 8 (Dirichlet on x and z), 6 (on y)
    "hex_p2_elast_n2_bc2": (2, 2, 1, 0, 0.15, 2, 0, 0),
    "hex_p2_elast_n2_bc3": (2, 2, 1, 0, 0.15, 3, 0, 0),
    "tet_p2_elast_n2_bc8": (2, 2, 1, 1, 0.15, 8, 0, 0),
    "hex_p1_elast_n2_bc6": (2, 1, 1, 0, 0.15, 6, 1, 0),
    "quad_p2_elast2d_n3_pert": (3, 2, 2, 0, 0.15, 1, 1, 0, 2),
    # (type 2 of TPZElasticity2D::ContributeBC crashes in the reference itself: sliced TPZMatLoadCasesBC copy, :220)
    "quad_p1_elast2d_stress_n3_bc3": (3, 1, 3, 0, 0.15, 3, 1, 0, 2),
    "tri_p2_elast2d_n3_pert": (3, 2, 2, 1, 0.15, 1, 1, 0, 2),
    "tri_p1_elast2d_stress_n2_bc3": (2, 1, 3, 1, 0.15, 3, 1, 0, 2),
    "quad_p2_poisson2d_n3_pert": (3, 2, 0, 0, 0.15, 1, 1, 0, 2),
    "tri_p2_poisson2d_n3_pert": (3, 2, 0, 1, 0.15, 1, 1, 0, 2),
    # plane quadrilaterals of order 3, 4 with oriented line boundary elements (scrambled node numbering)
    "quad_p3_elast2d_n3_pert_scr": (3, 3, 2, 0, 0.15, 1, 1, 9, 2),
    "quad_p4_poisson2d_n3_pert_scr": (3, 4, 0, 0, 0.15, 1, 1, 13, 2),
    "quad_p4_elast2d_stress_n2_bc3_scr": (2, 4, 3, 0, 0.15, 3, 1, 3, 2),
    # tet = 2: prisms (MMeshType::EPrismatic, triangular + quadrilateral boundary faces);
    # tet = 3: hexahedra + pyramids (MMeshType::EHexaPyrMixed: every other cell split into six pyramids around a centre node)
    "prism_p1_poisson_n2_pert": (2, 1, 0, 2, 0.15, 0, 1),
    "prism_p2_poisson_n2_pert": (2, 2, 0, 2, 0.15, 1, 1),
    "prism_p2_elast_n2_pert": (2, 2, 1, 2, 0.15, 1, 1),
    "prism_p1_elast_n3_bc2": (3, 1, 1, 2, 0.15, 2, 0),
    "hexpyr_p1_poisson_n2_pert": (2, 1, 0, 3, 0.15, 0, 1),
    "hexpyr_p2_poisson_n2_pert": (2, 2, 0, 3, 0.15, 1, 1),
    "hexpyr_p2_elast_n2_pert": (2, 2, 1, 3, 0.15, 1, 1),
    "hexpyr_p1_elast_n3": (3, 1, 1, 3, 0.0, 0, 0),
    # simplices of order 3, 4: oriented edges and triangular faces (scrambled node numbering), tetrahedron interior function
    "tet_p3_poisson_n2_pert_scr": (2, 3, 0, 1, 0.15, 1, 1, 7),
    "tet_p4_poisson_n2_pert_scr": (2, 4, 0, 1, 0.15, 1, 1, 11),
    "tet_p3_elast_n2_pert_scr": (2, 3, 1, 1, 0.15, 1, 1, 5),
    "tet_p4_elast_n1_pert": (1, 4, 1, 1, 0.15, 2, 1, 0),
    "tet_p3_poisson_n3": (3, 3, 0, 1, 0.0, 0, 0, 0),
    "tri_p3_elast2d_n3_pert_scr": (3, 3, 2, 1, 0.15, 1, 1, 9, 2),
    "tri_p4_poisson2d_n3_pert_scr": (3, 4, 0, 1, 0.15, 1, 1, 13, 2),
    "tri_p4_elast2d_stress_n2_bc3": (2, 4, 3, 1, 0.15, 3, 1, 0, 2),
    # boundary data from functions (TPZBndCondT::SetForcingFunctionBC): Dirichlet + Neumann (Poisson), Dirichlet + mixed (Elasticity3D)
    "hex_p2_poisson_n2_bcfunc": (2, 2, 0, 0, 0.15, 1, 1, 0, 3, 1),
    "tet_p2_elast_n2_bcfunc": (2, 2, 1, 1, 0.15, 2, 1, 0, 3, 1),
    "prism_p2_poisson_n2_bcfunc": (2, 2, 0, 2, 0.15, 1, 1, 0, 3, 1),
    # TPZMatPoisson::ContributeBC type 2 as the reference computes it (gradient penalty on the zmax face, TPZMatPoisson.cpp:104-118);
    # TPZElasticity3D::ContributeBC type 4 (stress field times the face normal, TPZElasticity3D.cpp:724-737)
    "hex_p2_poisson_n2_bc2": (2, 2, 0, 0, 0.15, 2, 1),
    "tet_p2_poisson_n2_bc2": (2, 2, 0, 1, 0.15, 2, 1),
    "hex_p2_elast_n2_bc4": (2, 2, 1, 0, 0.15, 4, 1),
    "tet_p2_elast_n2_bc4": (2, 2, 1, 1, 0.15, 4, 1),
    "hex_p3_poisson_n2_bc2_scr": (2, 3, 0, 0, 0.15, 2, 0, 7),
}


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/refdriver missing: run `make -f oracle/Makefile.ref -j8` first")
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        n, p, phys, tet, pert, bctype, elm = case[:7]
        scr = case[7] if len(case) > 7 else 0
        dim = case[8] if len(case) > 8 else 3
        bcfunc = case[9] if len(case) > 9 else 0
        with tempfile.TemporaryDirectory() as d:
            subprocess.check_call([DRIVER, "dump", d, str(n), str(p), str(phys), str(tet), repr(pert),
                                   str(bctype), str(elm), str(scr), str(dim), str(bcfunc)], stdout=subprocess.DEVNULL)
            arrays = {}
            for f in sorted(os.listdir(d)):
                if f.endswith(".npy"):
                    arrays[f[:-4]] = np.load(os.path.join(d, f))
            meta = json.load(open(os.path.join(d, "meta.json")))
            arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
            out = os.path.join(HERE, name + ".npz")
            np.savez_compressed(out, **arrays)
            print(name, "->", os.path.getsize(out) // 1024, "KiB", meta)


if __name__ == "__main__":
    main()
