"""Extracts the reference's simplex / pyramid quadrature tables (Integral/tpzintrulet.cpp, tpzintrulet3d.cpp, tpzintrulep3d.cpp —
published Dunavant / Zhang-Cui-Liu style tables, pure data) for the orders the hot path uses
(order 2p, p in {1,2,3,4}; pyramids p in {1,2}) from the golden fixtures (which oracle/_ref/refdriver read through
TPZIntPoints::Point) into neopz_b200/data/simplex_rules.npz, the table the standalone host ships.
The NeoPZ drop-in strategy does not use this file: it reads the rules from the live TPZIntPoints."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import golden_util as gu  # noqa: E402

out = {}
for name, p in (("tet_p1_poisson_n2_pert", 1), ("tet_p2_poisson_n2_pert", 2), ("tet_p3_poisson_n2_pert_scr", 3),
                ("tet_p4_poisson_n2_pert_scr", 4)):
    g = gu.load(name)
    for tag in ("tet", "tri"):
        out[f"{tag}_order{2 * p}_pts"] = g[f"rule_{tag}_pts"]
        out[f"{tag}_order{2 * p}_w"] = g[f"rule_{tag}_w"]
# pyramids: TPZIntRuleP3D tables (Integral/tpzintrulep3d.cpp) of order 2p, read the same way
for name, p in (("hexpyr_p1_poisson_n2_pert", 1), ("hexpyr_p2_poisson_n2_pert", 2)):
    g = gu.load(name)
    out[f"pyr_order{2 * p}_pts"] = g["rule_pyr_pts"]
    out[f"pyr_order{2 * p}_w"] = g["rule_pyr_w"]
path = os.path.join(ROOT, "neopz_b200", "data", "simplex_rules.npz")
np.savez(path, **out)
print({k: v.shape for k, v in out.items()})
