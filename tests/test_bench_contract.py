"""bench.py's host-side contract, checked without a GPU: the per-element work model of SURVEY.md 8(d), the kernel naming of the
default configurations, and the shape of the committed driver-format lines under profiles/ (what the judge reads)."""
import argparse
import glob
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (importing bench.py must not need a GPU: torch is imported inside main())


def _args(**kw):
    d = dict(engine=1, topo="hex", p=2, phys="poisson", perturb=0.1, variant=0, gpus=1, gather=0)
    d.update(kw)
    return argparse.Namespace(**d)


def test_work_model_of_the_headline_configuration():
    """C2: hexahedra p2 Poisson, 27 points x (7 n^2 + 2 n) + 27 x (18 n + 194) flops per element on the reference's arithmetic."""
    flops, byts = bench.algorithmic_work("hex", 2, "poisson", 2097152, 16974593, 546932609)
    assert flops == 27 * (7 * 27 * 27 + 2 * 27) + 27 * (18 * 27 + 194) == 157599
    assert abs(byts - 3963.135) < 0.01
    f5, _ = bench.algorithmic_work("hex", 2, "elasticity", 531441, 12992241, 1236613641)
    assert f5 == 27 * (57 * 27 * 27 + 12 * 27) + 27 * (18 * 27 + 194) == 1149039


@pytest.mark.parametrize("kw,needle", [({}, "sumfact_hex_p2_poisson_warp"), ({"phys": "elasticity"}, "gram_warp_elast"),
                                       ({"p": 4}, "gram_team"), ({"topo": "tet", "phys": "elasticity"}, "affine_simplex"),
                                       ({"perturb": 0.0}, "affine_hex"), ({"engine": 0}, "assemble_volume_kernel")])
def test_kernel_names_of_the_default_configurations(kw, needle):
    a = _args(**kw)
    name = bench.kernel_name(a)
    assert needle in name
    if a.engine == 1:
        assert bench.binding_resource(name), "every default kernel states what binds it"
        assert bench.executed_flops(a.topo, a.p, a.phys, name)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*gpu_all_configs.json"))))
def test_committed_bench_lines_have_the_contract_keys(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "configs"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["warmup"] >= 3
    assert d["gpu_launches"] > 0 and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for name in ("c3", "c4", "c5"):
        c = d["configs"][name]
        assert c["value"] > 0 and c["roofline"]["frac"] > 0 and c["dof"] > 0
    if d["n_gpus"] == 1:
        assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
        assert d["parity"]["ia_ja_equal"] and d["parity"]["relF_A"] <= 1e-12 and d["parity"]["relF_rhs"] <= 1e-12
    if d["n_gpus"] == 8:   # the north-star size: >= 100 M DOF of p2 elasticity on 8 GPUs at >= 0.6 of the per-element roofline
        assert d["configs"]["c5"]["dof"] >= 100_000_000 and d["configs"]["c5"]["roofline"]["frac"] >= 0.6
