#!/usr/bin/env python
"""Markdown table of the weak-scaling bench lines under profiles/ (r02_bench_{N}gpu_all_configs.json): elements/s, ms per step and
the efficiency against the one-GPU line, per configuration.  python tools/scaling_table.py > (pasted into DESIGN.md section 6)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(n):
    p = os.path.join(ROOT, "profiles", f"r02_bench_{n}gpu_all_configs.json")
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    runs = {n: load(n) for n in (1, 2, 4, 8)}
    runs = {n: d for n, d in runs.items() if d}
    cfgs = [("C2 hex p2 Poisson (general hexahedra)", lambda d: d), ("C5 hex p2 Elasticity3D (general hexahedra)", lambda d: d["configs"]["c5"]),
            ("C4 hex p4 Poisson", lambda d: d["configs"]["c4"]), ("C3 tet p2 Elasticity3D", lambda d: d["configs"]["c3"])]
    print("| configuration | " + " | ".join(f"{n} GPU" + ("s" if n > 1 else "") for n in runs) + " |")
    print("|---|" + "---|" * len(runs))
    for name, get in cfgs:
        cells = []
        base = None
        for n, d in runs.items():
            c = get(d)
            if "value" not in c:
                cells.append("—")
                continue
            per_gpu = c["value"] / n
            if n == 1:
                base = per_gpu
            dof = c.get("dof") or (c.get("config", {}).get("dof_per_gpu", 0) * n)
            eff = f", eff {per_gpu / base:.3f}" if base and n > 1 and "C2" in name else ""
            cells.append(f"{c['value'] / 1e6:.1f} M el/s, {c['ms_per_step']:.2f} ms, {dof / 1e6:.1f} M DOF{eff}")
        print(f"| {name} | " + " | ".join(cells) + " |")
    print()
    print("| end to end (C2, host buffers, D2H of the CSR values every step) | " + " | ".join(
        f"{d['e2e']['ms_per_step']:.1f} ms, {d['e2e']['d2h_GBps_per_gpu']:.1f} GB/s per GPU" for d in runs.values()) + " |")


if __name__ == "__main__":
    main()
