#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
launches, total time, DRAM bytes, share of the run.  Cold-cache, serialised launches: the SHARES are what agrees with the
CUDA-event timing of bench.py, not the absolute times.

    python tools/launch_summary.py gpurun_out/r02_launches_c2_n128.csv [steps]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    agg = collections.defaultdict(lambda: {"launches": set(), "time_ns": 0.0, "read": 0.0, "write": 0.0})
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        a = agg[name]
        a["launches"].add(r["ID"])
        v = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r["Metric Unit"], 1.0)
        if r["Metric Name"] == "gpu__time_duration.sum":
            a["time_ns"] += v
        elif r["Metric Name"] == "dram__bytes_read.sum":
            a["read"] += v
        elif r["Metric Name"] == "dram__bytes_write.sum":
            a["write"] += v
    total = sum(a["time_ns"] for a in agg.values()) or 1.0
    print("| kernel | launches | time ms | share | DRAM read GB | DRAM write GB |" + (" DRAM GB per step |" if steps else ""))
    print("|---|---|---|---|---|---|" + ("---|" if steps else ""))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_ns"]):
        line = f"| `{name[:90]}` | {len(a['launches'])} | {a['time_ns'] / 1e6:.3f} | {a['time_ns'] / total:.3f} | {a['read'] / 1e9:.3f} | {a['write'] / 1e9:.3f} |"
        if steps:
            line += f" {(a['read'] + a['write']) / 1e9 / steps:.3f} |"
        print(line)


if __name__ == "__main__":
    main()
