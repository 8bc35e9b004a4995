#!/usr/bin/env python
"""Static SASS summary of the kernels in neopz_b200/libb200asm.so (cuobjdump -sass; no GPU needed): per kernel the counts of the
instruction classes that matter for the roofline discussion - FP64 (DFMA / DMUL / DADD / DMMA), global reductions (RED / ATOM),
global and shared loads / stores, local-memory traffic (LDL / STL = register spills), barriers, TMA / tcgen05 opcodes (none: FP64
has no tcgen05 path).  Writes a markdown table.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "neopz_b200", "libb200asm.so")
CLASSES = [("DMMA", r"^DMMA"), ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("RED", r"^RED"), ("ATOM", r"^ATOM(G|S)?\b|^ATOMG|^ATOMS"),
           ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDC/LDCU", r"^LDCU?\b|^LDC"), ("LDL", r"^LDL"), ("STL", r"^STL"),
           ("BAR", r"^BAR"), ("SHFL", r"^SHFL"), ("TMA/tcgen05", r"^(UBLKCP|UTMA|TCGEN|UTC)")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True, check=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function ([^:]+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+).*STACK:(\d+).*SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for name, pat in CLASSES:
                if re.match(pat, op):
                    kernels[cur][name] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS summary of neopz_b200/libb200asm.so (sm_100a, cuobjdump; static instruction counts per kernel)\n")
    print("Spills = LDL + STL (local memory); `TMA/tcgen05` counts UBLKCP / UTMA* / TCGEN* opcodes (FP64 has no tcgen05 path: the tensor-core")
    print("instruction of these kernels is `DMMA`, i.e. `mma.sync.m8n8k4.f64`).\n")
    cols = [c for c, _ in CLASSES]
    print("| kernel | regs | stack | instr | " + " | ".join(cols) + " |")
    print("|---|---|---|---|" + "---|" * len(cols))
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*$", "", name)
        short = re.sub(r"^void ", "", short)
        if len(short) > 90:
            short = short[:87] + "..."
        r = regs.get(mangled, ("?", "?", "?"))
        print(f"| `{short}` | {r[0]} | {r[1]} | {cnt['total']} | " + " | ".join(str(cnt[c]) for c in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
