#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 assembly path (BASELINE.json metric).

A "step" = one full assembly (zero A and rhs, element matrices of every element, scatter-add into the
CSR values and the load vector) of the configured mesh, on a pattern created once — the same quantity
the reference times as the second TPZLinearAnalysis::Assemble() (Analysis/TPZLinearAnalysis.cpp:73-77).

    python bench.py --gpus 1 --steps 10 --warmup 3            # our arm   (CUDA, through the C ABI)
    python bench.py --impl reference --steps 3 --warmup 1     # reference (NeoPZ TPZStructMatrixOR on the host cores)

Default workload = BASELINE.json configs[1]: 3D Poisson, H1 p=2, 128^3 hexahedra (~17M DOF), one B200.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def algorithmic_work(topo, p, phys, nvol, neq, nnz):
    """FLOPs per volume element counted on the reference's arithmetic (full ek, mul/add = 1 flop each)
    and compulsory HBM bytes per element (node coordinates + destination indices + every stored CSR value and rhs
    entry written once + the scatter-map read of the element's stored entries)."""
    if topo == "hex":
        n, q, nodes, G = (p + 1) ** 3, int(0.51 * (2 * p + 2)) ** 3, 8, 194
    elif topo == "prism":   # TPZShapePrism, TPZIntPrism3D (line rule x triangle rule); grad x: 6 nodes x 9 x 2 + det / inverse
        n, q, nodes, G = {1: 6, 2: 18}[p], {1: 6, 2: 18}[p], 6, 158
    else:
        n, q, nodes, G = {1: 4, 2: 10, 3: 20, 4: 35}[p], {1: 4, 2: 14, 3: 24, 4: 46}[p], 4, 122   # tetrahedra: rules of order 2p (Zhang-Cui-Liu tables)
    if phys == "poisson":
        ndof, flops = n, q * (7 * n * n + 2 * n) + q * (18 * n + G)
    else:
        ndof, flops = 3 * n, q * (57 * n * n + 12 * n) + q * (18 * n + G)
    e_el = ndof * (ndof + 1) // 2
    byts = 8 * nodes * 3 + 4 * ndof + 8.0 * nnz / nvol + 8.0 * neq / nvol + 4 * e_el
    return float(flops), float(byts)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=128, help="grid divisions per direction (per GPU)")
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--phys", default="poisson", choices=["poisson", "elasticity"])
    ap.add_argument("--topo", default="hex", choices=["hex", "tet", "prism"], help="prism: one GPU only (two prisms per grid cell)")
    ap.add_argument("--cpu-n", type=int, default=0, help="grid size of the bounded CPU-baseline sample (0 = auto)")
    ap.add_argument("--engine", type=int, default=1, help="0: register-tile DFMA kernels, 1: DMMA panel kernels where available")
    ap.add_argument("--scatter", default="atomic", choices=["atomic", "colored"])
    ap.add_argument("--debug", type=int, default=0, help="profiling aid: 1 = drop the matrix scatter (wrong results, timing only)")
    ap.add_argument("--variant", type=int, default=0, help="tuning alternative of the DMMA kernels (0 = default)")
    ap.add_argument("--pattern", default="device", choices=["device", "host"],
                    help="one-off setup: CSR pattern built on the GPU (b200asm_build_pattern_device) or by the threaded host builder")
    ap.add_argument("--cg", type=int, default=0, help="also time N iterations of the device-resident CG (reported as extra keys)")
    ap.add_argument("--perturb", type=float, default=0.1, help="smooth node perturbation in units of h (default 0.1: general trilinear "
                    "hexahedra; 0 = the uniform grid of CreateGeoMeshOnGrid, whose parallelepiped cells take the closed-form kernel)")
    ap.add_argument("--locality", type=int, default=1, help="0: keep the mesh (lexicographic) element order on the device")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: interface rows pushed into the owner's memory over "
                    "NVLink while the interior is assembled (C ABI, b200asm_exchange_*), or NCCL send/recv after the kernels")
    ap.add_argument("--extra", default="c5,c4,c3", help="other BASELINE.json configurations measured after the headline and reported under "
                    "\"configs\" (c3 tetrahedra p2 elasticity, c4 hexahedra p4 Poisson, c5 hexahedra p2 elasticity; per-GPU slabs, weak scaling)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--gather", type=int, default=0, help="1: closed-form groups are assembled row by row in a fixed order (option gather of the C ABI; slower, deterministic)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the timing of TPZLinearAnalysis::Assemble() inside the unmodified reference")
    ap.add_argument("--dropin-n", type=int, default=48, help="grid of the drop-in timing mesh")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def kernel_name(a):
    if a.engine == 1:
        if a.topo == "tet" and a.variant == 0 and a.p <= 2:
            if a.gpus == 1 and a.gather:
                return "gat::gather_rows_kernel (closed-form element matrices of straight-sided tetrahedra, every CSR row written once by the warp that owns its node)"
            return "assemble_affine_simplex_kernel (closed-form element matrices of straight-sided tetrahedra, one warp per element)"
        if a.topo == "hex" and a.perturb == 0.0 and a.p <= 2:
            if a.gpus == 1 and a.gather:
                return "gat::gather_rows_kernel (closed-form element matrices of parallelepiped hexahedra, every CSR row written once by the warp that owns its node)"
            return "assemble_affine_hex_kernel (closed-form element matrices of parallelepiped hexahedra, one warp per element)"
        if a.topo == "hex" and a.phys == "poisson" and a.p == 2 and a.variant in (0, 20):
            return "assemble_sumfact_hex_p2_poisson_warp_kernel (sum factorisation, one warp per element, no block-wide barrier)"
        if a.topo == "hex" and a.phys == "poisson" and a.p == 2 and a.variant == 13:
            return "assemble_sumfact_hex_p2_poisson_kernel (sum factorisation with prefetch, one CTA of 64 threads per element)"
        if a.topo == "hex" and a.phys == "poisson" and a.p == 2:
            return "assemble_gram_mma_kernel (one warp per element, mma.sync.m8n8k4.f64)"
        if a.topo == "hex" and a.phys == "poisson" and a.p >= 3:
            return "assemble_gram_team_kernel (warp team per element, mma.sync.m8n8k4.f64)"
        if a.phys == "elasticity" and a.topo == "hex" and a.p == 2 and a.variant in (0, 31):
            return "assemble_gram_warp_elast_kernel (a pair of warps per element, whole panel in shared memory, mma.sync.m8n8k4.f64)"
        if a.phys == "elasticity" and a.topo == "hex":
            return "assemble_gram_team_kernel (warp team per element, mma.sync.m8n8k4.f64)"
    return "assemble_volume_kernel (register-tile DFMA)"


def workload_name(a, n=None):
    n = n or a.n
    return f"3D {a.phys} H1 p={a.p} {a.topo} {n}^3 grid, sym CSR (TPZSSpStructMatrix), Dirichlet on all faces"


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref, built from /root/reference by
# oracle/Makefile.ref) timed on this box's host cores, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_n(a):
    if a.cpu_n:
        return a.cpu_n
    # ~10-30 s of CPU work including mesh + Create(): sized from the survey's per-element costs
    if a.phys == "poisson":
        return {1: 64, 2: 48, 3: 14, 4: 8}.get(a.p, 6) if a.topo == "hex" else 20
    return {1: 24, 2: 14, 3: 8}.get(a.p, 6) if a.topo == "hex" else 14


def run_reference(a, steps, warmup, dumpdir=None):
    drv = os.path.join(ROOT, "oracle", "_ref", "refdriver")
    cores = os.cpu_count() or 1
    n = cpu_sample_n(a)
    if os.path.exists(drv):
        reps = max(1, steps)
        out = subprocess.run([drv, "time", str(n), str(a.p), "1" if a.phys == "elasticity" else "0",
                              {"hex": "0", "tet": "1", "prism": "2"}[a.topo], str(cores), str(reps + warmup), repr(float(a.perturb))] +
                             ([dumpdir] if dumpdir else []), capture_output=True, text=True, check=True).stdout
        line = [l for l in out.splitlines() if l.startswith("{")][-1]
        r = json.loads(line)
        sec = r["assemble_s_mean"]
        return {"value": r["vol_elements"] / sec, "dof_per_s": r["neq"] / sec, "unit": "elements/s", "cores": cores,
                "kind": "reference", "ms_per_step": sec * 1e3, "n": n,
                "sample": f"{workload_name(a, n)}: {r['vol_elements']} volume elements, {r['neq']} DOF, "
                          f"TPZStructMatrixOR SetNumThreads({cores}), mean of {reps + warmup} re-assemblies"}
    # fallback: the oracle port (scalar, 1 core)
    import numpy as np
    from neopz_b200 import gridmesh
    from tests.oracle_ref import oracle_assemble
    from tests.test_gpu_parity import materials_for
    from neopz_b200 import capi
    n = max(4, n // 2)
    mesh = gridmesh.grid_mesh(n, a.p, 3 if a.phys == "elasticity" else 1, tetrahedra=a.topo == "tet", prisms=a.topo == "prism")
    mats = materials_for(1 if a.phys == "elasticity" else 0)
    idx, graph = mesh.element_graph()
    ia, ja = capi.build_pattern(True, idx, graph, mesh.block_pos, mesh.block_size, 0)
    t0 = time.time()
    oracle_assemble(mesh, mats, True, ia, ja)
    sec = time.time() - t0
    nvol = len(mesh.blocks[0].elnodes)
    return {"value": nvol / sec, "dof_per_s": mesh.neq / sec, "unit": "elements/s", "cores": 1, "kind": "port",
            "ms_per_step": sec * 1e3, "n": n, "sample": f"{workload_name(a, n)}: oracle/oracle.c serial port, {nvol} elements"}


# ------------------------------------------------------------------------------------------------
def clocks_sampler(stop, samples, device):
    """SM clock / throttle reasons DURING the timed region: NVML in-process (a few hundred samples per second);
    nvidia-smi polling as the fallback."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = (("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap))
        while not stop.is_set():
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            samples.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx),
                            str(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                           ["Active" if r & b else "Not Active" for _n, b in bits])
            stop.wait(0.002)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(device)],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                samples.append([x.strip() for x in out.split(",")])
        except Exception:
            pass
        stop.wait(0.05)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    reasons = set()
    for s in samples:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "power_w_max": max(float(s[2]) for s in samples),
            "reasons": sorted(reasons), "samples": len(samples)}


# ------------------------------------------------------------------------------------------------
# BASELINE.json configurations measured next to the headline (reported under "configs" of the same JSON line):
# per-GPU grids; N > 1 stacks N slabs in z (weak scaling), e.g. C5 at N = 8: 81 x 81 x 648 hexahedra = 103.9 M DOF
EXTRA_CONFIGS = {
    "c3": {"topo": "tet", "p": 2, "phys": "elasticity", "n": 113, "n_multi": 96, "perturb": 0.1,
           "baseline": "configs[2]: 3D elasticity H1 p=2 on tetrahedra, ~30 M DOF on one B200"},
    "c4": {"topo": "hex", "p": 4, "phys": "poisson", "n": 64, "n_multi": 48, "perturb": 0.1,
           "baseline": "configs[3]: 3D Poisson H1 p=4 hexahedra (FP64 DMMA path), 1/2/4/8 B200"},
    "c5": {"topo": "hex", "p": 2, "phys": "elasticity", "n": 81, "n_multi": 81, "perturb": 0.1,
           "baseline": "configs[4]: 3D elasticity H1 p=2 hexahedra, >= 100 M DOF on 8 B200 (81 x 81 x 81 N)"},
}


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run (and first-touch the pinned host buffers) on the CPUs NVML reports as closest to the GPU, so that
    eight ranks do not download 36 GB through one memory controller."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class Case:
    """One workload on this rank's GPU: mesh slab, struct matrix, pattern; device-resident timing and roofline."""

    def __init__(self, a, rank, world, local_rank):
        import numpy as np
        import torch
        from neopz_b200 import distributed, gridmesh, strmatrix as sm
        self.a, self.rank, self.world = a, rank, world
        self.torch = torch
        t0 = time.time()
        ns = 3 if a.phys == "elasticity" else 1
        if a.topo == "prism":
            if world > 1:
                raise SystemExit("bench.py: --topo prism runs on one GPU (the z-slab partition covers hexahedra and tetrahedra)")
            import types
            pm = gridmesh.grid_mesh(a.n, a.p, ns, prisms=True, perturb=a.perturb)
            slab = types.SimpleNamespace(mesh=pm, nown=pm.neq)
        else:
            slab = distributed.slab_mesh(a.n, a.n * world, rank, world, a.p, ns, tetrahedra=a.topo == "tet", perturb=a.perturb)
        self.slab, self.mesh = slab, slab.mesh
        self.t_flat = time.time() - t0
        if a.phys == "poisson":
            mat = sm.TPZMatPoisson(1, 3)
            mat.SetForcingFunction(1.0)
            mats = {1: mat, -1: mat.CreateBC(-1, 0, [[0.0]], [0.0])}
        else:
            mat = sm.TPZElasticity3D(1, 1000.0, 0.3, (0.0, 0.0, -1.0))
            mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3))}
        self.mats = mats
        self.sharded = distributed.ShardedStructMatrix(slab, mats, symmetric=True, device=local_rank, engine=a.engine, scatter=a.scatter,
                                                       pattern=a.pattern, variant=a.variant, exchange=a.exchange) if world > 1 else None
        self.strmat = self.sharded.strmat if self.sharded else sm.TPZStructMatrixB200(self.mesh, mats, symmetric=True, device=local_rank,
                                                                                      engine=a.engine, scatter=a.scatter, variant=a.variant)
        self.stream = torch.cuda.current_stream()
        self.strmat.ctx.set_stream(self.stream.cuda_stream)
        if not a.locality:
            self.strmat.ctx.set_option("locality", 0)
        if a.debug:
            self.strmat.ctx.set_option("debug", a.debug)
        if a.gather:
            self.strmat.ctx.set_option("gather", 1)
        t0 = time.time()
        if self.sharded:
            ia, ja = self.sharded.Create()
        else:
            ia, ja = self.strmat.Create(on_device=a.pattern == "device", download=False)
        torch.cuda.synchronize()
        self.t_create = time.time() - t0
        self.nvol = len(self.mesh.blocks[0].elnodes)
        self.neq = slab.nown
        self.nnz = len(ja) if ja is not None else (self.sharded.nnz if self.sharded else self.strmat.nnz)
        self.step_async = self.sharded.AssembleDevice if self.sharded else self.strmat.ctx.assemble_async

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def time_device(self, steps, warmup, sampler=None):
        """W warm-up steps, then exactly K steps between CUDA events on the launching stream; max over ranks."""
        torch = self.torch
        for _ in range(max(3, warmup)):
            self.step_async()
        self.barrier()
        k0 = self.strmat.ctx.counters()[0]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if sampler:
            sampler[0].start()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for s0, s1 in evs:
            s0.record(self.stream)
            self.step_async()
            s1.record(self.stream)
        e1.record(self.stream)
        self.barrier()
        if sampler:
            sampler[1].set()
        # duration of the dominant kernel alone (the volume group's launches), CUDA events on the launching stream, taken
        # right after the timed region so that the step timing above carries no extra event records
        ctx = self.strmat.ctx
        ctx.set_option("timing", 1)
        vol_ms = []
        for _ in range(min(steps, 5)):
            self.step_async()
            vol_ms.append(sum(ctx.group_time_ms(g) for g in self.strmat.groups_of_block[0]))
        ctx.set_option("timing", 0)
        total_ms = e0.elapsed_time(e1)
        self.step_ms = [s0.elapsed_time(s1) for s0, s1 in evs]
        self.launches = ctx.counters()[0] - k0
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([total_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        self.ms_per_step = total_ms / steps
        self.kernel_ms = float(sum(vol_ms) / len(vol_ms))
        self.value = self.nvol * self.world / (self.ms_per_step * 1e-3)
        return self.value

    def roofline(self, peaks, fp64):
        a = self.a
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        fp64_peak = max(fp64.get("dfma_tflops", 0.0), fp64.get("dmma_tflops", 0.0)) or 37.0
        f_el, b_el = algorithmic_work(a.topo, a.p, a.phys, self.nvol, self.neq, self.nnz)
        flops, byts = f_el * self.nvol, b_el * self.nvol
        ach_tf = flops / (self.kernel_ms * 1e-3) / 1e12
        ach_gb = byts / (self.kernel_ms * 1e-3) / 1e9
        t_fp, t_hbm = flops / (fp64_peak * 1e12), byts / (hbm_peak * 1e9)
        kname = kernel_name(a)
        closed_form = "closed-form" in kname
        # closed-form kernels do not execute the reference's quadrature arithmetic (6-15x fewer flops): what bounds them is
        # the memory system, so their fraction is quoted on the HBM roofline of the compulsory bytes
        if t_fp >= t_hbm and not closed_form:
            r = {"bound": "tensor", "pipe": "fp64 (DMMA mma.sync.m8n8k4.f64 / DFMA; the larger measured peak)", "achieved": ach_tf,
                 "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                 "peak_source": "measured on this pool's B200 by tools/fp64_peak.cu (profiles/r01_fp64_peak.json): "
                                "MEASURED_PEAKS.json carries no FP64 figure"}
        else:
            r = {"bound": "hbm", "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
                 "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"}
            if closed_form:
                r["note"] = "closed-form element matrices: the FP64 roofline of the reference's arithmetic (%.2f of it) does not bound this kernel" % (ach_tf / fp64_peak)
        traffic = None
        try:  # DRAM bytes per assembly of the dominant kernel from the committed ncu launch list of this exact workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{a.topo},{a.p},{a.phys},{a.n},{a.scatter}")
            if tr and self.world == 1 and a.engine == 1:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        # executed arithmetic of the kernels (upper triangle of ek only, fused multiply-adds counted as 2): what the FP64
        # pipe actually does per element, next to the reference-arithmetic count `achieved` is quoted on
        x_el = executed_flops(a.topo, a.p, a.phys, kname)
        r.update({"traffic": traffic, "algorithmic_flops_per_element": f_el, "algorithmic_bytes_per_element": b_el,
                  "hbm_GBps_algorithmic": ach_gb, "hbm_frac_algorithmic": ach_gb / hbm_peak, "hbm_peak_GBps": hbm_peak,
                  "executed_flops_per_element": x_el,
                  "fp64_pipe_utilisation_executed": (x_el * self.nvol / (self.kernel_ms * 1e-3) / 1e12 / fp64_peak) if x_el else None,
                  "binding_resource": binding_resource(kname),
                  "kernel": kname, "kernel_ms": self.kernel_ms, "kernel_share_of_step": self.kernel_ms / self.ms_per_step,
                  "peaks": {"fp64_tflops": fp64_peak, "hbm_gbs": hbm_peak}})
        return r

    def close(self):
        if self.sharded:
            self.sharded.close()
        else:
            self.strmat.ctx.close()
        self.strmat = self.sharded = self.slab = self.mesh = None
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def binding_resource(kernel):
    """What the committed ncu captures show the kernel to be bound by (profiles/): none of the kernels is
    limited by the FP64 pipe or by HBM bandwidth any more, so `frac` (quoted on the reference's arithmetic, as north_star asks)
    can exceed 1 where the kernel executes fewer flops than the reference."""
    if "sumfact" in kernel and "warp" in kernel:
        return ("LSU wavefronts 81 % of peak (profiles/r02_ncu_sumfact_warp_raw.csv): shared-memory loads of the factor stages 54 %, "
                "378 red.global.add.f64 lanes per element at ~1.3 LSU cycles each (488 of the ~1000 cycles per element and SM)")
    if "warp_elast" in kernel:
        return ("3321 red.global.add.f64 lanes per element at ~1.3 LSU cycles each = 4284 of 7940 cycles per element and SM; tensor pipe 29 % busy "
                "(630 DMMA per element), 16 warps per SM (profiles/r02_ncu_elast_warp_raw.csv: the one-warp form)")
    if "closed-form" in kernel:
        return "red.global.add.f64 issue: ~1.3 LSU cycles per lane, 465 (tetrahedra p2 elasticity) / 378 / 3321 (hexahedra p2) lanes per element"
    if "team" in kernel:
        return "DMMA pipe (51 % in executed flops for p4) and the named barriers between the team's phases (profiles/r01_ncu_full_final_team_hexp4poisson.csv)"
    return None


def executed_flops(topo, p, phys, kernel):
    """FP64 operations the kernel families execute per element (FMA = 2), from their loop structure:
    quadrature kernels form the upper triangle of the Gram products of the (3 nq) x n panel (Poisson) or the nine sums
    per node pair (Elasticity3D) plus the panel build; the closed-form kernels contract a 3 x 3 Jacobian factor with the
    reference-element table.  None where no model is stated."""
    if topo == "hex":
        n, q = (p + 1) ** 3, int(0.51 * (2 * p + 2)) ** 3
    elif topo == "tet":
        n, q = {1: 4, 2: 10, 3: 20, 4: 35}[p], {1: 4, 2: 14, 3: 24, 4: 46}[p]
    else:
        return None
    pairs = n * (n + 1) // 2
    geom = q * (194 if topo == "hex" else 122) + q * 18 * n
    if "gather" in kernel:      # both triangles of the node pairs (the owner of a row recomputes its part of every element)
        n2 = n * n
        return 2.0 * 6 * n2 if phys == "poisson" else 2.0 * 54 * n2 + 30 * n2
    if "affine" in kernel:      # 54 FMA per node pair (9 Jinv products x 6 table rows) + block combination
        return 2.0 * 54 * pairs + (9 * 3 * pairs if phys == "elasticity" else 0) + 200
    if "sumfact" in kernel:
        return 2.0 * 21000
    if phys == "poisson":
        return 2.0 * 3 * q * pairs + geom
    return 2.0 * 9 * q * pairs + 9 * 3 * pairs + geom


def dropin_timing(a, local_rank):
    """an.Assemble() of the UNMODIFIED reference (tests/_bin/dropin_test: TPZLinearAnalysis + TPZSSpStructMatrix) on one
    TPZCompMesh, first with its TPZStructMatrixOR strategy on the host cores, then with TPZStructMatrixB200: the wall-clock
    times a NeoPZ user sees (mesh flattening, uploads, kernels, download into the TPZSYsmpMatrix included), and the
    comparison of the two results.  None when the binary was not built (needs /root/reference at build time)."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_bin", "dropin_test")
    if not os.path.exists(exe) or a.topo == "prism":
        return None
    threads = min(16, os.cpu_count() or 1)
    phys = 1 if a.phys == "elasticity" else 0
    n = a.dropin_n if a.phys == "poisson" and a.p <= 2 else max(8, a.dropin_n // 3)
    env = dict(os.environ, B200_SKIP_SERIAL="1", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
    # n p phys tet symmetric solve cpu_threads device_create equation_filter pin_host
    out = subprocess.run([exe, str(n), str(a.p), str(phys), "1" if a.topo == "tet" else "0", "1", "0", str(threads), "1", "0", "1"],
                         capture_output=True, text=True, timeout=900, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not lines:
        return {"error": (out.stdout[-300:] + out.stderr[-300:]).strip()}
    r = json.loads(lines[-1])
    nvol = r.get("vol_elements")
    d = {"mesh": workload_name(a, n), "api": "TPZLinearAnalysis::Assemble() of the unmodified reference, strategy TPZStructMatrixOR against TPZStructMatrixB200 "
         "(device-side Create(), page-locked TPZSYsmpMatrix values)", "cpu_threads": r.get("cpu_threads"),
         "cpu_threaded_assemble_s": r.get("cpu_threaded_assemble_s"), "gpu_first_assemble_s": r.get("gpu_first_assemble_s"),
         "gpu_second_assemble_s": r.get("gpu_second_assemble_s"), "ia_identical": r.get("ia_identical"), "ja_identical": r.get("ja_identical"),
         "relF_A": r.get("relF_A"), "relF_rhs": r.get("relF_rhs"), "ok": r.get("ok")}
    if nvol and r.get("gpu_second_assemble_s") and r.get("cpu_threaded_assemble_s"):
        d["volume_elements"] = nvol
        d["cpu_elements_per_s"] = nvol / r["cpu_threaded_assemble_s"]
        d["gpu_elements_per_s_second_assemble"] = nvol / r["gpu_second_assemble_s"]
    return d


def parity_against_reference(a, dumpdir, n, local_rank):
    """The GPU assembles the SAME mesh the CPU arm just timed (refdriver dumped its IA / JA / A / rhs): pattern equality and
    relative Frobenius differences.  The reference's arrays are only compared, never used by the GPU path."""
    import numpy as np
    from neopz_b200 import gridmesh, strmatrix as sm
    ns = 3 if a.phys == "elasticity" else 1
    mesh = gridmesh.grid_mesh(n, a.p, ns, tetrahedra=a.topo == "tet", prisms=a.topo == "prism", perturb=a.perturb)
    if a.phys == "poisson":
        mat = sm.TPZMatPoisson(1, 3)
        mat.SetForcingFunction(1.0)
        mats = {1: mat, -1: mat.CreateBC(-1, 0, [[0.0]], [0.0])}
    else:
        mat = sm.TPZElasticity3D(1, 1000.0, 0.3, (0.0, 0.0, -1.0))
        mats = {1: mat, -1: mat.CreateBC(-1, 0, np.zeros((3, 3)), np.zeros(3))}
    st = sm.TPZStructMatrixB200(mesh, mats, symmetric=True, device=local_rank, engine=a.engine, scatter=a.scatter, variant=a.variant)
    ia, ja = st.Create(on_device=True, download=True)
    av, rhs = st.Assemble()
    st.ctx.close()
    ia_r = np.fromfile(os.path.join(dumpdir, "ia.bin"), dtype=np.int64)
    ja_r = np.fromfile(os.path.join(dumpdir, "ja.bin"), dtype=np.int64)
    a_r = np.fromfile(os.path.join(dumpdir, "a.bin"), dtype=np.float64)
    rhs_r = np.fromfile(os.path.join(dumpdir, "rhs.bin"), dtype=np.float64)
    same = bool(len(ia) == len(ia_r) and len(ja) == len(ja_r) and np.array_equal(ia, ia_r) and np.array_equal(ja, ja_r))
    out = {"mesh": workload_name(a, n), "neq": int(mesh.neq), "nnz": int(len(ja)), "ia_ja_equal": same,
           "against": "the unmodified reference (oracle/_ref/refdriver: TPZSSpStructMatrix + TPZStructMatrixOR) on the same mesh in this run"}
    if same:
        out["relF_A"] = float(np.linalg.norm(av - a_r) / np.linalg.norm(a_r))
        out["relF_rhs"] = float(np.linalg.norm(rhs - rhs_r) / max(np.linalg.norm(rhs_r), 1e-300))
        # rows without penalty entries (the Frobenius norm is dominated by the Dirichlet big number)
        big = 1e10
        rows = np.repeat(np.arange(len(ia) - 1), np.diff(ia))
        dirty = np.zeros(len(ia) - 1, dtype=bool)
        dirty[rows[np.abs(a_r) > big]] = True
        dirty[ja_r[np.abs(a_r) > big]] = True
        keep = ~(dirty[rows] | dirty[ja_r])
        if keep.any():
            out["relF_A_rows_without_penalty"] = float(np.linalg.norm((av - a_r)[keep]) / np.linalg.norm(a_r[keep]))
        out["tolerance"] = 1e-12
        out["ok"] = bool(out["relF_A"] <= 1e-12 and out["relF_rhs"] <= 1e-12)
    else:
        out["ok"] = False
    return out


def main():
    a = parse()
    # stdout carries exactly ONE line (the JSON): anything libraries print there (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return
        if not a.cpu_n and a.topo == "hex" and a.p == 2 and a.phys == "poisson":
            a.cpu_n = 64   # the reference arm runs alone: a larger bounded sample of the headline workload (~2 min in all)
        r = run_reference(a, a.steps, a.warmup)
        line = {"impl": "reference", "metric": "assembled volume elements/s (Assemble on a created pattern)",
                "value": r["value"], "unit": "elements/s", "dof_per_s": r["dof_per_s"], "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(a), "sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "elements/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import copy
    import numpy as np
    import torch
    import torch.distributed as dist
    from neopz_b200 import gridmesh

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    peaks, fp64 = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        fp64 = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))
    except Exception:
        pass

    # ---- one-off setup (reported, not timed as assembly): mesh flatten, pattern, scatter maps ------------------
    # N > 1: ONE global mesh of n x n x (n*N) elements, z-slab per rank, rows owned by the rank that first touches
    # them, interface-row contributions pushed over NVLink every step (neopz_b200/distributed.py) -> weak scaling
    case = Case(a, rank, world, local_rank)
    strmat, sharded, mesh, stream = case.strmat, case.sharded, case.mesh, case.stream
    nvol, neq, nnz = case.nvol, case.neq, case.nnz
    step_async, barrier = case.step_async, case.barrier

    # ---- device-resident timing ---------------------------------------------------------------------
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
    value = case.time_device(a.steps, a.warmup, sampler=(th, stop) if rank == 0 else None)
    stop.set()
    ms_per_step, step_ms, launches, kernel_ms = case.ms_per_step, case.step_ms, case.launches, case.kernel_ms

    # ---- end to end through the C ABI with HOST buffers (H2D of the nodes, D2H of A and rhs, every step) ----
    e2e = None
    if not a.no_e2e:
        a_host = torch.empty(nnz, dtype=torch.float64).pin_memory()
        r_host = torch.empty(mesh.neq, dtype=torch.float64).pin_memory()
        x_host = torch.from_numpy(mesh.nodes.copy()).pin_memory()
        a_np, r_np, x_np = a_host.numpy(), r_host.numpy(), x_host.numpy()
        a_np[:] = 0.0   # first touch on this rank's NUMA node
        ksteps = max(2, min(a.steps, 5))
        def e2e_step():
            strmat.ctx.set_nodes(x_np)          # H2D: node coordinates (the geometry input of the step)
            if sharded and a.exchange == "nccl":
                sharded.AssembleDevice()        # kernels + NCCL interface exchange
                strmat.ctx.download(a_np, r_np)  # D2H of the CSR values and the load vector
            else:
                strmat.ctx.assemble(a_np, r_np)  # kernels (+ interface push) + D2H, overlapped
        def e2e_time():
            e2e_step()  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                e2e_step()
            barrier()
            return (time.perf_counter() - t0) / ksteps
        serial_s = None
        if not sharded:  # the same call without the download/assembly overlap, for the record
            strmat.ctx.set_option("overlap", 0)
            serial_s = e2e_time()
            strmat.ctx.set_option("overlap", 1)
        e2e_s = e2e_time()
        if world > 1:
            t = torch.tensor([e2e_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": nvol * world / e2e_s, "unit": "elements/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(x_np.nbytes), "d2h_bytes_per_step": int(a_np.nbytes + r_np.nbytes),
               "d2h_GBps_per_gpu": (a_np.nbytes + r_np.nbytes) / e2e_s / 1e9,
               "steps": ksteps, "note": "b200asm_set_nodes + b200asm_assemble(a_host, rhs_host), pinned host buffers; the D2H of "
               "finished CSR rows overlaps the assembly of later element chunks (option overlap)"}
        if numa_cpus:
            e2e["rank_cpu_affinity"] = f"{numa_cpus} CPUs closest to the GPU (NVML), pinned buffers first-touched there"
        if serial_s is not None:
            e2e["ms_per_step_without_overlap"] = serial_s * 1e3
        # the same call with the matrix LEFT on the device (a_host = NULL: what TPZB200CGSolver / b200asm_cg_solve consume): the
        # step then returns the load vector only, and the PCIe copy of the CSR values - the bound of `e2e` - disappears
        try:
            if not (sharded and a.exchange == "nccl"):
                def resident_step():
                    strmat.ctx.set_nodes(x_np)
                    strmat.ctx.assemble(None, r_np)
                resident_step()
                barrier()
                t0 = time.perf_counter()
                for _ in range(ksteps):
                    resident_step()
                barrier()
                res_s = (time.perf_counter() - t0) / ksteps
                if world > 1:
                    t = torch.tensor([res_s], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    res_s = float(t.item())
                e2e["matrix_resident"] = {"value": nvol * world / res_s, "unit": "elements/s", "ms_per_step": res_s * 1e3,
                                          "h2d_bytes_per_step": int(x_np.nbytes), "d2h_bytes_per_step": int(r_np.nbytes),
                                          "note": "b200asm_set_nodes + b200asm_assemble(NULL, rhs_host): the CSR values stay on the device for the device CG"}
        except Exception as ex:
            e2e["matrix_resident"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        del a_host, r_host, x_host, a_np, r_np

    # ---- the same mesh WITHOUT the node perturbation (the literal CreateGeoMeshOnGrid grid of BASELINE.json): every cell is
    # a parallelepiped, the context measures that on the device and switches the group to the closed-form kernel
    # (affine_hex.cuh).  Reported next to the headline, which stays on general (trilinear) hexahedra.
    uniform = None
    if world == 1 and a.topo == "hex" and a.p <= 2 and a.perturb != 0.0 and a.engine == 1 and not a.debug:
        moved = np.ascontiguousarray(mesh.nodes)
        strmat.ctx.set_nodes(gridmesh.grid_nodes(a.n, perturb=0.0))
        for _ in range(3):
            step_async()            # (the first one rebuilds the scatter maps in the closed-form kernel's layout)
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record(stream)
        for _ in range(a.steps):
            step_async()
        u1.record(stream)
        torch.cuda.synchronize()
        ums = u0.elapsed_time(u1) / a.steps
        f_el, b_el = algorithmic_work(a.topo, a.p, a.phys, nvol, neq, nnz)
        uniform = {"value": nvol / (ums * 1e-3), "unit": "elements/s", "ms_per_step": ums,
                   "kernel": ("gat::gather_rows_kernel (closed-form element matrices of parallelepiped hexahedra, every CSR row written once)" if a.gather else
                              "assemble_affine_hex_kernel (closed-form element matrices of parallelepiped hexahedra, one warp per element)"),
                   "hbm_GBps_algorithmic": b_el * nvol / (ums * 1e-3) / 1e9,
                   "hbm_frac_algorithmic": b_el * nvol / (ums * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                   "note": "same pattern and materials, unperturbed grid nodes; device-resident like `value`"}
        strmat.ctx.set_nodes(moved)
        step_async()
        torch.cuda.synchronize()

    cg = None
    if a.cg > 0 and world == 1:
        # the step after assembly on the same resident matrix: a.cg iterations of the reference's CG algorithm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _x, iters, resid = strmat.ctx.cg_solve(1, a.cg, 0.0, download=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        cg = {"iterations": iters, "relative_residual": resid, "ms_per_iteration": dt * 1e3 / max(iters, 1),
              "spmv_GBps_algorithmic": (12.0 * nnz + 24.0 * neq) * max(iters, 1) / dt / 1e9}

    roofline = case.roofline(peaks, fp64) if rank == 0 else None
    setup_s = {"flatten_mesh": case.t_flat, "pattern+upload+scatter_map": case.t_create, "pattern_builder": a.pattern}
    case.close()

    # ---- the other BASELINE.json configurations, device-resident, same timing rules (W >= 3, K steps, CUDA events, max over ranks)
    extras = {}
    for name in ([] if a.no_extra else [c for c in a.extra.split(",") if c]):
        spec = EXTRA_CONFIGS[name]
        b = copy.copy(a)
        b.topo, b.p, b.phys, b.perturb, b.variant = spec["topo"], spec["p"], spec["phys"], spec["perturb"], 0
        b.n = spec["n"] if world == 1 else spec["n_multi"]
        b.cpu_n = 0
        try:
            c = Case(b, rank, world, local_rank)
            st2, smp2 = threading.Event(), []
            th2 = threading.Thread(target=clocks_sampler, args=(st2, smp2, local_rank), daemon=True)
            ksteps = max(3, min(a.steps, 10))
            c.time_device(ksteps, a.warmup, sampler=(th2, st2) if rank == 0 else None)
            st2.set()
            if rank == 0:
                extras[name] = {"baseline_config": spec["baseline"],
                                "workload": workload_name(b) if world == 1 else workload_name(b).replace(f"{b.n}^3", f"{b.n}x{b.n}x{b.n * world}") + f", {world} z-slabs",
                                "value": c.value, "unit": "elements/s", "ms_per_step": c.ms_per_step, "steps": ksteps,
                                "dof": int(c.neq) * world if world == 1 else None, "dof_per_gpu": int(c.neq), "dof_per_s": None,
                                "volume_elements_per_gpu": int(c.nvol), "nnz_upper_per_gpu": int(c.nnz),
                                "roofline": c.roofline(peaks, fp64), "gpu_launches": c.launches, "clocks": summarize_clocks(smp2),
                                "setup_s": {"flatten_mesh": c.t_flat, "pattern+upload+scatter_map": c.t_create}}
            # total DOF of the global mesh = sum of the owned rows of all ranks
            if world > 1:
                t = torch.tensor([float(c.neq)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t)
                tot = int(t.item())
            else:
                tot = int(c.neq)
            if rank == 0:
                extras[name]["dof"] = tot
                extras[name]["dof_per_s"] = tot / (c.ms_per_step * 1e-3)
            c.close()
        except Exception as ex:  # an extra configuration never takes the headline down with it
            if rank == 0:
                extras[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            torch.cuda.empty_cache()
        if rank == 0 and not a.no_cpu_baseline and world == 1 and name in extras and "error" not in extras[name]:
            try:
                r = run_reference(b, 2, 1)
                extras[name]["cpu_baseline"] = {"value": r["value"], "unit": "elements/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as ex:
                extras[name]["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": str(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu, parity = None, None
    if not a.no_cpu_baseline and world == 1:   # (the CPU arm is timed on rank 0 at N = 1 only)
        import shutil
        import tempfile
        dump = tempfile.mkdtemp(prefix="b200asm_parity_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            r = run_reference(a, 3, 1, dumpdir=dump)
            cpu = {"value": r["value"], "unit": "elements/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "dof_per_s": r["dof_per_s"]}
            if r["kind"] == "reference":
                parity = parity_against_reference(a, dump, r["n"], local_rank)
        except Exception as ex:  # the baseline is reported, never the product path
            cpu = {"value": None, "unit": "elements/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(ex)[:200]}
        finally:
            shutil.rmtree(dump, ignore_errors=True)

    dropin = None
    if world == 1 and not a.no_dropin and not a.no_cpu_baseline:
        try:
            dropin = dropin_timing(a, local_rank)
        except Exception as ex:
            dropin = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}

    line = {"metric": "assembled volume elements/s (Assemble on a created pattern)", "value": value, "unit": "elements/s",
            "dof_per_s": neq * world / (ms_per_step * 1e-3),
            "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a) if world == 1 else workload_name(a).replace(f"{a.n}^3", f"{a.n}x{a.n}x{a.n * world}") +
                       f", {world} z-slabs, row-sharded CSR, interface rows pushed over NVLink ({a.exchange})", "volume_elements_per_gpu": nvol, "dof_per_gpu": neq, "nnz_upper_per_gpu": nnz,
                       "l2": "inputs larger than L2 (CSR values %.1f GB + scatter map rewritten every step)" % (nnz * 8 / 1e9),
                       "perturbed_nodes": a.perturb != 0.0, "element_order": "Morton curve inside 16 element chunks" if a.locality else "mesh order", "engine": "dmma" if a.engine == 1 else "dfma register tiles", "scatter": a.scatter, "setup_s": setup_s},
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": launches,
            "clocks": summarize_clocks(samples), "step_ms": step_ms}
    if extras:
        line["configs"] = extras
    if cg:
        line["device_cg"] = cg
    if uniform:
        line["uniform_grid"] = uniform
    if dropin:
        line["dropin"] = dropin
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
