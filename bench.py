#!/usr/bin/env python
"""bench.py -- stencil GLUP/s (f64 3D 7-point Jacobi sweep + fused ghost fill + update norm) on
N B200s, through librnmf_b200.so (C ABI).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--sub-iters S]
  python bench.py --impl reference ...     # the reference's CPU algorithm (oracle port) on host cores

step      = one solve_poisson-shaped call (poisson.rs:67-92): S = --sub-iters Jacobi sweeps, each
            followed by the ghost fill (fused), plus the L1 norm of the last update (main.rs:68-69);
            S defaults to 550 = n_sub_iter of examples/magnetostatics/magnetostatics.lua:12.
value     = lattice updates (valid cells x sweeps, all ranks) / time, inputs resident in HBM.
e2e       = the same call through rnmf_solve_poisson_host: pinned HOST frames in the reference's
            dense layout in, result frame out; H2D + S sweeps + D2H inside the timed region.
workloads = 3d7_768 (default; SURVEY 8d C5: 768^3 per GPU, weak scaling, mixed Dirichlet/Neumann),
            3d7_512 (C3), 3d7_1024 (C4, strong), 2d5_8192 (C2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_LUP = 24          # read psi (8) + read rhs (8) + write psi' (8)   
SEED_PSI, SEED_RHS = 0x5EED0001, 0x5EED0002

MIXED_3D = ([("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 0.25)],
            [("dirichlet", 1.0), ("neumann", -0.5), ("neumann", 0.5)])
DIRICHLET_2D = ([("dirichlet", 0.0), ("dirichlet", 0.0)], [("dirichlet", 1.0), ("dirichlet", 1.0)])

WORKLOADS = {
    # name: (per-GPU or global n, scaling, bc spec, description)
    "3d7_768": dict(n=(768, 768, 768), scaling="weak", bc=MIXED_3D,
                    desc="3D 7-point f64 Poisson Jacobi sweep, 768^3 cells per GPU (z-slabs), mixed Dirichlet/Neumann ghost fill"),
    "3d7_512": dict(n=(512, 512, 512), scaling="weak", bc=MIXED_3D,
                    desc="3D 7-point f64 Poisson Jacobi sweep, 512^3 per GPU, mixed Dirichlet/Neumann ghost fill"),
    "3d7_1024": dict(n=(1024, 1024, 1024), scaling="strong", bc=MIXED_3D,
                     desc="3D 7-point f64 Poisson Jacobi sweep, 1024^3 global, z-slab decomposed"),
    "2d5_8192": dict(n=(8192, 8192), scaling="weak", bc=DIRICHLET_2D,
                     desc="2D 5-point f64 Poisson Jacobi sweep, 8192^2 per GPU (y-slabs), Dirichlet"),
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="rnmf_b200", choices=["rnmf_b200", "reference"])
    p.add_argument("--workload", default="3d7_768", choices=sorted(WORKLOADS))
    p.add_argument("--sub-iters", type=int, default=550)
    p.add_argument("--e2e-steps", type=int, default=2, help="timed end-to-end calls (after 1 warm-up)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-others", action="store_true", help="skip the short extra runs of the other configs")
    p.add_argument("--variant", default=None, help="sweep kernel variant (measurement knob)")
    p.add_argument("--chunk", default=None, help="marching chunk per CTA (measurement knob)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,clocks.mem,clocks.max.mem")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[3 + i].lower().startswith("active")})
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        mem = sorted(float(s[7]) for s in self.samples if len(s) > 8 and s[7].replace(".", "").isdigit())
        mem_max = [float(s[8]) for s in self.samples if len(s) > 8 and s[8].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "power_w_max": max(pw) if pw else None,
                "mem_mhz": mem[len(mem) // 2] if mem else None, "mem_max_mhz": max(mem_max) if mem_max else None}


def copy_bandwidth_this_box():
    """The driver's own peak recipe (MEASURED_PEAKS.json `how`) repeated on THIS box, for context:
    torch b.copy_(a) over 1 Gi bf16 elements, read+write bytes, best of 10, CUDA events."""
    import torch
    n = 1 << 30
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    a.zero_()
    best = None
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    # the same copy back to back for ~2 s: what the box sustains once the 1 kW power cap bites
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int(2000.0 / best))
    e0.record()
    for _ in range(reps):
        b.copy_(a)
    e1.record()
    e1.synchronize()
    sustained = 2.0 * n * 2 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n * 2 / (best * 1e-3) / 1e9, sustained


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(key: str):
    """(dram bytes per launch of the dominant kernel, which capture they come from) from the committed ncu captures
    (profiles/traffic.json: `ncu --set full`, dram__bytes_read.sum + dram__bytes_write.sum of ONE launch); (None, None) if
    no capture exists for this workload / kernel.  A property of the kernel at this shape, not of the timed run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(key), (t.get("_sources") or {}).get(key)
    except Exception:
        return None, None


def build_roofline(main_res: dict, steps: int, sub_iters: int, workload: str, peak: float, peak_src: str) -> dict:
    """The `roofline` object of the JSON line from one timed region (pure; tested on the CPU).
    main_res: ms (timed region, max over ranks), doubles (double-sweep launches in it), multi_launches / multi_sweeps
    (temporally blocked launches and the sweeps they covered), local_cells, and -- when measured -- kernel_ms: the
    dominant kernel's average launch duration from CUDA events around `kernel_launches_timed` back-to-back launches."""
    sweep_ms = main_res["ms"] / (steps * sub_iters)                       # average time per sweep over the timed region
    double = main_res["doubles"] * 2 > steps * sub_iters // 2
    sweeps_per_launch = 2 if double else 1
    # K sweeps per launch: average depth of the temporally blocked launches of the region
    if main_res.get("multi_launches") and main_res.get("multi_sweeps", 0) * 2 > steps * sub_iters:
        depth = main_res["multi_sweeps"] / main_res["multi_launches"]
        double = True
        sweeps_per_launch = int(round(depth)) if abs(depth - round(depth)) < 1e-9 else depth
    step_share_ms = sweep_ms * sweeps_per_launch          # the launch's share of the step (an upper bound: the step also holds the
    #                                                       norm-carrying single sweep, the shell copy and the reduction)
    launch_ms = main_res.get("kernel_ms") or step_share_ms
    alg_bytes = BYTES_PER_LUP * main_res["local_cells"] * sweeps_per_launch   # 24 B per update x updates per launch
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9                           # GB/s per GPU (the slowest rank's clock)
    traffic, traffic_src = ncu_traffic(workload + ("_tb2" if sweeps_per_launch == 2 else f"_tb{sweeps_per_launch}" if double else ""))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": (f"{'double' if sweeps_per_launch == 2 else 'multi'} sweep ({sweeps_per_launch} sweeps + fused ghost fills per launch)"
                           if double else "jacobi sweep (fused ghost fill)"),
                "sweeps_per_launch": sweeps_per_launch, "double_sweep_launches": main_res["doubles"],
                "algorithmic_bytes_per_launch": alg_bytes,
                # the one-pass-per-sweep roofline allows sweeps_per_launch x 24 B-updates per byte pass: frac / K is the
                # fraction of the ceiling a K-sweeps-per-pass kernel could reach if DRAM were its only limit
                "frac_of_k_pass_ceiling": achieved / peak / sweeps_per_launch,
                "avg_launch_ms": launch_ms, "avg_launch_ms_from_step": step_share_ms,
                "launch_timing": (f"CUDA events around {main_res.get('kernel_launches_timed')} back-to-back launches on the library's stream, "
                                  "right after the timed steps (same clocks / power state)" if main_res.get("kernel_ms")
                                  else "the launch's share of the timed region"),
                "peak_source": peak_src + " -- of measured" if "MEASURED" in peak_src else peak_src}
    if double:
        roofline["note"] = ("achieved counts the ALGORITHMIC 24 B per update; the kernel performs several updates per byte pass "
                            "(temporal blocking), so frac can exceed 1. dram_frac = measured DRAM traffic per launch / launch time / peak.")
    if traffic:
        roofline["dram_gbs"] = traffic / (launch_ms * 1e-3) / 1e9
        roofline["dram_frac"] = roofline["dram_gbs"] / peak
    return roofline


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle's reference-faithful Gauss-Seidel sweep (poisson.rs:80-89 loop order)
# ------------------------------------------------------------------------------------------------
def cpu_sample_geom(workload: str):
    """Bounded sample of the workload for the CPU: same x/y extents (so the reference's
    stride-G0 column walk sees the same strides), a thin slab in the slowest direction."""
    w = WORKLOADS[workload]
    n = list(w["n"])
    n[-1] = 32 if len(n) == 3 else 512
    return n


def cpu_reference_run(workload: str, sweeps: int, reps: int, jacobi_omp: bool = False, jacobi_1core: bool = False):
    """returns (seconds per rep list, LUPs per rep)"""
    import numpy as np
    from oracle import oracle as ora
    w = WORKLOADS[workload]
    n = cpu_sample_geom(workload)
    g = ora.geom(n, hi=[float(v) for v in n])
    bc = ora.bcs(g.dim, [w["bc"]])
    nv = int(np.prod(n))
    psi, rhs = ora.zeros(g), ora.zeros(g)
    ora.valid_view(g, psi)[0] = ora.splitmix_field(nv, SEED_PSI).reshape(tuple(reversed(n)))
    ora.valid_view(g, rhs)[0] = ora.splitmix_field(nv, SEED_RHS).reshape(tuple(reversed(n)))
    ora.fill_bc(g, psi, bc)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        if jacobi_omp or jacobi_1core:
            ora.jacobi(g, psi, rhs, bc, sweeps, omp=jacobi_omp, want_norm=False)
        else:
            ora.gs(g, psi, rhs, bc, sweeps)
        times.append(time.perf_counter() - t0)
    return times, nv * sweeps


def run_reference_arm(args):
    """--impl reference: the reference's own CPU algorithm for the path.  The Rust crate cannot be
    built in this image (no cargo/rustc), so this is the oracle port (oracle/rnmf_oracle.c):
    lexicographic Gauss-Seidel in the reference's loop order, 1 thread (the reference is
    single-threaded; rayon is declared but unused)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sweeps = 1
    times, lups = cpu_reference_run(args.workload, sweeps, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = lups * len(timed) / total / 1e9
    n = cpu_sample_geom(args.workload)
    sample = (f"each step = {sweeps} reference-order Gauss-Seidel sweep (poisson.rs:80-89 loop order) over a "
              f"{'x'.join(map(str, n))} slab of the workload (same x/y extents, {n[-1]} slices), 1 thread")
    line = {
        "impl": "reference", "metric": "stencil GLUP/s (f64 3D 7-pt)" if len(n) == 3 else "stencil GLUP/s (f64 2D 5-pt)",
        "value": value, "unit": "GLUP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (splitmix64 field, same seeds as the GPU arm)",
        "config": {"workload": args.workload, "desc": w["desc"], "update_order": "lexicographic Gauss-Seidel (reference)",
                   "sub_iters": sweeps},
        "cpu_baseline": {"value": value, "unit": "GLUP/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "an UPPER bound for the reference: the thin slab shortens its stride-G0*G1 walk over the slowest index, "
                                 "and the C port has none of Rust's bounds checks / iterator overhead (SURVEY 8d)"},
        "e2e": {"value": value, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"nproc": os.cpu_count(), "cpu": _cpu_model()},
    }
    print(json.dumps(line), flush=True)


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_bcs(R, spec):
    mk = {"dirichlet": R.BcType.Dirichlet, "neumann": R.BcType.Neumann}
    return R.BCs([R.ComponentBCs([mk[k](v) for k, v in spec[0]], [mk[k](v) for k, v in spec[1]])])


def global_shape(workload: str, nranks: int):
    w = WORKLOADS[workload]
    n = list(w["n"])
    if w["scaling"] == "weak":
        n[-1] *= nranks
    return n


def timed_steps(ctx, psi, rhs, coef, sub_iters, steps, barrier):
    """K steps, device-timed on the context's stream; returns (ms, last norm)."""
    from rnmf_b200 import poisson
    barrier()
    ctx.sync()
    ctx.timer_begin()
    l1 = 0.0
    for _ in range(steps):
        l1 = poisson.jacobi_sweeps(psi, rhs, sub_iters, want_norm=True, coef=coef)
    ms = ctx.timer_end()
    ctx.sync()
    barrier()
    return ms, l1


def run_config(R, ctx, workload, nranks, sub_iters, steps, warmup, barrier, reduce_max, kernel_launches_timed=0):
    """One workload on the context's device(s): warm-up, K device-timed steps, and -- kernel_launches_timed > 0 -- a direct
    timing of the dominant kernel right afterwards (that many back-to-back launches, no norm, CUDA events on the library's stream)."""
    from rnmf_b200 import poisson
    w = WORKLOADS[workload]
    n = global_shape(workload, nranks)
    dim = len(n)
    mesh = R.CartesianMesh.new([0.0] * dim, [float(v) for v in n], n)       # dx = 1
    bc = make_bcs(R, w["bc"])
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx, slab=nranks > 1)
    rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx, slab=nranks > 1)
    psi.fill_synthetic(SEED_PSI)
    rhs.fill_synthetic(SEED_RHS)
    psi.fill_bc()
    if nranks > 1 and os.environ.get("RNMF_PEER_HALO", "1") != "0":
        psi.enable_peer_halo()       # fused compute + halo exchange over NVLink peer memory
    coef = poisson.poisson_coefs(dim, mesh.dx)
    for _ in range(warmup):
        poisson.jacobi_sweeps(psi, rhs, sub_iters, want_norm=True, coef=coef)
    launches0, doubles0, multi0 = ctx.kernel_launches, ctx.double_sweep_launches, ctx.multi_sweep_stats
    ms, l1 = timed_steps(ctx, psi, rhs, coef, sub_iters, steps, barrier)
    launches = ctx.kernel_launches - launches0
    doubles = ctx.double_sweep_launches - doubles0
    multi1 = ctx.multi_sweep_stats
    multi = (multi1[0] - multi0[0], multi1[1] - multi0[1])      # temporally blocked launches, sweeps they covered
    ms = reduce_max(ms)
    cells = 1
    for v in n:
        cells *= v
    lups = cells * sub_iters * steps
    checksum = psi.ctx.allreduce_sum(psi.abs_sum())             # of the field after (warmup + steps) x sub_iters sweeps
    res = dict(n_global=n, ms=ms, ms_per_step=ms / steps, glups=lups / (ms * 1e-3) / 1e9, launches=launches, doubles=doubles, multi_launches=multi[0], multi_sweeps=multi[1],
               l1_last=l1, abs_sum=checksum, local_cells=int(psi.n_local[0] * psi.n_local[1] * (psi.n_local[2] if dim == 3 else 1)))
    if kernel_launches_timed > 0:
        # the dominant kernel alone: a call of depth x L sweeps without the norm is L launches of the temporally blocked
        # kernel (depth sweeps each) behind one O(surface) shell copy -- timed with CUDA events on the library's stream
        depth = int(round(multi[1] / multi[0])) if multi[0] and multi[1] * 2 > sub_iters * steps else 1
        m0 = ctx.multi_sweep_stats
        barrier()
        ctx.timer_begin()
        poisson.jacobi_sweeps(psi, rhs, depth * kernel_launches_timed, want_norm=False, coef=coef)
        kms = reduce_max(ctx.timer_end())
        ctx.sync()
        barrier()
        m1 = ctx.multi_sweep_stats
        if depth == 1 or m1[0] - m0[0] == kernel_launches_timed:
            res["kernel_ms"] = kms / kernel_launches_timed
            res["kernel_launches_timed"] = kernel_launches_timed
    psi.close()
    rhs.close()
    return res


def run_c1_magnetostatics(R, ctx):
    """BASELINE config 1: examples/magnetostatics/magnetostatics.lua run whole (301^2, 10 outer iterations x 550 sweeps,
    main.rs:63-74) on the device, in the reference's update order (K6, bit-identical to the reference's lexicographic
    Gauss-Seidel) and in Jacobi order, next to the oracle's CPU time for the same case (1 thread)."""
    from rnmf_b200 import config as rconfig, poisson
    out = {}
    path = os.path.join(ROOT, "cases", "ferro_droplet.lua")   # same keys and values as examples/magnetostatics/magnetostatics.lua
    cfg = rconfig.read_lua(path)
    out["case"] = os.path.relpath(path, ROOT)
    out["n_cells"] = list(cfg.geom.n_cells)
    out["n_iter"], out["n_sub_iter"] = int(cfg.model.n_iter), int(cfg.model.n_sub_iter)
    for order, key in (("reference", "gpu_reference_order_s"), ("jacobi", "gpu_jacobi_order_s")):
        poisson.magnetostatics(cfg, order=order, ctx=ctx)[0].close()          # warm-up (module load, allocations)
        ctx.sync()
        t0 = time.perf_counter()
        psi, trace, it = poisson.magnetostatics(cfg, order=order, ctx=ctx)
        ctx.sync()
        out[key] = time.perf_counter() - t0
        out[key.replace("_s", "_sum_diff_last")] = trace[-1]
        out["outer_iterations"] = it
        psi.close()
    try:
        import ctypes as C
        import numpy as np
        from oracle import oracle as ora                     # cpu_baseline leg: the oracle as the timed CPU port
        m = cfg.model
        g = ora.geom(list(cfg.geom.n_cells), hi=list(cfg.geom.length))
        om = ora.model(h_far=tuple(m.h_far), relax=m.relax, n_iter=int(m.n_iter), n_sub_iter=int(m.n_sub_iter),
                       bubble_centre=tuple(m.bubble_centre), bubble_radius=m.bubble_radius, mu=tuple(m.mu), tol=m.tol)
        ref, tr, it = ora.zeros(g), np.zeros(int(m.n_iter)), C.c_int64()
        t0 = time.perf_counter()
        ora.lib().ora_magnetostatics_run(C.byref(g), ref, C.byref(om), 0, tr, C.byref(it))
        out["cpu_oracle_reference_order_s"] = time.perf_counter() - t0
        out["cpu_sum_diff_last"] = float(tr[it.value - 1])
        out["cpu_cores"] = 1
    except Exception as e:                                   # reported, never hidden
        out["cpu_error"] = str(e)
    return out


def run_e2e(R, ctx, workload, sub_iters, steps):
    """The reference-facing call with HOST frames (pinned), single rank."""
    import ctypes as C
    import numpy as np
    from rnmf_b200 import _capi
    from rnmf_b200.dataframe import _faces_array
    w = WORKLOADS[workload]
    n = list(w["n"])
    dim = len(n)
    L = _capi.lib()
    G = [v + 2 for v in n]
    count = 1
    for v in G:
        count *= v
    nbytes = count * 8
    bufs = []
    for _ in range(3):
        p = C.c_void_p()
        _capi.check(L.rnmf_host_alloc(nbytes, C.byref(p)))
        bufs.append(p)
    try:
        arrs = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(count,)) for p in bufs]
        psi_h, rhs_h, out_h = arrs
        # host frames in the reference's dense layout: deterministic, cheap to generate
        shape = tuple(reversed(G))
        rng = np.random.default_rng(7)
        blk = rng.standard_normal(shape[-1] * shape[-2])
        psi_h.reshape(-1, shape[-1] * shape[-2])[:] = blk
        rhs_h.reshape(-1, shape[-1] * shape[-2])[:] = blk[::-1]
        faces = _faces_array(make_bcs(R, w["bc"]), dim, 1)
        nn = (C.c_int64 * 3)(*n, *([1] * (3 - dim)))
        lo = (C.c_double * 3)(0.0, 0.0, 0.0)
        dx = (C.c_double * 3)(1.0, 1.0, 1.0)
        l1 = C.c_double()

        def call():
            _capi.check(L.rnmf_solve_poisson_host(ctx._h, dim, nn, lo, dx, 1, faces, bufs[0], bufs[1], sub_iters, 0,
                                                  bufs[2], C.byref(l1)))
        call()                                   # warm-up (allocates the cached device frames)
        ctx.sync()
        t0 = time.perf_counter()
        ctx.timer_begin()
        for _ in range(steps):
            call()
        ms_dev = ctx.timer_end()
        ctx.sync()
        wall = time.perf_counter() - t0
        cells = 1
        for v in n:
            cells *= v
        secs = max(wall, ms_dev * 1e-3)
        skew = int(os.environ.get("RNMF_E2E_SKEW", "0") or 0)          # opt-in time-skewed upload / download (capi.cu); same bits
        return {"value": cells * sub_iters * steps / secs / 1e9, "unit": "GLUP/s",
                "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes + 8,
                "ms_per_step": 1e3 * secs / steps, "steps": steps,
                "call": "rnmf_solve_poisson_host: pinned host psi+rhs (dense reference layout) -> H2D -> "
                        f"{sub_iters} sweeps(+fill_bc) + L1 norm -> D2H psi" + (f" (e2e_skew={skew})" if skew else ""),
                # of the last call's output frame: equal across e2e_skew settings (the field bit for bit, the norm to rounding)
                "checks": {"abs_sum_out": float(sum(float(np.abs(out_h[i:i + (1 << 24)]).sum()) for i in range(0, count, 1 << 24))),
                           "l1_update": l1.value}}
    finally:
        for p in bufs:
            L.rnmf_host_free(p)


def run_e2e_dist(R, ctx, workload, nranks, sub_iters, steps, barrier, reduce_max, reduce_sum):
    import ctypes as C
    import numpy as np
    from rnmf_b200 import _capi, poisson
    w = WORKLOADS[workload]
    n = global_shape(workload, nranks)
    dim = len(n)
    L = _capi.lib()
    mesh = R.CartesianMesh.new([0.0] * dim, [float(v) for v in n], n)
    bc = make_bcs(R, w["bc"])
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx, slab=True)
    rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx, slab=True)
    if os.environ.get("RNMF_PEER_HALO", "1") != "0":
        psi.enable_peer_halo()
    count = psi.n_nodes                               # this rank's dense grown slab
    nbytes = count * 8
    bufs = []
    for _ in range(3):
        p = C.c_void_p()
        _capi.check(L.rnmf_host_alloc(nbytes, C.byref(p)))
        bufs.append(p)
    try:
        arrs = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(count,)) for p in bufs]
        row = int(psi.n_grown[0] * psi.n_grown[1])
        blk = np.random.default_rng(7 + ctx.rank).standard_normal(row)
        arrs[0].reshape(-1, row)[:] = blk
        arrs[1].reshape(-1, row)[:] = blk[::-1]
        coef = poisson.poisson_coefs(dim, mesh.dx)

        coef4 = (C.c_double * 4)(*coef)
        l1c = C.c_double()

        def call():
            # the per-rank end-to-end call on slab frames: upload (chunked, overlapped with the first double sweeps; the wedges
            # next to the neighbours catch up once every rank's frames are up), halo-coupled sweeps, norm, download
            _capi.check(L.rnmf_jacobi_sweeps_host(psi._h, rhs._h, bufs[0], bufs[1], coef4, sub_iters, bufs[2], C.byref(l1c)))
            return l1c.value
        call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        ctx.sync()
        barrier()
        secs = reduce_max(time.perf_counter() - t0)
        cells = 1
        for v in n:
            cells *= v
        return {"value": cells * sub_iters * steps / secs / 1e9, "unit": "GLUP/s",
                "h2d_bytes_per_step": int(reduce_sum(2.0 * nbytes)), "d2h_bytes_per_step": int(reduce_sum(nbytes + 8.0)),
                "ms_per_step": 1e3 * secs / steps, "steps": steps,
                "h2d_gbs_per_rank": 2.0 * nbytes / 1e9 / (secs / steps), "d2h_gbs_per_rank": nbytes / 1e9 / (secs / steps),
                "call": "per rank: rnmf_jacobi_sweeps_host(psi slab, rhs slab): pinned host slabs (dense reference layout) -> chunked H2D "
                        f"overlapped with the first double sweeps -> {sub_iters} halo-coupled sweeps(+fill_bc) + all-reduced L1 norm -> D2H psi slab",
                "note": "h2d/d2h_gbs_per_rank = bytes / WHOLE call time (a lower bound of the copy rates: the copies overlap the sweeps)"}
    finally:
        for p in bufs:
            L.rnmf_host_free(p)
        psi.close()
        rhs.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rnmf_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import rnmf_b200 as R

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = R.Context.from_torch_distributed()

        def barrier():
            dist.barrier()
            torch.cuda.synchronize()

        def reduce_max(x):
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def reduce_sum(x):
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
    else:
        ctx = R.Context(local)

        def barrier():
            torch.cuda.synchronize()

        def reduce_max(x):
            return x

        def reduce_sum(x):
            return x

    if args.variant is not None:
        ctx.set_option("sweep_variant", args.variant)
    if args.chunk is not None:
        ctx.set_option("sweep_chunk", args.chunk)

    w = WORKLOADS[args.workload]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main_res = run_config(R, ctx, args.workload, world, args.sub_iters, args.steps, args.warmup, barrier, reduce_max,
                          kernel_launches_timed=60)
    if rank == 0:
        sampler.stop()
    launches_total = int(reduce_sum(float(main_res["launches"])))

    peak, peak_src = measured_peak()
    roofline = build_roofline(main_res, args.steps, args.sub_iters, args.workload, peak, peak_src)
    achieved = roofline["achieved"]
    if rank == 0:
        try:
            burst, sustained = copy_bandwidth_this_box()
            roofline["copy_gbs_this_box"] = burst
            roofline["copy_gbs_this_box_sustained_2s"] = sustained
            roofline["frac_of_this_box_copy"] = achieved / burst
            roofline["frac_of_this_box_copy_sustained"] = achieved / sustained
        except Exception as e:
            roofline["copy_gbs_this_box"] = None
            roofline["copy_error"] = str(e)

    # The other BASELINE configs, short: C3 (512^3, single GPU only), C2 (8192^2 per GPU, y-slabs at N > 1) and C4 (1024^3
    # GLOBAL, z-slabs: strong scaling).  Every line carries the checksums of its final field; C4's domain, seeds and sweep
    # count are the same at every N, so its checksums must agree across N = 1, 2, 4, 8 (to the rounding of the all-reduce).
    others = {}
    if not args.no_others:
        for name, sub, st in (("3d7_512", 101, 3), ("2d5_8192", 101, 3), ("3d7_1024", 61, 2)):     # odd: every sweep but the norm-carrying last one is temporally blocked
            if name == args.workload or (name == "3d7_512" and world > 1):
                continue
            try:
                r = run_config(R, ctx, name, world, sub, st, 1, barrier, reduce_max)
                ms_sweep = r["ms"] / (sub * st)
                others[name] = {"glups": r["glups"], "ms_per_sweep": ms_sweep, "scaling": WORKLOADS[name]["scaling"],
                                "n_global": r["n_global"], "sweeps": sub * (st + 1),
                                "hbm_frac_per_gpu": BYTES_PER_LUP * r["local_cells"] / (ms_sweep * 1e-3) / 1e9 / peak,
                                "abs_sum_psi": r["abs_sum"], "l1_update_last": r["l1_last"],
                                "temporally_blocked_launches": r["multi_launches"], "sweeps_in_them": r["multi_sweeps"]}
            except Exception as e:       # reported, never hidden
                others[name] = {"error": str(e)}
        if rank == 0 and world == 1:
            try:
                others["c1_magnetostatics_lua"] = run_c1_magnetostatics(R, ctx)
            except Exception as e:
                others["c1_magnetostatics_lua"] = {"error": str(e)}

    e2e = None
    if rank == 0 and not args.no_e2e:
        if world == 1:
            e2e = run_e2e(R, ctx, args.workload, args.sub_iters, max(1, args.e2e_steps))
    if world > 1 and not args.no_e2e:
        # multi-GPU end-to-end through the data-frame API: every rank uploads ITS SLAB of the global
        # host frames (pinned, dense reference layout), the ranks run the halo-coupled sweeps together,
        # every rank downloads its slab of the result.  Timed region = barrier .. barrier, max over ranks.
        e2e = run_e2e_dist(R, ctx, args.workload, world, args.sub_iters, max(1, args.e2e_steps), barrier, reduce_max, reduce_sum)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, lups = cpu_reference_run(args.workload, 1, 12)
        best = lups * len(times[2:]) / sum(times[2:]) / 1e9
        n = cpu_sample_geom(args.workload)
        cpu_baseline = {"value": best, "unit": "GLUP/s", "cores": 1, "kind": "port",
                        "sample": f"10 timed (2 warm-up) reference-order Gauss-Seidel sweeps (poisson.rs:80-89 loop order, oracle port) "
                                  f"over a {'x'.join(map(str, n))} slab of the workload, 1 thread; the Rust reference is single-threaded",
                        "note": "an UPPER bound for the reference: the thin slab shortens its stride-G0*G1 walk over the slowest index, and "
                                "the C port has none of Rust's bounds checks / iterator overhead (SURVEY 8d)",
                        "host": {"nproc": os.cpu_count(), "cpu": _cpu_model()}}
        try:
            t1, l1c = cpu_reference_run(args.workload, 4, 3, jacobi_1core=True)
            cpu_baseline["jacobi_same_expression_1_core"] = {"value": l1c * len(t1[1:]) / sum(t1[1:]) / 1e9, "unit": "GLUP/s", "cores": 1}
        except Exception as e:
            cpu_baseline["jacobi_same_expression_1_core"] = {"error": str(e)}
        try:
            from oracle import oracle as ora
            t2, l2 = cpu_reference_run(args.workload, 8, 4, jacobi_omp=True)      # 8 sweeps per call: amortises the ping-pong copies
            cpu_baseline["jacobi_all_cores_not_the_reference"] = {"value": l2 * len(t2[1:]) / sum(t2[1:]) / 1e9, "unit": "GLUP/s",
                                                                  "cores": ora.lib().ora_num_threads()}
        except Exception as e:
            cpu_baseline["jacobi_all_cores_not_the_reference"] = {"error": str(e)}

    if rank == 0:
        dim = len(w["n"])
        line = {
            "metric": f"stencil GLUP/s (f64 {'3D 7-pt' if dim == 3 else '2D 5-pt'})",
            "value": main_res["glups"], "unit": "GLUP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (splitmix64 field generated on device; random-free, deterministic)",
            "config": {"workload": args.workload, "desc": w["desc"], "n_global": main_res["n_global"],
                       "sub_iters": args.sub_iters, "step": "solve_poisson-shaped call: sub_iters Jacobi sweeps + fused fill_bc each + L1 update norm",
                       "n_ghost": 1, "bytes_per_lup": BYTES_PER_LUP,
                       "l2": "three 3.65 GB frames per GPU >> 126 MB L2: no flush needed" if args.workload == "3d7_768" else "frames >> L2",
                       "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
                       "halo": ("fused peer-memory stores from the sweep kernel (NVLink) + epoch flags" if os.environ.get("RNMF_PEER_HALO", "1") != "0"
                                else "NCCL send/recv overlapped with the interior") if world > 1 else "none",
                       "sweep_variant": args.variant or os.environ.get("RNMF_SWEEP_VARIANT", "auto")},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches_total,
            "clocks": sampler.summary(),
            "checks": {"l1_update_last": main_res["l1_last"], "abs_sum_psi": main_res["abs_sum"]},
            "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
