#!/usr/bin/env python3
"""GPU box (1 GPU): three against two sweeps per launch on THIN frames -- the per-GPU shapes of the strong-scaling leg (1024 x 1024 x
1024/N planes) without neighbours -- to see what chunk / wave quantisation does to either kernel.  Prints GLUP/s per (planes, variant)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnmf_b200 as R
from rnmf_b200 import poisson
import bench

ctx = R.Context(0)
for planes in [int(v) for v in os.environ.get("RNMF_PROBE_PLANES", "128,256,512").split(",")]:
    n = [1024, 1024, planes]
    mesh = R.CartesianMesh.new([0.0] * 3, [float(v) for v in n], n)
    bc = bench.make_bcs(R, bench.MIXED_3D)
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    psi.fill_synthetic(1)
    rhs.fill_synthetic(2)
    psi.fill_bc()
    coef = poisson.poisson_coefs(3, mesh.dx)
    chunks = [int(v) for v in os.environ.get("RNMF_PROBE_CHUNKS", "0").split(",")]       # 0 = the planner's choice
    for variant, chunk in [(v, c) for v in ("auto", 31) for c in chunks]:
        ctx.set_option("sweep_variant", variant)
        ctx.set_option("sweep_chunk", chunk)
        sweeps = 600
        poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=False, coef=coef)      # warm-up, brings the clocks to the sustained state
        ctx.sync()
        ctx.timer_begin()
        poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=False, coef=coef)
        ms = ctx.timer_end()
        print(f"planes {planes:4d} variant {variant} chunk {chunk:3d}: {n[0] * n[1] * n[2] * sweeps / (ms * 1e-3) / 1e9:7.1f} GLUP/s  ({ms / sweeps:.4f} ms per sweep)", flush=True)
    ctx.set_option("sweep_variant", "auto")
    ctx.set_option("sweep_chunk", 0)
    psi.close()
    rhs.close()
