#!/usr/bin/env python3
"""Shared-memory wavefronts per cell update of the temporally blocked kernels, from the kernel SOURCES run on the
CPU-side SIMT emulator with its bank model switched on (tests/emu/emu_cuda.hpp).  No GPU needed.  A static model:
LSU wavefronts of the consumer / halo warps' ld/st.shared only (TMA writes into shared memory are not counted),
on frames where interior tiles dominate.  Usage: python scripts/smem_wavefronts.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as ora  # noqa: E402  (checker of the run, as in the tests)
import test_emulated_kernels as T  # noqa: E402


def run(lib, n, hi, spec, variant, chunk, sweeps_per_launch):
    g, psi, rhs, want, kinds, vals, dx, nn, coef = T._setup(n, hi, spec, sweeps_per_launch, seed=7)
    out = np.zeros_like(psi)
    lib.emu_trace(1)
    rc = lib.emu_double_sweeps(len(n), nn, dx, kinds, vals, T._p(coef), T._p(psi), T._p(rhs), T._p(out), variant, chunk, 1)
    lib.emu_trace(0)
    assert rc == 0 and np.array_equal(out, want)
    t = (C.c_longlong * 4)()
    lib.emu_trace_read(t)
    updates = int(np.prod(n)) * sweeps_per_launch
    return [x / updates for x in t]


def main():
    lib = T.emu.__wrapped__() if hasattr(T.emu, "__wrapped__") else None
    if lib is None:
        import subprocess
        deps_ok = os.path.exists(T.OUT)
        if not deps_ok:
            subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emulated_kernels.py"), "-q", "-k", "emulator_reports"], check=True)
        lib = C.CDLL(T.OUT)
        lib.emu_set_timeout(60)
    rows = []
    # 2D: 3 strips of 512 (2 of 768 for the 768-column variants), one 96-row chunk
    for name, variant, depth, n in (("2D tb2 (30, default)", 30, 2, [1536, 96]), ("2D tbk K=2 (35)", 35, 2, [1536, 96]),
                                    ("2D tbk K=3 (32)", 32, 3, [1536, 96]), ("2D tbk K=3 de-interleaved (36)", 36, 3, [1536, 96]),
                                    ("2D tbk K=4 768 cols (34)", 34, 4, [1536, 96]), ("2D tbk K=4 768 cols de-interleaved (37)", 37, 4, [1536, 96]), ("2D tbk K=4 512 cols de-interleaved (38)", 38, 4, [1536, 96]), ("2D tbk K=5 768 cols de-interleaved (39)", 39, 5, [1536, 96]), ("2D tbk K=6 768 cols de-interleaved (40)", 40, 6, [1536, 96])):
        rows.append((name, run(lib, n, [1.0, 1.0], T.MIXED_2D, variant, 96, depth)))
    # 3D: 3 x 2 tiles of 128 x 12 (x 16 for variant 25), one 24-plane chunk
    for name, variant, n in (("3D tb2 128x12 (0/20, default)", 20, [384, 24, 24]), ("3D tb2x two rows per warp (24)", 24, [384, 24, 24]),
                             ("3D tb2x 128x16 (25)", 25, [384, 32, 24]), ("3D tb2 + interior loop (26)", 26, [384, 24, 24]), ("3D tb2x + interior loop (27)", 27, [384, 24, 24]), ("3D tb2x + interior loop + de-interleaved u1 (28)", 28, [384, 24, 24])):
        rows.append((name, run(lib, n, [1.0, 2.0, 3.0], T.MIXED_3D, variant, 24, 2)))
    print(f"{'kernel':50s} {'ld wf/upd':>10s} {'st wf/upd':>10s} {'total':>8s} {'ld instr/upd':>13s} {'st instr/upd':>13s}")
    for name, (lw, sw, li, si) in rows:
        print(f"{name:50s} {lw:10.4f} {sw:10.4f} {lw + sw:8.4f} {li:13.4f} {si:13.4f}")
    print("(warp-level; per cell update; multiply by 128 cells x levels for the per-row / per-plane figures of DESIGN.md)")


if __name__ == "__main__":
    main()
