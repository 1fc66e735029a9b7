#!/usr/bin/env python3
"""GPU box: rnmf_solve_poisson_host at 768^3 (550 sweeps + norm) over the knobs of the time-skewed host call
(e2e_skew_chunks / _pairs / _tail), one process, pinned buffers reused.  Prints GLUP/s end to end per setting."""
import ctypes as C
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rnmf_b200 as R
from rnmf_b200 import _capi
from rnmf_b200.dataframe import _faces_array
import bench

ctx = R.Context(0)
L = _capi.lib()
n = [768, 768, 768]
G = [v + 2 for v in n]
count = G[0] * G[1] * G[2]
bufs = []
for _ in range(3):
    p = C.c_void_p()
    _capi.check(L.rnmf_host_alloc(count * 8, C.byref(p)))
    bufs.append(p)
arrs = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(count,)) for p in bufs]
blk = np.random.default_rng(7).standard_normal(G[0] * G[1])
arrs[0].reshape(-1, G[0] * G[1])[:] = blk
arrs[1].reshape(-1, G[0] * G[1])[:] = blk[::-1]
faces = _faces_array(bench.make_bcs(R, bench.MIXED_3D), 3, 1)
nn = (C.c_int64 * 3)(*n)
lo = (C.c_double * 3)(0, 0, 0)
dx = (C.c_double * 3)(1, 1, 1)
l1 = C.c_double()
sub = 550


def call():
    _capi.check(L.rnmf_solve_poisson_host(ctx._h, 3, nn, lo, dx, 1, faces, bufs[0], bufs[1], sub, 0, bufs[2], C.byref(l1)))


def measure(reps=2):
    call()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    ctx.sync()
    return 768 ** 3 * sub * reps / (time.perf_counter() - t0) / 1e9


ctx.set_option("e2e_skew", 0)
print("plain", round(measure(), 1), flush=True)
ctx.set_option("e2e_skew", 2)
ref = None
# grid: RNMF_E2E_GRID="chunks;pairs;tail" (comma-separated values each), default = the round-2 grid around the defaults
grid = os.environ.get("RNMF_E2E_GRID", "4,6,8,12,16;40,56,72,96;24,32,48").split(";")
axes = [tuple(int(v) for v in a.split(",")) for a in grid]
for chunks, pairs, tail in itertools.chain([(8, 56, 32)], itertools.product(*axes)):
    ctx.set_option("e2e_skew_chunks", chunks)
    ctx.set_option("e2e_skew_pairs", pairs)
    ctx.set_option("e2e_skew_tail", tail)
    v = measure()
    print(f"chunks {chunks:2d} pairs {pairs:3d} tail {tail:2d}: {v:6.1f} GLUP/s  l1 {l1.value!r}", flush=True)
for p in bufs:
    L.rnmf_host_free(p)
