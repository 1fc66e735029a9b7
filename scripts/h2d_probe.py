"""Probe: host<->device copy rates, 1D (torch, pinned) vs the library's pitched 2D upload/download."""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rnmf_b200 as R
from rnmf_b200 import _capi
n = 768
ctx = R.Context(0)
mesh = R.CartesianMesh.new([0.0]*3, [float(n)]*3, [n]*3)
f = R.CartesianDataFrame.new_from(mesh, R.BCs.empty(), 1, 1, ctx)
count = (n+2)**3
L = _capi.lib()
p = C.c_void_p(); _capi.check(L.rnmf_host_alloc(count*8, C.byref(p)))
for rep in range(3):
    ctx.sync(); t=time.perf_counter(); _capi.check(L.rnmf_frame_upload(f._h, p)); ctx.sync(); up=time.perf_counter()-t
    t=time.perf_counter(); _capi.check(L.rnmf_frame_download(f._h, p)); ctx.sync(); dn=time.perf_counter()-t
    print(f"library pitched 2D: H2D {count*8/up/1e9:.1f} GB/s  D2H {count*8/dn/1e9:.1f} GB/s")
h = torch.empty(count, dtype=torch.float64).pin_memory()
d = torch.empty(count, dtype=torch.float64, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); up=time.perf_counter()-t
    t=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dn=time.perf_counter()-t
    print(f"torch 1D pinned:    H2D {count*8/up/1e9:.1f} GB/s  D2H {count*8/dn/1e9:.1f} GB/s")
