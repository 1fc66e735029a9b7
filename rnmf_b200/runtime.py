"""Device context: one CUDA device, optionally one rank of a multi-GPU (NCCL) job."""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

from . import _capi


class Context:
    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, nccl_id: Optional[bytes] = None):
        L = _capi.lib()
        h = C.c_void_p()
        if nranks > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("nccl_id (128 bytes from Context.unique_id() on rank 0) required")
            buf = C.create_string_buffer(nccl_id, 128)
            _capi.check(L.rnmf_ctx_create_dist(device, rank, nranks, C.cast(buf, C.c_void_p), C.byref(h)))
        else:
            _capi.check(L.rnmf_ctx_create(device, C.byref(h)))
        self._h = h
        self._frames = weakref.WeakSet()     # closed before the context (frames point into it)
        self.device, self.rank, self.nranks = device, rank, nranks

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _capi.check(_capi.lib().rnmf_nccl_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls) -> "Context":
        """One process per GPU under torchrun: the NCCL id travels through torch.distributed."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        local = int(os.environ.get("LOCAL_RANK", rank))
        if world == 1:
            return cls(local)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(local, rank, world, box[0])

    def sync(self) -> None:
        _capi.check(_capi.lib().rnmf_sync(self._h))

    def set_option(self, key: str, value) -> None:
        _capi.check(_capi.lib().rnmf_set_option(self._h, key.encode(), str(value).encode()))

    @property
    def kernel_launches(self) -> int:
        return int(_capi.lib().rnmf_kernel_launches(self._h))

    @property
    def double_sweep_launches(self) -> int:
        """launches of the 3D two-sweeps-per-pass kernel (kernels_sweep_tb2.cu)"""
        return int(_capi.lib().rnmf_double_sweep_launches(self._h))

    @property
    def multi_sweep_stats(self):
        """(launches, sweeps covered) of all temporally blocked launches: double-sweep kernels and the K-sweep 2D kernel"""
        a, b = C.c_int64(), C.c_int64()
        _capi.check(_capi.lib().rnmf_multi_sweep_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        _capi.check(_capi.lib().rnmf_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    def timer_begin(self) -> None:
        _capi.check(_capi.lib().rnmf_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_double()
        _capi.check(_capi.lib().rnmf_timer_end(self._h, C.byref(ms)))
        return ms.value

    def allreduce_sum(self, x: float) -> float:
        v = C.c_double(x)
        _capi.check(_capi.lib().rnmf_allreduce_sum(self._h, C.byref(v)))
        return v.value

    def close(self) -> None:
        if self._h:
            for f in list(self._frames):
                f.close()
            _capi.lib().rnmf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: Optional[Context] = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default


def slab_partition(n: int, nranks: int, rank: int):
    """Domain::new_rectangular's split (cartesian2d/mod.rs:64-73): (start, count) of block `rank`."""
    s, c = C.c_int64(), C.c_int64()
    _capi.check(_capi.lib().rnmf_slab_partition(n, nranks, rank, C.byref(s), C.byref(c)))
    return s.value, c.value
