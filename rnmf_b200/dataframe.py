"""Mirror of src/mesh/cartesian2d/dataframe.rs: a ghost-celled Cartesian data frame whose storage
lives in B200 HBM.  Same names and argument meaning as the reference; bodies call the C ABI.

Host access (`data`, `get_valid_cells`, `frame[(i, j, n)]`) goes through an explicit device->host
copy: per-cell indexing is kept off the hot path (it is where the reference spends its time).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _capi
from .boundary_conditions import BCs, BcType
from .mesh import CartesianMesh
from .runtime import Context, default_context


def _faces_array(bc: BCs, dim: int, n_comp: int):
    """BCs -> rnmf_bc_face[(comp*dim + d)*2 + side]"""
    if len(bc.bcs) != n_comp:
        raise ValueError(f"BCs has {len(bc.bcs)} components, frame has {n_comp}")
    arr = (_capi.BcFace * (n_comp * dim * 2))()
    keep = []
    for c, comp in enumerate(bc.bcs):
        if len(comp.lo) != dim or len(comp.hi) != dim:
            raise ValueError("ComponentBCs must give one BcType per direction for lo and hi")
        for d in range(dim):
            for side, t in enumerate((comp.lo[d], comp.hi[d])):
                f = arr[(c * dim + d) * 2 + side]
                f.kind, f.value = t.kind, t.value
                if t.kind == _capi.BC_PRESCRIBED:
                    vals = np.ascontiguousarray(t.values, dtype=np.float64)
                    keep.append(vals)
                    f.prescribed = vals.ctypes.data_as(C.POINTER(C.c_double))
                    f.n_prescribed = vals.size
    arr._keep = keep
    return arr


class CartesianDataFrame:
    """CartesianDataFrame2D<Real> (dataframe.rs:35-60); 3D by the extrapolation rule."""

    def __init__(self, handle, mesh: CartesianMesh, bc: BCs, n_comp: int, n_ghost: int, ctx: Context):
        self._h = handle
        self.underlying_mesh = mesh
        self.bc = bc
        self.n_comp = n_comp
        self.n_ghost = n_ghost
        self.ctx = ctx
        ctx._frames.add(self)
        info = _capi.FrameInfo()
        _capi.check(_capi.lib().rnmf_frame_info_get(self._h, C.byref(info)))
        self.info = info
        self.n_local = tuple(info.n[d] for d in range(mesh.dim))
        self.n_grown = tuple(n + 2 * n_ghost for n in self.n_local)
        self.n_nodes = int(np.prod(self.n_grown)) * n_comp

    # ---- construction ----------------------------------------------------------------------
    @classmethod
    def new_from(cls, m: CartesianMesh, bc: BCs, n_comp: int, n_ghost: int,
                 ctx: Optional[Context] = None, slab: bool = False) -> "CartesianDataFrame":
        """new_from(&mesh, bc, n_comp, n_ghost) (dataframe.rs:71-163); data zero-initialised.
        slab=True: the mesh is the GLOBAL mesh and this rank holds its slab (slowest direction)."""
        ctx = ctx or default_context()
        L = _capi.lib()
        h = C.c_void_p()
        n = (C.c_int64 * 3)(*m.n, *([1] * (3 - m.dim)))
        lo = (C.c_double * 3)(*m.lo, *([0.0] * (3 - m.dim)))
        dx = (C.c_double * 3)(*m.dx, *([1.0] * (3 - m.dim)))
        create = L.rnmf_frame_create_slab if slab else L.rnmf_frame_create
        _capi.check(create(ctx._h, m.dim, n, lo, dx, n_ghost, n_comp, C.byref(h)))
        f = cls(h, m, bc, n_comp, n_ghost, ctx)
        if bc.bcs:
            f.set_bc(bc)
        return f

    def set_bc(self, bc: BCs) -> None:
        self.bc = bc
        arr = _faces_array(bc, self.underlying_mesh.dim, self.n_comp)
        _capi.check(_capi.lib().rnmf_frame_set_bc(self._h, arr, len(arr)))

    def clone(self) -> "CartesianDataFrame":
        """Clone (poisson.rs:71)."""
        out = self._like()
        _capi.check(_capi.lib().rnmf_frame_copy(out._h, self._h))
        return out

    def _like(self, bc: Optional[BCs] = None) -> "CartesianDataFrame":
        f = CartesianDataFrame.new_from(self.underlying_mesh, bc or self.bc, self.n_comp, self.n_ghost, self.ctx,
                                        slab=bool(self.info.distributed))
        return f

    # ---- initial / boundary conditions -----------------------------------------------------
    def fill_ic(self, ic: Callable) -> None:
        """fill_ic(|x,y[,z],n| ..) (dataframe.rs:166-173): valid cells from cell centres
        (vectorised: the closure receives numpy arrays), ghosts untouched."""
        m = self.underlying_mesh
        off = [self.info.offset[d] for d in range(m.dim)]
        xs = [m.centres(d)[off[d]:off[d] + self.n_local[d]] for d in range(m.dim)]
        vals = np.empty((self.n_comp,) + tuple(reversed(self.n_local)), dtype=np.float64)
        for c in range(self.n_comp):
            if m.dim == 2:
                X, Y = np.meshgrid(xs[0], xs[1], indexing="xy")
                vals[c] = ic(X, Y, c)
            else:
                Z, Y, X = np.meshgrid(xs[2], xs[1], xs[0], indexing="ij")
                vals[c] = ic(X, Y, Z, c)
        _capi.check(_capi.lib().rnmf_frame_upload_valid(self._h, vals.ctypes.data_as(C.c_void_p)))
        self.ctx.sync()

    def fill_bc(self) -> None:
        """BoundaryCondition::fill_bc (boundary_conditions.rs:4-7, dataframe.rs:361-433)."""
        _capi.check(_capi.lib().rnmf_fill_bc(self._h))

    def halo_exchange(self) -> None:
        _capi.check(_capi.lib().rnmf_halo_exchange(self._h))

    def enable_peer_halo(self) -> None:
        """Fused compute + halo exchange over NVLink peer memory: swap CUDA IPC handles of the
        frame's buffers with the neighbouring ranks (through torch.distributed) and attach them.
        Collective over all ranks; slab frames (z-slabs in 3D, y-slabs in 2D)."""
        import torch.distributed as dist
        L = _capi.lib()
        mine = C.create_string_buffer(192)
        _capi.check(L.rnmf_frame_peer_export(self._h, C.cast(mine, C.c_void_p)))
        world, rank = dist.get_world_size(), dist.get_rank()
        everyone = [None] * world
        dist.all_gather_object(everyone, mine.raw)
        lo = C.create_string_buffer(everyone[rank - 1], 192) if rank > 0 else None
        hi = C.create_string_buffer(everyone[rank + 1], 192) if rank < world - 1 else None
        _capi.check(L.rnmf_frame_peer_attach(self._h, C.cast(lo, C.c_void_p) if lo else None,
                                             C.cast(hi, C.c_void_p) if hi else None))
        dist.barrier()

    def guards_intact(self) -> bool:
        """debug aid: False if some kernel stored past the end of one of the frame's buffers."""
        ok = C.c_int32()
        _capi.check(_capi.lib().rnmf_frame_check_guards(self._h, C.byref(ok)))
        return bool(ok.value)

    def fill_synthetic(self, seed: int) -> None:
        _capi.check(_capi.lib().rnmf_frame_fill_synthetic(self._h, seed))

    # ---- host access -----------------------------------------------------------------------
    @property
    def data(self) -> np.ndarray:
        """The reference's `data: Vec<S>` (dense grown layout, dataframe.rs:37), downloaded."""
        out = np.empty(self.n_nodes, dtype=np.float64)
        _capi.check(_capi.lib().rnmf_frame_download(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    @data.setter
    def data(self, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host, dtype=np.float64).reshape(-1)
        if host.size != self.n_nodes:
            raise ValueError(f"expected {self.n_nodes} values, got {host.size}")
        _capi.check(_capi.lib().rnmf_frame_upload(self._h, host.ctypes.data_as(C.c_void_p)))
        self.ctx.sync()

    def get_valid_cells(self) -> np.ndarray:
        """get_valid_cells (dataframe.rs:181-189)."""
        out = np.empty(int(np.prod(self.n_local)) * self.n_comp, dtype=np.float64)
        _capi.check(_capi.lib().rnmf_frame_download_valid(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def __getitem__(self, idx):
        """Index<(isize,isize,usize)> (dataframe.rs:221-233); low ghosts are negative."""
        *ijk, n = idx
        ijk = list(ijk) + [0] * (3 - len(ijk))
        v = C.c_double()
        _capi.check(_capi.lib().rnmf_frame_get_cell(self._h, ijk[0], ijk[1], ijk[2], n, C.byref(v)))
        return v.value

    def __setitem__(self, idx, value) -> None:
        """IndexMut<(isize,isize,usize)> (dataframe.rs:208-218)."""
        *ijk, n = idx
        ijk = list(ijk) + [0] * (3 - len(ijk))
        _capi.check(_capi.lib().rnmf_frame_set_cell(self._h, ijk[0], ijk[1], ijk[2], n, float(value)))

    def collect_side(self, direction: int, side: int, ghost: bool = False, comp: int = -1) -> np.ndarray:
        """collect_typed_cells / the side iterators (dataframe.rs:273-278, 709-817): the valid strip
        next to a face (ghost=False: Valid(North|NE|NW) ...) or the ghost cells of that side
        (ghost=True: Ghost(West|NW|SW) ...), flat order, corners included."""
        L = _capi.lib()
        n = C.c_int64()
        _capi.check(L.rnmf_frame_collect_side(self._h, direction, side, 1 if ghost else 0, comp, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        if n.value:
            _capi.check(L.rnmf_frame_collect_side(self._h, direction, side, 1 if ghost else 0, comp,
                                                  out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out

    # ---- whole-frame operators (dataframe.rs:628-691), ghosts included ---------------------
    def __rmul__(self, a: float) -> "CartesianDataFrame":            # Real * &df
        out = self._like()
        _capi.check(_capi.lib().rnmf_scale(out._h, float(a), self._h))
        return out

    def __add__(self, other: "CartesianDataFrame") -> "CartesianDataFrame":   # &df + &df (rhs's BCs)
        out = self._like(other.bc)
        _capi.check(_capi.lib().rnmf_add(out._h, self._h, other._h))
        return out

    def __mul__(self, other) -> "CartesianDataFrame":                # &df * &df (rhs's BCs)
        if not isinstance(other, CartesianDataFrame):
            return NotImplemented
        out = self._like(other.bc)
        _capi.check(_capi.lib().rnmf_mul(out._h, self._h, other._h))
        return out

    def axpby_(self, a: float, x: "CartesianDataFrame", b: float, y: "CartesianDataFrame") -> "CartesianDataFrame":
        """self = (a*x) + (b*y) in place, no allocation (fused form of the operator chains)."""
        _capi.check(_capi.lib().rnmf_axpby(self._h, float(a), x._h, float(b), y._h))
        return self

    def abs_sum(self) -> float:
        """sum (poisson.rs:94-100): sum |v| over this rank's valid cells."""
        v = C.c_double()
        _capi.check(_capi.lib().rnmf_abs_sum_valid(self._h, C.byref(v)))
        return v.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            _capi.lib().rnmf_frame_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


CartesianDataFrame2D = CartesianDataFrame
CartesianDataFrame3D = CartesianDataFrame
