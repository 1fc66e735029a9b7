"""Mirror of examples/magnetostatics/{poisson,model,main}.rs: the Poisson pieces of the
magnetostatics example, same names and argument meaning; bodies call the C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import _capi
from .boundary_conditions import BCs, BcType, ComponentBCs
from .dataframe import CartesianDataFrame
from .mesh import CartesianMesh
from .runtime import Context


@dataclass
class UserModel:
    """examples/magnetostatics/model.rs:11-35 (defaults of UserConfig::new)."""
    h_far: Tuple[float, float] = (0.0, 0.0)
    relax: float = 0.0
    n_iter: int = 0
    n_sub_iter: int = 0
    bubble_centre: Tuple[float, float] = (0.0, 0.0)
    bubble_radius: float = 0.0
    mu: Tuple[float, float] = (0.0, 0.0)
    tol: float = 0.4


@dataclass
class GeomConfig:
    """config.rs:34-39"""
    dim: int = 0
    length: Tuple[float, ...] = (0.0, 0.0)
    n_cells: Tuple[int, ...] = (0, 0)


@dataclass
class Config:
    """config.rs:15-31: Config<T>{geom, model, residual_iters}"""
    geom: GeomConfig = field(default_factory=GeomConfig)
    model: UserModel = field(default_factory=UserModel)
    residual_iters: int = 1


def _mu_params(config: Config) -> _capi.MuParams:
    p = _capi.MuParams()
    p.bubble_centre[:] = config.model.bubble_centre
    p.bubble_radius = config.model.bubble_radius
    p.mu[:] = config.model.mu
    return p


def poisson_coefs(dim: int, dx) -> List[float]:
    """alpha, beta, gamma of poisson.rs:72-76 (3D: a, a/dx2, a/dy2, a/dz2)."""
    out = (C.c_double * 4)()
    _capi.check(_capi.lib().rnmf_poisson_coefs(dim, (C.c_double * 3)(*dx, *([1.0] * (3 - dim))), out))
    return list(out)


def get_laplace(phi: CartesianDataFrame, config: Config) -> CartesianDataFrame:
    """poisson.rs:13-36"""
    z = BcType.Dirichlet(0.0)
    out = phi._like(BCs([ComponentBCs([z, z], [z, z])]))
    p = _mu_params(config)
    _capi.check(_capi.lib().rnmf_varcoef_laplace_2d(phi._h, C.byref(p), out._h))
    return out


def get_source(phi: CartesianDataFrame, config: Config, out: Optional[CartesianDataFrame] = None) -> CartesianDataFrame:
    """poisson.rs:38-44 (one fused kernel; pass `out` to reuse a frame)."""
    out = out or phi._like()
    p = _mu_params(config)
    _capi.check(_capi.lib().rnmf_poisson_source_2d(phi._h, C.byref(p), float(config.model.relax), out._h))
    return out


def solve_poisson(psi_init: CartesianDataFrame, rhs: CartesianDataFrame, config: Config, order: str = "jacobi",
                  out: Optional[CartesianDataFrame] = None) -> CartesianDataFrame:
    """poisson.rs:67-92: clone psi_init, n_sub_iter x { sweep; fill_bc }, return the new frame.
    order="jacobi": the throughput path (same expression, out-of-place);
    order="reference": the reference's lexicographic Gauss-Seidel order, bit-identical (2D)."""
    if out is None:
        psi = psi_init.clone()
    else:
        psi = out
        _capi.check(_capi.lib().rnmf_frame_copy(psi._h, psi_init._h))
    m = psi.underlying_mesh
    coef = (C.c_double * 4)(*poisson_coefs(m.dim, m.dx))
    n = int(config.model.n_sub_iter)
    L = _capi.lib()
    if order == "jacobi":
        _capi.check(L.rnmf_jacobi_sweeps(psi._h, rhs._h, coef, n, None))
    elif order == "reference":
        _capi.check(L.rnmf_gs_sweeps(psi._h, rhs._h, coef, n))
    else:
        raise ValueError("order must be 'jacobi' or 'reference'")
    return psi


def jacobi_sweeps(psi: CartesianDataFrame, rhs: CartesianDataFrame, n_sweeps: int, want_norm: bool = False,
                  coef=None) -> Optional[float]:
    """n_sweeps in-place Jacobi sweeps (+ fill_bc each); optionally the L1 norm of the last update."""
    m = psi.underlying_mesh
    coef = (C.c_double * 4)(*(coef if coef is not None else poisson_coefs(m.dim, m.dx)))
    l1 = C.c_double()
    _capi.check(_capi.lib().rnmf_jacobi_sweeps(psi._h, rhs._h, coef, int(n_sweeps), C.byref(l1) if want_norm else None))
    return l1.value if want_norm else None


def sum(psi: CartesianDataFrame) -> float:  # noqa: A001 - the reference's name (poisson.rs:94)
    """poisson.rs:94-100.  A slab (distributed) frame is summed over all ranks; a replicated frame on a multi-rank
    context (the reference-order GS path is "replicas only") holds the whole domain on every rank, so its sum is local:
    all-reducing it would multiply sum_diff by the number of ranks and change the `sum_diff > tol` test of main.rs:65."""
    local = psi.abs_sum()
    return psi.ctx.allreduce_sum(local) if psi.info.distributed else local


def magnetostatics(config: Config, order: str = "reference", ctx: Optional[Context] = None):
    """The body of examples/magnetostatics/main.rs:29-78 on the device.
    Returns (psi, sum_diff trace, iterations)."""
    mesh = CartesianMesh.new((0.0, 0.0), tuple(config.geom.length), tuple(config.geom.n_cells))
    mdl = config.model
    gx, gy = -mdl.h_far[0] / mesh.dx[0], -mdl.h_far[1] / mesh.dx[1]
    bc = BCs([ComponentBCs([BcType.Neumann(gx), BcType.Neumann(gy)], [BcType.Neumann(gx), BcType.Neumann(gy)])])
    psi = CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    psi.fill_ic(lambda x, y, _: -mdl.h_far[0] / mesh.dx[0] * x - mdl.h_far[1] / mesh.dx[1] * y)
    psi.fill_bc()
    source, psi_star, diff = psi._like(), psi._like(), psi._like()
    sum_diff, it, trace = 1.0, 0, []
    while it < mdl.n_iter and sum_diff > mdl.tol:                       # main.rs:65
        get_source(psi, config, out=source)                             # :66
        solve_poisson(psi, source, config, order, out=psi_star)         # :67
        diff.axpby_(1.0, psi_star, -1.0, psi)                           # :68  &psi_star + &(-1.0 * &psi)
        sum_diff = sum(diff)                                            # :69
        psi.axpby_(1.0, psi, mdl.relax, diff)                           # :70  &psi + &(relax * diff)
        psi.fill_bc()                                                   # :71
        trace.append(sum_diff)
        it += 1
    return psi, trace, it
