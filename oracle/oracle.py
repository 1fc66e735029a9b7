"""ctypes loader for the CPU oracle (oracle/rnmf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the rnmf_b200 package.  Arrays use the
reference's dense layout (cartesian2d/mod.rs:31-37): C-order numpy shape
(ncomp, [G2,] G1, G0) flattened, x fastest.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

DIRICHLET, NEUMANN, REFLECT, PRESCRIBED, HALO = 0, 1, 2, 3, 4
GHOST = 16
LOC = dict(Regular=0, North=1, South=2, East=3, West=4, NorthEast=5, NorthWest=6, SouthWest=7, SouthEast=8)


class Geom(C.Structure):
    _fields_ = [("dim", C.c_int32), ("ng", C.c_int32), ("ncomp", C.c_int32), ("_pad", C.c_int32),
                ("n", C.c_int64 * 3), ("lo", C.c_double * 3), ("dx", C.c_double * 3)]

    @property
    def grown(self):
        return tuple(self.n[d] + 2 * self.ng if d < self.dim else 1 for d in range(3))

    @property
    def n_nodes(self):
        g = self.grown
        return g[0] * g[1] * g[2] * self.ncomp

    @property
    def shape(self):
        g = self.grown
        return (self.ncomp, g[1], g[0]) if self.dim == 2 else (self.ncomp, g[2], g[1], g[0])


class BcFace(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("value", C.c_double),
                ("prescribed", C.POINTER(C.c_double)), ("n_prescribed", C.c_int64)]


class Model(C.Structure):
    _fields_ = [("h_far", C.c_double * 2), ("relax", C.c_double), ("n_iter", C.c_int64),
                ("n_sub_iter", C.c_int64), ("bubble_centre", C.c_double * 2),
                ("bubble_radius", C.c_double), ("mu", C.c_double * 2), ("tol", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rnmf_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    P = C.POINTER
    sig = {
        "ora_make_geom": (None, [P(Geom), C.c_int, P(C.c_int64), P(C.c_double), P(C.c_double), C.c_int, C.c_int]),
        "ora_flat_index": (C.c_int64, [P(Geom), C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
        "ora_get_ij_from_p": (C.c_int, [C.c_int64] * 5 + [P(C.c_int64)] * 3),
        "ora_classify_2d": (None, [P(Geom), np.ctypeslib.ndpointer(dtype=np.uint8)]),
        "ora_valid_cells": (C.c_int64, [P(Geom), _dp, _dp]),
        "ora_cell_centres_2d": (None, [P(Geom), _dp]),
        "ora_cell_centre": (C.c_double, [P(Geom), C.c_int, C.c_int64]),
        "ora_ghost_iter_2d": (C.c_int64, [P(Geom), _dp, np.ctypeslib.ndpointer(dtype=np.uint8), C.c_int,
                                          C.c_int64, _dp, np.ctypeslib.ndpointer(dtype=np.int64)]),
        "ora_fill_bc": (C.c_int, [P(Geom), _dp, P(BcFace)]),
        "ora_scale": (None, [C.c_int64, C.c_double, _dp, _dp]),
        "ora_add": (None, [C.c_int64, _dp, _dp, _dp]),
        "ora_mul": (None, [C.c_int64, _dp, _dp, _dp]),
        "ora_mu": (C.c_double, [C.c_int64, C.c_int64, P(C.c_double), P(Model)]),
        "ora_get_laplace_2d": (None, [P(Geom), _dp, P(Model), _dp]),
        "ora_get_source_2d": (None, [P(Geom), _dp, P(Model), _dp]),
        "ora_poisson_coefs": (None, [C.c_int, P(C.c_double), P(C.c_double)]),
        "ora_solve_poisson_gs": (C.c_int, [P(Geom), _dp, _dp, P(BcFace), C.c_int64]),
        "ora_solve_poisson_gs_wavefront_2d": (C.c_int, [P(Geom), _dp, _dp, P(BcFace), C.c_int64]),
        "ora_jacobi_sweeps": (C.c_int, [P(Geom), _dp, _dp, _dp, P(BcFace), C.c_int64, P(C.c_double)]),
        "ora_jacobi_sweeps_omp": (C.c_int, [P(Geom), _dp, _dp, _dp, P(BcFace), C.c_int64, P(C.c_double)]),
        "ora_sum_abs_valid": (C.c_double, [P(Geom), _dp]),
        "ora_sum_abs_valid_ld": (C.c_double, [P(Geom), _dp]),
        "ora_fill_ic_magnetostatics": (None, [P(Geom), _dp, P(Model)]),
        "ora_magnetostatics_run": (C.c_int, [P(Geom), _dp, P(Model), C.c_int, _dp, P(C.c_int64)]),
        "ora_get_h_2d": (None, [P(Geom), _dp, _dp]),
        "ora_num_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


# ---- convenience layer ------------------------------------------------------------------

def geom(n, lo=None, hi=None, ng=1, ncomp=1) -> Geom:
    """CartesianMesh2D::new(lo, hi, n) + new_from(.., ncomp, ng)  (mesh.rs:36, dataframe.rs:71)."""
    dim = len(n)
    lo = [0.0] * dim if lo is None else list(lo)
    hi = [float(v) for v in n] if hi is None else list(hi)
    g = Geom()
    lib().ora_make_geom(C.byref(g), dim, (C.c_int64 * dim)(*n), (C.c_double * dim)(*lo),
                        (C.c_double * dim)(*hi), ng, ncomp)
    return g


_KINDS = dict(dirichlet=DIRICHLET, neumann=NEUMANN, reflect=REFLECT, prescribed=PRESCRIBED, halo=HALO)


def bcs(dim, comps):
    """comps: per component (lo, hi), each a list over directions of (kind, value) with kind in
    dirichlet|neumann|reflect|prescribed|halo; prescribed's value is a sequence of floats.
    Mirrors BCs::new(vec![ComponentBCs::new(lo, hi)])  (boundary_conditions.rs:41-70)."""
    faces = (BcFace * (len(comps) * dim * 2))()
    keep = []
    for c, (lo, hi) in enumerate(comps):
        for d in range(dim):
            for side, spec in enumerate((lo[d], hi[d])):
                f = faces[(c * dim + d) * 2 + side]
                kind, val = spec
                f.kind = _KINDS[kind]
                if kind == "prescribed":
                    arr = np.ascontiguousarray(val, dtype=np.float64)
                    keep.append(arr)
                    f.prescribed = arr.ctypes.data_as(C.POINTER(C.c_double))
                    f.n_prescribed = arr.size
                    f.value = 0.0
                else:
                    f.value = 0.0 if val is None else float(val)
    faces._keep = keep
    return faces


def zeros(g: Geom) -> np.ndarray:
    return np.zeros(g.n_nodes, dtype=np.float64)


def as_grid(g: Geom, data: np.ndarray) -> np.ndarray:
    return data.reshape(g.shape)


def valid_view(g: Geom, data: np.ndarray) -> np.ndarray:
    a = as_grid(g, data)
    ng = g.ng
    if g.dim == 2:
        return a[:, ng:ng + g.n[1], ng:ng + g.n[0]]
    return a[:, ng:ng + g.n[2], ng:ng + g.n[1], ng:ng + g.n[0]]


def fill_ic(g: Geom, data: np.ndarray, fn) -> None:
    """fill_ic(|x,y[,z],n| ..)  (dataframe.rs:166-173): valid cells from cell-centre positions."""
    L = lib()
    xs = [np.array([L.ora_cell_centre(C.byref(g), d, i) for i in range(g.n[d])]) for d in range(g.dim)]
    v = valid_view(g, data)
    for c in range(g.ncomp):
        if g.dim == 2:
            X, Y = np.meshgrid(xs[0], xs[1], indexing="xy")
            v[c] = fn(X, Y, c)
        else:
            Z, Y, X = np.meshgrid(xs[2], xs[1], xs[0], indexing="ij")
            v[c] = fn(X, Y, Z, c)


def fill_bc(g, data, faces) -> int:
    return lib().ora_fill_bc(C.byref(g), data, faces)


def coefs(g: Geom) -> np.ndarray:
    out = (C.c_double * 4)()
    lib().ora_poisson_coefs(g.dim, g.dx, out)
    return np.array(list(out))


def jacobi(g, psi, rhs, faces, n_sweeps, omp=False, want_norm=True):
    """returns l1 norm of the last update (long-double accumulation; None if not wanted); psi updated in place."""
    scratch = np.empty_like(psi)
    l1 = C.c_double(0.0)
    fn = lib().ora_jacobi_sweeps_omp if omp else lib().ora_jacobi_sweeps
    rc = fn(C.byref(g), psi, scratch, rhs, faces, n_sweeps, C.byref(l1) if want_norm else None)
    if rc:
        raise RuntimeError(f"oracle jacobi rc={rc}")
    return l1.value if want_norm else None


def gs(g, psi, rhs, faces, n_sub_iter, wavefront=False):
    fn = lib().ora_solve_poisson_gs_wavefront_2d if wavefront else lib().ora_solve_poisson_gs
    rc = fn(C.byref(g), psi, rhs, faces, n_sub_iter)
    if rc:
        raise RuntimeError(f"oracle gs rc={rc}")


def model(h_far=(0.0, 1.0), relax=0.2, n_iter=10, n_sub_iter=550, bubble_centre=(150.5, 150.5),
          bubble_radius=15.05, mu=(3.2, 1.0), tol=0.1) -> Model:
    m = Model()
    m.h_far[:] = h_far
    m.relax, m.n_iter, m.n_sub_iter = relax, n_iter, n_sub_iter
    m.bubble_centre[:] = bubble_centre
    m.bubble_radius = bubble_radius
    m.mu[:] = mu
    m.tol = tol
    return m


def splitmix_field(n_valid: int, seed: int, offset: int = 0) -> np.ndarray:
    """Deterministic synthetic field: u[q] = splitmix64(seed ^ q) mapped to (-1, 1),
    q = global flat VALID index.  The CUDA library generates the same numbers on device."""
    q = (np.arange(offset, offset + n_valid, dtype=np.uint64)) ^ np.uint64(seed)
    with np.errstate(over="ignore"):
        z = q + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0
