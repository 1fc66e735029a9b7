"""ctypes binding of librnmf_b200.so (include/rnmf_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
The reference panics on every error of this path; RnmfError is the Python analogue.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librnmf_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_UNIMPLEMENTED = 0, -1, -2, -3, -4
BC_DIRICHLET, BC_NEUMANN, BC_REFLECT, BC_PRESCRIBED, BC_HALO = 0, 1, 2, 3, 4


class RnmfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rnmf_b200 error {code}: {msg}")
        self.code = code


class BcFace(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("value", C.c_double),
                ("prescribed", C.POINTER(C.c_double)), ("n_prescribed", C.c_int64)]


class MuParams(C.Structure):
    _fields_ = [("bubble_centre", C.c_double * 2), ("bubble_radius", C.c_double), ("mu", C.c_double * 2)]


class FrameInfo(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_ghost", C.c_int32), ("n_comp", C.c_int32), ("distributed", C.c_int32),
                ("n", C.c_int64 * 3), ("n_global", C.c_int64 * 3), ("offset", C.c_int64 * 3),
                ("pitch", C.c_int64), ("x_offset", C.c_int64), ("n_rows", C.c_int64),
                ("dx", C.c_double * 3), ("lo", C.c_double * 3)]


_P = C.POINTER
_vp = C.c_void_p
_dp = _P(C.c_double)
_i64p = _P(C.c_int64)

# name -> (restype, argtypes); every symbol include/rnmf_b200.h declares
SIGNATURES = {
    "rnmf_abi_version": (C.c_int32, []),
    "rnmf_last_error": (C.c_char_p, []),
    "rnmf_build_info": (C.c_char_p, []),
    "rnmf_ctx_create": (C.c_int32, [C.c_int32, _P(_vp)]),
    "rnmf_ctx_create_dist": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _vp, _P(_vp)]),
    "rnmf_nccl_unique_id": (C.c_int32, [_vp]),
    "rnmf_ctx_destroy": (C.c_int32, [_vp]),
    "rnmf_sync": (C.c_int32, [_vp]),
    "rnmf_ctx_stream": (C.c_int32, [_vp, _P(_vp)]),
    "rnmf_ctx_rank": (C.c_int32, [_vp, _P(C.c_int32), _P(C.c_int32)]),
    "rnmf_kernel_launches": (C.c_int64, [_vp]),
    "rnmf_double_sweep_launches": (C.c_int64, [_vp]),
    "rnmf_multi_sweep_stats": (C.c_int32, [_vp, _i64p, _i64p]),
    "rnmf_timer_begin": (C.c_int32, [_vp]),
    "rnmf_timer_end": (C.c_int32, [_vp, _dp]),
    "rnmf_set_option": (C.c_int32, [_vp, C.c_char_p, C.c_char_p]),
    "rnmf_slab_partition": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32, _i64p, _i64p]),
    "rnmf_frame_create": (C.c_int32, [_vp, C.c_int32, _i64p, _dp, _dp, C.c_int32, C.c_int32, _P(_vp)]),
    "rnmf_frame_create_slab": (C.c_int32, [_vp, C.c_int32, _i64p, _dp, _dp, C.c_int32, C.c_int32, _P(_vp)]),
    "rnmf_frame_destroy": (C.c_int32, [_vp]),
    "rnmf_frame_info_get": (C.c_int32, [_vp, _P(FrameInfo)]),
    "rnmf_frame_check_guards": (C.c_int32, [_vp, _P(C.c_int32)]),
    "rnmf_frame_set_bc": (C.c_int32, [_vp, _P(BcFace), C.c_int32]),
    "rnmf_frame_upload": (C.c_int32, [_vp, _vp]),
    "rnmf_frame_download": (C.c_int32, [_vp, _vp]),
    "rnmf_frame_download_valid": (C.c_int32, [_vp, _vp]),
    "rnmf_frame_upload_valid": (C.c_int32, [_vp, _vp]),
    "rnmf_frame_get_cell": (C.c_int32, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _dp]),
    "rnmf_frame_set_cell": (C.c_int32, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_double]),
    "rnmf_frame_collect_side": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _i64p]),
    "rnmf_frame_fill_zero": (C.c_int32, [_vp]),
    "rnmf_frame_copy": (C.c_int32, [_vp, _vp]),
    "rnmf_frame_fill_synthetic": (C.c_int32, [_vp, C.c_uint64]),
    "rnmf_frame_device_ptr": (C.c_int32, [_vp, _P(_vp)]),
    "rnmf_fill_bc": (C.c_int32, [_vp]),
    "rnmf_poisson_coefs": (C.c_int32, [C.c_int32, _dp, _dp]),
    "rnmf_jacobi_sweeps": (C.c_int32, [_vp, _vp, _dp, C.c_int64, _dp]),
    "rnmf_gs_sweeps": (C.c_int32, [_vp, _vp, _dp, C.c_int64]),
    "rnmf_varcoef_laplace_2d": (C.c_int32, [_vp, _P(MuParams), _vp]),
    "rnmf_poisson_source_2d": (C.c_int32, [_vp, _P(MuParams), C.c_double, _vp]),
    "rnmf_scale": (C.c_int32, [_vp, C.c_double, _vp]),
    "rnmf_add": (C.c_int32, [_vp, _vp, _vp]),
    "rnmf_mul": (C.c_int32, [_vp, _vp, _vp]),
    "rnmf_axpby": (C.c_int32, [_vp, C.c_double, _vp, C.c_double, _vp]),
    "rnmf_abs_sum_valid": (C.c_int32, [_vp, _dp]),
    "rnmf_allreduce_sum": (C.c_int32, [_vp, _dp]),
    "rnmf_halo_exchange": (C.c_int32, [_vp]),
    "rnmf_frame_peer_export": (C.c_int32, [_vp, _vp]),
    "rnmf_frame_peer_attach": (C.c_int32, [_vp, _vp, _vp]),
    "rnmf_gradient_h_2d": (C.c_int32, [_vp, _vp]),
    "rnmf_solve_poisson_host": (C.c_int32, [_vp, C.c_int32, _i64p, _dp, _dp, C.c_int32, _P(BcFace), _vp, _vp,
                                            C.c_int64, C.c_int32, _vp, _dp]),
    "rnmf_jacobi_sweeps_host": (C.c_int32, [_vp, _vp, _vp, _vp, _dp, C.c_int64, _vp, _dp]),
    "rnmf_host_alloc": (C.c_int32, [C.c_int64, _P(_vp)]),
    "rnmf_host_free": (C.c_int32, [_vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load librnmf_b200.so.  Raises ImportError (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C rnmf_b200/csrc`.  rnmf_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if L.rnmf_abi_version() != 1:
        raise ImportError("librnmf_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise RnmfError(rc, lib().rnmf_last_error().decode(errors="replace"))
