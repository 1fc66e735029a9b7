//! Raw binding of include/rnmf_b200.h: one declaration per exported symbol, same order as the header.
//! `tests/test_rust_binding_sync.py` checks names, argument counts and struct fields against the header.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct rnmf_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct rnmf_frame {
    _p: [u8; 0],
}

pub const RNMF_ABI_VERSION: i32 = 1;

pub const RNMF_OK: i32 = 0;
pub const RNMF_ERR_INVALID: i32 = -1;
pub const RNMF_ERR_CUDA: i32 = -2;
pub const RNMF_ERR_NCCL: i32 = -3;
pub const RNMF_ERR_UNIMPLEMENTED: i32 = -4;

pub const RNMF_BC_DIRICHLET: i32 = 0;
pub const RNMF_BC_NEUMANN: i32 = 1;
pub const RNMF_BC_REFLECT: i32 = 2;
pub const RNMF_BC_PRESCRIBED: i32 = 3;
pub const RNMF_BC_HALO: i32 = 4;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rnmf_bc_face {
    pub kind: i32,
    pub _pad: i32,
    pub value: f64,
    pub prescribed: *const f64,
    pub n_prescribed: i64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rnmf_mu_params {
    pub bubble_centre: [f64; 2],
    pub bubble_radius: f64,
    pub mu: [f64; 2],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rnmf_frame_info {
    pub dim: i32,
    pub n_ghost: i32,
    pub n_comp: i32,
    pub distributed: i32,
    pub n: [i64; 3],
    pub n_global: [i64; 3],
    pub offset: [i64; 3],
    pub pitch: i64,
    pub x_offset: i64,
    pub n_rows: i64,
    pub dx: [f64; 3],
    pub lo: [f64; 3],
}

extern "C" {
    // ---- library / context
    pub fn rnmf_abi_version() -> i32;
    pub fn rnmf_last_error() -> *const c_char;
    pub fn rnmf_build_info() -> *const c_char;
    pub fn rnmf_ctx_create(device: i32, out: *mut *mut rnmf_ctx) -> i32;
    pub fn rnmf_ctx_create_dist(device: i32, rank: i32, nranks: i32, nccl_id: *const c_void, out: *mut *mut rnmf_ctx) -> i32;
    pub fn rnmf_nccl_unique_id(id128: *mut c_void) -> i32;
    pub fn rnmf_ctx_destroy(ctx: *mut rnmf_ctx) -> i32;
    pub fn rnmf_sync(ctx: *mut rnmf_ctx) -> i32;
    pub fn rnmf_ctx_stream(ctx: *mut rnmf_ctx, cuda_stream: *mut *mut c_void) -> i32;
    pub fn rnmf_ctx_rank(ctx: *mut rnmf_ctx, rank: *mut i32, nranks: *mut i32) -> i32;
    pub fn rnmf_kernel_launches(ctx: *mut rnmf_ctx) -> i64;
    pub fn rnmf_double_sweep_launches(ctx: *mut rnmf_ctx) -> i64;
    pub fn rnmf_multi_sweep_stats(ctx: *mut rnmf_ctx, launches: *mut i64, sweeps: *mut i64) -> i32;
    pub fn rnmf_timer_begin(ctx: *mut rnmf_ctx) -> i32;
    pub fn rnmf_timer_end(ctx: *mut rnmf_ctx, elapsed_ms: *mut f64) -> i32;
    pub fn rnmf_set_option(ctx: *mut rnmf_ctx, key: *const c_char, value: *const c_char) -> i32;
    pub fn rnmf_slab_partition(n: i64, nranks: i32, rank: i32, start: *mut i64, count: *mut i64) -> i32;
    // ---- frames
    pub fn rnmf_frame_create(ctx: *mut rnmf_ctx, dim: i32, n: *const i64, lo: *const f64, dx: *const f64, n_ghost: i32, n_comp: i32, out: *mut *mut rnmf_frame) -> i32;
    pub fn rnmf_frame_create_slab(ctx: *mut rnmf_ctx, dim: i32, n_global: *const i64, lo: *const f64, dx: *const f64, n_ghost: i32, n_comp: i32, out: *mut *mut rnmf_frame) -> i32;
    pub fn rnmf_frame_destroy(f: *mut rnmf_frame) -> i32;
    pub fn rnmf_frame_info_get(f: *const rnmf_frame, out: *mut rnmf_frame_info) -> i32;
    pub fn rnmf_frame_check_guards(f: *mut rnmf_frame, ok: *mut i32) -> i32;
    pub fn rnmf_frame_set_bc(f: *mut rnmf_frame, faces: *const rnmf_bc_face, n_faces: i32) -> i32;
    pub fn rnmf_frame_upload(f: *mut rnmf_frame, host_dense: *const f64) -> i32;
    pub fn rnmf_frame_download(f: *mut rnmf_frame, host_dense: *mut f64) -> i32;
    pub fn rnmf_frame_download_valid(f: *mut rnmf_frame, host_valid: *mut f64) -> i32;
    pub fn rnmf_frame_upload_valid(f: *mut rnmf_frame, host_valid: *const f64) -> i32;
    pub fn rnmf_frame_get_cell(f: *mut rnmf_frame, i: i64, j: i64, k: i64, comp: i32, value: *mut f64) -> i32;
    pub fn rnmf_frame_set_cell(f: *mut rnmf_frame, i: i64, j: i64, k: i64, comp: i32, value: f64) -> i32;
    pub fn rnmf_frame_collect_side(f: *mut rnmf_frame, dir: i32, side: i32, which: i32, comp: i32, host_out: *mut f64, n_out: *mut i64) -> i32;
    pub fn rnmf_frame_fill_zero(f: *mut rnmf_frame) -> i32;
    pub fn rnmf_frame_copy(dst: *mut rnmf_frame, src: *const rnmf_frame) -> i32;
    pub fn rnmf_frame_fill_synthetic(f: *mut rnmf_frame, seed: u64) -> i32;
    pub fn rnmf_frame_device_ptr(f: *mut rnmf_frame, dev_ptr: *mut *mut c_void) -> i32;
    // ---- hot path
    pub fn rnmf_fill_bc(f: *mut rnmf_frame) -> i32;
    pub fn rnmf_poisson_coefs(dim: i32, dx: *const f64, coef4: *mut f64) -> i32;
    pub fn rnmf_jacobi_sweeps(psi: *mut rnmf_frame, rhs: *const rnmf_frame, coef4: *const f64, n_sweeps: i64, l1_update: *mut f64) -> i32;
    pub fn rnmf_gs_sweeps(psi: *mut rnmf_frame, rhs: *const rnmf_frame, coef4: *const f64, n_sweeps: i64) -> i32;
    pub fn rnmf_varcoef_laplace_2d(phi: *const rnmf_frame, mu: *const rnmf_mu_params, out: *mut rnmf_frame) -> i32;
    pub fn rnmf_poisson_source_2d(phi: *const rnmf_frame, mu: *const rnmf_mu_params, relax: f64, out: *mut rnmf_frame) -> i32;
    pub fn rnmf_scale(out: *mut rnmf_frame, a: f64, x: *const rnmf_frame) -> i32;
    pub fn rnmf_add(out: *mut rnmf_frame, x: *const rnmf_frame, y: *const rnmf_frame) -> i32;
    pub fn rnmf_mul(out: *mut rnmf_frame, x: *const rnmf_frame, y: *const rnmf_frame) -> i32;
    pub fn rnmf_axpby(out: *mut rnmf_frame, a: f64, x: *const rnmf_frame, b: f64, y: *const rnmf_frame) -> i32;
    pub fn rnmf_abs_sum_valid(f: *mut rnmf_frame, out: *mut f64) -> i32;
    pub fn rnmf_allreduce_sum(ctx: *mut rnmf_ctx, inout: *mut f64) -> i32;
    pub fn rnmf_halo_exchange(f: *mut rnmf_frame) -> i32;
    pub fn rnmf_frame_peer_export(f: *mut rnmf_frame, handles192: *mut c_void) -> i32;
    pub fn rnmf_frame_peer_attach(f: *mut rnmf_frame, lo_handles192: *const c_void, hi_handles192: *const c_void) -> i32;
    pub fn rnmf_gradient_h_2d(psi: *const rnmf_frame, h2: *mut rnmf_frame) -> i32;
    pub fn rnmf_jacobi_sweeps_host(psi: *mut rnmf_frame, rhs: *mut rnmf_frame, psi_in_dense: *const f64, rhs_in_dense: *const f64, coef: *const f64, n_sweeps: i64, psi_out_dense: *mut f64, l1_update: *mut f64) -> i32;
    pub fn rnmf_solve_poisson_host(ctx: *mut rnmf_ctx, dim: i32, n: *const i64, lo: *const f64, dx: *const f64, n_ghost: i32, faces: *const rnmf_bc_face, psi_init_dense: *const f64, rhs_dense: *const f64, n_sub_iter: i64, mode: i32, psi_out_dense: *mut f64, l1_update: *mut f64) -> i32;
    pub fn rnmf_host_alloc(bytes: i64, out: *mut *mut c_void) -> i32;
    pub fn rnmf_host_free(p: *mut c_void) -> i32;
}

/// The reference has no error returns on this path: it panics. So does the shim.
pub fn check(rc: i32, what: &str) {
    if rc != RNMF_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(rnmf_last_error()) }.to_string_lossy().into_owned();
        panic!("{} failed ({}): {}", what, rc, msg);
    }
}
