//! Safe wrappers over librnmf_b200's C ABI, shaped like the reference's own types so that the
//! bodies of `CartesianDataFrame2D::{new_from, fill_ic, get_valid_cells}`, `BoundaryCondition::fill_bc`,
//! the frame operators and the example's `solve_poisson / get_source / sum` can delegate to them
//! (INTEGRATION.md sections 3-5 show the replaced bodies).
//!
//! Ownership: a `Context` owns one CUDA device (and optionally one NCCL rank); a `DeviceFrame` owns one
//! device-resident data frame and borrows its context, so frames cannot outlive it.  Neither is `Send`:
//! the ABI's contract is one host thread per context at a time, which is what the reference's
//! `Rc<CartesianMesh>` already imposes on its frames.
//! Errors: the reference panics on this path (it has no error returns); `ffi::check` does the same.
pub mod ffi;

use std::ffi::CString;
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::ptr;

pub type Real = f64; // lib.rs:30-31 of the reference without the `single_precision` feature

/// boundary_conditions.rs:11-24
#[derive(Clone, Debug)]
pub enum BcType {
    Dirichlet(Real),
    Neumann(Real),
    Reflect,
    Prescribed(Vec<Real>),
}

/// boundary_conditions.rs:41-54
#[derive(Clone, Debug)]
pub struct ComponentBCs {
    pub lo: Vec<BcType>,
    pub hi: Vec<BcType>,
}

/// boundary_conditions.rs:26-39
#[derive(Clone, Debug)]
pub struct BCs {
    pub bcs: Vec<ComponentBCs>,
}

impl BCs {
    /// face array in the ABI's order: index = (comp*dim + d)*2 + side
    fn faces(&self, dim: usize) -> Vec<ffi::rnmf_bc_face> {
        let mut v = Vec::with_capacity(self.bcs.len() * dim * 2);
        for c in &self.bcs {
            for d in 0..dim {
                for t in [&c.lo[d], &c.hi[d]].iter() {
                    let (kind, value, p, n) = match t {
                        BcType::Dirichlet(x) => (ffi::RNMF_BC_DIRICHLET, *x, ptr::null(), 0),
                        BcType::Neumann(g) => (ffi::RNMF_BC_NEUMANN, *g, ptr::null(), 0),
                        BcType::Reflect => (ffi::RNMF_BC_REFLECT, 0.0, ptr::null(), 0),
                        BcType::Prescribed(vals) => (ffi::RNMF_BC_PRESCRIBED, 0.0, vals.as_ptr(), vals.len() as i64),
                    };
                    v.push(ffi::rnmf_bc_face { kind, _pad: 0, value, prescribed: p, n_prescribed: n });
                }
            }
        }
        v
    }
}

/// One CUDA device, optionally one rank of a multi-GPU job.
pub struct Context {
    raw: *mut ffi::rnmf_ctx,
    _not_send: PhantomData<*mut ()>,
}

impl Context {
    pub fn new(device: i32) -> Context {
        let mut raw = ptr::null_mut();
        ffi::check(unsafe { ffi::rnmf_ctx_create(device, &mut raw) }, "rnmf_ctx_create");
        Context { raw, _not_send: PhantomData }
    }
    /// `nccl_id`: the 128 bytes `Context::nccl_unique_id()` returned on rank 0, shipped by the launcher.
    pub fn new_dist(device: i32, rank: i32, nranks: i32, nccl_id: &[u8; 128]) -> Context {
        let mut raw = ptr::null_mut();
        ffi::check(
            unsafe { ffi::rnmf_ctx_create_dist(device, rank, nranks, nccl_id.as_ptr() as *const c_void, &mut raw) },
            "rnmf_ctx_create_dist",
        );
        Context { raw, _not_send: PhantomData }
    }
    pub fn nccl_unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        ffi::check(unsafe { ffi::rnmf_nccl_unique_id(id.as_mut_ptr() as *mut c_void) }, "rnmf_nccl_unique_id");
        id
    }
    pub fn sync(&self) {
        ffi::check(unsafe { ffi::rnmf_sync(self.raw) }, "rnmf_sync");
    }
    pub fn rank(&self) -> (i32, i32) {
        let (mut r, mut n) = (0, 1);
        ffi::check(unsafe { ffi::rnmf_ctx_rank(self.raw, &mut r, &mut n) }, "rnmf_ctx_rank");
        (r, n)
    }
    pub fn set_option(&self, key: &str, value: &str) {
        let (k, v) = (CString::new(key).unwrap(), CString::new(value).unwrap());
        ffi::check(unsafe { ffi::rnmf_set_option(self.raw, k.as_ptr(), v.as_ptr()) }, "rnmf_set_option");
    }
    pub fn kernel_launches(&self) -> i64 {
        unsafe { ffi::rnmf_kernel_launches(self.raw) }
    }
    pub fn double_sweep_launches(&self) -> i64 {
        unsafe { ffi::rnmf_double_sweep_launches(self.raw) }
    }
    /// (launches, sweeps covered) of all temporally blocked launches
    pub fn multi_sweep_stats(&self) -> (i64, i64) {
        let (mut l, mut s) = (0i64, 0i64);
        ffi::check(unsafe { ffi::rnmf_multi_sweep_stats(self.raw, &mut l, &mut s) }, "rnmf_multi_sweep_stats");
        (l, s)
    }
    /// sum of a host scalar over all ranks (identity on a single-GPU context)
    pub fn allreduce_sum(&self, v: Real) -> Real {
        let mut x = v;
        ffi::check(unsafe { ffi::rnmf_allreduce_sum(self.raw, &mut x) }, "rnmf_allreduce_sum");
        x
    }
    pub fn raw(&self) -> *mut ffi::rnmf_ctx {
        self.raw
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::rnmf_ctx_destroy(self.raw) };
    }
}

/// Domain::new_rectangular's split along one axis (cartesian2d/mod.rs:64-73): (start, count) of block `rank`.
pub fn slab_partition(n: i64, nranks: i32, rank: i32) -> (i64, i64) {
    let (mut s, mut c) = (0i64, 0i64);
    ffi::check(unsafe { ffi::rnmf_slab_partition(n, nranks, rank, &mut s, &mut c) }, "rnmf_slab_partition");
    (s, c)
}

/// Device-resident `CartesianDataFrame` (dataframe.rs:35-60).  Host buffers crossing this API use
/// the reference's dense layout (cartesian2d/mod.rs:31-37).
pub struct DeviceFrame<'c> {
    raw: *mut ffi::rnmf_frame,
    ctx: &'c Context,
    pub dim: usize,
    pub n_ghost: usize,
    pub n_comp: usize,
    // creation arguments (global mesh for slabs), kept so that `like` builds an identical frame
    shape: Vec<i64>,
    lo: Vec<Real>,
    dx: Vec<Real>,
    slab: bool,
}

impl<'c> DeviceFrame<'c> {
    fn create(ctx: &'c Context, n: &[i64], lo: &[Real], dx: &[Real], n_ghost: usize, n_comp: usize, slab: bool) -> DeviceFrame<'c> {
        let dim = n.len();
        assert!(dim == 2 || dim == 3, "dim must be 2 or 3");
        assert!(lo.len() == dim && dx.len() == dim);
        let (mut n3, mut lo3, mut dx3) = ([1i64; 3], [0.0; 3], [1.0; 3]);
        for d in 0..dim {
            n3[d] = n[d];
            lo3[d] = lo[d];
            dx3[d] = dx[d];
        }
        let mut raw = ptr::null_mut();
        let rc = unsafe {
            if slab {
                ffi::rnmf_frame_create_slab(ctx.raw, dim as i32, n3.as_ptr(), lo3.as_ptr(), dx3.as_ptr(), n_ghost as i32, n_comp as i32, &mut raw)
            } else {
                ffi::rnmf_frame_create(ctx.raw, dim as i32, n3.as_ptr(), lo3.as_ptr(), dx3.as_ptr(), n_ghost as i32, n_comp as i32, &mut raw)
            }
        };
        ffi::check(rc, "rnmf_frame_create");
        DeviceFrame { raw, ctx, dim, n_ghost, n_comp, shape: n.to_vec(), lo: lo.to_vec(), dx: dx.to_vec(), slab }
    }
    /// new_from (dataframe.rs:71-163): zero-initialised
    pub fn new(ctx: &'c Context, n: &[i64], lo: &[Real], dx: &[Real], bc: &BCs, n_comp: usize, n_ghost: usize) -> DeviceFrame<'c> {
        let mut f = DeviceFrame::create(ctx, n, lo, dx, n_ghost, n_comp, false);
        f.set_bc(bc);
        f
    }
    /// this rank's slab (slowest direction) of a frame over the GLOBAL mesh `n_global`
    pub fn new_slab(ctx: &'c Context, n_global: &[i64], lo: &[Real], dx: &[Real], bc: &BCs, n_comp: usize, n_ghost: usize) -> DeviceFrame<'c> {
        let mut f = DeviceFrame::create(ctx, n_global, lo, dx, n_ghost, n_comp, true);
        f.set_bc(bc);
        f
    }
    /// an empty frame of the same shape (operators return new frames, dataframe.rs:628-691)
    pub fn like(&self, bc: &BCs) -> DeviceFrame<'c> {
        DeviceFrame::create(self.ctx, &self.shape, &self.lo, &self.dx, self.n_ghost, self.n_comp, self.slab).with_bc(bc)
    }
    fn with_bc(mut self, bc: &BCs) -> DeviceFrame<'c> {
        self.set_bc(bc);
        self
    }
    pub fn set_bc(&mut self, bc: &BCs) {
        if bc.bcs.is_empty() {
            return; // BCs::empty(): fill_bc on such a frame is an error, as in the reference
        }
        let faces = bc.faces(self.dim);
        ffi::check(unsafe { ffi::rnmf_frame_set_bc(self.raw, faces.as_ptr(), faces.len() as i32) }, "rnmf_frame_set_bc");
    }
    pub fn info(&self) -> ffi::rnmf_frame_info {
        let mut out: ffi::rnmf_frame_info = unsafe { std::mem::zeroed() };
        ffi::check(unsafe { ffi::rnmf_frame_info_get(self.raw, &mut out) }, "rnmf_frame_info_get");
        out
    }
    /// number of doubles of the host mirror `data` (grown cells x components)
    pub fn n_nodes(&self) -> usize {
        let i = self.info();
        let g = 2 * i.n_ghost as i64;
        let mut v = (i.n[0] + g) * (i.n[1] + g);
        if i.dim == 3 {
            v *= i.n[2] + g;
        }
        (v * i.n_comp as i64) as usize
    }
    pub fn n_valid(&self) -> usize {
        let i = self.info();
        let mut v = i.n[0] * i.n[1];
        if i.dim == 3 {
            v *= i.n[2];
        }
        (v * i.n_comp as i64) as usize
    }
    /// host mirror -> device (`data: Vec<S>`, dataframe.rs:37)
    pub fn upload(&mut self, data: &[Real]) {
        assert_eq!(data.len(), self.n_nodes());
        ffi::check(unsafe { ffi::rnmf_frame_upload(self.raw, data.as_ptr()) }, "rnmf_frame_upload");
        // rnmf_frame_upload is asynchronous (rnmf_b200.h): the borrow of `data` ends with this call, so wait for the copy
        self.ctx.sync();
    }
    pub fn download(&self, data: &mut [Real]) {
        assert_eq!(data.len(), self.n_nodes());
        ffi::check(unsafe { ffi::rnmf_frame_download(self.raw, data.as_mut_ptr()) }, "rnmf_frame_download");
    }
    /// get_valid_cells (dataframe.rs:181-189)
    pub fn get_valid_cells(&self) -> Vec<Real> {
        let mut v = vec![0.0; self.n_valid()];
        ffi::check(unsafe { ffi::rnmf_frame_download_valid(self.raw, v.as_mut_ptr()) }, "rnmf_frame_download_valid");
        v
    }
    /// the result of fill_ic (dataframe.rs:166-173): valid cells in flat order, ghosts untouched
    pub fn upload_valid(&mut self, valid: &[Real]) {
        assert_eq!(valid.len(), self.n_valid());
        ffi::check(unsafe { ffi::rnmf_frame_upload_valid(self.raw, valid.as_ptr()) }, "rnmf_frame_upload_valid");
        self.ctx.sync();
    }
    /// Index<(isize, isize, usize)> (dataframe.rs:208-219); k = 0 in 2D
    pub fn get(&self, i: isize, j: isize, k: isize, comp: usize) -> Real {
        let mut v = 0.0;
        ffi::check(unsafe { ffi::rnmf_frame_get_cell(self.raw, i as i64, j as i64, k as i64, comp as i32, &mut v) }, "rnmf_frame_get_cell");
        v
    }
    /// IndexMut (dataframe.rs:221-233)
    pub fn set(&mut self, i: isize, j: isize, k: isize, comp: usize, value: Real) {
        ffi::check(unsafe { ffi::rnmf_frame_set_cell(self.raw, i as i64, j as i64, k as i64, comp as i32, value) }, "rnmf_frame_set_cell");
    }
    /// side lists (dataframe.rs:709-817): which = 0 valid strip next to the face, 1 = its ghost cells; comp < 0 = all
    pub fn collect_side(&self, dir: usize, hi_side: bool, which: i32, comp: i32) -> Vec<Real> {
        let mut n = 0i64;
        ffi::check(unsafe { ffi::rnmf_frame_collect_side(self.raw, dir as i32, hi_side as i32, which, comp, ptr::null_mut(), &mut n) }, "rnmf_frame_collect_side");
        let mut v = vec![0.0; n as usize];
        ffi::check(unsafe { ffi::rnmf_frame_collect_side(self.raw, dir as i32, hi_side as i32, which, comp, v.as_mut_ptr(), &mut n) }, "rnmf_frame_collect_side");
        v
    }
    /// BoundaryCondition::fill_bc (boundary_conditions.rs:4-7, dataframe.rs:361-433); Reflect panics like unimplemented!()
    pub fn fill_bc(&mut self) {
        ffi::check(unsafe { ffi::rnmf_fill_bc(self.raw) }, "fill_bc");
    }
    /// Clone (poisson.rs:71)
    pub fn copy_from(&mut self, src: &DeviceFrame) {
        ffi::check(unsafe { ffi::rnmf_frame_copy(self.raw, src.raw) }, "rnmf_frame_copy");
    }
    /// `Real * &df` (dataframe.rs:645-657) into self
    pub fn assign_scale(&mut self, a: Real, x: &DeviceFrame) {
        ffi::check(unsafe { ffi::rnmf_scale(self.raw, a, x.raw) }, "rnmf_scale");
    }
    /// `&x + &y` (dataframe.rs:662-674) into self
    pub fn assign_add(&mut self, x: &DeviceFrame, y: &DeviceFrame) {
        ffi::check(unsafe { ffi::rnmf_add(self.raw, x.raw, y.raw) }, "rnmf_add");
    }
    /// `&x * &y` (dataframe.rs:679-691) into self
    pub fn assign_mul(&mut self, x: &DeviceFrame, y: &DeviceFrame) {
        ffi::check(unsafe { ffi::rnmf_mul(self.raw, x.raw, y.raw) }, "rnmf_mul");
    }
    /// `(a*x) + (b*y)`, the operator chains of main.rs:68 and :70 in one pass, bit for bit
    pub fn assign_axpby(&mut self, a: Real, x: &DeviceFrame, b: Real, y: &DeviceFrame) {
        ffi::check(unsafe { ffi::rnmf_axpby(self.raw, a, x.raw, b, y.raw) }, "rnmf_axpby");
    }
    /// sum (poisson.rs:94-100) over this rank's valid cells
    pub fn abs_sum(&self) -> Real {
        let mut v = 0.0;
        ffi::check(unsafe { ffi::rnmf_abs_sum_valid(self.raw, &mut v) }, "rnmf_abs_sum_valid");
        v
    }
    pub fn halo_exchange(&mut self) {
        ffi::check(unsafe { ffi::rnmf_halo_exchange(self.raw) }, "rnmf_halo_exchange");
    }
    /// fused halo over NVLink peer memory: 192 bytes to ship to the neighbouring ranks
    pub fn peer_export(&mut self) -> [u8; 192] {
        let mut h = [0u8; 192];
        ffi::check(unsafe { ffi::rnmf_frame_peer_export(self.raw, h.as_mut_ptr() as *mut c_void) }, "rnmf_frame_peer_export");
        h
    }
    pub fn peer_attach(&mut self, lo: Option<&[u8; 192]>, hi: Option<&[u8; 192]>) {
        let p = |o: Option<&[u8; 192]>| o.map_or(ptr::null(), |b| b.as_ptr() as *const c_void);
        ffi::check(unsafe { ffi::rnmf_frame_peer_attach(self.raw, p(lo), p(hi)) }, "rnmf_frame_peer_attach");
    }
    pub fn guards_intact(&self) -> bool {
        let mut ok = 0;
        ffi::check(unsafe { ffi::rnmf_frame_check_guards(self.raw, &mut ok) }, "rnmf_frame_check_guards");
        ok != 0
    }
    pub fn context(&self) -> &'c Context {
        self.ctx
    }
    pub fn raw(&self) -> *mut ffi::rnmf_frame {
        self.raw
    }
}

impl<'c> Drop for DeviceFrame<'c> {
    fn drop(&mut self) {
        unsafe { ffi::rnmf_frame_destroy(self.raw) };
    }
}

/// alpha/beta/gamma of solve_poisson (poisson.rs:72-76); 3D: {a, a/dx2, a/dy2, a/dz2}
pub fn poisson_coefs(dx: &[Real]) -> [Real; 4] {
    let dim = dx.len();
    let mut dx3 = [1.0; 3];
    for d in 0..dim {
        dx3[d] = dx[d];
    }
    let mut c = [0.0; 4];
    ffi::check(unsafe { ffi::rnmf_poisson_coefs(dim as i32, dx3.as_ptr(), c.as_mut_ptr()) }, "rnmf_poisson_coefs");
    c
}

/// The sweep loop of solve_poisson (poisson.rs:80-89) as out-of-place Jacobi, `fill_bc` after every
/// sweep; returns sum|psi_last - psi_prev| over all ranks when `want_norm` (main.rs:68-69).
pub fn jacobi_sweeps(psi: &mut DeviceFrame, rhs: &DeviceFrame, coef: &[Real; 4], n_sweeps: usize, want_norm: bool) -> Option<Real> {
    let mut l1 = 0.0;
    let p = if want_norm { &mut l1 as *mut Real } else { ptr::null_mut() };
    ffi::check(unsafe { ffi::rnmf_jacobi_sweeps(psi.raw, rhs.raw, coef.as_ptr(), n_sweeps as i64, p) }, "rnmf_jacobi_sweeps");
    if want_norm {
        Some(l1)
    } else {
        None
    }
}

/// solve_poisson with HOST frames on existing device frames (single-GPU or slab): this rank's dense grown frames in, n_sweeps
/// Jacobi sweeps (+ fill_bc each), this rank's dense frame out; the copies overlap the sweeps (option "e2e_skew").  Collective
/// over the ranks of a slab frame.  Returns the L1 norm of the last update.
pub fn jacobi_sweeps_host(psi: &mut DeviceFrame, rhs: &mut DeviceFrame, psi_in: &[Real], rhs_in: &[Real], coef: &[Real; 4], n_sweeps: usize,
                          psi_out: &mut [Real]) -> Real {
    assert!(psi_in.len() == psi.n_nodes() && rhs_in.len() == psi_in.len() && psi_out.len() == psi_in.len());
    let mut l1 = 0.0;
    ffi::check(
        unsafe { ffi::rnmf_jacobi_sweeps_host(psi.raw, rhs.raw, psi_in.as_ptr(), rhs_in.as_ptr(), coef.as_ptr(), n_sweeps as i64, psi_out.as_mut_ptr(), &mut l1) },
        "rnmf_jacobi_sweeps_host",
    );
    l1
}

/// Same loop in the reference's exact (lexicographic Gauss-Seidel) order: bit-identical to poisson.rs:80-89. 2D.
pub fn gs_sweeps(psi: &mut DeviceFrame, rhs: &DeviceFrame, coef: &[Real; 4], n_sweeps: usize) {
    ffi::check(unsafe { ffi::rnmf_gs_sweeps(psi.raw, rhs.raw, coef.as_ptr(), n_sweeps as i64) }, "rnmf_gs_sweeps");
}

/// model.rs:11-20 fields used by `mu` (poisson.rs:47-64)
pub type MuParams = ffi::rnmf_mu_params;

/// get_laplace (poisson.rs:13-36)
pub fn varcoef_laplace_2d(phi: &DeviceFrame, mu: &MuParams, out: &mut DeviceFrame) {
    ffi::check(unsafe { ffi::rnmf_varcoef_laplace_2d(phi.raw, mu, out.raw) }, "rnmf_varcoef_laplace_2d");
}
/// get_source (poisson.rs:38-44)
pub fn poisson_source_2d(phi: &DeviceFrame, mu: &MuParams, relax: Real, out: &mut DeviceFrame) {
    ffi::check(unsafe { ffi::rnmf_poisson_source_2d(phi.raw, mu, relax, out.raw) }, "rnmf_poisson_source_2d");
}
/// get_h (model.rs:72-88) into a 2-component frame
pub fn gradient_h_2d(psi: &DeviceFrame, h2: &mut DeviceFrame) {
    ffi::check(unsafe { ffi::rnmf_gradient_h_2d(psi.raw, h2.raw) }, "rnmf_gradient_h_2d");
}

/// solve_poisson with HOST frames on both sides (dense layout): upload, sweeps (+fill_bc each), download.
/// `gs_order`: the reference's exact update order (2D) instead of Jacobi.
#[allow(clippy::too_many_arguments)]
pub fn solve_poisson_host(ctx: &Context, n: &[i64], lo: &[Real], dx: &[Real], n_ghost: usize, bc: &BCs, psi_init: &[Real], rhs: &[Real],
                          n_sub_iter: usize, gs_order: bool, psi_out: &mut [Real]) -> Real {
    let dim = n.len();
    let (mut n3, mut lo3, mut dx3) = ([1i64; 3], [0.0; 3], [1.0; 3]);
    for d in 0..dim {
        n3[d] = n[d];
        lo3[d] = lo[d];
        dx3[d] = dx[d];
    }
    assert!(psi_init.len() == rhs.len() && rhs.len() == psi_out.len());
    let faces = bc.faces(dim);
    let mut l1 = 0.0;
    ffi::check(
        unsafe {
            ffi::rnmf_solve_poisson_host(ctx.raw, dim as i32, n3.as_ptr(), lo3.as_ptr(), dx3.as_ptr(), n_ghost as i32, faces.as_ptr(),
                                         psi_init.as_ptr(), rhs.as_ptr(), n_sub_iter as i64, gs_order as i32, psi_out.as_mut_ptr(), &mut l1)
        },
        "rnmf_solve_poisson_host",
    );
    l1
}
