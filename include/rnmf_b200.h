/*
 * rnmf_b200.h -- C ABI of librnmf_b200.so: the B200-native (sm_100a) stencil sweep, ghost-cell
 * boundary fill and residual reduction that replace the bodies of wattrg/rnmf's data-frame /
 * boundary-condition / solver functions on the magnetostatics Poisson hot path.
 *
 * The reference exposes no FFI: its boundary is its public Rust API.  Every entry point below
 * names the reference item (file:line, relative to the reference root) whose body it replaces;
 * INTEGRATION.md shows the `extern "C"` block and the safe wrappers a maintainer adds on the
 * Rust side.  Plain pointers and sizes only; no exceptions cross this boundary.
 *
 * Conventions
 *  - every function returns int32_t: 0 = ok, negative = error (RNMF_ERR_*); the message of the
 *    last error on the calling thread is rnmf_last_error().  The Rust shim panics on nonzero,
 *    which is the reference's behaviour on this path (it has no error returns).
 *  - "host dense" buffers use the reference's layout exactly (cartesian2d/mod.rs:31-37):
 *        p = (i+ng) + G0*((j+ng) + G1*((k+ng) + G2*c)),   G = n + 2*ng,  i fastest,
 *    (2D: no k term), length G0*G1*[G2*]n_comp doubles.
 *  - device frames are pitched: each grown row starts 128-byte aligned, valid cell i=0 sits at
 *    row_base + 16 doubles (128 B), pitch = roundup(16 + n0 + ng, 16) doubles.
 *  - a context is driven by one host thread at a time (the reference is single-threaded; its
 *    frames hold an Rc and are !Send).  Calls enqueue work on the context's CUDA stream and
 *    return; calls that return a scalar (norms) or copy to pageable host memory synchronise.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    RNMF_ERR_CUDA.
 *  - thread safety: none beyond "one host thread per context at a time"; the library keeps small
 *    process-wide caches (NCCL entry points, TMA descriptors) that are not locked.
 */
#ifndef RNMF_B200_H
#define RNMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNMF_ABI_VERSION 1

enum {
    RNMF_OK = 0,
    RNMF_ERR_INVALID = -1,        /* bad argument / shape mismatch                                  */
    RNMF_ERR_CUDA = -2,           /* CUDA runtime/driver error (incl. "no device")                  */
    RNMF_ERR_NCCL = -3,           /* NCCL error or NCCL not loadable                                */
    RNMF_ERR_UNIMPLEMENTED = -4   /* BcType::Reflect: unimplemented!() in dataframe.rs:375,410      */
};

/* boundary_conditions.rs:11-24 BcType<S>.  HALO is not in the reference: it marks the internal
 * face of a slab (cartesian2d/mod.rs:131 notes such faces need different treatment); fill_bc
 * skips it and rnmf_halo_exchange fills it. */
enum {
    RNMF_BC_DIRICHLET = 0,
    RNMF_BC_NEUMANN = 1,
    RNMF_BC_REFLECT = 2,
    RNMF_BC_PRESCRIBED = 3,
    RNMF_BC_HALO = 4
};

/* One face of ComponentBCs{lo,hi} (boundary_conditions.rs:41-54).
 * Array order: index = (comp*dim + d)*2 + side, side 0 = lo[d], 1 = hi[d]. */
typedef struct rnmf_bc_face {
    int32_t kind;                 /* RNMF_BC_*                                                     */
    int32_t _pad;
    double  value;                /* Dirichlet(value) / Neumann(gradient)                          */
    const double* prescribed;     /* Prescribed(Vec<S>): host pointer, copied by set_bc            */
    int64_t n_prescribed;
} rnmf_bc_face;

/* examples/magnetostatics/model.rs:11-20 UserModel fields used by the variable-mu operator
 * (poisson.rs:47-64). */
typedef struct rnmf_mu_params {
    double bubble_centre[2];
    double bubble_radius;
    double mu[2];                 /* mu[0] inside the bubble, mu[1] outside                        */
} rnmf_mu_params;

typedef struct rnmf_frame_info {
    int32_t dim, n_ghost, n_comp, distributed;
    int64_t n[3];                 /* local valid cells                                             */
    int64_t n_global[3];
    int64_t offset[3];            /* global index of local valid cell (0,0,0)                      */
    int64_t pitch;                /* doubles per grown row on the device                           */
    int64_t x_offset;             /* row_base + x_offset = grown cell ii=0 (= 16 - ng)             */
    int64_t n_rows;               /* G1*G2*n_comp                                                  */
    double  dx[3], lo[3];
} rnmf_frame_info;

typedef struct rnmf_ctx rnmf_ctx;       /* one CUDA device (+ optionally one NCCL rank)            */
typedef struct rnmf_frame rnmf_frame;   /* device-resident CartesianDataFrame (dataframe.rs:35-60) */

/* ---- library / context ---------------------------------------------------------------------- */
int32_t     rnmf_abi_version(void);
const char* rnmf_last_error(void);
const char* rnmf_build_info(void);               /* arch, flags, kernel variants compiled in      */

/* Single-GPU context on `device`. */
int32_t rnmf_ctx_create(int32_t device, rnmf_ctx** out);
/* One rank of an nranks-GPU job (one process per GPU).  nccl_id: 128 bytes from
 * rnmf_nccl_unique_id() on rank 0, broadcast by the caller (torch.distributed, a file, ...).
 * Slabs follow Domain::new_rectangular's split (cartesian2d/mod.rs:64-73): floor(n/P) per
 * block, remainder to the last block; rank r is block r along the slowest direction. */
int32_t rnmf_ctx_create_dist(int32_t device, int32_t rank, int32_t nranks, const void* nccl_id,
                             rnmf_ctx** out);
int32_t rnmf_nccl_unique_id(void* id128);
/* Destroy every frame of the context first (frames point into it); synchronises the device. */
int32_t rnmf_ctx_destroy(rnmf_ctx* ctx);
int32_t rnmf_sync(rnmf_ctx* ctx);
int32_t rnmf_ctx_stream(rnmf_ctx* ctx, void** cuda_stream);  /* cudaStream_t the kernels run on   */
int32_t rnmf_ctx_rank(rnmf_ctx* ctx, int32_t* rank, int32_t* nranks);
/* number of this library's kernels launched on the context since creation (graph replays count
 * their kernel nodes) -- bench.py's gpu_launches */
int64_t rnmf_kernel_launches(rnmf_ctx* ctx);
/* how many of those were double-sweep kernels (two Jacobi sweeps + ghost fills per pass over HBM, 2D and 3D: the pairs
 * that the groups of three (3D) / four (2D) sweeps leave over, and 3D frames the three-sweep kernel does not take) */
int64_t rnmf_double_sweep_launches(rnmf_ctx* ctx);
/* all temporally blocked launches (two or more sweeps per pass over HBM: the three-sweep 3D kernel, the four-sweep 2D
 * kernel and the double-sweep kernels) and the number of sweeps they covered, since the context was created; either
 * pointer may be null -- bench.py's roofline accounting */
int32_t rnmf_multi_sweep_stats(rnmf_ctx* ctx, int64_t* launches, int64_t* sweeps);
/* CUDA-event timer on the context's stream: begin records, end records+synchronises. */
int32_t rnmf_timer_begin(rnmf_ctx* ctx);
int32_t rnmf_timer_end(rnmf_ctx* ctx, double* elapsed_ms);
/* tuning knobs for measurements (results never change, only which kernel / schedule produces them):
 *   "sweep_variant"  "auto" (default: 3D three, 2D four sweeps per launch) or a number from the one table in csrc/kernels_sweep.cu:
 *                    3D 31 = two sweeps per launch (+ single sweeps), 6 = one sweep per launch only;
 *                    2D 30 = two sweeps per launch (+ single sweeps), 20 = one sweep per launch only
 *   "sweep_chunk"    planes / rows marched per CTA (0 = planner)
 *   "resident"       1/0: one cooperative launch for all sweeps of an L2-resident frame
 *   "peer_halo", "peer_inkernel", "halo_overlap"   slab halo exchange: fused peer-memory path / in-kernel handshake / NCCL overlap
 *   "e2e_skew" (+ "e2e_skew_chunks", "e2e_skew_pairs", "e2e_skew_tail")  the host-frame entry points (rnmf_jacobi_sweeps_host,
 *                    rnmf_solve_poisson_host): 2 (default) = upload in chunks of z-planes overlapped with the first double
 *                    sweeps (also across slabs) and, on a single GPU, the last launches band-wise so that the download overlaps
 *                    them; 1 = the overlapped upload only; 0 = plain upload / sweeps / download */
int32_t rnmf_set_option(rnmf_ctx* ctx, const char* key, const char* value);

/* Domain::new_rectangular split rule (cartesian2d/mod.rs:64-73,109-116), 1-D along one axis.
 * Pure host function (no device needed). */
int32_t rnmf_slab_partition(int64_t n, int32_t nranks, int32_t rank, int64_t* start, int64_t* count);

/* ---- frames ---------------------------------------------------------------------------------- */
/* CartesianDataFrame2D::new_from(&mesh, bc, n_comp, n_ghost) (dataframe.rs:71-163): data
 * zero-initialised (:155).  n/lo/dx: mesh.rs:36-45 (dx = (hi-lo)/n, computed by the caller).
 * dim = 2 or 3 (3D by the extrapolation rule). */
int32_t rnmf_frame_create(rnmf_ctx* ctx, int32_t dim, const int64_t* n, const double* lo,
                          const double* dx, int32_t n_ghost, int32_t n_comp, rnmf_frame** out);
/* Same, but n is the GLOBAL mesh and the frame holds this rank's slab along the slowest
 * direction (z in 3D, y in 2D); internal faces become RNMF_BC_HALO in set_bc. */
int32_t rnmf_frame_create_slab(rnmf_ctx* ctx, int32_t dim, const int64_t* n_global, const double* lo,
                               const double* dx, int32_t n_ghost, int32_t n_comp, rnmf_frame** out);
int32_t rnmf_frame_destroy(rnmf_frame* f);
int32_t rnmf_frame_info_get(const rnmf_frame* f, rnmf_frame_info* out);
/* debug aid: every device buffer ends in a canary band; *ok = 0 if a kernel stored past the end */
int32_t rnmf_frame_check_guards(rnmf_frame* f, int32_t* ok);
/* the `bc: BCs<S>` field (dataframe.rs:59): n_faces must be n_comp*dim*2 */
int32_t rnmf_frame_set_bc(rnmf_frame* f, const rnmf_bc_face* faces, int32_t n_faces);
/* host mirror <-> device (replaces direct access to `data: Vec<S>`, dataframe.rs:37).
 * The two uploads are ASYNCHRONOUS: they return once the copy is queued on the context's stream.  With pageable host
 * memory the runtime has staged the source by then; with pinned memory (rnmf_host_alloc) the caller must not modify or
 * free the buffer before rnmf_sync(ctx) (or any synchronising call: a download, a norm) has returned.  The downloads
 * synchronise.  The safe wrappers of the bindings (C++ frame.hpp, Rust DeviceFrame::upload, Python `data =`) sync. */
int32_t rnmf_frame_upload(rnmf_frame* f, const double* host_dense);
int32_t rnmf_frame_download(rnmf_frame* f, double* host_dense);
/* get_valid_cells (dataframe.rs:181-189): valid cells, flat order, all components */
int32_t rnmf_frame_download_valid(rnmf_frame* f, double* host_valid);
/* fill_ic result (dataframe.rs:166-173): valid cells in flat order; ghosts untouched */
int32_t rnmf_frame_upload_valid(rnmf_frame* f, const double* host_valid);
/* Index / IndexMut<(isize,isize,usize)> (dataframe.rs:208-233): one cell by its valid-cell index
 * (low ghosts negative; k ignored in 2D).  For occasional host access, not for loops. */
int32_t rnmf_frame_get_cell(rnmf_frame* f, int64_t i, int64_t j, int64_t k, int32_t comp, double* value);
int32_t rnmf_frame_set_cell(rnmf_frame* f, int64_t i, int64_t j, int64_t k, int32_t comp, double value);
/* The side iterators (dataframe.rs:709-817) / collect_typed_cells (:273-278): the cells of one side
 * of direction `dir` in the frame's flat order, corners included, as the reference's side lists do.
 *   which = 0: the n_ghost-wide strip of VALID cells adjacent to the face (Valid(North|NE|NW) etc.:
 *              the halo PACK order pinned by dataframe.rs:1148-1201);
 *   which = 1: the GHOST cells of that side (Ghost(West|NW|SW) etc., the order fill_prescribed_bc uses).
 * comp < 0 = IterComp::All.  n_out receives the number of values written (host_out may be NULL to query). */
int32_t rnmf_frame_collect_side(rnmf_frame* f, int32_t dir, int32_t side, int32_t which, int32_t comp,
                                double* host_out, int64_t* n_out);
int32_t rnmf_frame_fill_zero(rnmf_frame* f);
int32_t rnmf_frame_copy(rnmf_frame* dst, const rnmf_frame* src);            /* Clone, poisson.rs:71 */
/* synthetic data for benchmarks, generated on the device: valid cell with GLOBAL flat valid
 * index q gets splitmix64(seed ^ q) mapped to (-1,1); ghosts zero.   */
int32_t rnmf_frame_fill_synthetic(rnmf_frame* f, uint64_t seed);
int32_t rnmf_frame_device_ptr(rnmf_frame* f, void** dev_ptr);               /* current buffer      */

/* ---- hot path -------------------------------------------------------------------------------- */
/* BoundaryCondition::fill_bc(&mut self) (boundary_conditions.rs:4-7; impl dataframe.rs:361-433):
 * Dirichlet :324-355, Neumann :289-322, Prescribed :378-398/:413-428, Reflect -> UNIMPLEMENTED. */
int32_t rnmf_fill_bc(rnmf_frame* f);

/* alpha/beta/gamma of solve_poisson (poisson.rs:72-76): coef = {alpha, beta, gamma, 0} in 2D,
 * {a, a/dx2, a/dy2, a/dz2} in 3D.  Host function, same IEEE operations as the reference. */
int32_t rnmf_poisson_coefs(int32_t dim, const double* dx, double* coef4);

/* The sweep of solve_poisson (poisson.rs:80-89) as out-of-place Jacobi:
 *   n_sweeps x { psi' = coef0*rhs + coef1*(E+W) + coef2*(N+S) [+ coef3*(U+D)] on valid cells;
 *                psi'.fill_bc() }        (association exactly as poisson.rs:83-85)
 * psi is updated in place from the caller's view (internally ping-pong; groups of sweeps that do not carry
 * the norm -- three in 3D, four in 2D -- run as ONE temporally blocked launch with bit-identical results).  On a distributed
 * context the slab halos are exchanged every sweep, overlapped with the interior update.
 * l1_update (nullable): sum |psi'_last - psi_prev| over valid cells (all ranks), i.e. the
 * quantity main.rs:68-69 forms with two frame operators and `sum`; requesting it synchronises. */
int32_t rnmf_jacobi_sweeps(rnmf_frame* psi, const rnmf_frame* rhs, const double* coef4,
                           int64_t n_sweeps, double* l1_update);

/* Same loop in the reference's exact update order (lexicographic Gauss-Seidel, i outer / j inner,
 * in place, poisson.rs:80-89) evaluated on hyperplanes i+j+2*sweep: bit-identical to the serial
 * reference loop.  2D, n_ghost=1, n_comp=1, Dirichlet/Neumann faces, single GPU. */
int32_t rnmf_gs_sweeps(rnmf_frame* psi, const rnmf_frame* rhs, const double* coef4, int64_t n_sweeps);

/* get_laplace (poisson.rs:13-36): variable-mu 5-point operator + Dirichlet(0) ghost fill. 2D. */
int32_t rnmf_varcoef_laplace_2d(const rnmf_frame* phi, const rnmf_mu_params* mu, rnmf_frame* out);
/* get_source (poisson.rs:38-44): (-(1-1/relax)/mu) * get_laplace(phi), over all data. 2D. */
int32_t rnmf_poisson_source_2d(const rnmf_frame* phi, const rnmf_mu_params* mu, double relax,
                               rnmf_frame* out);

/* frame operators, element-wise over the ENTIRE data vector, ghosts included
 * (dataframe.rs:628-691).  out may alias an input. */
int32_t rnmf_scale(rnmf_frame* out, double a, const rnmf_frame* x);                   /* Real * &df  :645-657 */
int32_t rnmf_add(rnmf_frame* out, const rnmf_frame* x, const rnmf_frame* y);          /* &df + &df   :662-674 */
int32_t rnmf_mul(rnmf_frame* out, const rnmf_frame* x, const rnmf_frame* y);          /* &df * &df   :679-691 */
/* out = (a*x) + (b*y), each product and the sum rounded separately (no FMA): equals the
 * operator chains of main.rs:68 (a=1,b=-1) and main.rs:70 (a=1,b=relax) bit for bit. */
int32_t rnmf_axpby(rnmf_frame* out, double a, const rnmf_frame* x, double b, const rnmf_frame* y);

/* sum (poisson.rs:94-100): sum |v| over valid cells of this rank; deterministic tree order.
 * Synchronises. */
int32_t rnmf_abs_sum_valid(rnmf_frame* f, double* out);
/* sum over ranks of a host scalar (norms); identity on a single-GPU context. Synchronises. */
int32_t rnmf_allreduce_sum(rnmf_ctx* ctx, double* inout);
/* fill the RNMF_BC_HALO ghost planes from the neighbouring slabs (halo unpack =
 * fill_prescribed_bc analogue, dataframe.rs:280-284; pack = collect_typed_cells :273-278). */
int32_t rnmf_halo_exchange(rnmf_frame* f);
/* Fused compute + halo exchange over NVLink peer memory (one process per GPU, one NVSwitch box):
 * export writes 3 CUDA IPC handles (3 x 64 bytes: the frame's two ping-pong buffers and the
 * context's epoch flags); the caller ships them to the neighbouring ranks (rank-1 / rank+1) and
 * attaches theirs.  Afterwards rnmf_jacobi_sweeps stores each slab's boundary planes straight
 * into the neighbours' ghost planes (rows in 2D) from inside the sweep kernel (no NCCL call per sweep).
 * lo handles: from rank-1 (NULL on rank 0); hi handles: from rank+1 (NULL on the last rank). */
int32_t rnmf_frame_peer_export(rnmf_frame* f, void* handles192);
int32_t rnmf_frame_peer_attach(rnmf_frame* f, const void* lo_handles192, const void* hi_handles192);

/* get_h (model.rs:72-88): h = -grad(psi) by central differences into a 2-component frame. 2D. */
int32_t rnmf_gradient_h_2d(const rnmf_frame* psi, rnmf_frame* h2);

/* solve_poisson(psi_init, rhs, config) -> frame (poisson.rs:67-92) with HOST frames in the
 * reference's dense layout on both sides: upload psi_init and rhs, n_sub_iter sweeps (each
 * followed by fill_bc), download the result.  mode 0 = Jacobi, 1 = reference-order GS (2D only).
 * Device frames are cached in the context between calls of the same shape.  l1_update nullable. */
int32_t rnmf_solve_poisson_host(rnmf_ctx* ctx, int32_t dim, const int64_t* n, const double* lo,
                                const double* dx, int32_t n_ghost, const rnmf_bc_face* faces,
                                const double* psi_init_dense, const double* rhs_dense,
                                int64_t n_sub_iter, int32_t mode, double* psi_out_dense,
                                double* l1_update);
/* The same on EXISTING device frames, single-GPU or slab (solve_poisson, poisson.rs:67-92, with the caller's own frames; the
 * per-rank call of a job decomposed as Domain::new_rectangular does, cartesian2d/mod.rs:58-143): psi_in / rhs_in / psi_out are THIS rank's frames in the dense grown layout (n + 2*n_ghost cells per
 * direction, cartesian2d/mod.rs:31-37); psi and rhs receive the uploads, psi ends as the result.  Equivalent, bit for bit,
 * to rnmf_frame_upload x 2 + rnmf_jacobi_sweeps + rnmf_frame_download; with option "e2e_skew" the copies overlap the sweeps.
 * Collective over the ranks of a slab frame.  Synchronises.  l1_update nullable. */
int32_t rnmf_jacobi_sweeps_host(rnmf_frame* psi, rnmf_frame* rhs, const double* psi_in_dense, const double* rhs_in_dense,
                                const double* coef, int64_t n_sweeps, double* psi_out_dense, double* l1_update);
/* pinned host allocation for the host-buffer entry points (optional; pageable works, slower) */
int32_t rnmf_host_alloc(int64_t bytes, void** out);
int32_t rnmf_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RNMF_B200_H */
