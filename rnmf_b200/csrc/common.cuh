// common.cuh -- internal types shared by the kernels and the C ABI of librnmf_b200.
// Device layout of a data frame (replaces `data: Vec<S>` of dataframe.rs:35-60):
//   element (ii,jj,kk,c) in GROWN indices lives at
//       base + xoff + ii + pitch*(jj + G1*(kk + G2*c)),      xoff = 16 - ng
//   so valid cell i=0 (ii=ng) is at row_base + 16 doubles = 128 B aligned; pitch % 16 == 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/rnmf_b200.h"

#define RNMF_ROW_ALIGN 16  // doubles (128 bytes)

#ifndef RNMF_EMULATE
#include <atomic>
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE opt-in and rnmf_ctx_create(device) allows several
// devices (and host threads) in one process: remember the opt-in per device, atomically (one flag per kernel instantiation).
struct PerDeviceOnce {
    std::atomic<unsigned long long> bits[2];
    PerDeviceOnce() { bits[0] = 0; bits[1] = 0; }
    static int dev() { int d = 0; cudaGetDevice(&d); return d; }
    bool done() const { const int d = dev(); return d < 128 && ((bits[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1ULL); }
    void mark() { const int d = dev(); if (d < 128) bits[d >> 6].fetch_or(1ULL << (d & 63), std::memory_order_release); }
};
#endif

struct FrameGeom {
    int dim, ng, ncomp;
    long long n[3];      // valid cells
    long long G[3];      // grown cells (G[2]=1 in 2D)
    long long pitch;     // doubles per grown row
    long long xoff;      // 16 - ng
    long long plane;     // pitch*G[1]
    long long comp;      // plane*G[2]
    long long total;     // comp*ncomp (doubles allocated)
    double dx[3], lo[3];
    long long nglob[3], off[3];
    int kg;              // ghost depth in z (ng in 3D, 0 in 2D)

    // pointer offset of valid-indexed cell (i,j,k,c)  (low ghosts negative, dataframe.rs:209-211)
    __host__ __device__ inline long long at(long long i, long long j, long long k, long long c) const
    {
        return xoff + (i + ng) + pitch * ((j + ng) + G[1] * ((k + kg) + G[2] * c));
    }
};

// One face of the fused / stand-alone Dirichlet-Neumann fill, precomputed on the host with the
// reference's own operations so the kernel only does one add/sub per ghost cell:
//   mode 1 (Dirichlet):  ghost = s - mirror,        s = 2.0*value                    (dataframe.rs:328,344)
//   mode 2 (Neumann lo): ghost = mirror - s_g,      s_g = (grad*(2.0*g-1.0))*dx      (:293-294)
//   mode 3 (Neumann hi): ghost = mirror + s_g,      s_g = (grad*(2.0*(g+1)-1.0))*dx  (:311-312)
//   mode 0: skip (HALO / Prescribed handled elsewhere)
// For ng>1 the Neumann scalar depends on the ghost layer; the kernels recompute it from
// grad and dx with the same association.
struct FaceBC {
    int mode;
    double value;   // Dirichlet value or Neumann gradient
    double s1;      // precomputed scalar for ghost layer 1 (the ng=1 fast path)
};

struct FaceSet {      // faces of ONE component: [d*2+side]
    FaceBC f[6];
};

__host__ __device__ inline double rnmf_ghost_value(const FaceBC& b, int side, double mirror, int layer /*0-based*/, double dxd)
{
    // layer = ghost_indx of the reference loops for Dirichlet/Neumann-high, ghost_indx-1 for Neumann-low
    if (b.mode == 1) {
#ifdef __CUDA_ARCH__
        return __dsub_rn(__dmul_rn(2.0, b.value), mirror);
#else
        return 2.0 * b.value - mirror;
#endif
    }
    double fac = 2.0 * (double)(layer + 1) - 1.0;
#ifdef __CUDA_ARCH__
    double s = __dmul_rn(__dmul_rn(b.value, fac), dxd);
    return side == 0 ? __dsub_rn(mirror, s) : __dadd_rn(mirror, s);
#else
    double s = b.value * fac * dxd;
    return side == 0 ? mirror - s : mirror + s;
#endif
}

// geometry of a frame with n valid cells per direction (capi.cu: rnmf_frame_create*)
inline void make_geom(FrameGeom& g, int dim, const int64_t* n, const double* lo, const double* dx, int ng, int ncomp)
{
    memset(&g, 0, sizeof(g));
    g.dim = dim; g.ng = ng; g.ncomp = ncomp;
    g.kg = dim == 3 ? ng : 0;
    for (int d = 0; d < 3; ++d) {
        g.n[d] = d < dim ? n[d] : 1;
        g.G[d] = d < dim ? n[d] + 2 * ng : 1;
        g.dx[d] = d < dim ? dx[d] : 1.0;
        g.lo[d] = d < dim ? lo[d] : 0.0;
        g.nglob[d] = g.n[d];
        g.off[d] = 0;
    }
    g.xoff = RNMF_ROW_ALIGN - ng;
    g.pitch = ((RNMF_ROW_ALIGN + g.n[0] + ng + RNMF_ROW_ALIGN - 1) / RNMF_ROW_ALIGN) * RNMF_ROW_ALIGN;
    g.plane = g.pitch * g.G[1];
    g.comp = g.plane * g.G[2];
    g.total = g.comp * ncomp;
}

// FaceBC for one raw face (HALO / Prescribed / Reflect -> mode 0 here)
inline FaceBC make_face(const rnmf_bc_face& f, int side, double dxd)
{
    FaceBC b;
    b.mode = 0; b.value = f.value; b.s1 = 0.0;
    if (f.kind == RNMF_BC_DIRICHLET) {
        b.mode = 1;
        b.s1 = 2.0 * f.value;                               // dataframe.rs:328,344
    } else if (f.kind == RNMF_BC_NEUMANN) {
        b.mode = side == 0 ? 2 : 3;
        b.s1 = f.value * (2.0 * 1.0 - 1.0) * dxd;           // :293-294 (g=1), :311-312 (g=0)
    }
    return b;
}

// error plumbing (capi.cu)
void rnmf_set_error(const char* fmt, ...);
#define RNMF_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            rnmf_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return RNMF_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)
#define RNMF_REQUIRE(cond, ...)                                                          \
    do {                                                                                 \
        if (!(cond)) { rnmf_set_error(__VA_ARGS__); return RNMF_ERR_INVALID; }          \
    } while (0)

struct SweepCoef { double c0, c1, c2, c3; };

// ---- kernel launchers (each returns the number of kernels it launched, or <0 on error) -----
int launch_fill_bc_faces(const FrameGeom& g, double* d, const FaceSet* faces_host, cudaStream_t s);
int launch_fill_prescribed(const FrameGeom& g, double* d, int comp, int dir, int side,
                           const double* vals_dev, long long nvals, cudaStream_t s);
long long collect_side_count(const FrameGeom& g, int dir, int which, int ncomp);
int launch_collect_side(const FrameGeom& g, const double* d, int dir, int side, int which, int comp0, int ncomp,
                        double* out_dev, cudaStream_t s);
int launch_copy_ghost_shell(const FrameGeom& g, const double* src, double* dst, cudaStream_t s);

// In-kernel epoch handshake of the fused peer-memory halo (kernels_sweep_tma.cu, kernels_sweep2d_tma.cu):
// only the CTAs whose chunk holds the slab's first / last slice take part.  They are scheduled LAST,
// wait (thread 0, bounded spin) until the neighbour has finished the boundary CTAs of its previous
// sweep (flag >= epoch: its halo stores into my input ghosts have landed and it no longer reads the
// ghost slice I am about to overwrite), and the last of them to finish publishes epoch+1 to the
// neighbour.  Dependency chain: my sweep e+1 <- neighbour's sweep e <- my sweep e-1 (complete):
// no cycle, and the waiter only depends on another GPU's progress.
// The double-sweep kernel (kernels_sweep_tb2.cu) uses the very same protocol, one epoch per launch: with deep halo
// planes a launch needs nothing from the neighbour that is produced during the launch.
struct PeerSync {
    const unsigned long long* wait_lo;   // my flag written by the lo neighbour (null: no lo neighbour / not used)
    const unsigned long long* wait_hi;
    unsigned long long* sig_lo;          // lo neighbour's flag for its hi side (peer-mapped)
    unsigned long long* sig_hi;
    unsigned long long* done;            // [2] local completion counters of the boundary CTAs (lo, hi)
    unsigned long long target_lo;        // counter value that marks the last boundary CTA of this launch
    unsigned long long target_hi;
    unsigned long long epoch;            // sweeps completed before this launch
    unsigned long long* status;          // set on spin timeout
    int enabled;
};

#ifdef __CUDACC__
__device__ __forceinline__ void peer_spin_wait(const unsigned long long* flag, unsigned long long target, unsigned long long* status)
{
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= target) break;
        if (clock64() - t0 > 20LL * 1000 * 1000 * 1000) { *status = target; break; }   // ~10 s: a dead neighbour
        __nanosleep(100);
    }
}
// called by ONE thread after a CTA-wide barrier that follows the CTA's last peer store
__device__ __forceinline__ void peer_publish(unsigned long long* done, unsigned long long target, unsigned long long* peer_flag,
                                             unsigned long long value)
{
    __threadfence_system();
    const unsigned long long old = atomicAdd(done, 1ULL);
    if (old + 1 == target) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"(value) : "memory");
    }
}
// boundary chunks last: chunk index -> 1, 2, .., n-2, 0, n-1
__device__ __forceinline__ unsigned peer_chunk_order(unsigned b, unsigned n)
{
    return n < 3 ? b : (b < n - 2 ? b + 1 : (b == n - 2 ? 0u : n - 1));
}
#elif defined(RNMF_EMULATE)
// host emulation (tests/emu/emu_cuda.hpp; test builds only): "devices" share the process's memory, flags are std atomics
inline void peer_spin_wait(const unsigned long long* flag, unsigned long long target, unsigned long long* status)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) >= target) break;
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(emu::g_wait_timeout_s.load())) { __atomic_store_n(status, target, __ATOMIC_RELEASE); break; }
        std::this_thread::yield();
    }
}
inline void peer_publish(unsigned long long* done, unsigned long long target, unsigned long long* peer_flag, unsigned long long value)
{
    const unsigned long long old = __atomic_fetch_add(done, 1ULL, __ATOMIC_ACQ_REL);
    if (old + 1 == target) __atomic_store_n(peer_flag, value, __ATOMIC_RELEASE);
}
inline unsigned peer_chunk_order(unsigned b, unsigned n)
{
    return n < 3 ? b : (b < n - 2 ? b + 1 : (b == n - 2 ? 0u : n - 1));
}
#endif

// Time-skewed start of rnmf_solve_poisson_host (capi.cu): the host frames arrive in C chunks of z-planes, chunk c
// making valid planes < Z(c+1) = n2*(c+1)/C available (the last one also the top ghost plane).  Double sweep number
// s = 1, 2, .. (u_{2s} from u_{2s-2}) needs its input two planes beyond its output, so what chunk c makes computable
// is the band [max(0, Z(c)-2s), max(0, Z(c+1)-2s)) (the last chunk: up to n2): bands of equal s tile [0, n2), and
// band (c, s) reads exactly what bands (<= c, s-1) wrote.  Ping-pong is safe: band (c, s+1) writes the buffer that
// band (c+1, s) still reads, but only below plane Z(c+1)-2s-2, the lowest plane that read touches.
inline void skew_region(long long n2, int n_chunks, int c, long long s, long long* lo, long long* hi)
{
    const long long zc = n2 * c / n_chunks - 2 * s, zn = n2 * (c + 1) / n_chunks - 2 * s;
    *lo = c == 0 ? 0 : (zc > 0 ? zc : 0);
    *hi = c == n_chunks - 1 ? n2 : (zn > 0 ? zn : 0);
}

// The mirror image at the end of the call: the last `levels` launches (double sweeps; the last one may be the single
// norm-carrying sweep, for which the two-plane margin is conservative) finish the frame chunk by chunk from the bottom,
// so that chunk c can go back to the host while chunk c+1 is still being swept.  Level s = 1..levels of chunk c covers
// [G_s(c-1), G_s(c)), G_s(c) = min(n2, Z(c+1) + 2*(levels - s)) (the last chunk: n2): after level `levels` exactly the
// planes below Z(c+1) are final.  Band (c, s) reads what bands (<= c, s-1) wrote, and the ping-pong partner is only
// overwritten below the lowest plane a later band still reads (the same argument as for skew_region, mirrored).
inline void skew_tail_region(long long n2, int n_chunks, int c, long long s, long long levels, long long* lo, long long* hi)
{
    const long long m = 2 * (levels - s);
    const long long gp = n2 * c / n_chunks + m, gc = n2 * (c + 1) / n_chunks + m;
    *lo = c == 0 ? 0 : (gp < n2 ? gp : n2);
    *hi = c == n_chunks - 1 ? n2 : (gc < n2 ? gc : n2);
}

struct SweepArgs {
    FrameGeom g;
    const double* in;
    const double* rhs;       // same geometry as in/out (pitch equal)
    double* out;
    SweepCoef c;
    FaceSet faces;           // comp 0 (sweeps are for n_comp = 1 frames)
    int fuse_fill;           // 1: write face ghosts of `out` (ng == 1)
    double* partials;        // nullable: per-block partial sums of |out-in|
    long long k_begin, k_end;// 3D: range of valid planes to update (slab interior/boundary split)
                             // 2D: range of valid rows
    int variant;
    int chunk_override;      // 0 = planner decides (measurement knob "sweep_chunk")
    double* peer_lo_plane;   // fused peer-memory halo: the lo neighbour's hi GHOST slice / the hi neighbour's lo GHOST slice of
    double* peer_hi_plane;   //   their `out` buffers, or null (deep slices follow behind / in front of them)
    PeerSync ps;             // in-kernel handshake (enabled = 0: none)
    // temporally blocked kernels across slabs (DEEP HALO, capi.cu: deep_elems): the in / out / rhs buffers carry `deep` extra
    // slices of the slowest direction in front of the lo ghost slice and behind the hi ghost slice; nb_lo / nb_hi: a slab
    // neighbour sits behind that face and its depth boundary slices are in my ghost + deep slices when the launch starts
    int deep;
    int nb_lo, nb_hi;
    // 3D double sweep only: the launch covers [k_begin, hole_begin) and [hole_end, k_end) (hole_end > hole_begin; 0, 0 = no hole):
    // the two boundary wedges of a slab in ONE launch, so that both neighbours see the usual one-epoch handshake
    long long hole_begin, hole_end;
};
bool sweep_supports_peer_halo(const SweepArgs& a);

// ONE table of sweep-kernel variants (kernels_sweep.cu; option "sweep_variant", 0 = auto = the defaults)
struct SweepVariantInfo {
    int id, dim;
    int depth;          // sweeps per launch of the temporally blocked kernel used for groups of sweeps (1: single sweeps only)
    bool pairs;         // depth > 2: sweeps left over by the groups are paired up by the two-sweep kernel
    const char* what;
};
const SweepVariantInfo* sweep_variant_info(int dim, int id);    // nullptr: no such variant

// Two sweeps per pass over HBM (kernels_sweep_tb2.cu), also across z-slabs (a.deep >= 1, a.nb_lo / a.nb_hi)
int launch_jacobi_double_sweep(const SweepArgs& a, cudaStream_t s);
int double_sweep_boundary_ctas(const SweepArgs& a);
// Three sweeps per pass over HBM (kernels_sweep_tb3.cu): frames with even n0 / n1; across z-slabs a.deep >= 2
bool triple_sweep_applicable(const FrameGeom& g);
int launch_jacobi_triple_sweep(const SweepArgs& a, cudaStream_t s);
int triple_sweep_boundary_ctas(const SweepArgs& a);
// 2D analogue (kernels_sweep2d_tb2.cu), single GPU: a.k_begin..a.k_end = output rows
int launch_jacobi_double_sweep_2d(const SweepArgs& a, cudaStream_t s);
// four sweeps per launch (kernels_sweep2d_tbk.cu), also across y-slabs (a.deep >= 3, a.nb_lo / a.nb_hi)
int launch_jacobi_multi_sweep_2d(const SweepArgs& a, cudaStream_t s);
int multi_sweep_2d_boundary_ctas(const SweepArgs& a);
int sweep_boundary_ctas(const SweepArgs& a);     // CTAs per boundary side (tiles across the slice) of the kernel that will run

// peer-memory halo handshake (kernels_peer.cu): epoch flags live in each rank's own memory and
// are written by the neighbours over NVLink
int launch_peer_signal(unsigned long long* peer_flag_a, unsigned long long* peer_flag_b, unsigned long long value, cudaStream_t s);
int launch_peer_wait(const unsigned long long* flag_a, const unsigned long long* flag_b, unsigned long long target,
                     unsigned long long* status, cudaStream_t s);
int sweep_num_partials(const SweepArgs& a);
int launch_jacobi_sweep(const SweepArgs& a, cudaStream_t s);
int launch_reduce_partials(const double* partials, int n, double* out_dev, cudaStream_t s);

int launch_abs_sum_valid(const FrameGeom& g, const double* d, double* partials, int max_partials,
                         double* out_dev, cudaStream_t s);
int abs_sum_num_partials(const FrameGeom& g);

int launch_scale(long long n, double a, const double* x, double* out, cudaStream_t s);
int launch_add(long long n, const double* x, const double* y, double* out, cudaStream_t s);
int launch_mul(long long n, const double* x, const double* y, double* out, cudaStream_t s);
int launch_axpby(long long n, double a, const double* x, double b, const double* y, double* out, cudaStream_t s);
int launch_fill_synthetic(const FrameGeom& g, double* d, unsigned long long seed, cudaStream_t s);
int launch_unpack_dense(const FrameGeom& g, const double* dense, double* frame, cudaStream_t s);
int launch_unpack_dense_rows(const FrameGeom& g, const double* dense, double* frame, double* shell_dst, long long r0, long long r1, cudaStream_t s);
int launch_pack_dense_rows(const FrameGeom& g, const double* frame, double* dense, long long r0, long long r1, cudaStream_t s);
int launch_pack_valid(const FrameGeom& g, const double* d, double* packed, cudaStream_t s);
int launch_unpack_valid(const FrameGeom& g, const double* packed, double* d, cudaStream_t s);

int launch_varcoef_2d(const FrameGeom& g, const double* phi, const rnmf_mu_params& mu,
                      int with_source, double relax, double* out, cudaStream_t s);
int launch_gradient_h_2d(const FrameGeom& g, const double* psi, const FrameGeom& gh, double* h, cudaStream_t s);

int launch_jacobi_resident(const FrameGeom& g, double* buf0, double* buf1, const double* rhs, SweepCoef c,
                           const FaceSet& faces, long long n_sweeps, cudaStream_t s);
int launch_gs_wavefront_2d(const FrameGeom& g, double* psi, const double* rhs, SweepCoef c,
                           const FaceSet& faces, long long n_sweeps, cudaStream_t s);
