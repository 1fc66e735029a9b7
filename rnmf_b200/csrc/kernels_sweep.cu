// kernels_sweep.cu -- K1: the Jacobi sweep of solve_poisson (poisson.rs:80-89) for 2D 5-point
// and 3D 7-point frames, with the per-sweep ghost fill (poisson.rs:88 -> dataframe.rs:361) and
// the update norm (main.rs:68-69) fused in.
//
// Update expression, association exactly as poisson.rs:83-85 (3D: "+ c3*(U+D)" appended,
// SURVEY 8a), every operation rounded separately (__dmul_rn/__dadd_rn: no FMA contraction):
//      new = ((c0*rhs + c1*(E+W)) + c2*(N+S)) [+ c3*(U+D)]
//
// This file holds the sweep dispatcher and the NON-TMA comparison kernels (variants 1-4, 10, 11 in
// 3D; 1, 2 in 2D).  The defaults are the TMA/mbarrier kernels in kernels_sweep_tma.cu (3D) and
// kernels_sweep2d_tma.cu (2D); small L2-resident frames take kernels_resident.cu.  Measured at
// 768^3 (sustained): zmarch 195-231 GLUP/s, TMA 259; at 8192^2: ymarch 244, TMA 270
// (profiles/r1_power.md).
//
// Bandwidth plan of the kernels below (24 B per lattice update: read psi, read rhs, write psi'):
//   3D "zmarch": a CTA owns a (2*TX x TY) xy-tile and marches ZC planes in z.  Each thread
//   owns two adjacent x cells (one 128-bit access per array per plane) and keeps the planes
//   k-1, k, k+1 of its own cells in registers, so psi is fetched from L2/HBM once per cell;
//   the N/S/E/W neighbours of plane k were fetched one iteration earlier by neighbouring
//   threads and are served by L1.
//   2D "ymarch": same idea marching in y, a CTA owns a 2*TX wide column strip.
#include "common.cuh"

namespace {

__device__ __forceinline__ double upd2(const SweepCoef& c, double r, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(c.c0, r), __dmul_rn(c.c1, __dadd_rn(e, w))),
                     __dmul_rn(c.c2, __dadd_rn(n, s)));
}
__device__ __forceinline__ double upd3(const SweepCoef& c, double r, double e, double w, double n, double s,
                                       double u, double d)
{
    return __dadd_rn(upd2(c, r, e, w, n, s), __dmul_rn(c.c3, __dadd_rn(u, d)));
}

// ng == 1 ghost value from the adjacent (new) valid cell; FaceBC.s1 precomputed on the host.
__device__ __forceinline__ double ghost1(const FaceBC& b, double m)
{
    return b.mode == 1 ? __dsub_rn(b.s1, m) : (b.mode == 2 ? __dsub_rn(m, b.s1) : __dadd_rn(m, b.s1));
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

template <int NT>
__device__ __forceinline__ void block_partial(double acc, double* partials, int linear_block)
{
    __shared__ double warp_sums[NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    if ((tid & 31) == 0) warp_sums[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {
        double v = tid < NT / 32 ? warp_sums[tid] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (tid == 0) partials[linear_block] = v;
    }
}

struct SweepParams {
    FrameGeom g;
    const double* in;
    const double* rhs;
    double* out;
    SweepCoef c;
    FaceSet faces;
    int fuse_fill;
    double* partials;
    long long m_begin, m_end;   // marching range (valid k in 3D, valid j in 2D)
    int chunk;                  // marching chunk per CTA
};

// ---------------------------------------------------------------------------------------------
// 3D 7-point, register z-marching
// ---------------------------------------------------------------------------------------------
template <int TX, int TY, bool NORM, int MINB = 1>
__global__ void __launch_bounds__(TX* TY, MINB) jacobi3d_zmarch_kernel(const SweepParams p)
{
    const FrameGeom& g = p.g;
    const long long i0 = 2 * ((long long)blockIdx.x * TX + threadIdx.x);
    const long long j = (long long)blockIdx.y * TY + threadIdx.y;
    const long long k0 = p.m_begin + (long long)blockIdx.z * p.chunk;
    const long long k1 = k0 + p.chunk < p.m_end ? k0 + p.chunk : p.m_end;
    const bool active = i0 < g.n[0] && j < g.n[1];
    const bool has2 = i0 + 1 < g.n[0];
    double acc = 0.0;

    if (active) {
        const long long pitch = g.pitch, plane = g.plane;
        long long off = g.at(i0, j, k0, 0);
        const double* __restrict__ in = p.in;
        const double* __restrict__ rhs = p.rhs;
        double* __restrict__ out = p.out;

        double2 below = ld2(in + off - plane);
        double2 cur = ld2(in + off);
        const bool x_lo = p.fuse_fill && i0 == 0 && p.faces.f[0].mode;
        const bool x_hi = p.fuse_fill && (i0 + 2 >= g.n[0]) && p.faces.f[1].mode;
        const bool y_lo = p.fuse_fill && j == 0 && p.faces.f[2].mode;
        const bool y_hi = p.fuse_fill && j == g.n[1] - 1 && p.faces.f[3].mode;

        for (long long k = k0; k < k1; ++k, off += plane) {
            const double2 above = ld2(in + off + plane);
            const double2 nn = ld2(in + off + pitch);
            const double2 ss = ld2(in + off - pitch);
            const double w = in[off - 1];
            const double e = in[off + 2];
            const double2 r = ld2(rhs + off);
            double2 nw;
            nw.x = upd3(p.c, r.x, cur.y, w, nn.x, ss.x, above.x, below.x);
            nw.y = upd3(p.c, r.y, e, cur.x, nn.y, ss.y, above.y, below.y);
            if (has2) st2(out + off, nw); else out[off] = nw.x;
            if (NORM) {
                acc += fabs(__dsub_rn(nw.x, cur.x));
                if (has2) acc += fabs(__dsub_rn(nw.y, cur.y));
            }
            if (p.fuse_fill) {
                if (x_lo) out[off - 1] = ghost1(p.faces.f[0], nw.x);
                if (x_hi) { if (has2) out[off + 2] = ghost1(p.faces.f[1], nw.y); else out[off + 1] = ghost1(p.faces.f[1], nw.x); }
                if (y_lo) {
                    if (has2) st2(out + off - pitch, make_double2(ghost1(p.faces.f[2], nw.x), ghost1(p.faces.f[2], nw.y)));
                    else out[off - pitch] = ghost1(p.faces.f[2], nw.x);
                }
                if (y_hi) {
                    if (has2) st2(out + off + pitch, make_double2(ghost1(p.faces.f[3], nw.x), ghost1(p.faces.f[3], nw.y)));
                    else out[off + pitch] = ghost1(p.faces.f[3], nw.x);
                }
                if (k == 0 && p.faces.f[4].mode) {
                    if (has2) st2(out + off - plane, make_double2(ghost1(p.faces.f[4], nw.x), ghost1(p.faces.f[4], nw.y)));
                    else out[off - plane] = ghost1(p.faces.f[4], nw.x);
                }
                if (k == g.n[2] - 1 && p.faces.f[5].mode) {
                    if (has2) st2(out + off + plane, make_double2(ghost1(p.faces.f[5], nw.x), ghost1(p.faces.f[5], nw.y)));
                    else out[off + plane] = ghost1(p.faces.f[5], nw.x);
                }
            }
            below = cur;
            cur = above;
        }
    }
    if (NORM) {
        const int lb = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        block_partial<TX * TY>(acc, p.partials, lb);
    }
}

// ---------------------------------------------------------------------------------------------
// 2D 5-point, register y-marching
// ---------------------------------------------------------------------------------------------
template <int TX, bool NORM>
__global__ void __launch_bounds__(TX) jacobi2d_ymarch_kernel(const SweepParams p)
{
    const FrameGeom& g = p.g;
    const long long i0 = 2 * ((long long)blockIdx.x * TX + threadIdx.x);
    const long long j0 = p.m_begin + (long long)blockIdx.y * p.chunk;
    const long long j1 = j0 + p.chunk < p.m_end ? j0 + p.chunk : p.m_end;
    const bool active = i0 < g.n[0];
    const bool has2 = i0 + 1 < g.n[0];
    double acc = 0.0;

    if (active) {
        const long long pitch = g.pitch;
        long long off = g.at(i0, j0, 0, 0);
        const double* __restrict__ in = p.in;
        const double* __restrict__ rhs = p.rhs;
        double* __restrict__ out = p.out;
        double2 below = ld2(in + off - pitch);
        double2 cur = ld2(in + off);
        const bool x_lo = p.fuse_fill && i0 == 0 && p.faces.f[0].mode;
        const bool x_hi = p.fuse_fill && (i0 + 2 >= g.n[0]) && p.faces.f[1].mode;

        for (long long j = j0; j < j1; ++j, off += pitch) {
            const double2 above = ld2(in + off + pitch);
            const double w = in[off - 1];
            const double e = in[off + 2];
            const double2 r = ld2(rhs + off);
            double2 nw;
            nw.x = upd2(p.c, r.x, cur.y, w, above.x, below.x);
            nw.y = upd2(p.c, r.y, e, cur.x, above.y, below.y);
            if (has2) st2(out + off, nw); else out[off] = nw.x;
            if (NORM) {
                acc += fabs(__dsub_rn(nw.x, cur.x));
                if (has2) acc += fabs(__dsub_rn(nw.y, cur.y));
            }
            if (p.fuse_fill) {
                if (x_lo) out[off - 1] = ghost1(p.faces.f[0], nw.x);
                if (x_hi) { if (has2) out[off + 2] = ghost1(p.faces.f[1], nw.y); else out[off + 1] = ghost1(p.faces.f[1], nw.x); }
                if (j == 0 && p.faces.f[2].mode) {
                    if (has2) st2(out + off - pitch, make_double2(ghost1(p.faces.f[2], nw.x), ghost1(p.faces.f[2], nw.y)));
                    else out[off - pitch] = ghost1(p.faces.f[2], nw.x);
                }
                if (j == g.n[1] - 1 && p.faces.f[3].mode) {
                    if (has2) st2(out + off + pitch, make_double2(ghost1(p.faces.f[3], nw.x), ghost1(p.faces.f[3], nw.y)));
                    else out[off + pitch] = ghost1(p.faces.f[3], nw.x);
                }
            }
            below = cur;
            cur = above;
        }
    }
    if (NORM) {
        const int lb = blockIdx.x + gridDim.x * blockIdx.y;
        block_partial<TX>(acc, p.partials, lb);
    }
}

// deterministic final reduction: one block, fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ partials, int n,
                                                               double* __restrict__ out)
{
    __shared__ double sm[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += partials[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

struct Plan {
    dim3 grid, block;
    int chunk;
};

// variant ids: 3D: 1 = 32x8 tile threads (64x8 cells), 2 = 32x4, 3 = 64x4, 4 = 16x16 ; 2D: 1 = 128 thr, 2 = 256 thr
Plan make_plan(const SweepArgs& a)
{
    Plan pl;
    const FrameGeom& g = a.g;
    const long long span = a.k_end - a.k_begin;
    if (g.dim == 3) {
        int tx = 32, ty = 8;
        switch (a.variant) {
        case 2: tx = 32; ty = 4; break;
        case 3: tx = 64; ty = 4; break;
        case 4: tx = 16; ty = 16; break;
        case 10: case 11: tx = 32; ty = 8; break;
        default: break;
        }
        pl.block = dim3(tx, ty, 1);
        const long long bx = (g.n[0] + 2 * tx - 1) / (2 * tx), by = (g.n[1] + ty - 1) / ty;
        // chunk: aim at >= ~8 CTAs per SM overall while keeping the per-chunk re-read (2 planes) small
        long long want = (148LL * 8 * 4 + bx * by - 1) / (bx * by);
        if (want < 1) want = 1;
        long long chunk = (span + want - 1) / want;
        if (chunk < 16) chunk = 16;
        if (a.chunk_override > 0) chunk = a.chunk_override;
        if (chunk > span) chunk = span > 0 ? span : 1;
        long long bz = (span + chunk - 1) / chunk;
        if (bz > 65535) { bz = 65535; chunk = (span + bz - 1) / bz; bz = (span + chunk - 1) / chunk; }
        pl.chunk = (int)chunk;
        pl.grid = dim3((unsigned)bx, (unsigned)by, (unsigned)(bz > 0 ? bz : 1));
    } else {
        int tx = a.variant == 2 ? 256 : 128;
        pl.block = dim3(tx, 1, 1);
        const long long bx = (g.n[0] + 2 * tx - 1) / (2 * tx);
        long long want = (148LL * 8 * 4 + bx - 1) / bx;
        long long chunk = (span + want - 1) / want;
        if (chunk < 32) chunk = 32;
        if (chunk > span) chunk = span > 0 ? span : 1;
        long long by = (span + chunk - 1) / chunk;
        if (by > 65535) { by = 65535; chunk = (span + by - 1) / by; by = (span + chunk - 1) / chunk; }
        pl.chunk = (int)chunk;
        pl.grid = dim3((unsigned)bx, (unsigned)(by > 0 ? by : 1), 1);
    }
    return pl;
}

template <bool NORM>
void launch3d(const Plan& pl, const SweepParams& p, cudaStream_t s, int variant)
{
    const int tx = pl.block.x, ty = pl.block.y;
    if (variant == 10) jacobi3d_zmarch_kernel<32, 8, NORM, 6><<<pl.grid, pl.block, 0, s>>>(p);
    else if (variant == 11) jacobi3d_zmarch_kernel<32, 8, NORM, 8><<<pl.grid, pl.block, 0, s>>>(p);
    else if (tx == 32 && ty == 8) jacobi3d_zmarch_kernel<32, 8, NORM><<<pl.grid, pl.block, 0, s>>>(p);
    else if (tx == 32 && ty == 4) jacobi3d_zmarch_kernel<32, 4, NORM><<<pl.grid, pl.block, 0, s>>>(p);
    else if (tx == 64 && ty == 4) jacobi3d_zmarch_kernel<64, 4, NORM><<<pl.grid, pl.block, 0, s>>>(p);
    else jacobi3d_zmarch_kernel<16, 16, NORM><<<pl.grid, pl.block, 0, s>>>(p);
}
template <bool NORM>
void launch2d(const Plan& pl, const SweepParams& p, cudaStream_t s)
{
    if (pl.block.x == 256) jacobi2d_ymarch_kernel<256, NORM><<<pl.grid, pl.block, 0, s>>>(p);
    else jacobi2d_ymarch_kernel<128, NORM><<<pl.grid, pl.block, 0, s>>>(p);
}

}  // namespace

int tma_sweep_num_partials(const SweepArgs& a);
int launch_jacobi_sweep_tma(const SweepArgs& a, cudaStream_t s);
// default (variant 0 = auto): the TMA kernel for 3D frames (measured 259 vs 195-231 GLUP/s at 768^3, sustained)
// (variants 20/21 select the double-sweep kernel for pairs of sweeps; their single sweeps are the default TMA kernel)
static inline bool use_tma(const SweepArgs& a) { return a.g.dim == 3 && (a.variant == 0 || (a.variant >= 5 && a.variant <= 9) || a.variant == 13 || a.variant == 20 || a.variant == 21 || a.variant == 22 || a.variant == 23 || a.variant == 24 || a.variant == 25 || a.variant == 26 || a.variant == 27 || a.variant == 28); }

int tma2d_sweep_num_partials(const SweepArgs& a);
int launch_jacobi_sweep_tma2d(const SweepArgs& a, cudaStream_t s);
// 2D default (variant 0 = auto, or 20): the TMA row-streaming kernel; variants 1/2 = jacobi2d_ymarch
// (variant 30 selects the 2D double-sweep kernel for pairs of sweeps; its single sweeps are this kernel)
// (variants 32..40 select the K-sweep kernel for groups of K sweeps, kernels_sweep2d_tbk.cu; same single sweeps)
static inline bool use_tma2d(const SweepArgs& a) { return a.g.dim == 2 && (a.variant == 0 || a.variant == 20 || a.variant == 30 || (a.variant >= 32 && a.variant <= 40)); }

bool sweep_supports_peer_halo(const SweepArgs& a) { return use_tma(a) || use_tma2d(a); }
int tma_sweep_boundary_ctas(const SweepArgs& a);
int tma2d_sweep_boundary_ctas(const SweepArgs& a);
int sweep_boundary_ctas(const SweepArgs& a) { return use_tma(a) ? tma_sweep_boundary_ctas(a) : (use_tma2d(a) ? tma2d_sweep_boundary_ctas(a) : 0); }

int sweep_num_partials(const SweepArgs& a)
{
    if (use_tma(a)) return tma_sweep_num_partials(a);
    if (use_tma2d(a)) return tma2d_sweep_num_partials(a);
    const Plan pl = make_plan(a);
    return (int)(pl.grid.x * pl.grid.y * pl.grid.z);
}

int launch_jacobi_sweep(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    if (use_tma(a)) return launch_jacobi_sweep_tma(a, s);
    if (use_tma2d(a)) return launch_jacobi_sweep_tma2d(a, s);
    const Plan pl = make_plan(a);
    SweepParams p;
    p.g = a.g; p.in = a.in; p.rhs = a.rhs; p.out = a.out; p.c = a.c; p.faces = a.faces;
    p.fuse_fill = a.fuse_fill; p.partials = a.partials;
    p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = pl.chunk;
    if (a.g.dim == 3) {
        if (a.partials) launch3d<true>(pl, p, s, a.variant); else launch3d<false>(pl, p, s, a.variant);
    } else {
        if (a.partials) launch2d<true>(pl, p, s); else launch2d<false>(pl, p, s);
    }
    return 1;
}

int launch_reduce_partials(const double* partials, int n, double* out_dev, cudaStream_t s)
{
    reduce_partials_kernel<<<1, 1024, 0, s>>>(partials, n, out_dev);
    return 1;
}
