// kernels_ops.cu -- K3 (variable-mu operator / source), K4 (whole-frame operators), K5 (valid-cell
// |.| sum), plus pack/unpack of valid cells and the synthetic-field generator.
#include "common.cuh"

namespace {

constexpr int kEwThreads = 256;

inline unsigned ew_blocks(long long n_vec)
{
    long long b = (n_vec + kEwThreads - 1) / kEwThreads;
    const long long cap = 148LL * 32;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- K4: element-wise over the ENTIRE data vector (dataframe.rs:628-691) -------------------
// The pitched buffer is processed as a flat array (padding included; it is never read back).
// total is a multiple of 16 doubles, so 128-bit accesses need no tail.
enum { OP_SCALE = 0, OP_ADD = 1, OP_MUL = 2, OP_AXPBY = 3 };

// No __restrict__: the API is used in place (psi = psi + relax*diff passes out == x); each element is read and
// written once by the same thread, so aliasing is well defined as long as the compiler is not told otherwise.
template <int OP>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(long long n2, double a, const double2* x,
                                                        double b, const double2* y,
                                                        double2* out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2;
         i += (long long)gridDim.x * blockDim.x) {
        const double2 xv = x[i];
        double2 yv = make_double2(0.0, 0.0);
        if (OP != OP_SCALE) yv = y[i];
        double2 r;
        if (OP == OP_SCALE) { r.x = __dmul_rn(a, xv.x); r.y = __dmul_rn(a, xv.y); }
        if (OP == OP_ADD) { r.x = __dadd_rn(xv.x, yv.x); r.y = __dadd_rn(xv.y, yv.y); }
        if (OP == OP_MUL) { r.x = __dmul_rn(xv.x, yv.x); r.y = __dmul_rn(xv.y, yv.y); }
        if (OP == OP_AXPBY) {
            r.x = __dadd_rn(__dmul_rn(a, xv.x), __dmul_rn(b, yv.x));
            r.y = __dadd_rn(__dmul_rn(a, xv.y), __dmul_rn(b, yv.y));
        }
        out[i] = r;
    }
}

// ---- synthetic field: splitmix64(seed ^ q) -> (-1,1), q = global flat valid index ---
__device__ __forceinline__ double splitmix_unit(unsigned long long seed, unsigned long long q)
{
    unsigned long long z = (q ^ seed) + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return __dsub_rn(__dmul_rn((double)(z >> 11), 2.0 / 9007199254740992.0), 1.0);
}

__global__ void __launch_bounds__(256) fill_synthetic_kernel(FrameGeom g, double* __restrict__ d, unsigned long long seed)
{
    const long long rows = g.n[1] * g.n[2] * g.ncomp;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long j = r % g.n[1];
        const long long k = (r / g.n[1]) % g.n[2];
        const long long c = r / (g.n[1] * g.n[2]);
        const long long base = g.at(0, j, k, c);
        const unsigned long long q0 =
            (unsigned long long)((g.off[0]) + g.nglob[0] * ((j + g.off[1]) + g.nglob[1] * ((k + g.off[2]) + g.nglob[2] * c)));
        for (long long i = threadIdx.x; i < g.n[0]; i += blockDim.x) d[base + i] = splitmix_unit(seed, q0 + (unsigned long long)i);
    }
}

// ---- pack / unpack of valid cells (get_valid_cells order, dataframe.rs:181-189) --------------
template <bool PACK>
__global__ void __launch_bounds__(256) pack_valid_kernel(FrameGeom g, double* __restrict__ frame, double* __restrict__ packed)
{
    const long long rows = g.n[1] * g.n[2] * g.ncomp;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long j = r % g.n[1];
        const long long k = (r / g.n[1]) % g.n[2];
        const long long c = r / (g.n[1] * g.n[2]);
        const long long base = g.at(0, j, k, c);
        for (long long i = threadIdx.x; i < g.n[0]; i += blockDim.x) {
            if (PACK) packed[r * g.n[0] + i] = frame[base + i];
            else frame[base + i] = packed[r * g.n[0] + i];
        }
    }
}

// ---- dense (reference layout, G0 doubles per row) -> pitched frame, all grown rows -----------
__global__ void __launch_bounds__(256) unpack_dense_kernel(FrameGeom g, const double* __restrict__ dense, double* __restrict__ frame)
{
    const long long rows = g.G[1] * g.G[2] * g.ncomp;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const double* src = dense + r * g.G[0];
        double* dst = frame + g.xoff + g.pitch * r;
        for (long long i = threadIdx.x; i < g.G[0]; i += blockDim.x) dst[i] = src[i];
    }
}

// ---- the same for grown rows [r0, r1) only (chunked upload, rnmf_solve_poisson_host); when `shell_dst` is given the
// ghost shell of those rows (copy_ghost_shell_kernel's cells: face, edge and corner ghosts) also goes to the frame's
// ping-pong partner, so no whole-frame shell copy is needed before the first sweeps
__global__ void __launch_bounds__(256) unpack_dense_rows_kernel(FrameGeom g, const double* __restrict__ dense, double* __restrict__ frame,
                                                                 double* __restrict__ shell_dst, long long r0, long long r1)
{
    const long long ng = g.ng;
    for (long long r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
        const double* src = dense + r * g.G[0];
        const long long base = g.xoff + g.pitch * r;
        for (long long i = threadIdx.x; i < g.G[0]; i += blockDim.x) frame[base + i] = src[i];
        if (shell_dst) {
            const long long jj = r % g.G[1];
            const long long kk = (r / g.G[1]) % g.G[2];
            const bool j_out = jj < ng || jj >= g.n[1] + ng;
            const bool k_out = g.dim == 3 && (kk < g.kg || kk >= g.n[2] + g.kg);
            if (j_out || k_out) {
                for (long long i = threadIdx.x; i < g.G[0]; i += blockDim.x) shell_dst[base + i] = src[i];
            } else {
                for (long long t = threadIdx.x; t < 2 * ng; t += blockDim.x) {
                    const long long ii = t < ng ? t : g.n[0] + t;
                    shell_dst[base + ii] = src[ii];
                }
            }
        }
    }
}

// ---- pitched frame -> dense, grown rows [r0, r1) (chunked download, rnmf_solve_poisson_host) ----
__global__ void __launch_bounds__(256) pack_dense_rows_kernel(FrameGeom g, const double* __restrict__ frame, double* __restrict__ dense,
                                                               long long r0, long long r1)
{
    for (long long r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
        const double* src = frame + g.xoff + g.pitch * r;
        double* dst = dense + r * g.G[0];
        for (long long i = threadIdx.x; i < g.G[0]; i += blockDim.x) dst[i] = src[i];
    }
}

// ---- K5: sum |v| over valid cells (poisson.rs:94-100), deterministic two-stage tree ---------
constexpr int kSumThreads = 256;
__global__ void __launch_bounds__(kSumThreads) abs_sum_valid_kernel(FrameGeom g, const double* __restrict__ d,
                                                                    double* __restrict__ partials)
{
    __shared__ double warp_sums[kSumThreads / 32];
    const long long rows = g.n[1] * g.n[2] * g.ncomp;
    double acc = 0.0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long j = r % g.n[1];
        const long long k = (r / g.n[1]) % g.n[2];
        const long long c = r / (g.n[1] * g.n[2]);
        const double* row = d + g.at(0, j, k, c);            // 128-B aligned
        const long long n2 = g.n[0] >> 1;
        for (long long i = threadIdx.x; i < n2; i += blockDim.x) {
            const double2 v = reinterpret_cast<const double2*>(row)[i];
            acc += fabs(v.x);
            acc += fabs(v.y);
        }
        if ((g.n[0] & 1) && threadIdx.x == 0) acc += fabs(row[g.n[0] - 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kSumThreads / 32 ? warp_sums[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}

// ---- K3: get_laplace / get_source (poisson.rs:13-64) ----------------------------------------
__device__ __forceinline__ double mu_at(long long i, long long j, const FrameGeom& g, const rnmf_mu_params& m)
{
    // poisson.rs:47-64: x=(i+0.5)*dx0, y=(j+0.5)*dx1; inside iff dx^2+dy^2 <= r*r  (powf(2.0) == x*x)
    const double x = __dmul_rn(__dadd_rn((double)i, 0.5), g.dx[0]);
    const double y = __dmul_rn(__dadd_rn((double)j, 0.5), g.dx[1]);
    const double xd = __dsub_rn(x, m.bubble_centre[0]);
    const double yd = __dsub_rn(y, m.bubble_centre[1]);
    const double d2 = __dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd));
    return d2 <= __dmul_rn(m.bubble_radius, m.bubble_radius) ? m.mu[0] : m.mu[1];
}

__device__ __forceinline__ double lap_at(long long i, long long j, const FrameGeom& g, const double* __restrict__ phi,
                                         const rnmf_mu_params& m)
{
    // poisson.rs:26-31, association as written there
    const long long p = g.at(i, j, 0, 0);
    const double mc = mu_at(i, j, g, m);
    const double me = mu_at(i + 1, j, g, m), mw = mu_at(i - 1, j, g, m);
    const double mn = mu_at(i, j + 1, g, m), ms = mu_at(i, j - 1, g, m);
    const double te = __ddiv_rn(__dmul_rn(phi[p + 1], __dadd_rn(me, mc)), 2.0);
    const double tw = __ddiv_rn(__dmul_rn(phi[p - 1], __dadd_rn(mw, mc)), 2.0);
    const double tn = __ddiv_rn(__dmul_rn(phi[p + g.pitch], __dadd_rn(mn, mc)), 2.0);
    const double ts = __ddiv_rn(__dmul_rn(phi[p - g.pitch], __dadd_rn(ms, mc)), 2.0);
    const double msum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(me, mw), mn), ms), __dmul_rn(4.0, mc));
    const double tc = __ddiv_rn(__dmul_rn(phi[p], msum), 2.0);
    return __dsub_rn(__dadd_rn(__dadd_rn(__dadd_rn(te, tw), tn), ts), tc);
}

// One thread per GROWN cell of `out`: valid -> lap (times source); face ghost -> the Dirichlet(0)
// fill of lap, 2.0*0.0 - lap(mirror) (poisson.rs:14-21,34), times the zero ghost of `source`
// (poisson.rs:43 multiplies over all data); corners: 0 (lap corner 0 [* 0]).
template <bool WITH_SOURCE>
__global__ void __launch_bounds__(256) varcoef_2d_kernel(FrameGeom g, const double* __restrict__ phi, rnmf_mu_params m,
                                                         double relax, double* __restrict__ out)
{
    const long long total = g.G[0] * g.G[1];
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total;
         w += (long long)gridDim.x * blockDim.x) {
        const long long ii = w % g.G[0], jj = w / g.G[0];
        const long long i = ii - g.ng, j = jj - g.ng;
        const bool xo = i < 0 || i >= g.n[0], yo = j < 0 || j >= g.n[1];
        double v;
        if (!xo && !yo) {
            v = lap_at(i, j, g, phi, m);
            if (WITH_SOURCE) {
                // poisson.rs:41: -(1.0 - 1.0/relax)/mu(i,j)
                const double s = __ddiv_rn(-__dsub_rn(1.0, __ddiv_rn(1.0, relax)), mu_at(i, j, g, m));
                v = __dmul_rn(s, v);
            }
        } else if (xo && yo) {
            v = WITH_SOURCE ? __dmul_rn(0.0, 0.0) : 0.0;
        } else {
            // mirror valid cell of this ghost layer (dataframe.rs:327-329, :343-346)
            long long mi = i, mj = j;
            if (xo) mi = i < 0 ? -i - 1 : 2 * g.n[0] - 1 - i;
            else mj = j < 0 ? -j - 1 : 2 * g.n[1] - 1 - j;
            const double gl = __dsub_rn(__dmul_rn(2.0, 0.0), lap_at(mi, mj, g, phi, m));
            v = WITH_SOURCE ? __dmul_rn(0.0, gl) : gl;
        }
        out[g.xoff + ii + g.pitch * jj] = v;
    }
}

// get_h (model.rs:72-88)
__global__ void __launch_bounds__(256) gradient_h_2d_kernel(FrameGeom g, const double* __restrict__ psi, FrameGeom gh,
                                                            double* __restrict__ h)
{
    const long long total = g.n[0] * g.n[1];
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total;
         w += (long long)gridDim.x * blockDim.x) {
        const long long i = w % g.n[0], j = w / g.n[0];
        const long long p = g.at(i, j, 0, 0);
        h[gh.at(i, j, 0, 0)] = __ddiv_rn(-__dsub_rn(psi[p + 1], psi[p - 1]), __dmul_rn(2.0, g.dx[0]));
        h[gh.at(i, j, 0, 1)] = __ddiv_rn(-__dsub_rn(psi[p + g.pitch], psi[p - g.pitch]), __dmul_rn(2.0, g.dx[1]));
    }
}

}  // namespace

int launch_scale(long long n, double a, const double* x, double* out, cudaStream_t s)
{
    ew_kernel<OP_SCALE><<<ew_blocks(n / 2), kEwThreads, 0, s>>>(n / 2, a, (const double2*)x, 0.0, nullptr, (double2*)out);
    return 1;
}
int launch_add(long long n, const double* x, const double* y, double* out, cudaStream_t s)
{
    ew_kernel<OP_ADD><<<ew_blocks(n / 2), kEwThreads, 0, s>>>(n / 2, 0.0, (const double2*)x, 0.0, (const double2*)y, (double2*)out);
    return 1;
}
int launch_mul(long long n, const double* x, const double* y, double* out, cudaStream_t s)
{
    ew_kernel<OP_MUL><<<ew_blocks(n / 2), kEwThreads, 0, s>>>(n / 2, 0.0, (const double2*)x, 0.0, (const double2*)y, (double2*)out);
    return 1;
}
int launch_axpby(long long n, double a, const double* x, double b, const double* y, double* out, cudaStream_t s)
{
    ew_kernel<OP_AXPBY><<<ew_blocks(n / 2), kEwThreads, 0, s>>>(n / 2, a, (const double2*)x, b, (const double2*)y, (double2*)out);
    return 1;
}

static unsigned row_blocks(long long rows)
{
    const long long cap = 148LL * 16;
    return (unsigned)(rows < 1 ? 1 : (rows > cap ? cap : rows));
}

int launch_fill_synthetic(const FrameGeom& g, double* d, unsigned long long seed, cudaStream_t s)
{
    fill_synthetic_kernel<<<row_blocks(g.n[1] * g.n[2] * g.ncomp), 256, 0, s>>>(g, d, seed);
    return 1;
}
int launch_unpack_dense(const FrameGeom& g, const double* dense, double* frame, cudaStream_t s)
{
    unpack_dense_kernel<<<row_blocks(g.G[1] * g.G[2] * g.ncomp), 256, 0, s>>>(g, dense, frame);
    return 1;
}
int launch_unpack_dense_rows(const FrameGeom& g, const double* dense, double* frame, double* shell_dst, long long r0, long long r1, cudaStream_t s)
{
    if (r1 <= r0) return 0;
    unpack_dense_rows_kernel<<<row_blocks(r1 - r0), 256, 0, s>>>(g, dense, frame, shell_dst, r0, r1);
    return 1;
}
int launch_pack_dense_rows(const FrameGeom& g, const double* frame, double* dense, long long r0, long long r1, cudaStream_t s)
{
    if (r1 <= r0) return 0;
    pack_dense_rows_kernel<<<row_blocks(r1 - r0), 256, 0, s>>>(g, frame, dense, r0, r1);
    return 1;
}
int launch_pack_valid(const FrameGeom& g, const double* d, double* packed, cudaStream_t s)
{
    pack_valid_kernel<true><<<row_blocks(g.n[1] * g.n[2] * g.ncomp), 256, 0, s>>>(g, const_cast<double*>(d), packed);
    return 1;
}
int launch_unpack_valid(const FrameGeom& g, const double* packed, double* d, cudaStream_t s)
{
    pack_valid_kernel<false><<<row_blocks(g.n[1] * g.n[2] * g.ncomp), 256, 0, s>>>(g, d, const_cast<double*>(packed));
    return 1;
}

int abs_sum_num_partials(const FrameGeom& g)
{
    return (int)row_blocks(g.n[1] * g.n[2] * g.ncomp);
}
int launch_reduce_partials(const double* partials, int n, double* out_dev, cudaStream_t s);
int launch_abs_sum_valid(const FrameGeom& g, const double* d, double* partials, int max_partials,
                         double* out_dev, cudaStream_t s)
{
    const int nb = abs_sum_num_partials(g);
    if (nb > max_partials) return -1;
    abs_sum_valid_kernel<<<nb, kSumThreads, 0, s>>>(g, d, partials);
    launch_reduce_partials(partials, nb, out_dev, s);
    return 2;
}

int launch_varcoef_2d(const FrameGeom& g, const double* phi, const rnmf_mu_params& mu, int with_source,
                      double relax, double* out, cudaStream_t s)
{
    const long long total = g.G[0] * g.G[1];
    const unsigned nb = ew_blocks(total);
    if (with_source) varcoef_2d_kernel<true><<<nb, 256, 0, s>>>(g, phi, mu, relax, out);
    else varcoef_2d_kernel<false><<<nb, 256, 0, s>>>(g, phi, mu, relax, out);
    return 1;
}

int launch_gradient_h_2d(const FrameGeom& g, const double* psi, const FrameGeom& gh, double* h, cudaStream_t s)
{
    gradient_h_2d_kernel<<<ew_blocks(g.n[0] * g.n[1]), 256, 0, s>>>(g, psi, gh, h);
    return 1;
}
