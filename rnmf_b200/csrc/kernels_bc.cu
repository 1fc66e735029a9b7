// kernels_bc.cu -- ghost-cell boundary fill (K2) and ghost-shell copy.
// Replaces CartesianDataFrame2D::fill_bc and its helpers (dataframe.rs:289-433); 3D by the
// extrapolation rule: per direction, face interiors only, edges/corners untouched.
#include "common.cuh"

// All Dirichlet/Neumann faces of all components in ONE launch.  The faces write disjoint ghost
// cells and read valid cells only, so the reference's comp->dim->lo,hi order is immaterial.
// Work item = (component, direction, side, tangential cell t1 (fast), t2, ghost layer).
struct FillFacesParams {
    FrameGeom g;
    FaceSet faces[4];          // up to 4 components per launch
    int comp0, ncomp;
    long long face_items[3];   // items per (side, comp) for direction d = n[d1]*n[d2]*ng
    long long cum[4];          // prefix over directions of 2*ncomp*face_items[d]
};

__global__ void __launch_bounds__(256) fill_bc_faces_kernel(FillFacesParams p, double* __restrict__ data)
{
    const FrameGeom& g = p.g;
    const long long total = p.cum[g.dim];
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total;
         w += (long long)gridDim.x * blockDim.x) {
        int d = 0;
        while (d + 1 < g.dim && w >= p.cum[d + 1]) ++d;
        long long r = w - p.cum[d];
        const long long per = p.face_items[d];
        const int sc = (int)(r / per);            // side + 2*comp_local
        r -= (long long)sc * per;
        const int side = sc & 1, cl = sc >> 1;
        const FaceBC b = p.faces[cl].f[d * 2 + side];
        if (b.mode == 0) continue;
        // tangential directions: for d=0 -> (y fast, z); d=1 -> (x fast, z); d=2 -> (x fast, y)
        const int d1 = d == 0 ? 1 : 0;
        const int d2 = d == 2 ? 1 : 2;
        const long long n1 = g.n[d1];
        const long long t1 = r % n1;
        r /= n1;
        const long long n2 = g.n[d2];
        const long long t2 = r % n2;
        const int layer = (int)(r / n2);          // 0-based ghost layer
        long long idx[3];
        idx[d1] = t1; idx[d2] = t2;
        const long long nd = g.n[d];
        const long long mirror = side == 0 ? layer : nd - 1 - layer;
        const long long ghost = side == 0 ? -layer - 1 : nd + layer;
        idx[d] = mirror;
        const double m = data[g.at(idx[0], idx[1], idx[2], p.comp0 + cl)];
        idx[d] = ghost;
        data[g.at(idx[0], idx[1], idx[2], p.comp0 + cl)] = rnmf_ghost_value(b, side, m, layer, g.dx[d]);
    }
}

int launch_fill_bc_faces(const FrameGeom& g, double* d, const FaceSet* faces_host, cudaStream_t s)
{
    int launches = 0;
    for (int c0 = 0; c0 < g.ncomp; c0 += 4) {
        FillFacesParams p;
        p.g = g;
        p.comp0 = c0;
        p.ncomp = g.ncomp - c0 < 4 ? g.ncomp - c0 : 4;
        for (int c = 0; c < p.ncomp; ++c) p.faces[c] = faces_host[c0 + c];
        p.cum[0] = 0;
        for (int dd = 0; dd < 3; ++dd) {
            if (dd < g.dim) {
                const int d1 = dd == 0 ? 1 : 0, d2 = dd == 2 ? 1 : 2;
                p.face_items[dd] = g.n[d1] * g.n[d2] * g.ng;
            } else {
                p.face_items[dd] = 0;
            }
            p.cum[dd + 1] = p.cum[dd] + 2 * p.ncomp * p.face_items[dd];
        }
        const long long total = p.cum[g.dim];
        if (total == 0) continue;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        fill_bc_faces_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, d);
        ++launches;
    }
    return launches;
}

// Prescribed(values) for one (component, direction, side): every grown cell whose coordinate in
// `dir` is below 0 (lo) or >= n (hi), corners included, in flat order, zipped with `vals`
// (dataframe.rs:378-398, :413-428 with the side lists {West,NW,SW} etc.; :280-284).
__global__ void __launch_bounds__(256) fill_prescribed_kernel(FrameGeom g, double* __restrict__ data, int comp,
                                                              int dir, int side, const double* __restrict__ vals,
                                                              long long nvals)
{
    const long long ng = g.ng;
    // extents of the side region in grown index space
    long long e[3] = { g.G[0], g.G[1], g.G[2] };
    e[dir] = dir < g.dim ? ng : 0;
    const long long total = e[0] * e[1] * e[2];
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total && m < nvals;
         m += (long long)gridDim.x * blockDim.x) {
        // flat order inside the region == flat order of the full frame restricted to the region
        long long ii = m % e[0];
        long long r = m / e[0];
        long long jj = r % e[1];
        long long kk = r / e[1];
        long long q[3] = { ii, jj, kk };
        if (side == 1) q[dir] += g.n[dir] + ng;
        data[g.xoff + q[0] + g.pitch * (q[1] + g.G[1] * (q[2] + g.G[2] * comp))] = vals[m];
    }
}

int launch_fill_prescribed(const FrameGeom& g, double* d, int comp, int dir, int side,
                           const double* vals_dev, long long nvals, cudaStream_t s)
{
    long long e[3] = { g.G[0], g.G[1], g.G[2] };
    e[dir] = g.ng;
    long long total = e[0] * e[1] * e[2];
    if (total > nvals) total = nvals;
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fill_prescribed_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, d, comp, dir, side, vals_dev, nvals);
    return 1;
}

// Copies the never-written part of the ghost shell (cells that are out of range in >= 2
// directions: corners in 2D, edges+corners in 3D) plus, when `faces_too`, the face ghosts, from
// src to dst.  Used to make the ping-pong buffer of a sweep agree with the frame on the cells a
// sweep + fill never writes (the reference keeps them untouched, dataframe.rs:289-355).
__global__ void __launch_bounds__(256) copy_ghost_shell_kernel(FrameGeom g, const double* __restrict__ src,
                                                               double* __restrict__ dst)
{
    // iterate over all grown rows; inside a row either the whole row is shell (jj or kk out of
    // range) or only the 2*ng end cells are.
    const long long rows = g.G[1] * g.G[2] * g.ncomp;
    const long long ng = g.ng;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long jj = r % g.G[1];
        const long long kk = (r / g.G[1]) % g.G[2];
        const bool j_out = jj < ng || jj >= g.n[1] + ng;
        const bool k_out = g.dim == 3 && (kk < g.kg || kk >= g.n[2] + g.kg);
        const long long base = g.xoff + g.pitch * r;
        if (j_out || k_out) {
            for (long long ii = threadIdx.x; ii < g.G[0]; ii += blockDim.x) dst[base + ii] = src[base + ii];
        } else {
            for (long long t = threadIdx.x; t < 2 * ng; t += blockDim.x) {
                const long long ii = t < ng ? t : g.n[0] + t;   // t>=ng -> n0+ng+(t-ng)
                dst[base + ii] = src[base + ii];
            }
        }
    }
}

int launch_copy_ghost_shell(const FrameGeom& g, const double* src, double* dst, cudaStream_t s)
{
    long long rows = g.G[1] * g.G[2] * g.ncomp;
    if (rows <= 0) return 0;
    long long blocks = rows < 148 * 8 ? rows : 148 * 8;
    copy_ghost_shell_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, src, dst);
    return 1;
}

// Side gather (collect_typed_cells, dataframe.rs:273-278; side iterators :709-817): cells of one side
// of direction `dir`, flat order inside each component, corners included.
//   which = 0: valid strip adjacent to the face: 0 <= idx[dir] < ng (lo) / n-ng <= idx[dir] < n (hi),
//              the other directions over their VALID range (the lists Valid(North|NE|NW) ...);
//   which = 1: ghost side: idx[dir] < 0 (lo) / >= n (hi), other directions over the GROWN range.
__global__ void __launch_bounds__(256) collect_side_kernel(FrameGeom g, const double* __restrict__ data, int dir, int side,
                                                           int which, int comp0, int ncomp, double* __restrict__ out)
{
    long long e[3], lo[3];
    for (int d = 0; d < 3; ++d) {
        const long long gh = d < g.dim ? g.ng : 0;
        if (which == 0) { e[d] = g.n[d]; lo[d] = gh; }            // valid range, in grown coordinates
        else { e[d] = g.G[d]; lo[d] = 0; }
    }
    e[dir] = g.ng;
    if (which == 0) lo[dir] = side == 0 ? g.ng : g.n[dir];        // grown index of the strip's first layer
    else lo[dir] = side == 0 ? 0 : g.n[dir] + g.ng;
    const long long per = e[0] * e[1] * e[2];
    const long long total = per * ncomp;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
        const long long c = m / per;
        long long r = m % per;
        const long long ii = r % e[0] + lo[0];
        r /= e[0];
        const long long jj = r % e[1] + lo[1];
        const long long kk = r / e[1] + lo[2];
        out[m] = data[g.xoff + ii + g.pitch * (jj + g.G[1] * (kk + g.G[2] * (comp0 + c)))];
    }
}

long long collect_side_count(const FrameGeom& g, int dir, int which, int ncomp)
{
    long long e[3];
    for (int d = 0; d < 3; ++d) e[d] = which == 0 ? g.n[d] : g.G[d];
    e[dir] = g.ng;
    return e[0] * e[1] * e[2] * ncomp;
}

int launch_collect_side(const FrameGeom& g, const double* d, int dir, int side, int which, int comp0, int ncomp,
                        double* out_dev, cudaStream_t s)
{
    const long long total = collect_side_count(g, dir, which, ncomp);
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    collect_side_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, d, dir, side, which, comp0, ncomp, out_dev);
    return 1;
}
