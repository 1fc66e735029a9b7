// kernels_resident.cu -- K1 for SMALL frames: many Jacobi sweeps in one cooperative kernel.
//
// A frame like the reference's shipped case (301 x 301, 0.7 MB) lives entirely in L2; one launch
// per sweep makes such runs launch-latency bound (~5 us per sweep for ~1 us of work).  This kernel
// keeps all CTAs resident, gives every thread a fixed set of cell pairs and runs S sweeps
// back to back, ping-ponging between the frame's two buffers with a grid barrier per sweep.
// Neighbours are read through L2 (ld.global.cg), never through a possibly stale L1 line.
// Same explicitly rounded update and fused ng=1 ghost fill as the streaming kernels
// (poisson.rs:83-85,88; dataframe.rs:289-355) => bit-identical results.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ double ghost1(const FaceBC& b, double m)
{
    return b.mode == 1 ? __dsub_rn(b.s1, m) : (b.mode == 2 ? __dsub_rn(m, b.s1) : __dadd_rn(m, b.s1));
}
__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stcg2(double* p, double2 v) { __stcg(reinterpret_cast<double2*>(p), v); }

struct ResidentParams {
    FrameGeom g;
    double* buf0;          // holds the input of sweep 0
    double* buf1;
    const double* rhs;
    SweepCoef c;
    FaceSet faces;
    long long n_sweeps;
};

__global__ void __launch_bounds__(256) jacobi_resident_kernel(const __grid_constant__ ResidentParams p)
{
    cg::grid_group grid = cg::this_grid();
    const FrameGeom& g = p.g;
    const long long n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    const long long pairs = (n0 + 1) / 2;
    const long long items = pairs * n1 * n2;
    const long long pitch = g.pitch, plane = g.plane;
    const bool is3d = g.dim == 3;
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;
    const double* in = p.buf0;
    double* out = p.buf1;
    for (long long sweep = 0; sweep < p.n_sweeps; ++sweep) {
        for (long long w = gtid; w < items; w += gsize) {
            const long long q = w % pairs;
            const long long r = w / pairs;
            const long long j = r % n1, k = r / n1;
            const long long i = 2 * q;
            const bool has2 = i + 1 < n0;
            const long long off = g.at(i, j, k, 0);
            const double2 cur = ldcg2(in + off);
            const double2 nn = ldcg2(in + off + pitch), ss = ldcg2(in + off - pitch);
            const double wv = __ldcg(in + off - 1), ev = __ldcg(in + off + 2);
            const double2 rr = *reinterpret_cast<const double2*>(p.rhs + off);
            double2 nw;
            nw.x = __dadd_rn(__dadd_rn(__dmul_rn(p.c.c0, rr.x), __dmul_rn(p.c.c1, __dadd_rn(cur.y, wv))), __dmul_rn(p.c.c2, __dadd_rn(nn.x, ss.x)));
            nw.y = __dadd_rn(__dadd_rn(__dmul_rn(p.c.c0, rr.y), __dmul_rn(p.c.c1, __dadd_rn(ev, cur.x))), __dmul_rn(p.c.c2, __dadd_rn(nn.y, ss.y)));
            if (is3d) {
                const double2 uu = ldcg2(in + off + plane), dd = ldcg2(in + off - plane);
                nw.x = __dadd_rn(nw.x, __dmul_rn(p.c.c3, __dadd_rn(uu.x, dd.x)));
                nw.y = __dadd_rn(nw.y, __dmul_rn(p.c.c3, __dadd_rn(uu.y, dd.y)));
            }
            if (has2) stcg2(out + off, nw); else __stcg(out + off, nw.x);
            // fused ghost fill (ng = 1): faces only, edges/corners untouched
            if (i == 0 && p.faces.f[0].mode) __stcg(out + off - 1, ghost1(p.faces.f[0], nw.x));
            if (i + 2 >= n0 && p.faces.f[1].mode) {
                if (has2) __stcg(out + off + 2, ghost1(p.faces.f[1], nw.y)); else __stcg(out + off + 1, ghost1(p.faces.f[1], nw.x));
            }
            if (j == 0 && p.faces.f[2].mode) {
                if (has2) stcg2(out + off - pitch, make_double2(ghost1(p.faces.f[2], nw.x), ghost1(p.faces.f[2], nw.y)));
                else __stcg(out + off - pitch, ghost1(p.faces.f[2], nw.x));
            }
            if (j == n1 - 1 && p.faces.f[3].mode) {
                if (has2) stcg2(out + off + pitch, make_double2(ghost1(p.faces.f[3], nw.x), ghost1(p.faces.f[3], nw.y)));
                else __stcg(out + off + pitch, ghost1(p.faces.f[3], nw.x));
            }
            if (is3d) {
                if (k == 0 && p.faces.f[4].mode) {
                    if (has2) stcg2(out + off - plane, make_double2(ghost1(p.faces.f[4], nw.x), ghost1(p.faces.f[4], nw.y)));
                    else __stcg(out + off - plane, ghost1(p.faces.f[4], nw.x));
                }
                if (k == n2 - 1 && p.faces.f[5].mode) {
                    if (has2) stcg2(out + off + plane, make_double2(ghost1(p.faces.f[5], nw.x), ghost1(p.faces.f[5], nw.y)));
                    else __stcg(out + off + plane, ghost1(p.faces.f[5], nw.x));
                }
            }
        }
        grid.sync();
        double* t = const_cast<double*>(in);
        in = out;
        out = t;
    }
}

}  // namespace

// Runs n_sweeps sweeps starting from buf0; the result is in (n_sweeps odd ? buf1 : buf0).
int launch_jacobi_resident(const FrameGeom& g, double* buf0, double* buf1, const double* rhs, SweepCoef c,
                           const FaceSet& faces, long long n_sweeps, cudaStream_t s)
{
    if (n_sweeps <= 0) return 0;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_resident_kernel, 256, 0) != cudaSuccess || per_sm < 1) return -1;
    if (per_sm > 4) per_sm = 4;
    const long long items = ((g.n[0] + 1) / 2) * g.n[1] * g.n[2];
    long long blocks = (long long)sms * per_sm;
    if ((items + 255) / 256 < blocks) blocks = (items + 255) / 256;
    if (blocks < 1) blocks = 1;
    ResidentParams p;
    p.g = g; p.buf0 = buf0; p.buf1 = buf1; p.rhs = rhs; p.c = c; p.faces = faces; p.n_sweeps = n_sweeps;
    void* args[] = { &p };
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)jacobi_resident_kernel, dim3((unsigned)blocks), dim3(256), args, 0, s);
    if (e != cudaSuccess) {
        rnmf_set_error("cudaLaunchCooperativeKernel(jacobi_resident): %s", cudaGetErrorString(e));
        return -1;
    }
    return 1;
}
