// kernels_gs.cu -- K6: the reference's exact update order on the GPU.
// solve_poisson (poisson.rs:80-89) updates psi IN PLACE with i outer / j inner, i.e. lexicographic
// Gauss-Seidel.  Cell (i,j) of sweep s needs the NEW values of (i-1,j),(i,j-1) and the OLD
// (sweep s-1) values of (i+1,j),(i,j+1); with time t = i + j + 2*s all four are exactly one
// step old, and the in-place store at time t cannot clobber anything still needed.  One
// cooperative kernel walks t = 0 .. (nx-1)+(ny-1)+2(S-1) with a grid barrier per step, so the
// result is bit-identical to the serial reference loop.  Face ghosts needed by sweep s>=1 are
// what fill_bc wrote after sweep s-1 = f(own old value) (dataframe.rs:289-355, ng=1), computed
// inline; sweep 0 reads the stored ghosts.  The final fill_bc is done by the caller.
// Parity mode: it is latency-bound by construction and is not the throughput path.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ double ghost1(const FaceBC& b, double m)
{
    return b.mode == 1 ? __dsub_rn(b.s1, m) : (b.mode == 2 ? __dsub_rn(m, b.s1) : __dadd_rn(m, b.s1));
}

__global__ void __launch_bounds__(256) gs_wavefront_2d_kernel(FrameGeom g, double* psi, const double* __restrict__ rhs,
                                                              SweepCoef c, FaceSet faces, long long S)
{
    cg::grid_group grid = cg::this_grid();
    const long long nx = g.n[0], ny = g.n[1];
    const long long t_end = (nx - 1) + (ny - 1) + 2 * (S - 1);
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;
    for (long long t = 0; t <= t_end; ++t) {
        long long s_lo = t - (nx + ny - 2);
        s_lo = s_lo > 0 ? (s_lo + 1) / 2 : 0;
        long long s_hi = t / 2;
        if (s_hi > S - 1) s_hi = S - 1;
        const long long items = (s_hi - s_lo + 1) * nx;
        for (long long w = gtid; w < items; w += gsize) {
            const long long s = s_lo + w / nx;
            const long long i = w % nx;
            const long long j = t - 2 * s - i;
            if (j < 0 || j >= ny) continue;
            const long long p = g.at(i, j, 0, 0);
            const double self = __ldcg(psi + p);
            double e = __ldcg(psi + p + 1), wv = __ldcg(psi + p - 1);
            double n = __ldcg(psi + p + g.pitch), so = __ldcg(psi + p - g.pitch);
            if (s > 0) {
                if (i == 0) wv = ghost1(faces.f[0], self);
                if (i == nx - 1) e = ghost1(faces.f[1], self);
                if (j == 0) so = ghost1(faces.f[2], self);
                if (j == ny - 1) n = ghost1(faces.f[3], self);
            }
            // poisson.rs:83-85
            const double v = __dadd_rn(__dadd_rn(__dmul_rn(c.c0, rhs[p]), __dmul_rn(c.c1, __dadd_rn(e, wv))),
                                       __dmul_rn(c.c2, __dadd_rn(n, so)));
            __stcg(psi + p, v);
        }
        grid.sync();
    }
}

}  // namespace

int launch_gs_wavefront_2d(const FrameGeom& g, double* psi, const double* rhs, SweepCoef c,
                           const FaceSet& faces, long long n_sweeps, cudaStream_t s)
{
    if (n_sweeps <= 0) return 0;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_wavefront_2d_kernel, 256, 0) != cudaSuccess) return -1;
    if (per_sm < 1) return -1;
    // measured on the 301^2 case: 0.079 / 0.081 / 0.092 s for 4 / 2 / 1 CTAs per SM
    { const char* e = getenv("RNMF_GS_CTAS_PER_SM"); const int want = e ? atoi(e) : 4; if (per_sm > want) per_sm = want; }
    // no more CTAs than the widest hyperplane family can use
    long long max_items = g.n[0] * ((g.n[0] + g.n[1]) / 2 + 1);
    long long blocks = (long long)sms * per_sm;
    if ((max_items + 255) / 256 < blocks) blocks = (max_items + 255) / 256;
    if (blocks < 1) blocks = 1;
    FrameGeom gg = g;
    long long S = n_sweeps;
    void* args[] = { &gg, &psi, &rhs, &c, const_cast<FaceSet*>(&faces), &S };
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)gs_wavefront_2d_kernel, dim3((unsigned)blocks), dim3(256),
                                                args, 0, s);
    if (e != cudaSuccess) {
        rnmf_set_error("cudaLaunchCooperativeKernel: %s", cudaGetErrorString(e));
        return -1;
    }
    return 1;
}
