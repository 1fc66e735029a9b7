// kernels_sweep.cu -- K1: the dispatcher of the Jacobi sweep of solve_poisson (poisson.rs:80-89) for 2D
// 5-point and 3D 7-point frames, the ONE table of sweep-kernel variants, and the final reduction of the
// fused update norm (main.rs:68-69).
//
// Update expression, association exactly as poisson.rs:83-85 (3D: "+ c3*(U+D)" appended,
// SURVEY 8a), every operation rounded separately (__dmul_rn/__dadd_rn: no FMA contraction):
//      new = ((c0*rhs + c1*(E+W)) + c2*(N+S)) [+ c3*(U+D)]
//
// Kernels (all TMA + mbarrier pipelines, 24 B per lattice update and pass over HBM):
//   one sweep per launch    kernels_sweep_tma.cu (3D), kernels_sweep2d_tma.cu (2D): odd sweeps, the norm-carrying
//                           sweep, frames with n_ghost > 1 / Prescribed sides (separate fill), the NCCL halo path
//   two sweeps per launch   kernels_sweep_tb2.cu (3D, also across z-slabs), kernels_sweep2d_tb2.cu (2D)
//   four sweeps per launch  kernels_sweep2d_tbk.cu (2D)
//   L2-resident frames      kernels_resident.cu (all sweeps of a call in one cooperative launch)
// The non-TMA comparison kernels of round 1 (register marching through L1, 195-244 GLUP/s) are gone: parity is
// checked against the CPU restatement of the reference, at BASELINE.json's full sizes too (tests/test_gpu_parity.py).
#include "common.cuh"

namespace {

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

// Option "sweep_variant" (a measurement knob; 0 = auto = what ships).  Adding or removing a kernel = one row here plus its
// launcher; rnmf_jacobi_sweeps (capi.cu) only reads `depth` / `pairs`.
const SweepVariantInfo kVariants[] = {
    // id dim depth pairs  what
    { 0, 3, 3, true, "3D default: three sweeps per launch (128x12 tile, quad mapping, kernels_sweep_tb3.cu), then a pair, then single sweeps" },
    { 31, 3, 2, false, "3D: two sweeps per launch (kernels_sweep_tb2.cu), single sweeps by the TMA kernel" },
    { 6, 3, 1, false, "3D: one sweep per launch only (TMA kernel, 128x8 tile, 4 stages)" },
    { 0, 2, 4, true, "2D default: four sweeps per launch (512-column strips, de-interleaved rings), then pairs, then single sweeps" },
    { 30, 2, 2, false, "2D: two sweeps per launch (kernels_sweep2d_tb2.cu), single sweeps by the TMA kernel" },
    { 20, 2, 1, false, "2D: one sweep per launch only (TMA row-streaming kernel)" },
};

}  // namespace

const SweepVariantInfo* sweep_variant_info(int dim, int id)
{
    for (const SweepVariantInfo& v : kVariants)
        if (v.dim == dim && v.id == id) return &v;
    return nullptr;
}

int tma_sweep_num_partials(const SweepArgs& a);
int launch_jacobi_sweep_tma(const SweepArgs& a, cudaStream_t s);
int tma_sweep_boundary_ctas(const SweepArgs& a);
int tma2d_sweep_num_partials(const SweepArgs& a);
int launch_jacobi_sweep_tma2d(const SweepArgs& a, cudaStream_t s);
int tma2d_sweep_boundary_ctas(const SweepArgs& a);

// every single-sweep kernel stores its boundary slices into the neighbours' ghost slices (fused peer-memory halo)
bool sweep_supports_peer_halo(const SweepArgs&) { return true; }
int sweep_boundary_ctas(const SweepArgs& a) { return a.g.dim == 3 ? tma_sweep_boundary_ctas(a) : tma2d_sweep_boundary_ctas(a); }
int sweep_num_partials(const SweepArgs& a) { return a.g.dim == 3 ? tma_sweep_num_partials(a) : tma2d_sweep_num_partials(a); }

int launch_jacobi_sweep(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    return a.g.dim == 3 ? launch_jacobi_sweep_tma(a, s) : launch_jacobi_sweep_tma2d(a, s);
}

int launch_reduce_partials(const double* partials, int n, double* out_dev, cudaStream_t s)
{
    reduce_partials_kernel<<<1, 1024, 0, s>>>(partials, n, out_dev);
    return 1;
}
