// kernels_sweep_tma.cu -- K1, Blackwell-native variant of the 3D 7-point Jacobi sweep
// (poisson.rs:80-89 as out-of-place Jacobi; fused ghost fill + update norm as in kernels_sweep.cu).
//
// Data movement is done entirely by TMA (cp.async.bulk.tensor, SASS UTMALDG):
//   * a CTA owns an (TX x TY) xy-tile and marches a chunk of z-planes;
//   * one producer warp streams, plane by plane, the psi tile WITH its xy halo
//     ((TX+4) x (TY+2) box: the x halo is two cells wide so the interior starts 16-byte aligned in
//     shared memory) and the rhs tile (TX x TY box) into an NS-stage shared-memory ring, completion
//     signalled on mbarriers (full[]); consumers hand stages back through empty[] mbarriers;
//   * TY consumer warps (one tile row each) keep planes k-1, k, k+1 of their own cells in registers
//     (register z-marching), read N/S neighbours and rhs from shared memory with conflict-free
//     128-bit loads and get E/W neighbours by warp shuffles, then store psi' with 128-bit
//     coalesced global stores.
// No consumer thread ever waits on a global load, so a handful of warps per SM keeps
// (NS-2) planes x ~18 KB x CTAs/SM of HBM traffic in flight.
//
// Arithmetic: same explicitly rounded sequence as kernels_sweep.cu => bit-identical results.
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include <stdlib.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ double upd3(const SweepCoef& c, double r, double e, double w, double n, double s,
                                       double u, double d)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.c0, r), __dmul_rn(c.c1, __dadd_rn(e, w))),
                               __dmul_rn(c.c2, __dadd_rn(n, s))),
                     __dmul_rn(c.c3, __dadd_rn(u, d)));
}
__device__ __forceinline__ double ghost1(const FaceBC& b, double m)
{
    return b.mode == 1 ? __dsub_rn(b.s1, m) : (b.mode == 2 ? __dsub_rn(m, b.s1) : __dadd_rn(m, b.s1));
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

struct TmaSweepParams {
    FrameGeom g;
    double* out;
    const double* in;        // only for nothing but debugging; loads go through the tensor maps
    SweepCoef c;
    FaceSet faces;
    int fuse_fill;
    double* partials;
    long long m_begin, m_end;
    int chunk;
    // fused halo exchange over NVLink peer memory: base (grown cell ii=0,jj=0) of the neighbour's
    // ghost plane that mirrors this slab's first / last valid plane, in the neighbour's `out`
    // buffer; null on physical faces / single GPU
    double* peer_lo_plane;
    double* peer_hi_plane;
};

template <int TX, int TY, int NS>
struct TmaCfg {
    static constexpr int BX = TX + 4;                     // psi box width (2-cell halo each side)
    static constexpr int BY = TY + 2;
    static constexpr int PSI_BYTES = BX * BY * 8;
    static constexpr int PSI_STAGE = (PSI_BYTES + 127) / 128 * 128;
    static constexpr int RHS_BYTES = TX * TY * 8;
    static constexpr int RHS_STAGE = (RHS_BYTES + 127) / 128 * 128;
    static constexpr int SMEM = NS * (PSI_STAGE + RHS_STAGE) + 2 * NS * 8 + 128;
    static constexpr int THREADS = 32 * TY + 32;          // TY consumer warps + 1 producer warp
};

// TX must be 128: a consumer warp covers a 128-cell row as two 64-cell halves, each lane owning
// the pairs (2*lane, 2*lane+1) and (64+2*lane, 65+2*lane).
// MINB: CTAs/SM the register allocation must allow (shared memory permitting)
template <int TY, int NS, bool NORM, int MINB>
__global__ void __launch_bounds__(32 * TY + 32, MINB) jacobi3d_tma_kernel(const __grid_constant__ CUtensorMap tm_psi,
                                                                    const __grid_constant__ CUtensorMap tm_rhs,
                                                                    const TmaSweepParams p)
{
    constexpr int TX = 128;
    using Cfg = TmaCfg<TX, TY, NS>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    double* psi_st = reinterpret_cast<double*>(sm);
    double* rhs_st = reinterpret_cast<double*>(sm + NS * Cfg::PSI_STAGE);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NS * (Cfg::PSI_STAGE + Cfg::RHS_STAGE));
    uint64_t* empty = full + NS;

    const FrameGeom& g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i0 = (long long)blockIdx.x * TX;
    const long long j0 = (long long)blockIdx.y * TY;
    const long long k0 = p.m_begin + (long long)blockIdx.z * p.chunk;
    const long long k1 = k0 + p.chunk < p.m_end ? k0 + p.chunk : p.m_end;
    const int n_planes = (int)(k1 - k0) + 2;              // psi planes k0-1 .. k1

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc = 0.0;
    if (warp == TY) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx_psi = (int)(g.xoff + g.ng + i0 - 2);
            const int cy_psi = (int)(g.ng + j0 - 1);
            const int cx_rhs = (int)(g.xoff + g.ng + i0);
            const int cy_rhs = (int)(g.ng + j0);
            for (int idx = 0; idx < n_planes; ++idx) {
                const int s = idx % NS, use = idx / NS;
                if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
                const long long plane = k0 - 1 + idx;
                const bool with_rhs = idx >= 1 && idx <= n_planes - 2;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(Cfg::PSI_BYTES + (with_rhs ? Cfg::RHS_BYTES : 0)));
                const int cz = (int)(g.kg + plane);
                tma_load_3d(reinterpret_cast<unsigned char*>(psi_st) + s * Cfg::PSI_STAGE, &tm_psi, cx_psi, cy_psi, cz, &full[s]);
                if (with_rhs)
                    tma_load_3d(reinterpret_cast<unsigned char*>(rhs_st) + s * Cfg::RHS_STAGE, &tm_rhs, cx_rhs, cy_rhs, cz, &full[s]);
            }
        }
    } else {
        // ------------------------------ consumer warps ------------------------------
        const int ty = warp;
        const long long j = j0 + ty;
        const bool row_ok = j < g.n[1];
        const long long ia = i0 + 2 * lane, ib = i0 + 64 + 2 * lane;       // first cell of the lane's two pairs
        const bool va0 = row_ok && ia < g.n[0], va1 = row_ok && ia + 1 < g.n[0];
        const bool vb0 = row_ok && ib < g.n[0], vb1 = row_ok && ib + 1 < g.n[0];
        // shared-memory element offsets inside a psi stage: cell (i0+c, j0+r) <-> (r+1)*BX + c + 2
        const int ea = (ty + 1) * Cfg::BX + 2 + 2 * lane;
        const int eb = ea + 64;
        const int ra = ty * TX + 2 * lane;                                 // rhs stage offsets
        const int rb = ra + 64;
        const long long pitch = g.pitch, plane_sz = g.plane;
        double* __restrict__ out = p.out;
        long long off_a = g.at(ia, j, k0, 0);

        auto stage_psi = [&](int idx) { return reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(psi_st) + (idx % NS) * Cfg::PSI_STAGE); };
        auto stage_rhs = [&](int idx) { return reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(rhs_st) + (idx % NS) * Cfg::RHS_STAGE); };

        // plane k0-1: centre only
        mbar_wait(&full[0], 0);
        double2 below_a = lds2(stage_psi(0) + ea), below_b = lds2(stage_psi(0) + eb);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[0]);
        // plane k0: centre
        mbar_wait(&full[1 % NS], (uint32_t)((1 / NS) & 1));
        double2 cur_a = lds2(stage_psi(1) + ea), cur_b = lds2(stage_psi(1) + eb);

        const bool f = p.fuse_fill != 0;
        const bool y_lo = f && j == 0 && p.faces.f[2].mode, y_hi = f && j == g.n[1] - 1 && p.faces.f[3].mode;

        for (int idx = 1; idx <= n_planes - 2; ++idx, off_a += plane_sz) {
            const long long k = k0 + idx - 1;
            // above plane (idx+1)
            mbar_wait(&full[(idx + 1) % NS], (uint32_t)(((idx + 1) / NS) & 1));
            const double* pa = stage_psi(idx + 1);
            const double2 above_a = lds2(pa + ea), above_b = lds2(pa + eb);
            const double* pc = stage_psi(idx);
            const double2 n_a = lds2(pc + ea + Cfg::BX), n_b = lds2(pc + eb + Cfg::BX);
            const double2 s_a = lds2(pc + ea - Cfg::BX), s_b = lds2(pc + eb - Cfg::BX);
            const double* pr = stage_rhs(idx);
            const double2 r_a = lds2(pr + ra), r_b = lds2(pr + rb);
            // E/W by shuffles; the tile-edge lanes read the halo from shared memory
            double w_a0 = __shfl_up_sync(0xffffffffu, cur_a.y, 1);
            double e_a1 = __shfl_down_sync(0xffffffffu, cur_a.x, 1);
            double w_b0 = __shfl_up_sync(0xffffffffu, cur_b.y, 1);
            double e_b1 = __shfl_down_sync(0xffffffffu, cur_b.x, 1);
            const double b0_first = __shfl_sync(0xffffffffu, cur_b.x, 0);
            const double a1_last = __shfl_sync(0xffffffffu, cur_a.y, 31);
            if (lane == 0) { w_a0 = pc[ea - 1]; w_b0 = a1_last; }
            if (lane == 31) { e_a1 = b0_first; e_b1 = pc[eb + 2]; }

            double2 na, nb;
            na.x = upd3(p.c, r_a.x, cur_a.y, w_a0, n_a.x, s_a.x, above_a.x, below_a.x);
            na.y = upd3(p.c, r_a.y, e_a1, cur_a.x, n_a.y, s_a.y, above_a.y, below_a.y);
            nb.x = upd3(p.c, r_b.x, cur_b.y, w_b0, n_b.x, s_b.x, above_b.x, below_b.x);
            nb.y = upd3(p.c, r_b.y, e_b1, cur_b.x, n_b.y, s_b.y, above_b.y, below_b.y);

            // stage idx is no longer needed by this warp (its centre lives on in registers)
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[idx % NS]);

            double* oa = out + off_a;
            double* ob = oa + 64;
            if (va1) *reinterpret_cast<double2*>(oa) = na; else if (va0) oa[0] = na.x;
            if (vb1) *reinterpret_cast<double2*>(ob) = nb; else if (vb0) ob[0] = nb.x;
            // compute + halo "send" in one kernel: the slab's boundary planes are also stored
            // straight into the neighbour GPU's ghost plane (peer-mapped pointer, NVLink)
            if ((k == 0 && p.peer_lo_plane) || (k == g.n[2] - 1 && p.peer_hi_plane)) {
                const long long inpl = g.xoff + (ia + g.ng) + pitch * (j + g.ng);
                if (k == 0 && p.peer_lo_plane) {
                    double* qa = p.peer_lo_plane + inpl;
                    if (va1) *reinterpret_cast<double2*>(qa) = na; else if (va0) qa[0] = na.x;
                    if (vb1) *reinterpret_cast<double2*>(qa + 64) = nb; else if (vb0) qa[64] = nb.x;
                }
                if (k == g.n[2] - 1 && p.peer_hi_plane) {
                    double* qa = p.peer_hi_plane + inpl;
                    if (va1) *reinterpret_cast<double2*>(qa) = na; else if (va0) qa[0] = na.x;
                    if (vb1) *reinterpret_cast<double2*>(qa + 64) = nb; else if (vb0) qa[64] = nb.x;
                }
            }
            if (NORM) {
                if (va0) acc += fabs(__dsub_rn(na.x, cur_a.x));
                if (va1) acc += fabs(__dsub_rn(na.y, cur_a.y));
                if (vb0) acc += fabs(__dsub_rn(nb.x, cur_b.x));
                if (vb1) acc += fabs(__dsub_rn(nb.y, cur_b.y));
            }
            if (f) {
                // x faces
                if (p.faces.f[0].mode && va0 && ia == 0) oa[-1] = ghost1(p.faces.f[0], na.x);
                if (p.faces.f[1].mode) {
                    const long long last = g.n[0] - 1;
                    if (va0 && ia == last) oa[1] = ghost1(p.faces.f[1], na.x);
                    if (va1 && ia + 1 == last) oa[2] = ghost1(p.faces.f[1], na.y);
                    if (vb0 && ib == last) ob[1] = ghost1(p.faces.f[1], nb.x);
                    if (vb1 && ib + 1 == last) ob[2] = ghost1(p.faces.f[1], nb.y);
                }
                // y and z faces: the ghost row/plane of every valid cell of this row
                const bool z_lo = k == 0 && p.faces.f[4].mode, z_hi = k == g.n[2] - 1 && p.faces.f[5].mode;
                if (y_lo | y_hi | z_lo | z_hi) {
                    auto put = [&](double* base, const FaceBC& b) {
                        if (va1) *reinterpret_cast<double2*>(base) = make_double2(ghost1(b, na.x), ghost1(b, na.y));
                        else if (va0) base[0] = ghost1(b, na.x);
                        if (vb1) *reinterpret_cast<double2*>(base + 64) = make_double2(ghost1(b, nb.x), ghost1(b, nb.y));
                        else if (vb0) base[64] = ghost1(b, nb.x);
                    };
                    if (y_lo) put(oa - pitch, p.faces.f[2]);
                    if (y_hi) put(oa + pitch, p.faces.f[3]);
                    if (z_lo) put(oa - plane_sz, p.faces.f[4]);
                    if (z_hi) put(oa + plane_sz, p.faces.f[5]);
                }
            }
            below_a = cur_a; below_b = cur_b;
            cur_a = above_a; cur_b = above_b;
        }
    }
    if (NORM) {
        __shared__ double warp_sums[TY + 1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) warp_sums[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < TY; ++w) v += warp_sums[w];
            p.partials[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)] = v;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        rnmf_set_error("cuTensorMapEncodeTiled not available from the driver");
        return nullptr;
    }
    fn = (EncodeTiledFn)p;
    return fn;
}

int encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by)
{
    EncodeTiledFn enc = get_encoder();
    if (!enc) return -1;
    cuuint64_t dims[3] = { (cuuint64_t)g.pitch, (cuuint64_t)g.G[1], (cuuint64_t)(g.G[2] * g.ncomp) };
    cuuint64_t strides[2] = { (cuuint64_t)g.pitch * 8, (cuuint64_t)g.plane * 8 };
    cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    static int promo = -1;
    if (promo < 0) { const char* e = getenv("RNMF_TMA_L2PROMO"); promo = e ? atoi(e) : 3; }
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rnmf_set_error("cuTensorMapEncodeTiled failed (%d): pitch=%lld G1=%lld G2=%lld box=%dx%d", (int)r, g.pitch, g.G[1], g.G[2], bx, by);
        return -1;
    }
    return 0;
}

template <int TY, int NS, int MINB = 1>
struct Launcher {
    using Cfg = TmaCfg<128, TY, NS>;
    static int plan(const SweepArgs& a, dim3& grid, int& chunk)
    {
        const FrameGeom& g = a.g;
        const long long span = a.k_end - a.k_begin;
        const long long bx = (g.n[0] + 127) / 128, by = (g.n[1] + TY - 1) / TY;
        // Short chunks keep the last partial wave short (measured at 768^3: 258 GLUP/s for 32..64
        // planes, 244 for 192); each chunk re-reads 2 planes, so stay >= 16 unless the xy plane has
        // too few tiles to fill the SMs otherwise.
        long long want = (148LL * 3 * 8 + bx * by - 1) / (bx * by);
        if (want < 1) want = 1;
        long long c = (span + want - 1) / want;
        if (c > 48) c = 48;
        if (c < 16) c = 16;
        if (a.chunk_override > 0) c = a.chunk_override;
        if (c > span) c = span > 0 ? span : 1;
        long long bz = (span + c - 1) / c;
        if (bz > 65535) { bz = 65535; c = (span + bz - 1) / bz; bz = (span + c - 1) / c; }
        chunk = (int)c;
        grid = dim3((unsigned)bx, (unsigned)by, (unsigned)(bz > 0 ? bz : 1));
        return (int)(bx * by * bz);
    }
    static int launch(const SweepArgs& a, cudaStream_t s)
    {
        dim3 grid;
        int chunk;
        plan(a, grid, chunk);
        CUtensorMap tm_psi, tm_rhs;
        if (encode_map(&tm_psi, a.g, a.in, Cfg::BX, Cfg::BY)) return -1;
        if (encode_map(&tm_rhs, a.g, a.rhs, 128, TY)) return -1;
        TmaSweepParams p;
        p.g = a.g; p.out = a.out; p.in = a.in; p.c = a.c; p.faces = a.faces; p.fuse_fill = a.fuse_fill;
        p.partials = a.partials; p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk;
        p.peer_lo_plane = a.peer_lo_plane; p.peer_hi_plane = a.peer_hi_plane;
        static bool attr_set[2] = { false, false };
        if (a.partials) {
            if (!attr_set[1]) {
                if (cudaFuncSetAttribute(jacobi3d_tma_kernel<TY, NS, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
                cudaFuncSetAttribute(jacobi3d_tma_kernel<TY, NS, true, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                attr_set[1] = true;
            }
            jacobi3d_tma_kernel<TY, NS, true, MINB><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(tm_psi, tm_rhs, p);
        } else {
            if (!attr_set[0]) {
                if (cudaFuncSetAttribute(jacobi3d_tma_kernel<TY, NS, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
                cudaFuncSetAttribute(jacobi3d_tma_kernel<TY, NS, false, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                attr_set[0] = true;
            }
            jacobi3d_tma_kernel<TY, NS, false, MINB><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(tm_psi, tm_rhs, p);
        }
        return 1;
    }
};

}  // namespace

// variants: 0/6 = 128x8 tile, 4 stages (default); 5 = 128x8, 5 stages; 7 = 128x16, 4 stages; 8 = 128x4, 6 stages
int tma_sweep_num_partials(const SweepArgs& a)
{
    dim3 grid;
    int chunk;
    switch (a.variant) {
    case 5: return Launcher<8, 5, 2>::plan(a, grid, chunk);
    case 7: return Launcher<16, 4>::plan(a, grid, chunk);
    case 8: return Launcher<4, 6, 3>::plan(a, grid, chunk);
    case 9: return Launcher<8, 3, 2>::plan(a, grid, chunk);
    case 13: return Launcher<16, 3, 1>::plan(a, grid, chunk);
    default: return Launcher<8, 4, 2>::plan(a, grid, chunk);
    }
}

int launch_jacobi_sweep_tma(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    switch (a.variant) {
    case 5: return Launcher<8, 5, 2>::launch(a, s);
    case 7: return Launcher<16, 4>::launch(a, s);
    case 8: return Launcher<4, 6, 3>::launch(a, s);
    case 9: return Launcher<8, 3, 2>::launch(a, s);
    case 13: return Launcher<16, 3, 1>::launch(a, s);
    default: return Launcher<8, 4, 2>::launch(a, s);
    }
}
