// kernels_sweep2d_tma.cu -- K1 for 2D frames: the 5-point Jacobi sweep of solve_poisson
// (poisson.rs:80-89, out of place) with the fused ghost fill and update norm, TMA/mbarrier variant.
//
// Same structure as the 3D kernel (kernels_sweep_tma.cu) with "plane" -> "row":
//   * a CTA owns a 512-column strip and marches a chunk of rows in y;
//   * the producer warp streams one row per stage: the psi row segment with a 2-cell x halo
//     (512+4 doubles, three TMA boxes 256+256+4) and the rhs segment (two boxes), NS = 8 stages;
//   * four consumer warps own 128 cells each (lane l: pairs 2l,2l+1 and 64+2l,65+2l) and keep rows
//     j-1, j, j+1 of their cells in registers: the N/S neighbours of the stencil ARE the marching
//     direction, so shared memory only serves the next row's centre, rhs and the E/W neighbours.
// Chunks of <= 64 rows per CTA (measured best; see plan2d).
#include <cuda.h>

#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ double2 lds2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds1(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
// association exactly as poisson.rs:83-85
__device__ __forceinline__ double upd2(double c0, double c1, double c2, double r, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, __dadd_rn(e, w))), __dmul_rn(c2, __dadd_rn(n, s)));
}
struct Face1 {
    double add;
    bool neg;
    bool on;
};
__device__ __forceinline__ Face1 make_face1(const FaceBC& b, bool fuse)
{
    Face1 f;
    f.on = fuse && b.mode != 0;
    f.neg = b.mode == 1;
    f.add = b.mode == 2 ? -b.s1 : b.s1;
    return f;
}
__device__ __forceinline__ double ghost1(const Face1& f, double m) { return __dadd_rn(f.neg ? -m : m, f.add); }
__device__ __forceinline__ void store4(double* a, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
    if (full) {
        *reinterpret_cast<double2*>(a) = va;
        *reinterpret_cast<double2*>(a + 64) = vb;
    } else {
        if (a1) *reinterpret_cast<double2*>(a) = va; else if (a0) a[0] = va.x;
        if (b1) *reinterpret_cast<double2*>(a + 64) = vb; else if (b0) a[64] = vb.x;
    }
}

struct Tma2dParams {
    FrameGeom g;
    double* out;
    SweepCoef c;
    FaceSet faces;
    int fuse_fill;
    double* partials;
    long long m_begin, m_end;   // valid rows to update
    int chunk;
    // fused halo exchange over NVLink peer memory (y-slabs): base (grown cell ii=0) of the
    // neighbour's ghost row mirroring this slab's first / last valid row, in its `out` buffer
    double* peer_lo_row;
    double* peer_hi_row;
    PeerSync ps;
};

constexpr int NW = 4;                                  // consumer warps: 128 columns each
constexpr int TXC = 128 * NW;                          // 512 columns per CTA
constexpr int PSI_BYTES = (TXC + 4) * 8;               // 4128
constexpr int PSI_STAGE = (PSI_BYTES + 127) / 128 * 128;
constexpr int RHS_BYTES = TXC * 8;
constexpr int RHS_STAGE = RHS_BYTES;

template <int NS>
struct Cfg2d {
    static constexpr int BAR_OFF = NS * (PSI_STAGE + RHS_STAGE);
    static constexpr int SMEM = BAR_OFF + 2 * NS * 8 + 128;
    static constexpr int THREADS = 32 * NW + 32;
};

template <int NS, bool NORM>
__global__ void __launch_bounds__(32 * NW + 32, 3) jacobi2d_tma_kernel(const __grid_constant__ CUtensorMap tm_psi256,
                                                                       const __grid_constant__ CUtensorMap tm_psi4,
                                                                       const __grid_constant__ CUtensorMap tm_rhs256,
                                                                       const __grid_constant__ Tma2dParams p)
{
    using Cfg = Cfg2d<NS>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t psi0 = sb, rhs0 = sb + NS * PSI_STAGE, full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i0 = (long long)blockIdx.x * TXC;
    const unsigned yc = p.ps.enabled ? peer_chunk_order(blockIdx.y, gridDim.y) : blockIdx.y;
    const long long j0 = p.m_begin + (long long)yc * p.chunk;
    const long long j1 = j0 + p.chunk < p.m_end ? j0 + p.chunk : p.m_end;
    const int n_rows = (int)(j1 - j0) + 2;               // psi rows j0-1 .. j1   (row idx <-> stage idx % NS)
    const bool sync_lo = p.ps.enabled && p.peer_lo_row && j0 == 0;
    const bool sync_hi = p.ps.enabled && p.peer_hi_row && j1 == p.g.n[1];

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
    }
    __syncthreads();

    double acc = 0.0;
    if (warp == NW) {
        if (lane == 0) {
            const int cx = (int)(p.g.xoff + p.g.ng + i0 - 2);
            int cy = (int)(p.g.ng + j0 - 1);
            uint32_t wrap = 0;
            for (int base = 0; base < n_rows; base += NS, ++wrap) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int idx = base + s;
                    if (idx < n_rows) {
                        if (wrap > 0) mbar_wait(empty0 + 8 * s, (wrap - 1) & 1);
                        const bool with_rhs = idx >= 1 && idx <= n_rows - 2;
                        mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(PSI_BYTES + (with_rhs ? RHS_BYTES : 0)));
                        const uint32_t d = psi0 + s * PSI_STAGE;
                        tma_load_2d(d, &tm_psi256, cx, cy, full0 + 8 * s);
                        tma_load_2d(d + 2048, &tm_psi256, cx + 256, cy, full0 + 8 * s);
                        tma_load_2d(d + 4096, &tm_psi4, cx + 512, cy, full0 + 8 * s);
                        if (with_rhs) {
                            const uint32_t r = rhs0 + s * RHS_STAGE;
                            tma_load_2d(r, &tm_rhs256, cx + 2, cy, full0 + 8 * s);
                            tma_load_2d(r + 2048, &tm_rhs256, cx + 2 + 256, cy, full0 + 8 * s);
                        }
                        ++cy;
                    }
                }
            }
        }
    } else {
        const long long n0 = p.g.n[0], n1 = p.g.n[1];
        const long long ia = i0 + 128 * warp + 2 * lane, ib = ia + 64;
        const bool va0 = ia < n0, va1 = ia + 1 < n0, vb0 = ib < n0, vb1 = ib + 1 < n0;
        const bool full = vb1;
        const uint32_t pa = psi0 + (uint32_t)((128 * warp + 2 + 2 * lane) * 8);
        const uint32_t ra = rhs0 + (uint32_t)((128 * warp + 2 * lane) * 8);
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2;
        const long long pitch = p.g.pitch;
        double* oa = p.out + p.g.at(ia, j0, 0, 0);

        const bool fuse = p.fuse_fill != 0;
        const Face1 fx0 = make_face1(p.faces.f[0], fuse), fx1 = make_face1(p.faces.f[1], fuse);
        const bool do_xlo = fx0.on && va0 && ia == 0;
        const long long last = n0 - 1;
        const int xhi_sel = !fx1.on ? -1 : (va0 && ia == last) ? 0 : (va1 && ia + 1 == last) ? 1 : (vb0 && ib == last) ? 2 : (vb1 && ib + 1 == last) ? 3 : -1;
        const bool any_x = do_xlo || xhi_sel >= 0;
        const int idx_ylo = (j0 == 0 && ((fuse && p.faces.f[2].mode) || p.peer_lo_row)) ? 1 : -1;
        const int idx_yhi = (j1 == n1 && ((fuse && p.faces.f[3].mode) || p.peer_hi_row)) ? n_rows - 2 : -1;
        const long long in_row = p.g.xoff + (ia + p.g.ng);

        mbar_wait(full0, 0);
        double2 below_a = lds2(pa), below_b = lds2(pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8 * (1 % NS), (uint32_t)(1 / NS));
        double2 cur_a = lds2(pa + (1 % NS) * PSI_STAGE), cur_b = lds2(pa + (1 % NS) * PSI_STAGE + 512);

        uint32_t wrap = 0;
        for (int base = 0; base < n_rows - 1; base += NS, ++wrap) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int idx = base + s;
                if (idx >= 1 && idx <= n_rows - 2) {
                    const int sn = (s + 1) % NS;
                    mbar_wait(full0 + 8 * sn, (wrap + (s + 1 == NS ? 1u : 0u)) & 1u);
                    const uint32_t pn = pa + sn * PSI_STAGE, pc = pa + s * PSI_STAGE, pr = ra + s * RHS_STAGE;
                    const double2 above_a = lds2(pn), above_b = lds2(pn + 512);
                    const double2 r_a = lds2(pr), r_b = lds2(pr + 512);
                    const double w_a0 = lds1(pc - 8), e_a1 = lds1(pc + 16);
                    const double w_b0 = lds1(pc + 512 - 8), e_b1 = lds1(pc + 512 + 16);
                    double2 na, nb;
                    na.x = upd2(c0, c1, c2, r_a.x, cur_a.y, w_a0, above_a.x, below_a.x);
                    na.y = upd2(c0, c1, c2, r_a.y, e_a1, cur_a.x, above_a.y, below_a.y);
                    nb.x = upd2(c0, c1, c2, r_b.x, cur_b.y, w_b0, above_b.x, below_b.x);
                    nb.y = upd2(c0, c1, c2, r_b.y, e_b1, cur_b.x, above_b.y, below_b.y);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8 * s);

                    store4(oa, na, nb, full, va0, va1, vb0, vb1);
                    if (NORM) {
                        double t = 0.0;
                        if (va0) t += fabs(__dsub_rn(na.x, cur_a.x));
                        if (va1) t += fabs(__dsub_rn(na.y, cur_a.y));
                        if (vb0) t += fabs(__dsub_rn(nb.x, cur_b.x));
                        if (vb1) t += fabs(__dsub_rn(nb.y, cur_b.y));
                        acc += t;
                    }
                    if (any_x) {
                        if (do_xlo) oa[-1] = ghost1(fx0, na.x);
                        if (xhi_sel >= 0) {
                            const double m = xhi_sel == 0 ? na.x : xhi_sel == 1 ? na.y : xhi_sel == 2 ? nb.x : nb.y;
                            oa[(xhi_sel >= 2 ? 64 : 0) + (xhi_sel & 1) + 1] = ghost1(fx1, m);
                        }
                    }
                    if (idx == idx_ylo) {
                        const Face1 fy = make_face1(p.faces.f[2], fuse);
                        if (fy.on)
                            store4(oa - pitch, make_double2(ghost1(fy, na.x), ghost1(fy, na.y)),
                                   make_double2(ghost1(fy, nb.x), ghost1(fy, nb.y)), full, va0, va1, vb0, vb1);
                        // compute + halo "send" in one kernel: boundary row stored into the neighbour GPU's ghost row
                        if (p.peer_lo_row) store4(p.peer_lo_row + in_row, na, nb, full, va0, va1, vb0, vb1);
                    }
                    if (idx == idx_yhi) {
                        const Face1 fy = make_face1(p.faces.f[3], fuse);
                        if (fy.on)
                            store4(oa + pitch, make_double2(ghost1(fy, na.x), ghost1(fy, na.y)),
                                   make_double2(ghost1(fy, nb.x), ghost1(fy, nb.y)), full, va0, va1, vb0, vb1);
                        if (p.peer_hi_row) store4(p.peer_hi_row + in_row, na, nb, full, va0, va1, vb0, vb1);
                    }
                    below_a = cur_a; below_b = cur_b;
                    cur_a = above_a; cur_b = above_b;
                    oa += pitch;
                }
            }
        }
    }
    if (sync_lo || sync_hi) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.ps.sig_lo, p.ps.epoch + 1);
            if (sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.ps.sig_hi, p.ps.epoch + 1);
        }
    }
    if (NORM) {
        __shared__ double warp_sums[NW + 1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) warp_sums[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < NW; ++w) v += warp_sums[w];
            p.partials[blockIdx.x + gridDim.x * blockIdx.y] = v;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder2d()
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

struct MapKey2 { const double* base; long long pitch, g1n; int bx; };
struct MapSlot2 { MapKey2 k; CUtensorMap tm; bool used; };
static MapSlot2 g_map_cache2[32];
static int g_map_next2 = 0;
static std::mutex g_map_mutex2;     // contexts on different host threads share the cache

int encode_map2d_uncached(CUtensorMap* tm, const FrameGeom& g, const double* base, long long rows, int bx);

// cached: see kernels_sweep_tma.cu.  deep > 0: the buffer has `deep` extra rows in front of the y-lo ghost row and behind the
// y-hi ghost row (capi.cu), and the tensor covers them
int encode_map2d(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int deep)
{
    const double* b0 = base - (long long)deep * g.pitch;
    const long long rows = g.G[1] * g.ncomp + 2 * deep;
    const MapKey2 k = { b0, g.pitch, rows, bx };
    std::lock_guard<std::mutex> lock(g_map_mutex2);
    for (auto& sl : g_map_cache2)
        if (sl.used && sl.k.base == k.base && sl.k.pitch == k.pitch && sl.k.g1n == k.g1n && sl.k.bx == k.bx) {
            *tm = sl.tm;
            return 0;
        }
    if (encode_map2d_uncached(tm, g, b0, rows, bx)) return -1;
    MapSlot2& sl = g_map_cache2[g_map_next2];
    g_map_next2 = (g_map_next2 + 1) % 32;
    sl.k = k; sl.tm = *tm; sl.used = true;
    return 0;
}

int encode_map2d_uncached(CUtensorMap* tm, const FrameGeom& g, const double* base, long long rows, int bx)
{
    EncodeTiledFn enc = get_encoder2d();
    if (!enc) return -1;
    cuuint64_t dims[2] = { (cuuint64_t)g.pitch, (cuuint64_t)rows };
    cuuint64_t strides[1] = { (cuuint64_t)g.pitch * 8 };
    cuuint32_t box[2] = { (cuuint32_t)bx, 1 };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rnmf_set_error("cuTensorMapEncodeTiled (2D) failed (%d): pitch=%lld G1=%lld box=%d", (int)r, g.pitch, g.G[1], bx);
        return -1;
    }
    return 0;
}

constexpr int kNS2d = 8;

int plan2d(const SweepArgs& a, dim3& grid, int& chunk)
{
    const FrameGeom& g = a.g;
    const long long span = a.k_end - a.k_begin;
    const long long bx = (g.n[0] + TXC - 1) / TXC;
    // 64-row chunks: measured best at 8192^2 (267 GLUP/s sustained; 264 / 260 / 249 / 223 for
    // 96 / 128 / 304 / 1024 rows): many waves of short CTAs keep HBM busier than one wave of long
    // ones although each chunk re-reads 2 rows.  Small frames get shorter chunks (>= 4 rows) so
    // that they still spread over the SMs.
    long long rows_chunks = (148LL * 3 * 2 + bx - 1) / bx;
    long long c = (span + rows_chunks - 1) / rows_chunks;
    if (c > 64) c = 64;
    if (c < 4) c = 4;
    if (a.chunk_override > 0) c = a.chunk_override;
    if (c > span) c = span > 0 ? span : 1;
    long long by = (span + c - 1) / c;
    if (by > 65535) { by = 65535; c = (span + by - 1) / by; by = (span + c - 1) / c; }
    chunk = (int)c;
    grid = dim3((unsigned)bx, (unsigned)(by > 0 ? by : 1), 1);
    return (int)(bx * by);
}

template <bool NORM>
int launch2d_one(const dim3& grid, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const Tma2dParams& p, cudaStream_t s)
{
    using Cfg = Cfg2d<kNS2d>;
    static PerDeviceOnce attr_set;
    if (!attr_set.done()) {
        if (cudaFuncSetAttribute(jacobi2d_tma_kernel<kNS2d, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
        cudaFuncSetAttribute(jacobi2d_tma_kernel<kNS2d, NORM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set.mark();
    }
    jacobi2d_tma_kernel<kNS2d, NORM><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(a, b, c, p);
    return 1;
}

}  // namespace

int tma2d_encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx) { return encode_map2d(tm, g, base, bx, 0); }
int tma2d_encode_map_deep(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int deep) { return encode_map2d(tm, g, base, bx, deep); }

int tma2d_sweep_num_partials(const SweepArgs& a)
{
    dim3 grid;
    int chunk;
    return plan2d(a, grid, chunk);
}

int tma2d_sweep_boundary_ctas(const SweepArgs& a)
{
    dim3 grid;
    int chunk;
    plan2d(a, grid, chunk);
    return (int)grid.x;
}

int launch_jacobi_sweep_tma2d(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    dim3 grid;
    int chunk;
    plan2d(a, grid, chunk);
    CUtensorMap tm_psi256, tm_psi4, tm_rhs256;
    if (encode_map2d(&tm_psi256, a.g, a.in, 256, 0)) return -1;
    if (encode_map2d(&tm_psi4, a.g, a.in, 4, 0)) return -1;
    if (encode_map2d(&tm_rhs256, a.g, a.rhs, 256, 0)) return -1;
    Tma2dParams p;
    p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces; p.fuse_fill = a.fuse_fill;
    p.partials = a.partials; p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk;
    p.peer_lo_row = a.peer_lo_plane; p.peer_hi_row = a.peer_hi_plane;
    p.ps = a.ps;
    return a.partials ? launch2d_one<true>(grid, tm_psi256, tm_psi4, tm_rhs256, p, s)
                      : launch2d_one<false>(grid, tm_psi256, tm_psi4, tm_rhs256, p, s);
}
