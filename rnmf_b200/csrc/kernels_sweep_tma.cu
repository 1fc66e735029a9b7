// kernels_sweep_tma.cu -- K1, Blackwell-native variant of the 3D 7-point Jacobi sweep
// (poisson.rs:80-89 as out-of-place Jacobi; fused ghost fill + update norm as in kernels_sweep.cu).
//
// Data movement is done entirely by TMA (cp.async.bulk.tensor, SASS UTMALDG):
//   * a CTA owns a (128 x TY) xy-tile and marches a chunk of z-planes;
//   * one producer warp streams, plane by plane, the psi tile WITH its xy halo
//     ((128+4) x (TY+2) box: the x halo is two cells wide so the interior starts 16-byte aligned
//     in shared memory) and the rhs tile (128 x TY box) into an NS-stage shared-memory ring,
//     completion signalled on mbarriers (full[]); consumers hand stages back through empty[];
//   * TY consumer warps (one tile row each) keep planes k-1, k, k+1 of their own cells in registers
//     (register z-marching) and read the N/S/E/W neighbours and rhs of plane k from shared memory
//     (conflict-free 128-bit loads for N/S/rhs/centre), then store psi' with 128-bit coalesced
//     global stores.
// No consumer thread ever waits on a global load, so a handful of warps per SM keeps
// (NS-2) planes x ~18 KB x CTAs/SM of HBM traffic in flight.
//
// The consumer loop is unrolled NS times so that every shared-memory address is a compile-time
// offset from two per-thread base registers, and all boundary handling (fused ghost fill, peer
// halo stores) is reduced to a few precomputed per-thread predicates: at 1 kW the B200 power cap
// makes instruction count a first-order cost (profiles/r1_power.md).
//
// Arithmetic: same explicitly rounded sequence as kernels_sweep.cu => bit-identical results.
#include <stdlib.h>

#include <mutex>

#include "tma_util.cuh"

namespace {
using namespace tmau;

struct TmaSweepParams {
    FrameGeom g;
    double* out;
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
    int l2_hints;            // 1: evict_first on rhs loads and psi' stores
    PeerSync ps;
};

template <int TX, int TY, int NS>
struct TmaCfg {
    static constexpr int BX = TX + 4;                     // psi box width (2-cell halo each side)
    static constexpr int BY = TY + 2;
    static constexpr int PSI_BYTES = BX * BY * 8;
    static constexpr int PSI_STAGE = (PSI_BYTES + 127) / 128 * 128;
    static constexpr int RHS_BYTES = TX * TY * 8;
    static constexpr int RHS_STAGE = (RHS_BYTES + 127) / 128 * 128;
    static constexpr int BAR_OFF = NS * (PSI_STAGE + RHS_STAGE);
    static constexpr int SMEM = BAR_OFF + 2 * NS * 8 + 128;
    static constexpr int THREADS = 32 * TY + 32;          // TY consumer warps + 1 producer warp
};

// A consumer warp covers a 128-cell row as two 64-cell halves; lane l owns the pairs
// (2l, 2l+1) and (64+2l, 65+2l).  MINB: CTAs/SM the register allocation must allow.
template <int TY, int NS, bool NORM, int MINB>
__global__ void __launch_bounds__(32 * TY + 32, MINB) jacobi3d_tma_kernel(const __grid_constant__ CUtensorMap tm_psi,
                                                                          const __grid_constant__ CUtensorMap tm_rhs,
                                                                          const __grid_constant__ TmaSweepParams p)
{
    constexpr int TX = 128;
    using Cfg = TmaCfg<TX, TY, NS>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t psi0 = sb, rhs0 = sb + NS * Cfg::PSI_STAGE, full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i0 = (long long)blockIdx.x * TX;
    const long long j0 = (long long)blockIdx.y * TY;
    const unsigned zc = p.ps.enabled ? peer_chunk_order(blockIdx.z, gridDim.z) : blockIdx.z;
    const long long k0 = p.m_begin + (long long)zc * p.chunk;
    const long long k1 = k0 + p.chunk < p.m_end ? k0 + p.chunk : p.m_end;
    const int n_planes = (int)(k1 - k0) + 2;              // psi planes k0-1 .. k1  (plane idx <-> stage idx % NS)
    // fused-halo handshake: CTAs holding the slab's first / last plane
    const bool sync_lo = p.ps.enabled && p.peer_lo_plane && k0 == 0;
    const bool sync_hi = p.ps.enabled && p.peer_hi_plane && k1 == p.g.n[2];

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, TY);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
    }
    __syncthreads();

    double acc = 0.0;
    if (warp == TY) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx_psi = (int)(p.g.xoff + p.g.ng + i0 - 2);
            const int cy_psi = (int)(p.g.ng + j0 - 1);
            const int cx_rhs = cx_psi + 2, cy_rhs = cy_psi + 1;
            int cz = (int)(p.g.kg + k0 - 1);
            const bool hints = p.l2_hints != 0;
            const uint64_t pol = policy_evict_first();
            uint32_t wrap = 0;
            for (int base = 0; base < n_planes; base += NS, ++wrap) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int idx = base + s;
                    if (idx < n_planes) {
                        if (wrap > 0) mbar_wait(empty0 + 8 * s, (wrap - 1) & 1);
                        const bool with_rhs = idx >= 1 && idx <= n_planes - 2;
                        mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(Cfg::PSI_BYTES + (with_rhs ? Cfg::RHS_BYTES : 0)));
                        tma_load_3d(psi0 + s * Cfg::PSI_STAGE, &tm_psi, cx_psi, cy_psi, cz, full0 + 8 * s);
                        if (with_rhs) {
                            if (hints) tma_load_3d_hint(rhs0 + s * Cfg::RHS_STAGE, &tm_rhs, cx_rhs, cy_rhs, cz, full0 + 8 * s, pol);
                            else tma_load_3d(rhs0 + s * Cfg::RHS_STAGE, &tm_rhs, cx_rhs, cy_rhs, cz, full0 + 8 * s);
                        }
                        ++cz;
                    }
                }
            }
        }
    } else {
        // ------------------------------ consumer warps ------------------------------
        const long long j = j0 + warp;
        const long long n0 = p.g.n[0], n1 = p.g.n[1], n2 = p.g.n[2];
        const bool row_ok = j < n1;
        const long long ia = i0 + 2 * lane, ib = ia + 64;
        const bool va0 = row_ok && ia < n0, va1 = row_ok && ia + 1 < n0;
        const bool vb0 = row_ok && ib < n0, vb1 = row_ok && ib + 1 < n0;
        const bool full = vb1;                                            // all four cells valid
        // per-thread shared-memory byte offsets: cell (i0+c, j0+r) of a psi stage is at ((r+1)*BX + c + 2)*8
        const uint32_t pa = psi0 + (uint32_t)(((warp + 1) * Cfg::BX + 2 + 2 * lane) * 8);
        const uint32_t ra = rhs0 + (uint32_t)((warp * TX + 2 * lane) * 8);
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2, c3 = p.c.c3;
        const long long pitch = p.g.pitch, plane_sz = p.g.plane;
        double* oa = p.out + p.g.at(ia, j, k0, 0);

        const bool hints = p.l2_hints != 0;
        const uint64_t pol = policy_evict_first();
        // boundary work reduced to per-thread predicates
        const bool fuse = p.fuse_fill != 0;
        const Face1 fx0 = make_face1(p.faces.f[0], fuse), fx1 = make_face1(p.faces.f[1], fuse);
        const Face1 fy0 = make_face1(p.faces.f[2], fuse), fy1 = make_face1(p.faces.f[3], fuse);
        const bool do_xlo = fx0.on && va0 && ia == 0;
        const long long last = n0 - 1;
        const int xhi_sel = !fx1.on ? -1 : (va0 && ia == last) ? 0 : (va1 && ia + 1 == last) ? 1 : (vb0 && ib == last) ? 2 : (vb1 && ib + 1 == last) ? 3 : -1;
        const bool do_ylo = fy0.on && j == 0, do_yhi = fy1.on && j == n1 - 1;
        const bool any_xy = do_xlo || xhi_sel >= 0 || do_ylo || do_yhi;
        // z faces / peer halo planes only concern the first and last plane of the slab
        const bool z_lo_here = k0 == 0 && ((fuse && p.faces.f[4].mode) || p.peer_lo_plane);
        const bool z_hi_here = k1 == n2 && ((fuse && p.faces.f[5].mode) || p.peer_hi_plane);
        const int idx_zlo = z_lo_here ? 1 : -1, idx_zhi = z_hi_here ? n_planes - 2 : -1;

        // plane k0-1: centre only; plane k0: centre
        mbar_wait(full0, 0);
        double2 below_a = lds2(pa), below_b = lds2(pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8 * (1 % NS), (uint32_t)(1 / NS));
        double2 cur_a = lds2(pa + (1 % NS) * Cfg::PSI_STAGE), cur_b = lds2(pa + (1 % NS) * Cfg::PSI_STAGE + 512);

        uint32_t wrap = 0;
        for (int base = 0; base < n_planes - 1; base += NS, ++wrap) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int idx = base + s;                       // plane being updated (stage s)
                if (idx >= 1 && idx <= n_planes - 2) {
                    const int sn = (s + 1) % NS;                // stage of plane idx+1 (constant after unrolling)
                    mbar_wait(full0 + 8 * sn, (wrap + (s + 1 == NS ? 1u : 0u)) & 1u);
                    const uint32_t pn = pa + sn * Cfg::PSI_STAGE, pc = pa + s * Cfg::PSI_STAGE, pr = ra + s * Cfg::RHS_STAGE;
                    const double2 above_a = lds2(pn), above_b = lds2(pn + 512);
                    const double2 n_a = lds2(pc + Cfg::BX * 8), n_b = lds2(pc + Cfg::BX * 8 + 512);
                    const double2 s_a = lds2(pc - Cfg::BX * 8), s_b = lds2(pc - Cfg::BX * 8 + 512);
                    const double2 r_a = lds2(pr), r_b = lds2(pr + 512);
                    // the row is contiguous in shared memory (halo | 64 cells | 64 cells | halo), so the
                    // E/W neighbours of the pair ends are plain 8-byte loads for every lane
                    const double w_a0 = lds1(pc - 8), e_a1 = lds1(pc + 16);
                    const double w_b0 = lds1(pc + 512 - 8), e_b1 = lds1(pc + 512 + 16);

                    double2 na, nb;
                    na.x = upd3(c0, c1, c2, c3, r_a.x, cur_a.y, w_a0, n_a.x, s_a.x, above_a.x, below_a.x);
                    na.y = upd3(c0, c1, c2, c3, r_a.y, e_a1, cur_a.x, n_a.y, s_a.y, above_a.y, below_a.y);
                    nb.x = upd3(c0, c1, c2, c3, r_b.x, cur_b.y, w_b0, n_b.x, s_b.x, above_b.x, below_b.x);
                    nb.y = upd3(c0, c1, c2, c3, r_b.y, e_b1, cur_b.x, n_b.y, s_b.y, above_b.y, below_b.y);

                    // stage s is no longer needed by this warp (its centre lives on in registers)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8 * s);

                    if (full && hints) { st2_hint(oa, na, pol); st2_hint(oa + 64, nb, pol); }
                    else store4(oa, na, nb, full, va0, va1, vb0, vb1);
                    if (NORM) {
                        double t = 0.0;
                        if (va0) t += fabs(__dsub_rn(na.x, cur_a.x));
                        if (va1) t += fabs(__dsub_rn(na.y, cur_a.y));
                        if (vb0) t += fabs(__dsub_rn(nb.x, cur_b.x));
                        if (vb1) t += fabs(__dsub_rn(nb.y, cur_b.y));
                        acc += t;
                    }
                    if (any_xy) {
                        if (do_xlo) oa[-1] = ghost1(fx0, na.x);
                        if (xhi_sel >= 0) {
                            const double m = xhi_sel == 0 ? na.x : xhi_sel == 1 ? na.y : xhi_sel == 2 ? nb.x : nb.y;
                            oa[(xhi_sel >= 2 ? 64 : 0) + (xhi_sel & 1) + 1] = ghost1(fx1, m);
                        }
                        if (do_ylo)
                            store4(oa - pitch, make_double2(ghost1(fy0, na.x), ghost1(fy0, na.y)),
                                   make_double2(ghost1(fy0, nb.x), ghost1(fy0, nb.y)), full, va0, va1, vb0, vb1);
                        if (do_yhi)
                            store4(oa + pitch, make_double2(ghost1(fy1, na.x), ghost1(fy1, na.y)),
                                   make_double2(ghost1(fy1, nb.x), ghost1(fy1, nb.y)), full, va0, va1, vb0, vb1);
                    }
                    if (idx == idx_zlo || idx == idx_zhi) {
                        for (int pass = 0; pass < 2; ++pass) {          // a 1-plane slab does both
                            const bool is_lo = pass == 0;
                            if (is_lo ? idx != idx_zlo : idx != idx_zhi) continue;
                            const Face1 fz = is_lo ? make_face1(p.faces.f[4], fuse) : make_face1(p.faces.f[5], fuse);
                            if (fz.on)
                                store4(is_lo ? oa - plane_sz : oa + plane_sz, make_double2(ghost1(fz, na.x), ghost1(fz, na.y)),
                                       make_double2(ghost1(fz, nb.x), ghost1(fz, nb.y)), full, va0, va1, vb0, vb1);
                            // compute + halo "send" in one kernel: the slab's boundary plane is also stored
                            // straight into the neighbour GPU's ghost plane (peer-mapped pointer, NVLink)
                            double* peer = is_lo ? p.peer_lo_plane : p.peer_hi_plane;
                            if (peer) store4(peer + p.g.xoff + (ia + p.g.ng) + pitch * (j + p.g.ng), na, nb, full, va0, va1, vb0, vb1);
                        }
                    }
                    below_a = cur_a; below_b = cur_b;
                    cur_a = above_a; cur_b = above_b;
                    oa += plane_sz;
                }
            }
        }
    }
    if (sync_lo || sync_hi) {
        __syncthreads();                                   // every peer store of this CTA has been issued
        if (threadIdx.x == 0) {
            if (sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.ps.sig_lo, p.ps.epoch + 1);
            if (sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.ps.sig_hi, p.ps.epoch + 1);
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

// tensor maps depend only on (buffer, geometry, box): a frame ping-pongs between two buffers, so a
// small cache removes the encoder from the per-sweep path (it matters for small, launch-bound grids)
struct MapKey { const double* base; long long pitch, g1, g2n; int bx, by; };
struct MapSlot { MapKey k; CUtensorMap tm; bool used; };
static MapSlot g_map_cache[32];
static int g_map_next = 0;
static std::mutex g_map_mutex;      // contexts on different host threads share the cache

int encode_map_uncached(CUtensorMap* tm, const double* base, long long pitch, long long g1, long long planes, long long plane_stride, int bx, int by);

// deep > 0: the buffer has `deep` extra planes in front of the z-lo ghost plane and behind the z-hi ghost plane (capi.cu)
int encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by, int deep)
{
    const double* b0 = base - (long long)deep * g.plane;
    const long long planes = g.G[2] * g.ncomp + 2 * deep;
    const MapKey k = { b0, g.pitch, g.G[1], planes, bx, by };
    std::lock_guard<std::mutex> lock(g_map_mutex);
    for (auto& sl : g_map_cache)
        if (sl.used && sl.k.base == k.base && sl.k.pitch == k.pitch && sl.k.g1 == k.g1 && sl.k.g2n == k.g2n && sl.k.bx == k.bx && sl.k.by == k.by) {
            *tm = sl.tm;
            return 0;
        }
    if (encode_map_uncached(tm, b0, g.pitch, g.G[1], planes, g.plane, bx, by)) return -1;
    MapSlot& sl = g_map_cache[g_map_next];
    g_map_next = (g_map_next + 1) % 32;
    sl.k = k; sl.tm = *tm; sl.used = true;
    return 0;
}

int encode_map_uncached(CUtensorMap* tm, const double* base, long long pitch, long long g1, long long planes, long long plane_stride, int bx, int by)
{
    EncodeTiledFn enc = get_encoder();
    if (!enc) return -1;
    cuuint64_t dims[3] = { (cuuint64_t)pitch, (cuuint64_t)g1, (cuuint64_t)planes };
    cuuint64_t strides[2] = { (cuuint64_t)pitch * 8, (cuuint64_t)plane_stride * 8 };
    cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    static int promo = -1;
    if (promo < 0) { const char* e = getenv("RNMF_TMA_L2PROMO"); promo = e ? atoi(e) : 2; }   // 128 B: same throughput as 256 B with 20 % fewer L2 read sectors (profiles/r1_power.md)
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rnmf_set_error("cuTensorMapEncodeTiled failed (%d): pitch=%lld G1=%lld planes=%lld box=%dx%d", (int)r, pitch, g1, planes, bx, by);
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
        // planes, 244 for 192); each chunk re-reads 2 planes.  Small frames get shorter chunks (>= 4
        // planes) so that they still spread over the SMs.
        long long want = (148LL * 3 * 8 + bx * by - 1) / (bx * by);
        if (want < 1) want = 1;
        long long c = (span + want - 1) / want;
        if (c > 48) c = 48;
        if (c < 4) c = 4;
        if (a.chunk_override > 0) c = a.chunk_override;
        if (c > span) c = span > 0 ? span : 1;
        long long bz = (span + c - 1) / c;
        if (bz > 65535) { bz = 65535; c = (span + bz - 1) / bz; bz = (span + c - 1) / c; }
        chunk = (int)c;
        grid = dim3((unsigned)bx, (unsigned)by, (unsigned)(bz > 0 ? bz : 1));
        return (int)(bx * by * bz);
    }
    template <bool NORM>
    static int launch_one(const dim3& grid, const CUtensorMap& tm_psi, const CUtensorMap& tm_rhs, const TmaSweepParams& p, cudaStream_t s)
    {
        static PerDeviceOnce attr_set;
        if (!attr_set.done()) {
            if (cudaFuncSetAttribute(jacobi3d_tma_kernel<TY, NS, NORM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
            cudaFuncSetAttribute(jacobi3d_tma_kernel<TY, NS, NORM, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            attr_set.mark();
        }
        jacobi3d_tma_kernel<TY, NS, NORM, MINB><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(tm_psi, tm_rhs, p);
        return 1;
    }
    static int launch(const SweepArgs& a, cudaStream_t s)
    {
        dim3 grid;
        int chunk;
        plan(a, grid, chunk);
        CUtensorMap tm_psi, tm_rhs;
        if (encode_map(&tm_psi, a.g, a.in, Cfg::BX, Cfg::BY, 0)) return -1;
        if (encode_map(&tm_rhs, a.g, a.rhs, 128, TY, 0)) return -1;
        TmaSweepParams p;
        p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces; p.fuse_fill = a.fuse_fill;
        p.partials = a.partials; p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk;
        p.peer_lo_plane = a.peer_lo_plane; p.peer_hi_plane = a.peer_hi_plane;
        static int hints = -1;
        if (hints < 0) { const char* e = getenv("RNMF_L2_HINTS"); hints = e ? atoi(e) : 0; }
        p.l2_hints = hints;
        p.ps = a.ps;
        return a.partials ? launch_one<true>(grid, tm_psi, tm_rhs, p, s) : launch_one<false>(grid, tm_psi, tm_rhs, p, s);
    }
};

}  // namespace

int tma3d_encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by) { return encode_map(tm, g, base, bx, by, 0); }
int tma3d_encode_map_deep(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by, int deep) { return encode_map(tm, g, base, bx, by, deep); }

// 128x8 tile, 4 stages, 2 CTAs per SM (the other tile shapes tried in round 1 -- 128x8x5, 128x16x4, 128x4x6, 128x8x3,
// 128x16x3 -- were within 2 % or slower and are gone)
using DefaultLauncher = Launcher<8, 4, 2>;

int tma_sweep_num_partials(const SweepArgs& a)
{
    dim3 grid;
    int chunk;
    return DefaultLauncher::plan(a, grid, chunk);
}

int tma_sweep_boundary_ctas(const SweepArgs& a)
{
    dim3 grid;
    int chunk;
    DefaultLauncher::plan(a, grid, chunk);
    return (int)(grid.x * grid.y);
}

int launch_jacobi_sweep_tma(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    return DefaultLauncher::launch(a, s);
}
