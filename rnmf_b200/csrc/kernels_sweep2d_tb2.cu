// kernels_sweep2d_tb2.cu -- K1 for 2D frames, temporally blocked: TWO 5-point Jacobi sweeps (each
// followed by the ng=1 ghost fill, poisson.rs:80-89 + dataframe.rs:289-355) per pass over HBM.
// The 2D analogue of kernels_sweep_tb2.cu with "plane" -> "row":
//   * a CTA owns a 512-column strip and marches a chunk of rows in y; the producer warp streams one
//     row of psi (u0) and one of rhs per stage, both 512+4 doubles wide (TMA boxes 256+256+4);
//   * four consumer warps own 128 cells each (lane l: pairs 2l,2l+1 and 64+2l,65+2l).  The N/S
//     neighbours of the stencil ARE the marching direction, so both levels keep a three-row window of
//     their own cells in registers: level 1 over u0, level 2 over u1 = sweep(u0).  Shared memory
//     serves the next u0 row, rhs, and the E/W neighbours of both levels; u1 rows live in a two-slot
//     ring (plus the x-face ghosts fill_bc would give u1), the y-face ghosts of u1 are formed in
//     registers;
//   * a fifth consumer warp computes u1 in the two halo columns i0-1 and i0+512 (two lanes): the
//     redundant work is 2/512 per row, the chunk re-reads 4 rows;
//   * one wait phase per row; level 2's x part (c0*rhs + c1*(E+W)) is formed before level 1, the
//     y term c2*(N+S) is added once u1 of the new row exists: the reference's association.
// u2 is bit-identical to two single sweeps.  Single GPU only (y-slabs use the single-sweep kernel).
#include <stdlib.h>

#include "tma_util.cuh"

namespace {
using namespace tmau;

struct Tb2dParams {
    FrameGeom g;
    double* out;
    SweepCoef c;
    FaceSet faces;
    long long m_begin, m_end;   // output rows
    int chunk;
};

constexpr int NW2 = 4;                         // consumer warps with 128 columns each
constexpr int TXC2 = 128 * NW2;                // 512 columns per CTA
constexpr int ROW2 = TXC2 + 4;                 // a staged row: columns i0-2 .. i0+513
constexpr int ROW2_BYTES = ROW2 * 8;           // 4128
constexpr int ROW2_STAGE = (ROW2_BYTES + 127) / 128 * 128;

template <int NS>
struct Cfg2 {
    static constexpr int RHS_OFF = NS * ROW2_STAGE;
    static constexpr int U1_OFF = RHS_OFF + NS * ROW2_STAGE;
    static constexpr int BAR_OFF = U1_OFF + 2 * ROW2_STAGE;
    static constexpr int SMEM = BAR_OFF + (2 * NS + 2) * 8 + 128;
    static constexpr int NCONS = NW2 + 1;      // + the halo-column warp
    static constexpr int THREADS = 32 * (NW2 + 2);
};

struct Row2K {
    uint32_t pa, ra, ua;
    uint32_t full0, empty0, u1full0;
    double c0, c1, c2;
    bool full, va0, va1, vb0, vb1;
    bool do_xlo, any_x;
    int xhi_sel;
    uint32_t xhi_off;
    Face1 fx0, fx1, fy0, fy1;
    long long pitch;
    int lane;
    int it_ylo;
};
struct Row2St {
    double2 below_a, below_b, cur_a, cur_b;      // u0 rows a+it-1, a+it
    double2 u1lo_a, u1lo_b, u1c_a, u1c_b;        // u1 rows q-1, q   (q = a+it-1)
    double2 rp_a, rp_b;                          // rhs row q
    double* oa;
    uint32_t sc, sn, phn;
};

// poisson.rs:83-85: (c0*r + c1*(e+w)) + c2*(n+s)
__device__ __forceinline__ double upd2(double c0, double c1, double c2, double r, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, __dadd_rn(e, w))), __dmul_rn(c2, __dadd_rn(n, s)));
}
__device__ __forceinline__ double part1(double c0, double c1, double r, double e, double w)
{
    return __dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, __dadd_rn(e, w)));
}
__device__ __forceinline__ double fin1(double part, double c2, double n, double s) { return __dadd_rn(part, __dmul_rn(c2, __dadd_rn(n, s))); }

template <int NS, bool DO_L1, bool DO_L2>
__device__ __forceinline__ void row2_iter(const Row2K& k, Row2St& st, const int it)
{
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 above_a = zero2, above_b = zero2, u1n_a = zero2, u1n_b = zero2, rc_a = zero2, rc_b = zero2, P_a = zero2, P_b = zero2;

    if (DO_L1) mbar_wait(k.full0 + 8 * st.sn, st.phn);
    if (DO_L2 || it >= 2) mbar_wait(k.u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);

    if (DO_L2) {
        const uint32_t up = k.ua + (uint32_t)((it - 1) & 1) * ROW2_STAGE;
        const double uw_a0 = lds1(up - 8), ue_a1 = lds1(up + 16);
        const double uw_b0 = lds1(up + 512 - 8), ue_b1 = lds1(up + 512 + 16);
        P_a.x = part1(k.c0, k.c1, st.rp_a.x, st.u1c_a.y, uw_a0);
        P_a.y = part1(k.c0, k.c1, st.rp_a.y, ue_a1, st.u1c_a.x);
        P_b.x = part1(k.c0, k.c1, st.rp_b.x, st.u1c_b.y, uw_b0);
        P_b.y = part1(k.c0, k.c1, st.rp_b.y, ue_b1, st.u1c_b.x);
    }
    if (DO_L1) {
        const uint32_t pn = k.pa + st.sn * ROW2_STAGE, pc = k.pa + st.sc * ROW2_STAGE, pr = k.ra + st.sc * ROW2_STAGE;
        above_a = lds2(pn); above_b = lds2(pn + 512);
        rc_a = lds2(pr); rc_b = lds2(pr + 512);
        const double w_a0 = lds1(pc - 8), e_a1 = lds1(pc + 16);
        const double w_b0 = lds1(pc + 512 - 8), e_b1 = lds1(pc + 512 + 16);
        // ---------------- level 1: u1 of row a+it ----------------
        u1n_a.x = upd2(k.c0, k.c1, k.c2, rc_a.x, st.cur_a.y, w_a0, above_a.x, st.below_a.x);
        u1n_a.y = upd2(k.c0, k.c1, k.c2, rc_a.y, e_a1, st.cur_a.x, above_a.y, st.below_a.y);
        u1n_b.x = upd2(k.c0, k.c1, k.c2, rc_b.x, st.cur_b.y, w_b0, above_b.x, st.below_b.x);
        u1n_b.y = upd2(k.c0, k.c1, k.c2, rc_b.y, e_b1, st.cur_b.x, above_b.y, st.below_b.y);
        // odd n0: the E neighbour of the last valid cell at level 2 is the pair's second register
        if (k.xhi_sel == 0) u1n_a.y = ghost1(k.fx1, u1n_a.x);
        if (k.xhi_sel == 2) u1n_b.y = ghost1(k.fx1, u1n_b.x);
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.empty0 + 8 * st.sc);
        const uint32_t us = k.ua + (uint32_t)(it & 1) * ROW2_STAGE;
        sstore4(us, u1n_a, u1n_b, k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (k.any_x) {
            if (k.do_xlo) sts1(us - 8, ghost1(k.fx0, u1n_a.x));
            if (k.xhi_sel >= 0) {
                const double m = k.xhi_sel == 0 ? u1n_a.x : k.xhi_sel == 1 ? u1n_a.y : k.xhi_sel == 2 ? u1n_b.x : u1n_b.y;
                sts1(us + k.xhi_off, ghost1(k.fx1, m));
            }
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.u1full0 + 8 * (it & 1));
    }
    if (DO_L2) {
        // ---------------- level 2: u2 of row q = a+it-1 ----------------
        double2 dn_a = st.u1lo_a, dn_b = st.u1lo_b, up_a = u1n_a, up_b = u1n_b;
        const bool ylo = it == k.it_ylo;
        if (ylo) { dn_a = ghost1(k.fy0, st.u1c_a); dn_b = ghost1(k.fy0, st.u1c_b); }
        if (!DO_L1) { up_a = ghost1(k.fy1, st.u1c_a); up_b = ghost1(k.fy1, st.u1c_b); }      // the frame's last row
        double2 na, nb;
        na.x = fin1(P_a.x, k.c2, up_a.x, dn_a.x);
        na.y = fin1(P_a.y, k.c2, up_a.y, dn_a.y);
        nb.x = fin1(P_b.x, k.c2, up_b.x, dn_b.x);
        nb.y = fin1(P_b.y, k.c2, up_b.y, dn_b.y);
        double* oa = st.oa;
        store4(oa, na, nb, k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (k.any_x) {
            if (k.do_xlo) oa[-1] = ghost1(k.fx0, na.x);
            if (k.xhi_sel >= 0) {
                const double m = k.xhi_sel == 0 ? na.x : k.xhi_sel == 1 ? na.y : k.xhi_sel == 2 ? nb.x : nb.y;
                oa[(k.xhi_sel >= 2 ? 64 : 0) + (k.xhi_sel & 1) + 1] = ghost1(k.fx1, m);
            }
        }
        if (ylo) store4(oa - k.pitch, ghost1(k.fy0, na), ghost1(k.fy0, nb), k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (!DO_L1) store4(oa + k.pitch, ghost1(k.fy1, na), ghost1(k.fy1, nb), k.full, k.va0, k.va1, k.vb0, k.vb1);
        st.oa = oa + k.pitch;
    }
    if (DO_L1) {
        st.below_a = st.cur_a; st.below_b = st.cur_b;
        st.cur_a = above_a; st.cur_b = above_b;
        st.sc = st.sn;
        if (++st.sn == NS) { st.sn = 0; st.phn ^= 1u; }
    }
    st.u1lo_a = st.u1c_a; st.u1lo_b = st.u1c_b;
    st.u1c_a = u1n_a; st.u1c_b = u1n_b;
    st.rp_a = rc_a; st.rp_b = rc_b;
}

template <int NS>
__global__ void __launch_bounds__(32 * (NW2 + 2), 2) jacobi2d_tb2_kernel(const __grid_constant__ CUtensorMap tm_u256,
                                                                         const __grid_constant__ CUtensorMap tm_u4,
                                                                         const __grid_constant__ CUtensorMap tm_r256,
                                                                         const __grid_constant__ CUtensorMap tm_r4,
                                                                         const __grid_constant__ Tb2dParams p)
{
    using Cfg = Cfg2<NS>;
    RNMF_DYNAMIC_SMEM(smem_raw);
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t u0_0 = sb, rhs_0 = sb + Cfg::RHS_OFF, u1_0 = sb + Cfg::U1_OFF;
    const uint32_t full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8, u1full0 = empty0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = p.g.n[0], n1 = p.g.n[1];
    const long long i0 = (long long)blockIdx.x * TXC2;
    const long long j0 = p.m_begin + (long long)blockIdx.y * p.chunk;
    const long long j1 = j0 + p.chunk < p.m_end ? j0 + p.chunk : p.m_end;
    const long long a = j0 > 0 ? j0 - 1 : 0;              // level-1 rows a..b
    const long long b = j1 < n1 ? j1 : n1 - 1;
    const int nL1 = (int)(b - a) + 1;
    const int n_rows = nL1 + 2;                           // u0 rows a-1..b+1 (row a-1+idx <-> stage idx % NS)
    const bool top = b < j1;                              // the chunk ends with the frame's last row
    const int it_l2_first = (int)(j0 - a) + 1;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, Cfg::NCONS);
        }
        mbar_init(u1full0, Cfg::NCONS);
        mbar_init(u1full0 + 8, Cfg::NCONS);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NW2 + 1) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx = (int)(p.g.xoff + p.g.ng + i0 - 2);
            int cy = (int)(p.g.ng + a - 1);
            uint32_t s = 0, wrap = 0;
            for (int idx = 0; idx < n_rows; ++idx) {
                if (wrap > 0) mbar_wait(empty0 + 8 * s, (wrap - 1) & 1u);
                const bool with_rhs = idx >= 1 && idx <= nL1;
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(ROW2_BYTES * (with_rhs ? 2 : 1)));
                const uint32_t d = u0_0 + s * ROW2_STAGE;
                tma_load_2d(d, &tm_u256, cx, cy, full0 + 8 * s);
                tma_load_2d(d + 2048, &tm_u256, cx + 256, cy, full0 + 8 * s);
                tma_load_2d(d + 4096, &tm_u4, cx + 512, cy, full0 + 8 * s);
                if (with_rhs) {
                    const uint32_t r = rhs_0 + s * ROW2_STAGE;
                    tma_load_2d(r, &tm_r256, cx, cy, full0 + 8 * s);
                    tma_load_2d(r + 2048, &tm_r256, cx + 256, cy, full0 + 8 * s);
                    tma_load_2d(r + 4096, &tm_r4, cx + 512, cy, full0 + 8 * s);
                }
                ++cy;
                if (++s == NS) { s = 0; ++wrap; }
            }
        }
    } else if (warp == NW2) {
        // ------------------------------ halo-column warp (level 1 only) ------------------------------
        // lane 0: column i0-1; lane 1: column i0+512
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2;
        const int col = lane == 1 ? TXC2 : -1;
        const long long i = i0 + col;
        const bool ok = lane < 2 && i >= 0 && i < n0;
        const uint32_t hp = u0_0 + (uint32_t)((col + 2) * 8);
        const uint32_t hr = rhs_0 + (uint32_t)((col + 2) * 8);
        const uint32_t hu = u1_0 + (uint32_t)((col + 2) * 8);
        mbar_wait(full0, 0);
        double below = lds1(hp);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        double cur = lds1(hp + ROW2_STAGE);
        uint32_t sc = 1, sn = 2, phn = 0;
        for (int it = 0; it < nL1; ++it) {
            mbar_wait(full0 + 8 * sn, phn);
            if (it >= 2) mbar_wait(u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
            const uint32_t pc = hp + sc * ROW2_STAGE;
            const double above = lds1(hp + sn * ROW2_STAGE);
            const double ee = lds1(pc + 8), ww = lds1(pc - 8);
            const double rr = lds1(hr + sc * ROW2_STAGE);
            const double v = upd2(c0, c1, c2, rr, ee, ww, above, below);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * sc);
            if (ok) sts1(hu + (uint32_t)(it & 1) * ROW2_STAGE, v);
            __syncwarp();
            if (lane == 0) mbar_arrive(u1full0 + 8 * (it & 1));
            below = cur;
            cur = above;
            sc = sn;
            if (++sn == NS) { sn = 0; phn ^= 1u; }
        }
    } else {
        // ------------------------------ consumer warps ------------------------------
        const long long ia = i0 + 128 * warp + 2 * lane, ib = ia + 64;
        Row2K k;
        k.va0 = ia < n0; k.va1 = ia + 1 < n0; k.vb0 = ib < n0; k.vb1 = ib + 1 < n0;
        k.full = k.vb1;
        // column c of the strip sits at (c+2)*8 in a staged u0 / rhs row and in a u1 slot
        k.pa = u0_0 + (uint32_t)((128 * warp + 2 + 2 * lane) * 8);
        k.ra = rhs_0 + (uint32_t)((128 * warp + 2 + 2 * lane) * 8);
        k.ua = u1_0 + (uint32_t)((128 * warp + 2 + 2 * lane) * 8);
        k.full0 = full0; k.empty0 = empty0; k.u1full0 = u1full0;
        k.c0 = p.c.c0; k.c1 = p.c.c1; k.c2 = p.c.c2;
        k.pitch = p.g.pitch;
        k.lane = lane;
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fy0 = make_face1(p.faces.f[2], true); k.fy1 = make_face1(p.faces.f[3], true);
        k.do_xlo = k.fx0.on && k.va0 && ia == 0;
        const long long last = n0 - 1;
        k.xhi_sel = !k.fx1.on ? -1 : (k.va0 && ia == last) ? 0 : (k.va1 && ia + 1 == last) ? 1 : (k.vb0 && ib == last) ? 2 : (k.vb1 && ib + 1 == last) ? 3 : -1;
        k.xhi_off = (uint32_t)((k.xhi_sel >= 2 ? 512 : 0) + ((k.xhi_sel & 1) + 1) * 8);
        k.any_x = k.do_xlo || k.xhi_sel >= 0;
        k.it_ylo = j0 == 0 ? 1 : -1;

        Row2St st;
        const double2 zero2 = make_double2(0.0, 0.0);
        mbar_wait(full0, 0);
        st.below_a = lds2(k.pa); st.below_b = lds2(k.pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        st.cur_a = lds2(k.pa + ROW2_STAGE); st.cur_b = lds2(k.pa + ROW2_STAGE + 512);
        st.u1lo_a = zero2; st.u1lo_b = zero2; st.u1c_a = zero2; st.u1c_b = zero2; st.rp_a = zero2; st.rp_b = zero2;
        st.oa = p.out + p.g.at(ia, j0, 0, 0);
        st.sc = 1; st.sn = 2; st.phn = 0;

        int it = 0;
        for (; it < it_l2_first; ++it) row2_iter<NS, true, false>(k, st, it);
        for (; it < nL1; it += 3) {
            row2_iter<NS, true, true>(k, st, it);
            if (it + 1 >= nL1) break;
            row2_iter<NS, true, true>(k, st, it + 1);
            if (it + 2 >= nL1) break;
            row2_iter<NS, true, true>(k, st, it + 2);
        }
        if (top) row2_iter<NS, false, true>(k, st, nL1);
    }
}

constexpr int kNS2 = 10;

void plan2(const SweepArgs& a, dim3& grid, int& chunk)
{
    const FrameGeom& g = a.g;
    const long long span = a.k_end - a.k_begin;
    const long long bx = (g.n[0] + TXC2 - 1) / TXC2;
    // chunk length: whole waves of two CTAs per SM, each CTA paying chunk+3 row iterations
    long long best_c = 1, best_cost = -1;
    static const int cand[] = { 4, 8, 16, 24, 32, 48, 64, 96, 128, 192, 256 };
    for (int c : cand) {
        const long long cc = c > span ? span : c;
        if (cc < 1) break;
        const long long chunks = (span + cc - 1) / cc;
        if (chunks <= 60000) {
            const long long waves = (bx * chunks + 295) / 296;
            const long long cost = waves * (cc + 3);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_c = cc; }
        }
        if (c >= span) break;
    }
    if (a.chunk_override > 0) best_c = a.chunk_override > span ? span : a.chunk_override;
    if (best_c < 1) best_c = 1;
    chunk = (int)best_c;
    grid = dim3((unsigned)bx, (unsigned)((span + best_c - 1) / best_c), 1);
}

}  // namespace

int launch_jacobi_double_sweep_2d(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    using Cfg = Cfg2<kNS2>;
    dim3 grid;
    int chunk;
    plan2(a, grid, chunk);
    CUtensorMap tm_u256, tm_u4, tm_r256, tm_r4;
    if (tma2d_encode_map(&tm_u256, a.g, a.in, 256)) return -1;
    if (tma2d_encode_map(&tm_u4, a.g, a.in, 4)) return -1;
    if (tma2d_encode_map(&tm_r256, a.g, a.rhs, 256)) return -1;
    if (tma2d_encode_map(&tm_r4, a.g, a.rhs, 4)) return -1;
    Tb2dParams p;
    p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces;
    p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk;
#ifndef RNMF_EMULATE
    static PerDeviceOnce attr_set;
    if (!attr_set.done()) {
        if (cudaFuncSetAttribute(jacobi2d_tb2_kernel<kNS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
        cudaFuncSetAttribute(jacobi2d_tb2_kernel<kNS2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set.mark();
    }
#endif
    RNMF_KERNEL_LAUNCH(jacobi2d_tb2_kernel<kNS2>, grid, Cfg::THREADS, Cfg::SMEM, s, tm_u256, tm_u4, tm_r256, tm_r4, p);
    return 1;
}
