// kernels_sweep2d_tbk.cu -- K1 for 2D frames, temporally blocked over K = 2 .. 6 sweeps: K 5-point
// Jacobi sweeps (each followed by the ng=1 ghost fill, poisson.rs:80-89 + dataframe.rs:289-355) per
// pass over HBM, i.e. 24/K algorithmic bytes per update.  Generalises kernels_sweep2d_tb2.cu (K = 2);
// OPT-IN (sweep variants 32..40) until it has been measured on the device.
//   * a CTA owns a strip of 128*NW columns and marches a chunk of rows in y; the producer warp streams
//     one row of psi (u0) and one of rhs per stage, the strip plus PAD >= K halo columns either side;
//   * NW consumer warps own 128 cells each (lane l: pairs 2l,2l+1 and 64+2l,65+2l).  Level l = 1..K
//     (u_l = sweep(u_{l-1}) + its ghost fill) runs one row behind level l-1: in iteration `it` level l
//     forms row a+it-(l-1), its N neighbour being the value level l-1 formed in the same iteration and
//     its centre / S neighbour the two before it -- a two-row register window per level, so N/S never
//     touch shared memory.  Shared memory serves the next u0 row, rhs, and the E/W neighbours: rows of
//     u_1..u_{K-1} live in two-slot rings (with the x-face ghosts fill_bc would give them); the y-face
//     ghosts of u_1..u_{K-1} are formed in registers; c0*rhs of the last K-1 rows is a register queue
//     (the product is rounded once and reused by every level); warps that own no face / ragged cell run
//     a copy of the loop without ghost code and store masks; with DI the ring rows are de-interleaved;
//   * one more consumer warp computes the halo columns: level l on the K-l columns either side of the
//     strip, one lane per (level, column), levels chained through the rings with __syncwarp
//     (redundant work K(K-1)/(128*NW) per row; a chunk re-reads 2K rows);
//   * one barrier per ring slot: every consumer warp arrives once per iteration after its stores, and
//     waits for the previous iteration's arrivals before it loads (that wait is also the
//     write-after-read guard of the slot it overwrites).
// Every operation is explicitly rounded in the reference's association (poisson.rs:83-85):
// u_K is bit-identical to K single sweeps.
// y-slabs (one GPU per slab), DEEP HALO as in kernels_sweep_tb2.cu: the frame buffers carry K-1 extra rows in front of the y-lo
// ghost row and behind the y-hi ghost row.  With the neighbour's K boundary rows in ghost + deep rows a slab face towards a
// neighbour is a chunk boundary (no face rule, no clipping of the lower levels); u_K of my K rows next to the face goes into the
// neighbour's ghost + deep rows of its `out` buffer (whole grown rows, x-face ghosts included) under the single-sweep kernels'
// epoch handshake, one epoch per launch.
#include <stdlib.h>

#include "tma_util.cuh"

namespace {
using namespace tmau;

struct TbkParams {
    FrameGeom g;
    double* out;
    SweepCoef c;
    FaceSet faces;
    long long m_begin, m_end;   // output rows
    int chunk;
    int n_chunks;
    int nb_lo, nb_hi;           // a slab neighbour behind the y-lo / y-hi face
    int yshift;                 // the tensor maps start this many rows before the y-lo ghost row (deep rows present)
    // slabs: u_K of my rows [0, K) also goes to the lo neighbour's y-hi ghost + deep rows, of my rows [n1-K, n1) to the hi
    // neighbour's deep + y-lo ghost rows (its `out` buffer).  Both are a CONSTANT element offset from where the row goes in my own
    // `out` buffer (same pitch, same x offsets): destination = own address + delta
    int has_peer_lo, has_peer_hi;
    long long peer_lo_delta, peer_hi_delta;
    PeerSync ps;
};

// DI: ring rows are stored de-interleaved -- even positions first, odd positions from ODD_OFF -- so that the E/W loads
// (the pair's left neighbour is an odd position, its right neighbour an even one) are dense 8-byte accesses, two
// shared-memory wavefronts per warp instead of the four that 8-byte accesses 16 bytes apart cost; the pair stores
// become two dense 8-byte stores each (the same four wavefronts as one 16-byte store)
template <int K, int NW, int NS, bool DI = false>
struct CfgK {
    static constexpr int PAD = (K + 1) & ~1;               // even (pairs stay 16-byte aligned in a staged row), >= K
    static constexpr int TXC = 128 * NW;
    static constexpr int ROW = TXC + 2 * PAD;              // a staged row: columns i0-PAD .. i0+TXC+PAD-1
    static constexpr int ROW_BYTES = ROW * 8;
    static constexpr int STAGE = (ROW_BYTES + 127) / 128 * 128;
    static constexpr int RHS_OFF = NS * STAGE;
    static constexpr int ODD_OFF = (ROW / 2 * 8 + 127) / 128 * 128;     // DI: byte offset of the odd positions in a ring row
    static constexpr int RSLOT = DI ? 2 * ODD_OFF : STAGE;              // bytes per ring slot
    static constexpr int RING_OFF = 2 * NS * STAGE;        // ring of level s (1..K-1), slot t: RING_OFF + (2(s-1)+t)*RSLOT
    static constexpr int BAR_OFF = RING_OFF + 2 * (K - 1) * RSLOT;
    static constexpr int SMEM = BAR_OFF + (2 * NS + 2) * 8 + 128;
    static constexpr int NCONS = NW + 1;                   // + the halo-column warp
    static constexpr int THREADS = 32 * (NW + 2);
};

// iterations in which level l+1 is active in this chunk, and the ones in which it meets a y face
template <int K>
struct LevelPlan {
    int first[K], last[K], it_ylo[K], it_yhi[K];
    int n_it;           // iterations of the chunk
    int s0, s1;         // steady iterations [s0, s1]: every level active, no y face
    long long a;        // first level-1 row
    int nL1;            // level-1 rows
};
template <int K>
__device__ __forceinline__ LevelPlan<K> make_plan(long long j0, long long j1, long long n1, bool nb_lo, bool nb_hi)
{
    LevelPlan<K> p;
    long long lo[K], hi[K];
    // level l covers the output rows grown by K-l rows either side, clipped at a PHYSICAL y face (its ghost rule replaces the
    // missing neighbour row); towards a slab neighbour the deep rows make the grown range computable
#pragma unroll
    for (int l = 1; l <= K; ++l) {
        lo[l - 1] = (j0 - (K - l) > 0 || nb_lo) ? j0 - (K - l) : 0;
        hi[l - 1] = (j1 - 1 + (K - l) < n1 - 1 || nb_hi) ? j1 - 1 + (K - l) : n1 - 1;
    }
    p.a = lo[0];
    p.nL1 = (int)(hi[0] - lo[0]) + 1;
#pragma unroll
    for (int l = 1; l <= K; ++l) {
        p.first[l - 1] = (int)(lo[l - 1] - p.a) + (l - 1);
        p.last[l - 1] = (int)(hi[l - 1] - p.a) + (l - 1);
        p.it_ylo[l - 1] = (!nb_lo && lo[l - 1] == 0) ? p.first[l - 1] : -1;
        p.it_yhi[l - 1] = (!nb_hi && hi[l - 1] == n1 - 1) ? p.last[l - 1] : -1;
    }
    p.n_it = p.last[K - 1] + 1;
    p.s0 = p.first[K - 1] + ((!nb_lo && lo[K - 1] == 0) ? 1 : 0);
    p.s1 = p.last[0];
    // rows that also go to a slab neighbour (the first / last K rows of the slab) are stored by the general iterations:
    // level K forms row r in iteration r - a + K - 1
    if (nb_lo && j0 < K) { const int e = (int)(K - 1 - p.a) + K; if (e > p.s0) p.s0 = e; }
    if (nb_hi && j1 > n1 - K) { const int b = (int)(n1 - K - p.a) + K - 2; if (b < p.s1) p.s1 = b; }
    return p;
}

template <int K>
struct LaneK {
    uint32_t pa, ra, ua;            // the lane's first pair in stage 0 of u0 / rhs, in slot 0 of the level-1 ring
    uint32_t full0, empty0, ufull0;
    double c0, c1, c2;
    bool full, va0, va1, vb0, vb1;
    bool do_xlo, any_x;
    int xhi_sel;
    uint32_t xhi_off, xlo_off;      // ring-row offsets of the x-hi / x-lo ghost from the lane's first cell
    Face1 fx0, fx1, fy0, fy1;
    long long pitch;
    int lane;
    LevelPlan<K> lp;
};
template <int K>
struct StateK {
    double2 below_a, below_b, cur_a, cur_b;              // u0 rows a+it-1, a+it
    double2 lo_a[K - 1], lo_b[K - 1], c_a[K - 1], c_b[K - 1];   // [s-1]: the last two rows level s formed
    double2 rq_a[K - 1], rq_b[K - 1];                    // [d]: c0 * rhs of row a+it-1-d
    double* oa;
    uint32_t sc, sn, phn;
};

// ---- ring rows: position q = column - i0 + PAD ----
template <class Cfg, bool DI>
__device__ __forceinline__ uint32_t ring_pos(int q) { return DI ? (uint32_t)((q & 1) * Cfg::ODD_OFF + (q >> 1) * 8) : (uint32_t)(q * 8); }
// E/W neighbours of the lane's two pairs in the ring row where the lane's first cell sits at `u`
template <class Cfg, bool DI>
__device__ __forceinline__ void ring_ld_ew(uint32_t u, double& w_a, double& e_a, double& w_b, double& e_b)
{
    if (DI) { w_a = lds1(u + Cfg::ODD_OFF - 8); e_a = lds1(u + 8); w_b = lds1(u + Cfg::ODD_OFF + 256 - 8); e_b = lds1(u + 256 + 8); }
    else { w_a = lds1(u - 8); e_a = lds1(u + 16); w_b = lds1(u + 512 - 8); e_b = lds1(u + 512 + 16); }
}
template <class Cfg, bool DI>
__device__ __forceinline__ void ring_st4(uint32_t u, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
    if (!DI) { sstore4(u, va, vb, full, a0, a1, b0, b1); return; }
    if (full) {
        sts1(u, va.x); sts1(u + Cfg::ODD_OFF, va.y); sts1(u + 256, vb.x); sts1(u + Cfg::ODD_OFF + 256, vb.y);
    } else {
        if (a0) sts1(u, va.x);
        if (a1) sts1(u + Cfg::ODD_OFF, va.y);
        if (b0) sts1(u + 256, vb.x);
        if (b1) sts1(u + Cfg::ODD_OFF + 256, vb.y);
    }
}

// poisson.rs:83-85: (c0*r + c1*(e+w)) + c2*(n+s); p0 = c0*r is rounded once per cell and row and reused by every
// level (the same IEEE operation on the same operands, so the result is the reference's)
__device__ __forceinline__ double updk(double c1, double c2, double p0, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(p0, __dmul_rn(c1, __dadd_rn(e, w))), __dmul_rn(c2, __dadd_rn(n, s)));
}
__device__ __forceinline__ double2 mul2(double c, const double2& v) { return make_double2(__dmul_rn(c, v.x), __dmul_rn(c, v.y)); }

// x-face ghosts of a row of u_1..u_{K-1} just stored at `us` (lanes next to a face only: any_x): the ghost
// fill_bc would give the level goes into the ring beside the row; with odd n0 the E neighbour of the last valid
// cell at the next level is the pair's second register, which takes the ghost too
template <int K>
__device__ __forceinline__ void xface_ring(const LaneK<K>& k, uint32_t us, double2& v_a, double2& v_b)
{
    if (k.do_xlo) sts1(us + k.xlo_off, ghost1(k.fx0, v_a.x));
    if (k.xhi_sel >= 0) {
        const double m = k.xhi_sel == 0 ? v_a.x : k.xhi_sel == 1 ? v_a.y : k.xhi_sel == 2 ? v_b.x : v_b.y;
        const double gv = ghost1(k.fx1, m);
        sts1(us + k.xhi_off, gv);
        if (k.xhi_sel == 0) v_a.y = gv;
        if (k.xhi_sel == 2) v_b.y = gv;
    }
}

// u_K of the lane's four cells into one grown row: the valid cells and the x-face ghosts fill_bc gives them
template <int K, bool XF>
__device__ __forceinline__ void row_store(const LaneK<K>& k, double* oa, const double2& o_a, const double2& o_b)
{
    store4(oa, o_a, o_b, !XF || k.full, k.va0, k.va1, k.vb0, k.vb1);
    if (XF && k.any_x) {
        if (k.do_xlo) oa[-1] = ghost1(k.fx0, o_a.x);
        if (k.xhi_sel >= 0) {
            const double m = k.xhi_sel == 0 ? o_a.x : k.xhi_sel == 1 ? o_a.y : k.xhi_sel == 2 ? o_b.x : o_b.y;
            oa[(k.xhi_sel >= 2 ? 64 : 0) + (k.xhi_sel & 1) + 1] = ghost1(k.fx1, m);
        }
    }
}

// XF: an edge warp -- it owns a cell next to an x face, or columns past the frame (masked stores).  Warp-uniform, so
// the ghost code and the store masks are compiled out of the other warps' loop.
template <int K, int NW, int NS, bool DI, bool STEADY, bool XF>
__device__ __forceinline__ void rowk_iter(const TbkParams& p, const LaneK<K>& k, StateK<K>& st, const int it)
{
    using Cfg = CfgK<K, NW, NS, DI>;
    const double2 zero2 = make_double2(0.0, 0.0);
    const bool act1 = STEADY || it <= k.lp.last[0];

    if (act1) mbar_wait(k.full0 + 8 * st.sn, st.phn);
    if (it >= 1) mbar_wait(k.ufull0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);

    // ---- every shared-memory load of the iteration, issued together ----
    const uint32_t pslot = (uint32_t)((it - 1) & 1) * Cfg::RSLOT, slot = (uint32_t)(it & 1) * Cfg::RSLOT;
    double uw_a[K - 1], ue_a[K - 1], uw_b[K - 1], ue_b[K - 1];      // [s-1]: E/W of the row level s formed in the previous iteration
    // (the 512-column K = 4 form has 170 registers per thread: there the ring loads are issued level by level instead)
    constexpr bool LAZY = K >= 4 && NW <= 4;
#pragma unroll
    for (int s = 1; s < K; ++s) {
        const bool act = STEADY || (it >= k.lp.first[s] && it <= k.lp.last[s]);       // level s+1
        uw_a[s - 1] = ue_a[s - 1] = uw_b[s - 1] = ue_b[s - 1] = 0.0;
        if (act && !LAZY) {
            const uint32_t up = k.ua + (uint32_t)(2 * (s - 1)) * Cfg::RSLOT + pslot;
            ring_ld_ew<Cfg, DI>(up, uw_a[s - 1], ue_a[s - 1], uw_b[s - 1], ue_b[s - 1]);
        }
    }
    double2 above_a = zero2, above_b = zero2, rc_a = zero2, rc_b = zero2, f_a = zero2, f_b = zero2;

    // ---------------- level 1: u1 of row a+it ----------------
    if (act1) {
        const uint32_t pn = k.pa + st.sn * Cfg::STAGE, pc = k.pa + st.sc * Cfg::STAGE, pr = k.ra + st.sc * Cfg::STAGE;
        above_a = lds2(pn); above_b = lds2(pn + 512);
        rc_a = mul2(k.c0, lds2(pr)); rc_b = mul2(k.c0, lds2(pr + 512));
        const double w_a0 = lds1(pc - 8), e_a1 = lds1(pc + 16);
        const double w_b0 = lds1(pc + 512 - 8), e_b1 = lds1(pc + 512 + 16);
        f_a.x = updk(k.c1, k.c2, rc_a.x, st.cur_a.y, w_a0, above_a.x, st.below_a.x);
        f_a.y = updk(k.c1, k.c2, rc_a.y, e_a1, st.cur_a.x, above_a.y, st.below_a.y);
        f_b.x = updk(k.c1, k.c2, rc_b.x, st.cur_b.y, w_b0, above_b.x, st.below_b.x);
        f_b.y = updk(k.c1, k.c2, rc_b.y, e_b1, st.cur_b.x, above_b.y, st.below_b.y);
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.empty0 + 8 * st.sc);
        const uint32_t us = k.ua + slot;
        ring_st4<Cfg, DI>(us, f_a, f_b, !XF || k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (XF && k.any_x) xface_ring(k, us, f_a, f_b);
    }

    // ---------------- levels 2..K: u_l of row a+it-(l-1) ----------------
    bool prev_act = act1;
#pragma unroll
    for (int l = 2; l <= K; ++l) {
        const int s = l - 1;                                   // the source level
        const bool act = STEADY || (it >= k.lp.first[l - 1] && it <= k.lp.last[l - 1]);
        double2 o_a = zero2, o_b = zero2;
        if (act) {
            const double2 ce_a = st.c_a[s - 1], ce_b = st.c_b[s - 1];
            double2 dn_a = st.lo_a[s - 1], dn_b = st.lo_b[s - 1], up_a = f_a, up_b = f_b;
            if (!STEADY) {
                if (it == k.lp.it_ylo[l - 1]) { dn_a = ghost1(k.fy0, ce_a); dn_b = ghost1(k.fy0, ce_b); }
                if (it == k.lp.it_yhi[l - 1]) { up_a = ghost1(k.fy1, ce_a); up_b = ghost1(k.fy1, ce_b); }
            }
            if (LAZY) {
                const uint32_t up = k.ua + (uint32_t)(2 * (s - 1)) * Cfg::RSLOT + pslot;
                ring_ld_ew<Cfg, DI>(up, uw_a[s - 1], ue_a[s - 1], uw_b[s - 1], ue_b[s - 1]);
            }
            const double2 r_a = st.rq_a[l - 2], r_b = st.rq_b[l - 2];
            o_a.x = updk(k.c1, k.c2, r_a.x, ce_a.y, uw_a[s - 1], up_a.x, dn_a.x);
            o_a.y = updk(k.c1, k.c2, r_a.y, ue_a[s - 1], ce_a.x, up_a.y, dn_a.y);
            o_b.x = updk(k.c1, k.c2, r_b.x, ce_b.y, uw_b[s - 1], up_b.x, dn_b.x);
            o_b.y = updk(k.c1, k.c2, r_b.y, ue_b[s - 1], ce_b.x, up_b.y, dn_b.y);
            if (l < K) {
                const uint32_t us = k.ua + (uint32_t)(2 * (l - 1)) * Cfg::RSLOT + slot;
                ring_st4<Cfg, DI>(us, o_a, o_b, !XF || k.full, k.va0, k.va1, k.vb0, k.vb1);
                if (XF && k.any_x) xface_ring(k, us, o_a, o_b);
            } else {
                double* oa = st.oa;
                row_store<K, XF>(k, oa, o_a, o_b);
                if (!STEADY) {
                    if (it == k.lp.it_ylo[K - 1]) store4(oa - k.pitch, ghost1(k.fy0, o_a), ghost1(k.fy0, o_b), !XF || k.full, k.va0, k.va1, k.vb0, k.vb1);
                    if (it == k.lp.it_yhi[K - 1]) store4(oa + k.pitch, ghost1(k.fy1, o_a), ghost1(k.fy1, o_b), !XF || k.full, k.va0, k.va1, k.vb0, k.vb1);
                    // rows next to a slab neighbour: the same grown row into its ghost / deep rows (NVLink stores)
                    const long long r = k.lp.a - (K - 1) + it;          // the row level K forms in this iteration
                    if (p.has_peer_lo && r < K) row_store<K, XF>(k, oa + p.peer_lo_delta, o_a, o_b);
                    if (p.has_peer_hi && r >= p.g.n[1] - K) row_store<K, XF>(k, oa + p.peer_hi_delta, o_a, o_b);
                }
                st.oa = oa + k.pitch;
            }
        }
        if (prev_act) {                                        // level s formed a row: its window moves up
            st.lo_a[s - 1] = st.c_a[s - 1]; st.lo_b[s - 1] = st.c_b[s - 1];
            st.c_a[s - 1] = f_a; st.c_b[s - 1] = f_b;
        }
        f_a = o_a; f_b = o_b;
        prev_act = act;
    }
    __syncwarp();
    if (k.lane == 0) mbar_arrive(k.ufull0 + 8 * (it & 1));

    if (act1) {
        st.below_a = st.cur_a; st.below_b = st.cur_b;
        st.cur_a = above_a; st.cur_b = above_b;
        st.sc = st.sn;
        if (++st.sn == NS) { st.sn = 0; st.phn ^= 1u; }
    }
#pragma unroll
    for (int d = K - 2; d >= 1; --d) { st.rq_a[d] = st.rq_a[d - 1]; st.rq_b[d] = st.rq_b[d - 1]; }
    st.rq_a[0] = rc_a; st.rq_b[0] = rc_b;
}

// the iterations of a chunk: ramp-up (levels start one after the other, y-lo faces), the steady part unrolled by
// three (the three-row windows rotate by renaming), drain (y-hi faces)
template <int K, int NW, int NS, bool DI, bool XF>
__device__ __forceinline__ void rowk_chunk(const TbkParams& p, const LaneK<K>& k, StateK<K>& st)
{
    int it = 0;
    for (; it < k.lp.s0 && it < k.lp.n_it; ++it) rowk_iter<K, NW, NS, DI, false, XF>(p, k, st, it);
    for (; it + 2 <= k.lp.s1; it += 3) {
        rowk_iter<K, NW, NS, DI, true, XF>(p, k, st, it);
        rowk_iter<K, NW, NS, DI, true, XF>(p, k, st, it + 1);
        rowk_iter<K, NW, NS, DI, true, XF>(p, k, st, it + 2);
    }
    for (; it < k.lp.n_it; ++it) rowk_iter<K, NW, NS, DI, false, XF>(p, k, st, it);
}

template <int K, int NW, int NS, int MINB, bool DI>
__global__ void __launch_bounds__(32 * (NW + 2), MINB) jacobi2d_tbk_kernel(const __grid_constant__ CUtensorMap tm_u256,
                                                                             const __grid_constant__ CUtensorMap tm_upad,
                                                                             const __grid_constant__ CUtensorMap tm_r256,
                                                                             const __grid_constant__ CUtensorMap tm_rpad,
                                                                             const __grid_constant__ TbkParams p)
{
    using Cfg = CfgK<K, NW, NS, DI>;
    constexpr int PAD = Cfg::PAD, TXC = Cfg::TXC, STAGE = Cfg::STAGE, RSLOT = Cfg::RSLOT;
    RNMF_DYNAMIC_SMEM(smem_raw);
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t u0_0 = sb, rhs_0 = sb + Cfg::RHS_OFF, ring_0 = sb + Cfg::RING_OFF;
    const uint32_t full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8, ufull0 = empty0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = p.g.n[0], n1 = p.g.n[1];
    const long long i0 = (long long)blockIdx.x * TXC;
    const unsigned yc = p.ps.enabled ? peer_chunk_order(blockIdx.y, (unsigned)p.n_chunks) : blockIdx.y;      // boundary chunks last
    const long long j0 = p.m_begin + (long long)yc * p.chunk;
    const long long j1 = j0 + p.chunk < p.m_end ? j0 + p.chunk : p.m_end;
    const LevelPlan<K> lp = make_plan<K>(j0, j1, n1, p.nb_lo != 0, p.nb_hi != 0);
    const int n_rows = lp.nL1 + 2;                          // u0 rows a-1..b+1 (row a-1+idx <-> stage idx % NS)
    // slab handshake roles: the chunk reads the neighbour's rows and stores into the neighbour's buffer
    const bool sync_lo = p.ps.enabled && p.nb_lo && j0 == 0;
    const bool sync_hi = p.ps.enabled && p.nb_hi && j1 == n1;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, Cfg::NCONS);
        }
        mbar_init(ufull0, Cfg::NCONS);
        mbar_init(ufull0 + 8, Cfg::NCONS);
        mbar_fence_init();
        if (sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
    }
    __syncthreads();

    if (warp == NW + 1) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx = (int)(p.g.xoff + p.g.ng + i0 - PAD);
            int cy = (int)(p.g.ng + lp.a - 1) + p.yshift;
            uint32_t s = 0, wrap = 0;
            for (int idx = 0; idx < n_rows; ++idx) {
                if (wrap > 0) mbar_wait(empty0 + 8 * s, (wrap - 1) & 1u);
                const bool with_rhs = idx >= 1 && idx <= lp.nL1;
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(Cfg::ROW_BYTES * (with_rhs ? 2 : 1)));
                const uint32_t d = u0_0 + s * STAGE;
#pragma unroll
                for (int b = 0; b < TXC / 256; ++b) tma_load_2d(d + 2048 * b, &tm_u256, cx + 256 * b, cy, full0 + 8 * s);
                tma_load_2d(d + 8 * TXC, &tm_upad, cx + TXC, cy, full0 + 8 * s);
                if (with_rhs) {
                    const uint32_t r = rhs_0 + s * STAGE;
#pragma unroll
                    for (int b = 0; b < TXC / 256; ++b) tma_load_2d(r + 2048 * b, &tm_r256, cx + 256 * b, cy, full0 + 8 * s);
                    tma_load_2d(r + 8 * TXC, &tm_rpad, cx + TXC, cy, full0 + 8 * s);
                }
                ++cy;
                if (++s == NS) { s = 0; ++wrap; }
            }
        }
    } else if (warp == NW) {
        // ------------------------------ halo-column warp ------------------------------
        // one lane per (level L = 1..K-1, column): level L on the K-L columns either side of the strip
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2;
        int L = 0, col = 0;
        {
            int base = 0;
#pragma unroll
            for (int l = 1; l < K; ++l) {
                const int w = K - l;
                if (L == 0 && lane < base + 2 * w) {
                    const int j = lane - base;
                    L = l;
                    col = j < w ? j - w : TXC + (j - w);
                }
                base += 2 * w;
            }
        }
        const long long i = i0 + col;
        const bool ok = L > 0 && i >= 0 && i < n0;
        const Face1 fx1 = make_face1(p.faces.f[1], true), fy0 = make_face1(p.faces.f[2], true), fy1 = make_face1(p.faces.f[3], true);
        const bool xhi = ok && fx1.on && i == n0 - 1;         // the x-hi ghost of my level sits right of my column
        const uint32_t off = (uint32_t)((col + PAD) * 8);
        const uint32_t hp = u0_0 + off, hr = rhs_0 + off;
        // my column and its two neighbours in a ring row
        const uint32_t hu = ring_0 + ring_pos<Cfg, DI>(col + PAD), hue = ring_0 + ring_pos<Cfg, DI>(col + PAD + 1), huw = ring_0 + ring_pos<Cfg, DI>(col + PAD - 1);
        double below = 0.0, cur = 0.0, wlo = 0.0, wc = 0.0;
        double q[K - 1];
#pragma unroll
        for (int d = 0; d < K - 1; ++d) q[d] = 0.0;
        mbar_wait(full0, 0);
        below = lds1(hp);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        cur = lds1(hp + STAGE);
        uint32_t sc = 1, sn = 2, phn = 0;
        for (int it = 0; it < lp.n_it; ++it) {
            const bool act1 = it <= lp.last[0];
            if (act1) mbar_wait(full0 + 8 * sn, phn);
            if (it >= 1) mbar_wait(ufull0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
            const uint32_t pslot = (uint32_t)((it - 1) & 1) * RSLOT, slot = (uint32_t)(it & 1) * RSLOT;
            double rr = 0.0, above = 0.0;
            if (act1) {
                rr = __dmul_rn(c0, lds1(hr + sc * STAGE));
                above = lds1(hp + sn * STAGE);
                if (L == 1) {
                    const uint32_t pc = hp + sc * STAGE;
                    const double ee = lds1(pc + 8), ww = lds1(pc - 8);
                    const double v = updk(c1, c2, rr, ee, ww, above, below);
                    if (ok) {
                        sts1(hu + slot, v);
                        if (xhi) sts1(hue + slot, ghost1(fx1, v));
                    }
                }
            }
            __syncwarp();
            if (act1 && lane == 0) mbar_arrive(empty0 + 8 * sc);
#pragma unroll
            for (int l = 2; l < K; ++l) {
                const bool actl = it >= lp.first[l - 1] && it <= lp.last[l - 1];
                const bool actp = it >= lp.first[l - 2] && it <= lp.last[l - 2];   // level l-1 stored a row in this iteration
                if (L == l) {
                    const uint32_t lsrc = (uint32_t)(2 * (l - 2)) * RSLOT;
                    double fresh = 0.0;
                    if (actp && ok) fresh = lds1(hu + lsrc + slot);
                    if (actl) {
                        double dn = wlo, up = fresh;
                        if (it == lp.it_ylo[l - 1]) dn = ghost1(fy0, wc);
                        if (it == lp.it_yhi[l - 1]) up = ghost1(fy1, wc);
                        const double ee = lds1(hue + lsrc + pslot), ww = lds1(huw + lsrc + pslot);
                        const double v = updk(c1, c2, q[l - 2], ee, ww, up, dn);
                        if (ok) {
                            const uint32_t ldst = (uint32_t)(2 * (l - 1)) * RSLOT + slot;
                            sts1(hu + ldst, v);
                            if (xhi) sts1(hue + ldst, ghost1(fx1, v));
                        }
                    }
                    if (actp) { wlo = wc; wc = fresh; }
                }
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ufull0 + 8 * (it & 1));
            if (act1) {
                below = cur;
                cur = above;
                sc = sn;
                if (++sn == NS) { sn = 0; phn ^= 1u; }
            }
#pragma unroll
            for (int d = K - 2; d >= 1; --d) q[d] = q[d - 1];
            q[0] = rr;
        }
    } else {
        // ------------------------------ consumer warps ------------------------------
        const long long ia = i0 + 128 * warp + 2 * lane, ib = ia + 64;
        LaneK<K> k;
        k.va0 = ia < n0; k.va1 = ia + 1 < n0; k.vb0 = ib < n0; k.vb1 = ib + 1 < n0;
        k.full = k.vb1;
        // column c of the strip sits at (c+PAD)*8 in a staged u0 / rhs row and in a ring slot
        const uint32_t off = (uint32_t)((128 * warp + PAD + 2 * lane) * 8);
        k.pa = u0_0 + off; k.ra = rhs_0 + off; k.ua = ring_0 + ring_pos<Cfg, DI>(128 * warp + PAD + 2 * lane);
        k.full0 = full0; k.empty0 = empty0; k.ufull0 = ufull0;
        k.c0 = p.c.c0; k.c1 = p.c.c1; k.c2 = p.c.c2;
        k.pitch = p.g.pitch;
        k.lane = lane;
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fy0 = make_face1(p.faces.f[2], true); k.fy1 = make_face1(p.faces.f[3], true);
        k.do_xlo = k.fx0.on && k.va0 && ia == 0;
        const long long last = n0 - 1;
        k.xhi_sel = !k.fx1.on ? -1 : (k.va0 && ia == last) ? 0 : (k.va1 && ia + 1 == last) ? 1 : (k.vb0 && ib == last) ? 2 : (k.vb1 && ib + 1 == last) ? 3 : -1;
        // the x-hi ghost sits one position right of the last valid cell, the x-lo ghost one left of the lane's first cell
        k.xhi_off = DI ? (uint32_t)((k.xhi_sel >= 2 ? 256 : 0) + ((k.xhi_sel & 1) ? 8 : Cfg::ODD_OFF))
                       : (uint32_t)((k.xhi_sel >= 2 ? 512 : 0) + ((k.xhi_sel & 1) + 1) * 8);
        k.xlo_off = DI ? (uint32_t)(Cfg::ODD_OFF - 8) : (uint32_t)-8;
        k.any_x = k.do_xlo || k.xhi_sel >= 0;
        k.lp = lp;

        StateK<K> st;
        const double2 zero2 = make_double2(0.0, 0.0);
        mbar_wait(full0, 0);
        st.below_a = lds2(k.pa); st.below_b = lds2(k.pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        st.cur_a = lds2(k.pa + STAGE); st.cur_b = lds2(k.pa + STAGE + 512);
#pragma unroll
        for (int s = 0; s < K - 1; ++s) {
            st.lo_a[s] = zero2; st.lo_b[s] = zero2; st.c_a[s] = zero2; st.c_b[s] = zero2;
            st.rq_a[s] = zero2; st.rq_b[s] = zero2;
        }
        st.oa = p.out + p.g.at(ia, j0, 0, 0);
        st.sc = 1; st.sn = 2; st.phn = 0;

        // warp-uniform: does this warp's 128-column span touch an x face or reach past the frame?
        const long long w0 = i0 + 128 * warp;
        const bool warp_x = (k.fx0.on && w0 == 0) || (k.fx1.on && last >= w0 && last < w0 + 128) || w0 + 128 > n0;
        if (warp_x) rowk_chunk<K, NW, NS, DI, true>(p, k, st);
        else rowk_chunk<K, NW, NS, DI, false>(p, k, st);
    }
    if (sync_lo || sync_hi) {
        __syncthreads();                                       // every peer store of this CTA has been issued
        if (threadIdx.x == 0) {
            if (sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.ps.sig_lo, p.ps.epoch + 1);
            if (sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.ps.sig_hi, p.ps.epoch + 1);
        }
    }
}

template <int K, int NW, int NS, int MINB, bool DI = false>
void plan_tbk(const SweepArgs& a, dim3& grid, int& chunk)
{
    using Cfg = CfgK<K, NW, NS, DI>;
    const FrameGeom& g = a.g;
    const long long span = a.k_end - a.k_begin;
    const long long bx = (g.n[0] + Cfg::TXC - 1) / Cfg::TXC;
    // chunk length: W whole waves of 148*MINB CTAs, each CTA paying chunk + 2(K-1) + 1 row iterations.  For every W the
    // largest number of chunks that still fits W waves gives the shortest chunk; take the cheapest W (8192^2, 768-column strips:
    // 40 chunks of 205 rows = 440 CTAs = 2.97 waves, 4 % above the wave-free ideal; the fixed candidate list gave 10 %).
    long long best_c = span, best_cost = -1;
    const long long per_wave = 148LL * MINB, ramp = 2 * (K - 1) + 1;
    for (long long W = 1; W <= 64; ++W) {
        long long chunks = W * per_wave / bx;
        if (chunks < 1) continue;
        if (chunks > span) chunks = span;
        if (chunks > 60000) chunks = 60000;
        long long cc = (span + chunks - 1) / chunks;
        if (cc < 16 * K) cc = 16 * K < span ? 16 * K : span;        // a chunk re-reads 2K rows: keep that below ~12 % of its traffic
        const long long ctas = bx * ((span + cc - 1) / cc);
        const long long cost = ((ctas + per_wave - 1) / per_wave) * (cc + ramp);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = cc; }
        if (chunks == span) break;
    }
    // Measured at 8192^2 (profiles/r2_notes.md section 7): the wave model's pick, 37 chunks of 222 rows = exactly two waves of
    // 296 CTAs, runs 589 GLUP/s; chunks of 96-128 rows run 679 (48-64: 664-668, 160-256: 635-639, 32: 585) although the model
    // charges them a fourth wave -- the two CTAs of an SM then ramp up / drain at different times.  Cap the chunk at 128 rows.
    if (best_c > 128) best_c = 128;
    if (a.chunk_override > 0) best_c = a.chunk_override > span ? span : a.chunk_override;
    if (best_c < 1) best_c = 1;
    // slabs: the first / last chunk holds all K rows that go to the neighbour
    if (a.nb_lo || a.nb_hi) {
        if (best_c < K) best_c = span < K ? span : K;
        while ((span + best_c - 1) / best_c > 1 && span - ((span + best_c - 1) / best_c - 1) * best_c < K) ++best_c;
    }
    chunk = (int)best_c;
    grid = dim3((unsigned)bx, (unsigned)((span + best_c - 1) / best_c), 1);
}

template <int K, int NW, int NS, int MINB, bool DI = false>
int launch_tbk(const SweepArgs& a, cudaStream_t s)
{
    using Cfg = CfgK<K, NW, NS, DI>;
    const FrameGeom& g = a.g;
    dim3 grid;
    int chunk;
    plan_tbk<K, NW, NS, MINB, DI>(a, grid, chunk);
    if ((a.nb_lo || a.nb_hi) && a.deep < K - 1) { rnmf_set_error("K-sweep kernel across slabs needs K-1 deep halo rows"); return -1; }
    const int deep = a.deep >= K - 1 ? K - 1 : 0;           // rows in front of the y-lo ghost row that the tensor maps include
    CUtensorMap tm_u256, tm_upad, tm_r256, tm_rpad;
    if (tma2d_encode_map_deep(&tm_u256, g, a.in, 256, deep)) return -1;
    if (tma2d_encode_map_deep(&tm_upad, g, a.in, 2 * Cfg::PAD, deep)) return -1;
    if (tma2d_encode_map_deep(&tm_r256, g, a.rhs, 256, deep)) return -1;
    if (tma2d_encode_map_deep(&tm_rpad, g, a.rhs, 2 * Cfg::PAD, deep)) return -1;
    TbkParams p;
    p.g = g; p.out = a.out; p.c = a.c; p.faces = a.faces;
    p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk; p.n_chunks = (int)grid.y;
    p.nb_lo = a.nb_lo; p.nb_hi = a.nb_hi; p.yshift = deep;
    // my row r sits at out + pitch*(r + ng) (+ x); the lo neighbour's hi ghost row (a.peer_lo_plane) receives row 0, the hi
    // neighbour's lo ghost row (a.peer_hi_plane) receives row n1-1: element offsets between the two address spaces
    p.has_peer_lo = a.peer_lo_plane != nullptr; p.has_peer_hi = a.peer_hi_plane != nullptr;
    p.peer_lo_delta = p.has_peer_lo ? (long long)(((intptr_t)a.peer_lo_plane - (intptr_t)a.out) / 8) - g.pitch * g.ng : 0;
    p.peer_hi_delta = p.has_peer_hi ? (long long)(((intptr_t)a.peer_hi_plane - (intptr_t)a.out) / 8) - g.pitch * (g.n[1] - 1 + g.ng) : 0;
    p.ps = a.ps;
    auto kern = jacobi2d_tbk_kernel<K, NW, NS, MINB, DI>;
#ifndef RNMF_EMULATE
    static PerDeviceOnce attr_set;                      // one per instantiation
    if (!attr_set.done()) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set.mark();
    }
#endif
    RNMF_KERNEL_LAUNCH(kern, grid, Cfg::THREADS, Cfg::SMEM, s, tm_u256, tm_upad, tm_r256, tm_rpad, p);
    return 1;
}

}  // namespace

// Four sweeps (+ ghost fills) per launch, a.k_begin..a.k_end = output rows; every x / y face must carry a Dirichlet or Neumann
// rule (the caller checks).  512-column strips, 4 consumer warps, 8 stages, two CTAs per SM, de-interleaved ring rows (dense
// E/W loads).  Measured at 8192^2 (sustained, round 2): 636 GLUP/s; the other shapes of the template that were tried --
// K = 2 / 3 / 5 / 6, 768-column strips, interleaved rings -- gave 443 .. 614 and are no longer instantiated.
int launch_jacobi_multi_sweep_2d(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    return launch_tbk<4, 4, 8, 2, true>(a, s);
}

// CTAs per boundary side (strips across a row): the first / last chunk of the launch takes part in the handshake
int multi_sweep_2d_boundary_ctas(const SweepArgs& a)
{
    dim3 grid;
    int chunk;
    plan_tbk<4, 4, 8, 2, true>(a, grid, chunk);
    return (int)grid.x;
}
