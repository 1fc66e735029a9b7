// kernels_sweep_tb2.cu -- K1, temporally blocked: TWO Jacobi sweeps (each followed by the ng=1
// ghost fill, poisson.rs:80-89 + dataframe.rs:289-355) per pass over HBM.
//
// The single-sweep TMA kernel moves 24 B per update and sits at the DRAM limit of the part; the
// only way past that limit is fewer bytes per update.  Here a CTA streams psi (u0) and rhs once,
// forms u1 = sweep(u0) in shared memory and writes only u2 = sweep(u1): 24 B per TWO updates.
//
//   * CTA tile: 128 x TY cells of u2, a chunk of z-planes.  u1 is needed on the tile grown by one
//     cell in x and y (no corners) and by one plane in z, u0 on the tile grown by two.
//   * producer warp: per plane one TMA box of u0 ((128+4) x (TY+4), the two-cell x halo keeps the
//     interior 16-byte aligned in shared memory) and one of rhs ((128+4) x (TY+2)) into an NS-stage
//     ring (full/empty mbarriers), as in kernels_sweep_tma.cu; ring indices are run-time values.
//   * TY "row" warps own one tile row each (lane l: cell pairs 2l,2l+1 and 64+2l,65+2l), two more
//     row warps own the halo rows j0-1 and j0+TY (level 1 only), one warp owns the two halo columns
//     i0-1 and i0+128 (level 1 only).  Level 1 is register z-marched over u0 exactly like the
//     single-sweep kernel; its result goes to a 2-slot shared-memory ring of u1 planes, together
//     with the ghost values fill_bc would give u1 on the physical x/y faces.  Level 2 is register
//     z-marched over the lane's own u1 values (the z ghosts of u1 are formed in registers) and
//     reads the N/S/E/W neighbours of u1 from the ring.
//   * one wait phase per plane iteration: full[stage of the next u0 plane] and u1full[slot of u1
//     plane it-1] (one mbarrier per slot; every level-1 warp arrives after writing plane it into slot
//     it%2).  The second wait is the read-after-write guard of level 2 AND the write-after-read guard
//     of the slot level 1 is about to overwrite: a warp reads plane it-2 at the top of iteration it-1,
//     before it arrives for plane it-1 (release), so once all have arrived nobody reads it any more.
//   * behind the waits all shared-memory loads of both levels are issued together; level 2's x/y part
//     (c0*rhs + c1*(E+W)) + c2*(N+S) is formed BEFORE level 1 (it does not depend on the new u1
//     plane), its z term + c3*(U+D) is added once level 1 has produced that plane -- the
//     reference's association (poisson.rs:83-88), so u2 does not change.
//   * the main loop makes three plane iterations per trip: the three-plane register windows of both
//     levels rotate by renaming; level-1-only iterations (planes before the chunk's first output,
//     and the halo-row warps) and the level-2-only last plane of the slab are peeled.
//   * u2 is stored with the fused ghost fill of the output frame, as in the single-sweep kernel.
//
// Slabs (one GPU per z-slab): u2 of the slab's first/last plane needs the neighbour's u1 boundary
// plane.  "A" chunks cover the planes that do not (1..m-2); the A chunks next to a neighbour store
// u1 of plane 0 / m-1 into that neighbour's scratch plane (peer memory, NVLink) and publish the
// half step; one-plane "B" chunks, scheduled last, wait for the neighbour's half step, recompute
// u1 of their two planes, take the missing u1 plane from the scratch and store u2 both locally and
// into the neighbour's ghost plane, then publish the full step (see PeerSync in common.cuh).
//
// Arithmetic: the same explicitly rounded sequence as every other sweep kernel: u2 is bit-identical
// to two single sweeps.
#include <stdlib.h>

#include "tma_util.cuh"

namespace {
using namespace tmau;

struct Tb2Params {
    FrameGeom g;
    double* out;
    SweepCoef c;
    FaceSet faces;
    long long m_begin, m_end;    // output planes of the A chunks
    int chunk;
    int n_a;                     // A chunks: blockIdx.z < n_a; then the B chunks (lo first)
    int b_lo, b_hi;              // one-plane B chunk for the slab's first / last plane
    // slab neighbours (all null on a single GPU)
    double* peer_u1_lo;          // lo neighbour's scratch plane that receives my u1(0)
    double* peer_u1_hi;          // hi neighbour's scratch plane that receives my u1(m-1)
    const double* u1_in_lo;      // my scratch plane holding the lo neighbour's last u1 plane
    const double* u1_in_hi;      // my scratch plane holding the hi neighbour's first u1 plane
    double* peer_lo_plane;       // neighbour ghost planes of its `out` buffer (my u2(0) / u2(m-1))
    double* peer_hi_plane;
    PeerSync ps;                 // full-step flags / A-boundary counters
    const unsigned long long* half_wait_lo;   // my half-step flags, written by the neighbours' A chunks
    const unsigned long long* half_wait_hi;
    unsigned long long* half_sig_lo;          // neighbours' half-step flags (peer-mapped)
    unsigned long long* half_sig_hi;
    unsigned long long* done_b;               // [2] completion counters of the B chunks
    unsigned long long target_b_lo, target_b_hi;
};

template <int TY, int NS_>
struct Tb2Cfg {
    static constexpr int TX = 128;
    static constexpr int NS = NS_;                        // u0/rhs stages
    static constexpr int NU1 = 2;                         // u1 slots
    static constexpr int BX = TX + 4;
    static constexpr int U0_ROWS = TY + 4;
    static constexpr int R_ROWS = TY + 2;
    static constexpr int U0_BYTES = BX * U0_ROWS * 8;
    static constexpr int U0_STAGE = (U0_BYTES + 127) / 128 * 128;
    static constexpr int RHS_BYTES = BX * R_ROWS * 8;
    static constexpr int RHS_STAGE = (RHS_BYTES + 127) / 128 * 128;
    static constexpr int U1_SLOT = RHS_STAGE;
    static constexpr int RHS_OFF = NS * U0_STAGE;
    static constexpr int U1_OFF = RHS_OFF + NS * RHS_STAGE;
    static constexpr int BAR_OFF = U1_OFF + NU1 * U1_SLOT;
    static constexpr int SMEM = BAR_OFF + (2 * NS + NU1) * 8 + 128;
    static constexpr int NCONS = TY + 3;                  // TY row warps + 2 halo-row warps + 1 halo-column warp
    static constexpr int THREADS = 32 * (TY + 4);         // + producer warp
    static_assert(NS >= 3 && SMEM <= 227 * 1024, "shared-memory budget");
};

// per-thread constants of a row warp
struct RowK {
    uint32_t pa, ra, ua;                 // byte address of the lane's first cell in u0 stage 0 / rhs stage 0 / u1 slot 0
    uint32_t full0, empty0, u1full0;
    double c0, c1, c2, c3;
    bool full, va0, va1, vb0, vb1;       // validity of the lane's four cells
    bool do_xlo, any_x, any_y, do_ylo, do_yhi;
    int xhi_sel;                         // which of the four cells is i == n0-1 (-1: none)
    uint32_t xhi_off;                    // byte offset of the x-hi ghost from the lane's first cell
    Face1 fx0, fx1, fy0, fy1, fz0, fz1;
    long long pitch, plane_sz, xy_off;
    int lane;
    int it_zlo;                          // iteration whose level 2 produces the slab's first plane (-1: not in this chunk)
    int it_u1lo, it_u1hi;                // level-1 iterations whose plane goes to a slab neighbour (-1: none)
};

// rotating per-thread state of a row warp
struct RowSt {
    double2 below_a, below_b, cur_a, cur_b;      // u0 of planes a+it-1, a+it        (own cells)
    double2 u1lo_a, u1lo_b, u1c_a, u1c_b;        // u1 of planes q-1, q, q = a+it-1  (own cells)
    double2 rp_a, rp_b;                          // rhs of plane q
    double* oa;                                  // where u2 of plane q goes
    uint32_t sc, sn, phn;                        // ring: stage of plane a+it, of plane a+it+1, parity its full barrier completes with
};

// first two terms of poisson.rs:83-88 (the z term is added last: same association as upd3)
__device__ __forceinline__ double part2(double c0, double c1, double c2, double r, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, __dadd_rn(e, w))), __dmul_rn(c2, __dadd_rn(n, s)));
}
__device__ __forceinline__ double fin2(double part, double c3, double u, double d) { return __dadd_rn(part, __dmul_rn(c3, __dadd_rn(u, d))); }

__device__ __forceinline__ void load4(const double* q, double2& a, double2& b, const RowK& k)
{
    if (k.full) { a = *reinterpret_cast<const double2*>(q); b = *reinterpret_cast<const double2*>(q + 64); }
    else {
        if (k.va0) a.x = q[0];
        if (k.va1) a.y = q[1];
        if (k.vb0) b.x = q[64];
        if (k.vb1) b.y = q[65];
    }
}

// One plane iteration of a row warp: level 1 of plane a+it (DO_L1) and level 2 of plane q = a+it-1 (DO_L2).
// All shared-memory loads of both levels are issued behind ONE wait phase; level 2's x/y part is formed
// before level 1's result exists (part2), its z term is added once u1 of plane a+it is known (fin2).
// EDGE (specialised kernels only): 2 = general; 0 = the warp owns no cell next to an x / y face, no cell past the frame, and the
// iteration meets neither the z-lo face nor a slab neighbour -- the ghost code and the store masks are compiled out;
// 1 = as 0, but some lane sits next to an x face of a whole (even-width) tile: only the x-ghost code stays.
template <int TY, int NS, bool DO_L1, bool DO_L2, int EDGE = 2>
__device__ __forceinline__ void row_iter(const Tb2Params& p, const RowK& k, RowSt& st, const int it)
{
    using Cfg = Tb2Cfg<TY, NS>;
    constexpr int BX = Cfg::BX;
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 above_a = zero2, above_b = zero2, u1n_a = zero2, u1n_b = zero2, rc_a = zero2, rc_b = zero2, P_a = zero2, P_b = zero2;
    double2 n_a, n_b, s_a, s_b;
    double w_a0, e_a1, w_b0, e_b1;

    // ---- wait phase --------------------------------------------------------------------------
    if (DO_L1) mbar_wait(k.full0 + 8 * st.sn, st.phn);                    // u0 plane a+it+1 (plane a+it was waited for one iteration ago)
    // u1 plane a+it-1 complete: level 2 reads it; and every warp has finished reading plane a+it-2,
    // whose slot level 1 overwrites below (reads precede the arrival in program order)
    if (DO_L2 || it >= 2) mbar_wait(k.u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);

    // ---- loads ---------------------------------------------------------------------------------
    if (DO_L2) {
        const uint32_t up = k.ua + (uint32_t)((it - 1) & 1) * Cfg::U1_SLOT;
        const double2 un_a = lds2(up + BX * 8), un_b = lds2(up + BX * 8 + 512);
        const double2 us_a = lds2(up - BX * 8), us_b = lds2(up - BX * 8 + 512);
        const double uw_a0 = lds1(up - 8), ue_a1 = lds1(up + 16);
        const double uw_b0 = lds1(up + 512 - 8), ue_b1 = lds1(up + 512 + 16);
        P_a.x = part2(k.c0, k.c1, k.c2, st.rp_a.x, st.u1c_a.y, uw_a0, un_a.x, us_a.x);
        P_a.y = part2(k.c0, k.c1, k.c2, st.rp_a.y, ue_a1, st.u1c_a.x, un_a.y, us_a.y);
        P_b.x = part2(k.c0, k.c1, k.c2, st.rp_b.x, st.u1c_b.y, uw_b0, un_b.x, us_b.x);
        P_b.y = part2(k.c0, k.c1, k.c2, st.rp_b.y, ue_b1, st.u1c_b.x, un_b.y, us_b.y);
    }
    if (DO_L1) {
        const uint32_t pn = k.pa + st.sn * Cfg::U0_STAGE, pc = k.pa + st.sc * Cfg::U0_STAGE, pr = k.ra + st.sc * Cfg::RHS_STAGE;
        above_a = lds2(pn); above_b = lds2(pn + 512);
        n_a = lds2(pc + BX * 8); n_b = lds2(pc + BX * 8 + 512);
        s_a = lds2(pc - BX * 8); s_b = lds2(pc - BX * 8 + 512);
        rc_a = lds2(pr); rc_b = lds2(pr + 512);
        w_a0 = lds1(pc - 8); e_a1 = lds1(pc + 16);
        w_b0 = lds1(pc + 512 - 8); e_b1 = lds1(pc + 512 + 16);

        // ---------------- level 1: u1 of plane a+it ----------------
        u1n_a.x = upd3(k.c0, k.c1, k.c2, k.c3, rc_a.x, st.cur_a.y, w_a0, n_a.x, s_a.x, above_a.x, st.below_a.x);
        u1n_a.y = upd3(k.c0, k.c1, k.c2, k.c3, rc_a.y, e_a1, st.cur_a.x, n_a.y, s_a.y, above_a.y, st.below_a.y);
        u1n_b.x = upd3(k.c0, k.c1, k.c2, k.c3, rc_b.x, st.cur_b.y, w_b0, n_b.x, s_b.x, above_b.x, st.below_b.x);
        u1n_b.y = upd3(k.c0, k.c1, k.c2, k.c3, rc_b.y, e_b1, st.cur_b.x, n_b.y, s_b.y, above_b.y, st.below_b.y);
        // odd n0: the last valid cell is the first of a pair, whose E neighbour at level 2 is the pair's
        // second register: it must hold the x-hi ghost of u1, not a stencil value
        if (EDGE == 2 && k.xhi_sel == 0) u1n_a.y = ghost1(k.fx1, u1n_a.x);
        if (EDGE == 2 && k.xhi_sel == 2) u1n_b.y = ghost1(k.fx1, u1n_b.x);

        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.empty0 + 8 * st.sc);              // this warp is done with the stage of plane a+it
        const uint32_t us = k.ua + (uint32_t)(it & 1) * Cfg::U1_SLOT;
        sstore4(us, u1n_a, u1n_b, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        // what fill_bc gives u1 on the physical x / y faces (read by level 2 of the edge cells)
        if (EDGE != 0 && k.any_x) {
            if (k.do_xlo) sts1(us - 8, ghost1(k.fx0, u1n_a.x));
            if (k.xhi_sel >= 0) {
                const double m = k.xhi_sel == 0 ? u1n_a.x : k.xhi_sel == 1 ? u1n_a.y : k.xhi_sel == 2 ? u1n_b.x : u1n_b.y;
                sts1(us + k.xhi_off, ghost1(k.fx1, m));
            }
        }
        if (EDGE == 2 && k.any_y) {
            if (k.do_ylo) sstore4(us - BX * 8, ghost1(k.fy0, u1n_a), ghost1(k.fy0, u1n_b), EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
            if (k.do_yhi) sstore4(us + BX * 8, ghost1(k.fy1, u1n_a), ghost1(k.fy1, u1n_b), EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.u1full0 + 8 * (it & 1));
        // slab neighbours get u1 of my boundary plane (half step) over NVLink
        if (EDGE == 2 && it == k.it_u1lo) store4(p.peer_u1_lo + k.xy_off, u1n_a, u1n_b, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (EDGE == 2 && it == k.it_u1hi) store4(p.peer_u1_hi + k.xy_off, u1n_a, u1n_b, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
    }
    if (DO_L2) {
        // ---------------- level 2: u2 of plane q = a+it-1 ----------------
        // z neighbours of u1: registers, the face's ghost rule, or the neighbour slab's plane
        double2 dn_a = st.u1lo_a, dn_b = st.u1lo_b, up_a = u1n_a, up_b = u1n_b;
        const bool zlo = EDGE == 2 && it == k.it_zlo;
        if (zlo) {
            if (k.fz0.on) { dn_a = ghost1(k.fz0, st.u1c_a); dn_b = ghost1(k.fz0, st.u1c_b); }
            else if (p.u1_in_lo) load4(p.u1_in_lo + k.xy_off, dn_a, dn_b, k);
        }
        if (!DO_L1) {                                                     // the slab's last plane (no plane above)
            if (k.fz1.on) { up_a = ghost1(k.fz1, st.u1c_a); up_b = ghost1(k.fz1, st.u1c_b); }
            else if (p.u1_in_hi) load4(p.u1_in_hi + k.xy_off, up_a, up_b, k);
        }
        double2 na, nb;
        na.x = fin2(P_a.x, k.c3, up_a.x, dn_a.x);
        na.y = fin2(P_a.y, k.c3, up_a.y, dn_a.y);
        nb.x = fin2(P_b.x, k.c3, up_b.x, dn_b.x);
        nb.y = fin2(P_b.y, k.c3, up_b.y, dn_b.y);

        double* oa = st.oa;
        store4(oa, na, nb, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (EDGE != 0 && k.any_x) {
            if (k.do_xlo) oa[-1] = ghost1(k.fx0, na.x);
            if (k.xhi_sel >= 0) {
                const double m = k.xhi_sel == 0 ? na.x : k.xhi_sel == 1 ? na.y : k.xhi_sel == 2 ? nb.x : nb.y;
                oa[(k.xhi_sel >= 2 ? 64 : 0) + (k.xhi_sel & 1) + 1] = ghost1(k.fx1, m);
            }
        }
        if (EDGE == 2 && k.any_y) {
            if (k.do_ylo) store4(oa - k.pitch, ghost1(k.fy0, na), ghost1(k.fy0, nb), EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
            if (k.do_yhi) store4(oa + k.pitch, ghost1(k.fy1, na), ghost1(k.fy1, nb), EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        }
        if (zlo) {
            if (k.fz0.on) store4(oa - k.plane_sz, ghost1(k.fz0, na), ghost1(k.fz0, nb), EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
            if (p.peer_lo_plane) store4(p.peer_lo_plane + k.xy_off, na, nb, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        }
        if (!DO_L1) {
            if (k.fz1.on) store4(oa + k.plane_sz, ghost1(k.fz1, na), ghost1(k.fz1, nb), EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
            if (p.peer_hi_plane) store4(p.peer_hi_plane + k.xy_off, na, nb, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        }
        st.oa = oa + k.plane_sz;
    }
    // ---- rotate --------------------------------------------------------------------------------
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

// ===============================================================================================
// QUAD mapping of the row warps (template flag QUAD of the kernel below): a warp owns TWO adjacent tile rows x 64 columns
// instead of one row x 128 columns; lane l holds the 2 x 2 patch {rows jA, jA+1} x {columns ia, ia+1}, ia = i0 + 64*xh + 2l.
// Still four cells per lane and level (the same register budget, the same 16 warps), but the N neighbour of row jA is the
// lane's own row jA+1 and vice versa, at both levels: per warp and plane the N/S loads cost 8 + 8 instead of 16 + 16
// shared-memory wavefronts, 72 instead of 88 in total (the shared-memory pipe bounds this kernel: DESIGN.md section 4).
// Ring protocol, arithmetic and association are those of row_iter: the result is bit-identical.
// ===============================================================================================
struct QuadK {
    uint32_t pa, ra, ua;                 // byte address of the lane's first cell (row A) in u0 stage 0 / rhs stage 0 / u1 slot 0
    uint32_t full0, empty0, u1full0;
    double c0, c1, c2, c3;
    bool full, va0, va1, vb0, vb1;       // validity: row A cells ia, ia+1; row B cells ia, ia+1
    bool do_xlo, any_x, any_y, do_ylo;
    int xhi_sel;                         // 0 / 1: cell ia / ia+1 is i == n0-1 (-1: neither)
    int yhi_row;                         // 1: row B is j == n1-1; 0: row A is (row B lies outside the frame); -1: neither
    uint32_t xhi_off;                    // byte offset of the x-hi ghost from the lane's first cell
    Face1 fx0, fx1, fy0, fy1, fz0, fz1;
    long long pitch, plane_sz, xy_off;
    int lane;
    int it_zlo, it_u1lo, it_u1hi;
};

// the lane's 2 x 2 patch: pair `a` at q, pair `b` one row (rs elements) further
__device__ __forceinline__ void qstore(double* q, long long rs, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
#ifdef RNMF_EMULATE
    if ((full || a1 || b1) && (reinterpret_cast<uintptr_t>(q) % 16 != 0 || (rs & 1))) emu::fault("misaligned 16-byte global store", reinterpret_cast<uintptr_t>(q), 16);
#endif
    if (full) {
        *reinterpret_cast<double2*>(q) = va;
        *reinterpret_cast<double2*>(q + rs) = vb;
    } else {
        if (a1) *reinterpret_cast<double2*>(q) = va; else if (a0) q[0] = va.x;
        if (b1) *reinterpret_cast<double2*>(q + rs) = vb; else if (b0) q[rs] = vb.x;
    }
}
__device__ __forceinline__ void pstore(double* q, const double2& v, bool full, bool c0, bool c1)
{
    if (full || c1) *reinterpret_cast<double2*>(q) = v; else if (c0) q[0] = v.x;
}
__device__ __forceinline__ void qsstore(uint32_t a, uint32_t rb, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
    if (full) {
        sts2(a, va);
        sts2(a + rb, vb);
    } else {
        if (a1) sts2(a, va); else if (a0) sts1(a, va.x);
        if (b1) sts2(a + rb, vb); else if (b0) sts1(a + rb, vb.x);
    }
}
__device__ __forceinline__ void psstore(uint32_t a, const double2& v, bool full, bool c0, bool c1)
{
    if (full || c1) sts2(a, v); else if (c0) sts1(a, v.x);
}
__device__ __forceinline__ void qload(const double* q, long long rs, double2& a, double2& b, const QuadK& k)
{
    if (k.full) { a = *reinterpret_cast<const double2*>(q); b = *reinterpret_cast<const double2*>(q + rs); }
    else {
        if (k.va0) a.x = q[0];
        if (k.va1) a.y = q[1];
        if (k.vb0) b.x = q[rs];
        if (k.vb1) b.y = q[rs + 1];
    }
}

// One plane iteration of a quad warp; same structure, template flags and EDGE levels as row_iter.
template <int TY, int NS, bool DO_L1, bool DO_L2, int EDGE = 2>
__device__ __forceinline__ void quad_iter(const Tb2Params& p, const QuadK& k, RowSt& st, const int it)
{
    using Cfg = Tb2Cfg<TY, NS>;
    constexpr uint32_t RB = Cfg::BX * 8;                                  // one tile row in shared memory
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 above_a = zero2, above_b = zero2, u1n_a = zero2, u1n_b = zero2, rc_a = zero2, rc_b = zero2, P_a = zero2, P_b = zero2;

    // ---- wait phase (as row_iter) ----------------------------------------------------------------
    if (DO_L1) mbar_wait(k.full0 + 8 * st.sn, st.phn);
    if (DO_L2 || it >= 2) mbar_wait(k.u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);

    // ---- loads -----------------------------------------------------------------------------------
    if (DO_L2) {
        const uint32_t up = k.ua + (uint32_t)((it - 1) & 1) * Cfg::U1_SLOT;
        const double2 un_b = lds2(up + 2 * RB);                           // N of row B; N of row A is row B (registers)
        const double2 us_a = lds2(up - RB);                               // S of row A; S of row B is row A
        const double uw_a0 = lds1(up - 8), ue_a1 = lds1(up + 16);
        const double uw_b0 = lds1(up + RB - 8), ue_b1 = lds1(up + RB + 16);
        P_a.x = part2(k.c0, k.c1, k.c2, st.rp_a.x, st.u1c_a.y, uw_a0, st.u1c_b.x, us_a.x);
        P_a.y = part2(k.c0, k.c1, k.c2, st.rp_a.y, ue_a1, st.u1c_a.x, st.u1c_b.y, us_a.y);
        P_b.x = part2(k.c0, k.c1, k.c2, st.rp_b.x, st.u1c_b.y, uw_b0, un_b.x, st.u1c_a.x);
        P_b.y = part2(k.c0, k.c1, k.c2, st.rp_b.y, ue_b1, st.u1c_b.x, un_b.y, st.u1c_a.y);
    }
    if (DO_L1) {
        const uint32_t pn = k.pa + st.sn * Cfg::U0_STAGE, pc = k.pa + st.sc * Cfg::U0_STAGE, pr = k.ra + st.sc * Cfg::RHS_STAGE;
        above_a = lds2(pn); above_b = lds2(pn + RB);
        const double2 n_b = lds2(pc + 2 * RB);
        const double2 s_a = lds2(pc - RB);
        rc_a = lds2(pr); rc_b = lds2(pr + RB);
        const double w_a0 = lds1(pc - 8), e_a1 = lds1(pc + 16);
        const double w_b0 = lds1(pc + RB - 8), e_b1 = lds1(pc + RB + 16);

        // ---------------- level 1: u1 of plane a+it ----------------
        u1n_a.x = upd3(k.c0, k.c1, k.c2, k.c3, rc_a.x, st.cur_a.y, w_a0, st.cur_b.x, s_a.x, above_a.x, st.below_a.x);
        u1n_a.y = upd3(k.c0, k.c1, k.c2, k.c3, rc_a.y, e_a1, st.cur_a.x, st.cur_b.y, s_a.y, above_a.y, st.below_a.y);
        u1n_b.x = upd3(k.c0, k.c1, k.c2, k.c3, rc_b.x, st.cur_b.y, w_b0, n_b.x, st.cur_a.x, above_b.x, st.below_b.x);
        u1n_b.y = upd3(k.c0, k.c1, k.c2, k.c3, rc_b.y, e_b1, st.cur_b.x, n_b.y, st.cur_a.y, above_b.y, st.below_b.y);
        // registers that level 2 reads as a neighbour but that are not stencil values of the frame:
        // odd n0: the pair's second register must hold the x-hi ghost of u1; odd row count: row B must hold the y-hi ghost row
        if (EDGE == 2 && k.xhi_sel == 0) { u1n_a.y = ghost1(k.fx1, u1n_a.x); u1n_b.y = ghost1(k.fx1, u1n_b.x); }
        if (EDGE == 2 && k.yhi_row == 0) u1n_b = ghost1(k.fy1, u1n_a);

        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.empty0 + 8 * st.sc);
        const uint32_t us = k.ua + (uint32_t)(it & 1) * Cfg::U1_SLOT;
        qsstore(us, RB, u1n_a, u1n_b, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (EDGE != 0 && k.any_x) {
            if (k.do_xlo) {
                sts1(us - 8, ghost1(k.fx0, u1n_a.x));
                if (EDGE != 2 || k.vb0) sts1(us + RB - 8, ghost1(k.fx0, u1n_b.x));
            }
            if (k.xhi_sel >= 0) {
                sts1(us + k.xhi_off, ghost1(k.fx1, k.xhi_sel == 0 ? u1n_a.x : u1n_a.y));
                if (EDGE != 2 || k.vb0) sts1(us + RB + k.xhi_off, ghost1(k.fx1, k.xhi_sel == 0 ? u1n_b.x : u1n_b.y));
            }
        }
        if (EDGE == 2 && k.any_y) {
            if (k.do_ylo) psstore(us - RB, ghost1(k.fy0, u1n_a), k.full, k.va0, k.va1);
            if (k.yhi_row == 1) psstore(us + 2 * RB, ghost1(k.fy1, u1n_b), k.full, k.vb0, k.vb1);
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.u1full0 + 8 * (it & 1));
        if (EDGE == 2 && it == k.it_u1lo) qstore(p.peer_u1_lo + k.xy_off, k.pitch, u1n_a, u1n_b, k.full, k.va0, k.va1, k.vb0, k.vb1);
        if (EDGE == 2 && it == k.it_u1hi) qstore(p.peer_u1_hi + k.xy_off, k.pitch, u1n_a, u1n_b, k.full, k.va0, k.va1, k.vb0, k.vb1);
    }
    if (DO_L2) {
        // ---------------- level 2: u2 of plane q = a+it-1 ----------------
        double2 dn_a = st.u1lo_a, dn_b = st.u1lo_b, up_a = u1n_a, up_b = u1n_b;
        const bool zlo = EDGE == 2 && it == k.it_zlo;
        if (zlo) {
            if (k.fz0.on) { dn_a = ghost1(k.fz0, st.u1c_a); dn_b = ghost1(k.fz0, st.u1c_b); }
            else if (p.u1_in_lo) qload(p.u1_in_lo + k.xy_off, k.pitch, dn_a, dn_b, k);
        }
        if (!DO_L1) {
            if (k.fz1.on) { up_a = ghost1(k.fz1, st.u1c_a); up_b = ghost1(k.fz1, st.u1c_b); }
            else if (p.u1_in_hi) qload(p.u1_in_hi + k.xy_off, k.pitch, up_a, up_b, k);
        }
        double2 na, nb;
        na.x = fin2(P_a.x, k.c3, up_a.x, dn_a.x);
        na.y = fin2(P_a.y, k.c3, up_a.y, dn_a.y);
        nb.x = fin2(P_b.x, k.c3, up_b.x, dn_b.x);
        nb.y = fin2(P_b.y, k.c3, up_b.y, dn_b.y);

        double* oa = st.oa;
        const bool full = EDGE != 2 || k.full;
        qstore(oa, k.pitch, na, nb, full, k.va0, k.va1, k.vb0, k.vb1);
        if (EDGE != 0 && k.any_x) {
            if (k.do_xlo) {
                oa[-1] = ghost1(k.fx0, na.x);
                if (EDGE != 2 || k.vb0) oa[k.pitch - 1] = ghost1(k.fx0, nb.x);
            }
            if (k.xhi_sel >= 0) {
                oa[k.xhi_sel + 1] = ghost1(k.fx1, k.xhi_sel == 0 ? na.x : na.y);
                if (EDGE != 2 || k.vb0) oa[k.pitch + k.xhi_sel + 1] = ghost1(k.fx1, k.xhi_sel == 0 ? nb.x : nb.y);
            }
        }
        if (EDGE == 2 && k.any_y) {
            if (k.do_ylo) pstore(oa - k.pitch, ghost1(k.fy0, na), k.full, k.va0, k.va1);
            if (k.yhi_row == 1) pstore(oa + 2 * k.pitch, ghost1(k.fy1, nb), k.full, k.vb0, k.vb1);
            if (k.yhi_row == 0) pstore(oa + k.pitch, ghost1(k.fy1, na), false, k.va0, k.va1);
        }
        if (zlo) {
            if (k.fz0.on) qstore(oa - k.plane_sz, k.pitch, ghost1(k.fz0, na), ghost1(k.fz0, nb), full, k.va0, k.va1, k.vb0, k.vb1);
            if (p.peer_lo_plane) qstore(p.peer_lo_plane + k.xy_off, k.pitch, na, nb, full, k.va0, k.va1, k.vb0, k.vb1);
        }
        if (!DO_L1) {
            if (k.fz1.on) qstore(oa + k.plane_sz, k.pitch, ghost1(k.fz1, na), ghost1(k.fz1, nb), full, k.va0, k.va1, k.vb0, k.vb1);
            if (p.peer_hi_plane) qstore(p.peer_hi_plane + k.xy_off, k.pitch, na, nb, full, k.va0, k.va1, k.vb0, k.vb1);
        }
        st.oa = oa + k.plane_sz;
    }
    // ---- rotate ----------------------------------------------------------------------------------
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

// SPEC (variant 26, opt-in): row warps that own no face / ragged cell run an interior-only copy of the plane loop
// QUAD: the TY tile-row warps use the 2-rows x 64-columns mapping (quad_iter above); the halo warps are unchanged
template <int TY, int NS, bool SPEC = false, bool QUAD = false>
__global__ void __launch_bounds__(32 * (TY + 4), 1) jacobi3d_tb2_kernel(const __grid_constant__ CUtensorMap tm_u0,
                                                                         const __grid_constant__ CUtensorMap tm_rhs,
                                                                         const __grid_constant__ Tb2Params p)
{
    using Cfg = Tb2Cfg<TY, NS>;
    constexpr int TX = Cfg::TX, BX = Cfg::BX;
    static_assert(TY <= 16, "the halo-column warp maps 16 rows per side");
    static_assert(!QUAD || TY % 2 == 0, "QUAD: row pairs");
    RNMF_DYNAMIC_SMEM(smem_raw);
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t u0_0 = sb, rhs_0 = sb + Cfg::RHS_OFF, u1_0 = sb + Cfg::U1_OFF;
    const uint32_t full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8, u1full0 = empty0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = p.g.n[0], n1 = p.g.n[1], n2 = p.g.n[2];
    const long long i0 = (long long)blockIdx.x * TX;
    const long long j0 = (long long)blockIdx.y * TY;

    // ---- which planes: A chunk (blockIdx.z < n_a) or one-plane B chunk -------------------------
    const int bz = (int)blockIdx.z;
    const bool is_b = bz >= p.n_a;
    long long k0, k1;
    if (!is_b) {
        const unsigned zc = p.ps.enabled ? peer_chunk_order((unsigned)bz, (unsigned)p.n_a) : (unsigned)bz;
        k0 = p.m_begin + (long long)zc * p.chunk;
        k1 = k0 + p.chunk < p.m_end ? k0 + p.chunk : p.m_end;
    } else {
        k0 = (p.b_lo && bz == p.n_a) ? 0 : n2 - 1;
        k1 = k0 + 1;
    }
    const long long a = k0 > 0 ? k0 - 1 : 0;              // level-1 planes a..b
    const long long b = k1 < n2 ? k1 : n2 - 1;
    const int nL1 = (int)(b - a) + 1;
    const int n_planes = nL1 + 2;                         // u0 planes a-1..b+1  (plane a-1+idx <-> stage idx % NS)
    const bool top = b < k1;                              // the chunk ends with the slab's last plane: one level-2-only iteration
    const int it_l2_first = (int)(k0 - a) + 1;            // iteration it: level 1 of plane a+it, level 2 of plane a+it-1

    // slab handshake roles of this CTA
    const bool a_sync_lo = !is_b && p.peer_u1_lo && k0 == p.m_begin;     // computes u1(0), reads ghost plane -1 of u0
    const bool a_sync_hi = !is_b && p.peer_u1_hi && k1 == p.m_end;
    const bool b_is_lo = is_b && k0 == 0 && p.b_lo;
    const bool b_is_hi = is_b && !b_is_lo;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, Cfg::NCONS);
        }
        mbar_init(u1full0, Cfg::NCONS);
        mbar_init(u1full0 + 8, Cfg::NCONS);
        mbar_fence_init();
        if (a_sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (a_sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
        if (b_is_lo) peer_spin_wait(p.half_wait_lo, p.ps.epoch + 1, p.ps.status);
        if (b_is_hi) peer_spin_wait(p.half_wait_hi, p.ps.epoch + 1, p.ps.status);
    }
    __syncthreads();

    if (warp == TY + 3) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx = (int)(p.g.xoff + p.g.ng + i0 - 2);
            const int cy_u0 = (int)(p.g.ng + j0 - 2);
            const int cy_rhs = cy_u0 + 1;
            int cz = (int)(p.g.kg + a - 1);
            uint32_t s = 0, wrap = 0;
            for (int idx = 0; idx < n_planes; ++idx) {
                if (wrap > 0) mbar_wait(empty0 + 8 * s, (wrap - 1) & 1u);
                const bool with_rhs = idx >= 1 && idx <= nL1;
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(Cfg::U0_BYTES + (with_rhs ? Cfg::RHS_BYTES : 0)));
                tma_load_3d(u0_0 + s * Cfg::U0_STAGE, &tm_u0, cx, cy_u0, cz, full0 + 8 * s);
                if (with_rhs) tma_load_3d(rhs_0 + s * Cfg::RHS_STAGE, &tm_rhs, cx, cy_rhs, cz, full0 + 8 * s);
                ++cz;
                if (++s == NS) { s = 0; ++wrap; }
            }
        }
    } else if (warp == TY + 2) {
        // ------------------------------ halo-column warp (level 1 only) ------------------------------
        // lanes 0..15: column i0-1, rows j0+lane; lanes 16..31: column i0+128, rows j0+lane-16
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2, c3 = p.c.c3;
        const int side = lane >> 4, r = lane & 15;
        const bool act = r < TY;
        const int rc = act ? r : 0;
        const int col = side ? TX : -1;
        const long long i = i0 + col, j = j0 + rc;
        const bool ok = act && i >= 0 && i < n0 && j < n1;
        const uint32_t hp = u0_0 + (uint32_t)(((rc + 2) * BX + col + 2) * 8);
        const uint32_t hr = rhs_0 + (uint32_t)(((rc + 1) * BX + col + 2) * 8);
        const uint32_t hu = u1_0 + (uint32_t)(((rc + 1) * BX + col + 2) * 8);

        mbar_wait(full0, 0);
        double below = lds1(hp);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        double cur = lds1(hp + Cfg::U0_STAGE);
        uint32_t sc = 1, sn = 2, phn = 0;
        for (int it = 0; it < nL1; ++it) {
            mbar_wait(full0 + 8 * sn, phn);
            if (it >= 2) mbar_wait(u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
            const uint32_t pc = hp + sc * Cfg::U0_STAGE;
            const double above = lds1(hp + sn * Cfg::U0_STAGE);
            const double nn = lds1(pc + BX * 8), ss = lds1(pc - BX * 8);
            const double ee = lds1(pc + 8), ww = lds1(pc - 8);
            const double rr = lds1(hr + sc * Cfg::RHS_STAGE);
            const double v = upd3(c0, c1, c2, c3, rr, ee, ww, nn, ss, above, below);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * sc);
            if (ok) sts1(hu + (uint32_t)(it & 1) * Cfg::U1_SLOT, v);
            __syncwarp();
            if (lane == 0) mbar_arrive(u1full0 + 8 * (it & 1));
            below = cur;
            cur = above;
            sc = sn;
            if (++sn == NS) { sn = 0; phn ^= 1u; }
        }
    } else if (QUAD && warp < TY) {
        // ------------------------------ quad warps: rows j0+2*rp, j0+2*rp+1 x columns i0+64*xh .. +63 ------------------------------
        const int xh = warp & 1, rp = warp >> 1;
        const int ra = 2 * rp;                                  // tile row of row A
        const long long jA = j0 + ra, jB = jA + 1;
        const bool okA = jA < n1, okB = jB < n1;
        const long long ia = i0 + 64 * xh + 2 * lane;
        QuadK k;
        k.va0 = okA && ia < n0; k.va1 = okA && ia + 1 < n0;
        k.vb0 = okB && ia < n0; k.vb1 = okB && ia + 1 < n0;
        k.full = k.vb1;
        const int col = 2 + 64 * xh + 2 * lane;                 // position of cell ia in a tile row (two-cell x halo)
        k.pa = u0_0 + (uint32_t)(((ra + 2) * BX + col) * 8);
        k.ra = rhs_0 + (uint32_t)(((ra + 1) * BX + col) * 8);
        k.ua = u1_0 + (uint32_t)(((ra + 1) * BX + col) * 8);
        k.full0 = full0; k.empty0 = empty0; k.u1full0 = u1full0;
        k.c0 = p.c.c0; k.c1 = p.c.c1; k.c2 = p.c.c2; k.c3 = p.c.c3;
        k.pitch = p.g.pitch; k.plane_sz = p.g.plane;
        k.xy_off = p.g.xoff + (ia + p.g.ng) + k.pitch * (jA + p.g.ng);
        k.lane = lane;
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fy0 = make_face1(p.faces.f[2], true); k.fy1 = make_face1(p.faces.f[3], true);
        k.fz0 = make_face1(p.faces.f[4], true); k.fz1 = make_face1(p.faces.f[5], true);
        k.do_xlo = k.fx0.on && k.va0 && ia == 0;
        const long long last = n0 - 1;
        k.xhi_sel = !k.fx1.on ? -1 : (k.va0 && ia == last) ? 0 : (k.va1 && ia + 1 == last) ? 1 : -1;
        k.xhi_off = (uint32_t)((k.xhi_sel + 1) * 8);
        k.do_ylo = k.fy0.on && jA == 0;
        k.yhi_row = !k.fy1.on ? -1 : (jB == n1 - 1) ? 1 : (jA == n1 - 1) ? 0 : -1;
        k.any_x = k.do_xlo || k.xhi_sel >= 0;
        k.any_y = k.do_ylo || k.yhi_row >= 0;
        k.it_zlo = k0 == 0 ? 1 : -1;
        k.it_u1lo = a_sync_lo ? 0 : -1;
        k.it_u1hi = a_sync_hi ? nL1 - 1 : -1;

        RowSt st;
        const double2 zero2 = make_double2(0.0, 0.0);
        constexpr uint32_t RB = BX * 8;
        mbar_wait(full0, 0);
        st.below_a = lds2(k.pa); st.below_b = lds2(k.pa + RB);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        st.cur_a = lds2(k.pa + Cfg::U0_STAGE); st.cur_b = lds2(k.pa + Cfg::U0_STAGE + RB);
        st.u1lo_a = zero2; st.u1lo_b = zero2; st.u1c_a = zero2; st.u1c_b = zero2; st.rp_a = zero2; st.rp_b = zero2;
        st.oa = p.out + p.g.at(ia, jA, k0, 0);
        st.sc = 1; st.sn = 2; st.phn = 0;

        int it = 0;
        // warp-uniform: both rows and all 64 columns are valid cells, no face of the frame touches the warp's patch, no slab role
        const long long iw = i0 + 64 * xh;
        const bool whole = okB && iw + 64 <= n0 && !k.do_ylo && k.yhi_row < 0 && !a_sync_lo && !a_sync_hi && !is_b;
        const bool interior = SPEC && whole && !(k.fx0.on && iw == 0) && !(k.fx1.on && iw + 64 == n0);
        const bool xface_only = SPEC && whole && !interior;
        if (interior) {
            for (; it < it_l2_first; ++it) quad_iter<TY, NS, true, false, 0>(p, k, st, it);
            if (it < nL1) { quad_iter<TY, NS, true, true, 2>(p, k, st, it); ++it; }      // may meet the z-lo face: general code, once
            for (; it < nL1; it += 3) {
                quad_iter<TY, NS, true, true, 0>(p, k, st, it);
                if (it + 1 >= nL1) break;
                quad_iter<TY, NS, true, true, 0>(p, k, st, it + 1);
                if (it + 2 >= nL1) break;
                quad_iter<TY, NS, true, true, 0>(p, k, st, it + 2);
            }
        } else if (xface_only) {
            for (; it < it_l2_first; ++it) quad_iter<TY, NS, true, false, 1>(p, k, st, it);
            if (it < nL1) { quad_iter<TY, NS, true, true, 2>(p, k, st, it); ++it; }
            for (; it < nL1; it += 3) {
                quad_iter<TY, NS, true, true, 1>(p, k, st, it);
                if (it + 1 >= nL1) break;
                quad_iter<TY, NS, true, true, 1>(p, k, st, it + 1);
                if (it + 2 >= nL1) break;
                quad_iter<TY, NS, true, true, 1>(p, k, st, it + 2);
            }
        } else {
            for (; it < it_l2_first; ++it) quad_iter<TY, NS, true, false, 2>(p, k, st, it);
            for (; it < nL1; it += 3) {
                quad_iter<TY, NS, true, true, 2>(p, k, st, it);
                if (it + 1 >= nL1) break;
                quad_iter<TY, NS, true, true, 2>(p, k, st, it + 1);
                if (it + 2 >= nL1) break;
                quad_iter<TY, NS, true, true, 2>(p, k, st, it + 2);
            }
        }
        if (top) quad_iter<TY, NS, false, true, 2>(p, k, st, nL1);
    } else {
        // ------------------------------ row warps ------------------------------
        // warps 0..TY-1: tile rows (levels 1 and 2); warp TY: halo row j0-1; warp TY+1: halo row j0+TY (level 1 only)
        // (QUAD: only the two halo-row warps come here)
        const bool do_l2 = warp < TY;
        const int rr = do_l2 ? warp : (warp == TY ? -1 : TY);
        const long long j = j0 + rr;
        const bool row_ok = j >= 0 && j < n1;
        const long long ia = i0 + 2 * lane, ib = ia + 64;
        RowK k;
        k.va0 = row_ok && ia < n0; k.va1 = row_ok && ia + 1 < n0;
        k.vb0 = row_ok && ib < n0; k.vb1 = row_ok && ib + 1 < n0;
        k.full = k.vb1;
        // byte offsets: cell (i0+c, j0+r) sits at ((r+2)*BX + c+2)*8 in a u0 stage and at ((r+1)*BX + c+2)*8
        // in an rhs stage / u1 slot
        k.pa = u0_0 + (uint32_t)(((rr + 2) * BX + 2 + 2 * lane) * 8);
        k.ra = rhs_0 + (uint32_t)(((rr + 1) * BX + 2 + 2 * lane) * 8);
        k.ua = u1_0 + (uint32_t)(((rr + 1) * BX + 2 + 2 * lane) * 8);
        k.full0 = full0; k.empty0 = empty0; k.u1full0 = u1full0;
        k.c0 = p.c.c0; k.c1 = p.c.c1; k.c2 = p.c.c2; k.c3 = p.c.c3;
        k.pitch = p.g.pitch; k.plane_sz = p.g.plane;
        // offset of the lane's first cell inside a grown xy plane (scratch / peer ghost planes)
        k.xy_off = p.g.xoff + (ia + p.g.ng) + k.pitch * (j + p.g.ng);
        k.lane = lane;
        // ghost rules of the six faces (a face with mode 0 is a slab neighbour)
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fy0 = make_face1(p.faces.f[2], true); k.fy1 = make_face1(p.faces.f[3], true);
        k.fz0 = make_face1(p.faces.f[4], true); k.fz1 = make_face1(p.faces.f[5], true);
        k.do_xlo = do_l2 && k.fx0.on && k.va0 && ia == 0;
        const long long last = n0 - 1;
        k.xhi_sel = !(do_l2 && k.fx1.on) ? -1 : (k.va0 && ia == last) ? 0 : (k.va1 && ia + 1 == last) ? 1 : (k.vb0 && ib == last) ? 2 : (k.vb1 && ib + 1 == last) ? 3 : -1;
        k.xhi_off = (uint32_t)((k.xhi_sel >= 2 ? 512 : 0) + ((k.xhi_sel & 1) + 1) * 8);
        k.do_ylo = do_l2 && k.fy0.on && j == 0; k.do_yhi = do_l2 && k.fy1.on && j == n1 - 1;
        k.any_x = k.do_xlo || k.xhi_sel >= 0;
        k.any_y = k.do_ylo || k.do_yhi;
        k.it_zlo = k0 == 0 ? 1 : -1;
        k.it_u1lo = (a_sync_lo && do_l2) ? 0 : -1;
        k.it_u1hi = (a_sync_hi && do_l2) ? nL1 - 1 : -1;

        RowSt st;
        const double2 zero2 = make_double2(0.0, 0.0);
        mbar_wait(full0, 0);
        st.below_a = lds2(k.pa); st.below_b = lds2(k.pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        st.cur_a = lds2(k.pa + Cfg::U0_STAGE); st.cur_b = lds2(k.pa + Cfg::U0_STAGE + 512);
        st.u1lo_a = zero2; st.u1lo_b = zero2; st.u1c_a = zero2; st.u1c_b = zero2; st.rp_a = zero2; st.rp_b = zero2;
        st.oa = p.out + p.g.at(ia, j, k0, 0);
        st.sc = 1; st.sn = 2; st.phn = 0;

        // level 1 only: the planes before the chunk's first output plane; every plane for the halo rows
        const int n_l1_only = do_l2 ? it_l2_first : nL1;
        int it = 0;
        // warp-uniform: the whole 128-cell row is valid and away from the x / y faces, and the chunk has no slab role
        const bool interior = SPEC && row_ok && i0 + TX <= n0 && !(k.fx0.on && i0 == 0) && !(k.fx1.on && i0 + TX == n0) &&
                              !k.any_y && !a_sync_lo && !a_sync_hi && !is_b;
        // the same tile shape with an x face on one side (n0 a multiple of the tile width): only the x-ghost code is kept
        const bool xface_only = SPEC && !interior && row_ok && i0 + TX <= n0 && !k.any_y && !a_sync_lo && !a_sync_hi && !is_b;
        if (SPEC && interior) {
            for (; it < n_l1_only; ++it) row_iter<TY, NS, true, false, 0>(p, k, st, it);
            if (do_l2) {
                // the first two-level iteration may meet the z-lo face: general code, once
                if (it < nL1) { row_iter<TY, NS, true, true, 2>(p, k, st, it); ++it; }
                for (; it < nL1; it += 3) {
                    row_iter<TY, NS, true, true, 0>(p, k, st, it);
                    if (it + 1 >= nL1) { ++it; break; }
                    row_iter<TY, NS, true, true, 0>(p, k, st, it + 1);
                    if (it + 2 >= nL1) { it += 2; break; }
                    row_iter<TY, NS, true, true, 0>(p, k, st, it + 2);
                }
                if (top) row_iter<TY, NS, false, true>(p, k, st, nL1);
            }
        } else if (SPEC && xface_only) {
            for (; it < n_l1_only; ++it) row_iter<TY, NS, true, false, 1>(p, k, st, it);
            if (do_l2) {
                if (it < nL1) { row_iter<TY, NS, true, true, 2>(p, k, st, it); ++it; }
                for (; it < nL1; it += 3) {
                    row_iter<TY, NS, true, true, 1>(p, k, st, it);
                    if (it + 1 >= nL1) { ++it; break; }
                    row_iter<TY, NS, true, true, 1>(p, k, st, it + 1);
                    if (it + 2 >= nL1) { it += 2; break; }
                    row_iter<TY, NS, true, true, 1>(p, k, st, it + 2);
                }
                if (top) row_iter<TY, NS, false, true>(p, k, st, nL1);
            }
        } else {
        for (; it < n_l1_only; ++it) row_iter<TY, NS, true, false>(p, k, st, it);
        if (do_l2) {
            // both levels; three plane iterations per trip so that the three-plane register windows rotate by renaming
            for (; it < nL1; it += 3) {
                row_iter<TY, NS, true, true>(p, k, st, it);
                if (it + 1 >= nL1) { ++it; break; }
                row_iter<TY, NS, true, true>(p, k, st, it + 1);
                if (it + 2 >= nL1) { it += 2; break; }
                row_iter<TY, NS, true, true>(p, k, st, it + 2);
            }
            if (top) row_iter<TY, NS, false, true>(p, k, st, nL1);
        }
        }
    }
    if (a_sync_lo || a_sync_hi || is_b) {
        if (p.ps.enabled) {
            __syncthreads();                                   // every peer store of this CTA has been issued
            if (threadIdx.x == 0) {
                if (a_sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.half_sig_lo, p.ps.epoch + 1);
                if (a_sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.half_sig_hi, p.ps.epoch + 1);
                if (b_is_lo) peer_publish(p.done_b + 0, p.target_b_lo, p.ps.sig_lo, p.ps.epoch + 2);
                if (b_is_hi) peer_publish(p.done_b + 1, p.target_b_hi, p.ps.sig_hi, p.ps.epoch + 2);
            }
        }
    }
}

// ===============================================================================================
// Two rows per row warp (variants 24 / 25; not the default).  The shared-memory pipe bounds the kernel above
// (88 wavefronts per row warp and plane: DESIGN.md section 9); when one warp owns two ADJACENT rows the
// N/S neighbour of one row is the other row's register, so N/S costs 8 instead of 16 wavefronts per row and
// level, and half as many warps wait on the barriers.  Same ring protocol, same arithmetic (bit-identical).
//   MERGE = true : TY/2 two-row warps + ONE helper warp (both halo rows, then the two halo columns) + producer
//                  (128x12 tile: 8 warps, up to 255 registers per thread)
//   MERGE = false: TY/2 two-row warps + halo-row warp + halo-column warp + producer (128x16 tile: 11 warps)
// ===============================================================================================
struct C4 { double2 a, b; };                  // the four cells of a lane in one row
struct EW4 { double w_a0, e_a1, w_b0, e_b1; };

__device__ __forceinline__ C4 ld4(uint32_t a) { C4 v; v.a = lds2(a); v.b = lds2(a + 512); return v; }
__device__ __forceinline__ EW4 ldew(uint32_t pc)
{
    EW4 e;
    e.w_a0 = lds1(pc - 8); e.e_a1 = lds1(pc + 16);
    e.w_b0 = lds1(pc + 512 - 8); e.e_b1 = lds1(pc + 512 + 16);
    return e;
}
__device__ __forceinline__ C4 upd3_4(const SweepCoef& c, const C4& r, const C4& cur, const EW4& e, const C4& n, const C4& s, const C4& u, const C4& d)
{
    C4 o;
    o.a.x = upd3(c.c0, c.c1, c.c2, c.c3, r.a.x, cur.a.y, e.w_a0, n.a.x, s.a.x, u.a.x, d.a.x);
    o.a.y = upd3(c.c0, c.c1, c.c2, c.c3, r.a.y, e.e_a1, cur.a.x, n.a.y, s.a.y, u.a.y, d.a.y);
    o.b.x = upd3(c.c0, c.c1, c.c2, c.c3, r.b.x, cur.b.y, e.w_b0, n.b.x, s.b.x, u.b.x, d.b.x);
    o.b.y = upd3(c.c0, c.c1, c.c2, c.c3, r.b.y, e.e_b1, cur.b.x, n.b.y, s.b.y, u.b.y, d.b.y);
    return o;
}
__device__ __forceinline__ C4 part2_4(const SweepCoef& c, const C4& r, const C4& cur, const EW4& e, const C4& n, const C4& s)
{
    C4 o;
    o.a.x = part2(c.c0, c.c1, c.c2, r.a.x, cur.a.y, e.w_a0, n.a.x, s.a.x);
    o.a.y = part2(c.c0, c.c1, c.c2, r.a.y, e.e_a1, cur.a.x, n.a.y, s.a.y);
    o.b.x = part2(c.c0, c.c1, c.c2, r.b.x, cur.b.y, e.w_b0, n.b.x, s.b.x);
    o.b.y = part2(c.c0, c.c1, c.c2, r.b.y, e.e_b1, cur.b.x, n.b.y, s.b.y);
    return o;
}
__device__ __forceinline__ C4 fin2_4(const C4& part, double c3, const C4& u, const C4& d)
{
    C4 o;
    o.a.x = fin2(part.a.x, c3, u.a.x, d.a.x); o.a.y = fin2(part.a.y, c3, u.a.y, d.a.y);
    o.b.x = fin2(part.b.x, c3, u.b.x, d.b.x); o.b.y = fin2(part.b.y, c3, u.b.y, d.b.y);
    return o;
}
__device__ __forceinline__ C4 ghost4(const Face1& f, const C4& m) { C4 o; o.a = ghost1(f, m.a); o.b = ghost1(f, m.b); return o; }

// ---- rows of the u1 ring.  DI (variant 28): a row of BX = 132 positions is stored de-interleaved, the 66 even positions first
// and the 66 odd ones behind (U1_ODD bytes on): a pair's left neighbour is an odd position and its right neighbour an even one,
// so the E/W loads become dense 8-byte accesses (2 shared-memory wavefronts per warp instead of 4); pair loads / stores become two
// dense 8-byte accesses each (the same 4 wavefronts as one 16-byte access).  `a` = where the lane's first cell sits in the row.
constexpr uint32_t U1_ODD = (128 + 4) / 2 * 8;
template <bool DI>
__device__ __forceinline__ uint32_t u1_pos(int q) { return DI ? (uint32_t)((q & 1) * U1_ODD + (q >> 1) * 8) : (uint32_t)(q * 8); }
template <bool DI>
__device__ __forceinline__ C4 ldu4(uint32_t a)
{
    if (!DI) return ld4(a);
    C4 v;
    v.a.x = lds1(a); v.a.y = lds1(a + U1_ODD); v.b.x = lds1(a + 256); v.b.y = lds1(a + U1_ODD + 256);
    return v;
}
template <bool DI>
__device__ __forceinline__ EW4 lduew(uint32_t a)
{
    if (!DI) return ldew(a);
    EW4 e;
    e.w_a0 = lds1(a + U1_ODD - 8); e.e_a1 = lds1(a + 8);
    e.w_b0 = lds1(a + U1_ODD + 256 - 8); e.e_b1 = lds1(a + 256 + 8);
    return e;
}
template <bool DI>
__device__ __forceinline__ void stu4(uint32_t a, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
    if (!DI) { sstore4(a, va, vb, full, a0, a1, b0, b1); return; }
    if (full) { sts1(a, va.x); sts1(a + U1_ODD, va.y); sts1(a + 256, vb.x); sts1(a + U1_ODD + 256, vb.y); }
    else {
        if (a0) sts1(a, va.x);
        if (a1) sts1(a + U1_ODD, va.y);
        if (b0) sts1(a + 256, vb.x);
        if (b1) sts1(a + U1_ODD + 256, vb.y);
    }
}

template <int TY, int NS, bool MERGE>
struct Tb2xCfg : Tb2Cfg<TY, NS> {
    static_assert(TY % 2 == 0, "two rows per warp");
    static constexpr int NFULL = TY / 2;
    static constexpr int NCONS = NFULL + (MERGE ? 1 : 2);
    static constexpr int THREADS = 32 * (NCONS + 1);
};

struct RowK2 {
    uint32_t pa, ra, ua;                 // the lane's first cell of the warp's FIRST row (second row: + BX*8)
    uint32_t full0, empty0, u1full0;
    SweepCoef c;
    bool xv0, xv1, xv2, xv3;             // x validity of the lane's four cells
    bool ok[2];                          // the warp's rows are rows of the frame
    bool do_xlo, any_x, do_ylo;
    int xhi_sel;
    uint32_t xhi_off, xlo_off;           // u1-ring offsets of the x-hi / x-lo ghost from the lane's first cell
    int yhi_row;                         // which of the two rows is the frame's last row (-1: none)
    Face1 fx0, fx1, fy0, fy1, fz0, fz1;
    long long pitch, plane_sz, xy_off;
    int lane;
    int it_zlo, it_u1lo, it_u1hi;
};
struct RowSt2 {
    C4 below[2], cur[2];                 // u0 planes a+it-1, a+it
    C4 u1lo[2], u1c[2];                  // u1 planes q-1, q
    C4 rp[2];                            // rhs plane q
    double* oa;
    uint32_t sc, sn, phn;
};

template <bool MASKED = true>
__device__ __forceinline__ void st4g(double* a, const C4& v, const RowK2& k, int r)
{
    store4(a, v.a, v.b, !MASKED || (k.ok[r] && k.xv3), k.ok[r] && k.xv0, k.ok[r] && k.xv1, k.ok[r] && k.xv2, k.ok[r] && k.xv3);
}
template <bool MASKED = true, bool DI = false>
__device__ __forceinline__ void st4s(uint32_t a, const C4& v, const RowK2& k, int r)
{
    stu4<DI>(a, v.a, v.b, !MASKED || (k.ok[r] && k.xv3), k.ok[r] && k.xv0, k.ok[r] && k.xv1, k.ok[r] && k.xv2, k.ok[r] && k.xv3);
}
__device__ __forceinline__ double xhi_pick(const C4& v, int sel) { return sel == 0 ? v.a.x : sel == 1 ? v.a.y : sel == 2 ? v.b.x : v.b.y; }
__device__ __forceinline__ void ld4g(const double* q, C4& v, const RowK2& k, int r)
{
    if (k.ok[r] && k.xv3) { v.a = *reinterpret_cast<const double2*>(q); v.b = *reinterpret_cast<const double2*>(q + 64); }
    else if (k.ok[r]) {
        if (k.xv0) v.a.x = q[0];
        if (k.xv1) v.a.y = q[1];
        if (k.xv2) v.b.x = q[64];
        if (k.xv3) v.b.y = q[65];
    }
}

// One plane iteration of a two-row warp (rows r0 = 2w and r0+1 of the tile): see row_iter for the protocol.
// EDGE: 2 = general, 0 = interior-specialised copy, 1 = x-ghost code only (see row_iter)
template <int TY, int NS, bool DO_L1, bool DO_L2, int EDGE = 2, bool DI = false>
__device__ __forceinline__ void row2_iter(const Tb2Params& p, const RowK2& k, RowSt2& st, const int it)
{
    using Cfg = Tb2Cfg<TY, NS>;
    constexpr uint32_t RB = Cfg::BX * 8;                                  // bytes per staged row
    const double2 zero2 = make_double2(0.0, 0.0);
    const C4 zero4 = { zero2, zero2 };
    C4 above[2] = { zero4, zero4 }, u1n[2] = { zero4, zero4 }, rc[2] = { zero4, zero4 }, P[2] = { zero4, zero4 };

    if (DO_L1) mbar_wait(k.full0 + 8 * st.sn, st.phn);
    if (DO_L2 || it >= 2) mbar_wait(k.u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);

    if (DO_L2) {
        const uint32_t up = k.ua + (uint32_t)((it - 1) & 1) * Cfg::U1_SLOT;
        // row 0: S from the ring (row below), N = the warp's second row; row 1: N from the ring, S = the first row
        const C4 s0 = ldu4<DI>(up - RB), n1 = ldu4<DI>(up + 2 * RB);
        const EW4 e0 = lduew<DI>(up), e1 = lduew<DI>(up + RB);
        P[0] = part2_4(k.c, st.rp[0], st.u1c[0], e0, st.u1c[1], s0);
        P[1] = part2_4(k.c, st.rp[1], st.u1c[1], e1, n1, st.u1c[0]);
    }
    if (DO_L1) {
        const uint32_t pn = k.pa + st.sn * Cfg::U0_STAGE, pc = k.pa + st.sc * Cfg::U0_STAGE, pr = k.ra + st.sc * Cfg::RHS_STAGE;
        above[0] = ld4(pn); above[1] = ld4(pn + RB);
        rc[0] = ld4(pr); rc[1] = ld4(pr + RB);
        const C4 s0 = ld4(pc - RB), n1 = ld4(pc + 2 * RB);
        const EW4 e0 = ldew(pc), e1 = ldew(pc + RB);
        // ---------------- level 1: u1 of plane a+it, both rows ----------------
        u1n[0] = upd3_4(k.c, rc[0], st.cur[0], e0, st.cur[1], s0, above[0], st.below[0]);
        u1n[1] = upd3_4(k.c, rc[1], st.cur[1], e1, n1, st.cur[0], above[1], st.below[1]);
        // ghosts of u1 that level 2 takes from REGISTERS: odd n0 (E neighbour of the last valid cell is the pair's
        // second register) and the y-hi ghost row when the frame's last row is the warp's first row
        if (EDGE == 2) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (k.xhi_sel == 0) u1n[r].a.y = ghost1(k.fx1, u1n[r].a.x);
                if (k.xhi_sel == 2) u1n[r].b.y = ghost1(k.fx1, u1n[r].b.x);
            }
            if (k.yhi_row == 0) u1n[1] = ghost4(k.fy1, u1n[0]);
        }

        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.empty0 + 8 * st.sc);
        const uint32_t us = k.ua + (uint32_t)(it & 1) * Cfg::U1_SLOT;
        st4s<EDGE == 2, DI>(us, u1n[0], k, 0);
        st4s<EDGE == 2, DI>(us + RB, u1n[1], k, 1);
        if (EDGE != 0 && k.any_x) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (k.ok[r]) {
                    if (k.do_xlo) sts1(us + r * RB + k.xlo_off, ghost1(k.fx0, u1n[r].a.x));
                    if (k.xhi_sel >= 0) sts1(us + r * RB + k.xhi_off, ghost1(k.fx1, xhi_pick(u1n[r], k.xhi_sel)));
                }
        }
        if (EDGE == 2 && k.do_ylo) st4s<true, DI>(us - RB, ghost4(k.fy0, u1n[0]), k, 0);
        if (EDGE == 2 && k.yhi_row == 1) st4s<true, DI>(us + 2 * RB, ghost4(k.fy1, u1n[1]), k, 1);
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.u1full0 + 8 * (it & 1));
        if (EDGE == 2 && it == k.it_u1lo) { st4g(p.peer_u1_lo + k.xy_off, u1n[0], k, 0); st4g(p.peer_u1_lo + k.xy_off + k.pitch, u1n[1], k, 1); }
        if (EDGE == 2 && it == k.it_u1hi) { st4g(p.peer_u1_hi + k.xy_off, u1n[0], k, 0); st4g(p.peer_u1_hi + k.xy_off + k.pitch, u1n[1], k, 1); }
    }
    if (DO_L2) {
        // ---------------- level 2: u2 of plane q = a+it-1, both rows ----------------
        C4 dn[2] = { st.u1lo[0], st.u1lo[1] }, upv[2] = { u1n[0], u1n[1] };
        const bool zlo = EDGE == 2 && it == k.it_zlo;
        if (zlo) {
            if (k.fz0.on) { dn[0] = ghost4(k.fz0, st.u1c[0]); dn[1] = ghost4(k.fz0, st.u1c[1]); }
            else if (p.u1_in_lo) { ld4g(p.u1_in_lo + k.xy_off, dn[0], k, 0); ld4g(p.u1_in_lo + k.xy_off + k.pitch, dn[1], k, 1); }
        }
        if (!DO_L1) {
            if (k.fz1.on) { upv[0] = ghost4(k.fz1, st.u1c[0]); upv[1] = ghost4(k.fz1, st.u1c[1]); }
            else if (p.u1_in_hi) { ld4g(p.u1_in_hi + k.xy_off, upv[0], k, 0); ld4g(p.u1_in_hi + k.xy_off + k.pitch, upv[1], k, 1); }
        }
        double* oa = st.oa;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const C4 u2 = fin2_4(P[r], k.c.c3, upv[r], dn[r]);
            double* o = oa + r * k.pitch;
            st4g<EDGE == 2>(o, u2, k, r);
            if (EDGE != 0 && k.any_x && k.ok[r]) {
                if (k.do_xlo) o[-1] = ghost1(k.fx0, u2.a.x);
                if (k.xhi_sel >= 0) o[(k.xhi_sel >= 2 ? 64 : 0) + (k.xhi_sel & 1) + 1] = ghost1(k.fx1, xhi_pick(u2, k.xhi_sel));
            }
            if (EDGE == 2 && r == 0 && k.do_ylo) st4g(o - k.pitch, ghost4(k.fy0, u2), k, 0);
            if (EDGE == 2 && k.yhi_row == r) st4g(o + k.pitch, ghost4(k.fy1, u2), k, r);
            if (zlo) {
                if (k.fz0.on) st4g(o - k.plane_sz, ghost4(k.fz0, u2), k, r);
                if (p.peer_lo_plane) st4g(p.peer_lo_plane + k.xy_off + r * k.pitch, u2, k, r);
            }
            if (!DO_L1) {
                if (k.fz1.on) st4g(o + k.plane_sz, ghost4(k.fz1, u2), k, r);
                if (p.peer_hi_plane) st4g(p.peer_hi_plane + k.xy_off + r * k.pitch, u2, k, r);
            }
        }
        st.oa = oa + k.plane_sz;
    }
    if (DO_L1) {
        st.below[0] = st.cur[0]; st.below[1] = st.cur[1];
        st.cur[0] = above[0]; st.cur[1] = above[1];
        st.sc = st.sn;
        if (++st.sn == NS) { st.sn = 0; st.phn ^= 1u; }
    }
    st.u1lo[0] = st.u1c[0]; st.u1lo[1] = st.u1c[1];
    st.u1c[0] = u1n[0]; st.u1c[1] = u1n[1];
    st.rp[0] = rc[0]; st.rp[1] = rc[1];
}

template <int TY, int NS, bool MERGE, bool SPEC = false, bool DI = false>
__global__ void __launch_bounds__(32 * (TY / 2 + (MERGE ? 2 : 3)), 1) jacobi3d_tb2x_kernel(const __grid_constant__ CUtensorMap tm_u0,
                                                                                           const __grid_constant__ CUtensorMap tm_rhs,
                                                                                           const __grid_constant__ Tb2Params p)
{
    using Cfg = Tb2xCfg<TY, NS, MERGE>;
    constexpr int TX = Cfg::TX, BX = Cfg::BX, NFULL = Cfg::NFULL;
    constexpr uint32_t RB = BX * 8;
    static_assert(TY <= 16, "the halo-column lanes map 16 rows per side");
    RNMF_DYNAMIC_SMEM(smem_raw);
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t u0_0 = sb, rhs_0 = sb + Cfg::RHS_OFF, u1_0 = sb + Cfg::U1_OFF;
    const uint32_t full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8, u1full0 = empty0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = p.g.n[0], n1 = p.g.n[1], n2 = p.g.n[2];
    const long long i0 = (long long)blockIdx.x * TX;
    const long long j0 = (long long)blockIdx.y * TY;

    // chunk roles exactly as in jacobi3d_tb2_kernel
    const int bz = (int)blockIdx.z;
    const bool is_b = bz >= p.n_a;
    long long k0, k1;
    if (!is_b) {
        const unsigned zc = p.ps.enabled ? peer_chunk_order((unsigned)bz, (unsigned)p.n_a) : (unsigned)bz;
        k0 = p.m_begin + (long long)zc * p.chunk;
        k1 = k0 + p.chunk < p.m_end ? k0 + p.chunk : p.m_end;
    } else {
        k0 = (p.b_lo && bz == p.n_a) ? 0 : n2 - 1;
        k1 = k0 + 1;
    }
    const long long a = k0 > 0 ? k0 - 1 : 0;
    const long long b = k1 < n2 ? k1 : n2 - 1;
    const int nL1 = (int)(b - a) + 1;
    const int n_planes = nL1 + 2;
    const bool top = b < k1;
    const int it_l2_first = (int)(k0 - a) + 1;
    const bool a_sync_lo = !is_b && p.peer_u1_lo && k0 == p.m_begin;
    const bool a_sync_hi = !is_b && p.peer_u1_hi && k1 == p.m_end;
    const bool b_is_lo = is_b && k0 == 0 && p.b_lo;
    const bool b_is_hi = is_b && !b_is_lo;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, Cfg::NCONS);
        }
        mbar_init(u1full0, Cfg::NCONS);
        mbar_init(u1full0 + 8, Cfg::NCONS);
        mbar_fence_init();
        if (a_sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (a_sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
        if (b_is_lo) peer_spin_wait(p.half_wait_lo, p.ps.epoch + 1, p.ps.status);
        if (b_is_hi) peer_spin_wait(p.half_wait_hi, p.ps.epoch + 1, p.ps.status);
    }
    __syncthreads();

    const int w_rows = NFULL, w_cols = MERGE ? NFULL : NFULL + 1, w_prod = Cfg::NCONS;
    const bool do_rows = warp == w_rows, do_cols = warp == w_cols;

    if (warp == w_prod) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx = (int)(p.g.xoff + p.g.ng + i0 - 2);
            const int cy_u0 = (int)(p.g.ng + j0 - 2);
            const int cy_rhs = cy_u0 + 1;
            int cz = (int)(p.g.kg + a - 1);
            uint32_t s = 0, wrap = 0;
            for (int idx = 0; idx < n_planes; ++idx) {
                if (wrap > 0) mbar_wait(empty0 + 8 * s, (wrap - 1) & 1u);
                const bool with_rhs = idx >= 1 && idx <= nL1;
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(Cfg::U0_BYTES + (with_rhs ? Cfg::RHS_BYTES : 0)));
                tma_load_3d(u0_0 + s * Cfg::U0_STAGE, &tm_u0, cx, cy_u0, cz, full0 + 8 * s);
                if (with_rhs) tma_load_3d(rhs_0 + s * Cfg::RHS_STAGE, &tm_rhs, cx, cy_rhs, cz, full0 + 8 * s);
                ++cz;
                if (++s == NS) { s = 0; ++wrap; }
            }
        }
    } else if (do_rows || do_cols) {
        // ------------------------------ helper warp(s): halo rows j0-1 and j0+TY, halo columns i0-1 and i0+128 (level 1 only) ---
        const SweepCoef c = p.c;
        // rows: h = 0 -> tile row -1, h = 1 -> tile row TY; same lane <-> cell mapping as the row warps
        const long long ia = i0 + 2 * lane, ib = ia + 64;
        const bool xv0 = ia < n0, xv1 = ia + 1 < n0, xv2 = ib < n0, xv3 = ib + 1 < n0;
        const bool rok[2] = { do_rows && j0 - 1 >= 0, do_rows && j0 + TY < n1 };
        const int rrow[2] = { -1, TY };
        uint32_t hpa[2], hra[2], hua[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            hpa[h] = u0_0 + (uint32_t)(((rrow[h] + 2) * BX + 2 + 2 * lane) * 8);
            hra[h] = rhs_0 + (uint32_t)(((rrow[h] + 1) * BX + 2 + 2 * lane) * 8);
            hua[h] = u1_0 + (uint32_t)((rrow[h] + 1) * BX * 8) + u1_pos<DI>(2 + 2 * lane);
        }
        // columns: lanes 0..15: column i0-1, rows j0+lane; lanes 16..31: column i0+128, rows j0+lane-16
        const int side = lane >> 4, rcl = (lane & 15) < TY ? (lane & 15) : 0;
        const int col = side ? TX : -1;
        const bool cok = do_cols && (lane & 15) < TY && i0 + col >= 0 && i0 + col < n0 && j0 + rcl < n1;
        const uint32_t cp = u0_0 + (uint32_t)(((rcl + 2) * BX + col + 2) * 8);
        const uint32_t cr = rhs_0 + (uint32_t)(((rcl + 1) * BX + col + 2) * 8);
        const uint32_t cu = u1_0 + (uint32_t)((rcl + 1) * BX * 8) + u1_pos<DI>(col + 2);

        mbar_wait(full0, 0);
        C4 below[2] = { ld4(hpa[0]), ld4(hpa[1]) };
        double cbelow = lds1(cp);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        C4 cur[2] = { ld4(hpa[0] + Cfg::U0_STAGE), ld4(hpa[1] + Cfg::U0_STAGE) };
        double ccur = lds1(cp + Cfg::U0_STAGE);
        uint32_t sc = 1, sn = 2, phn = 0;
        for (int it = 0; it < nL1; ++it) {
            mbar_wait(full0 + 8 * sn, phn);
            if (it >= 2) mbar_wait(u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
            C4 above[2], v[2];
            double cabove = 0.0, cv = 0.0;
            if (do_rows) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t pc = hpa[h] + sc * Cfg::U0_STAGE;
                    above[h] = ld4(hpa[h] + sn * Cfg::U0_STAGE);
                    const C4 nn = ld4(pc + RB), ss = ld4(pc - RB), rr = ld4(hra[h] + sc * Cfg::RHS_STAGE);
                    const EW4 e = ldew(pc);
                    v[h] = upd3_4(c, rr, cur[h], e, nn, ss, above[h], below[h]);
                }
            }
            if (do_cols) {
                const uint32_t pc = cp + sc * Cfg::U0_STAGE;
                cabove = lds1(cp + sn * Cfg::U0_STAGE);
                const double nn = lds1(pc + RB), ss = lds1(pc - RB), ee = lds1(pc + 8), ww = lds1(pc - 8);
                const double rr = lds1(cr + sc * Cfg::RHS_STAGE);
                cv = upd3(c.c0, c.c1, c.c2, c.c3, rr, ee, ww, nn, ss, cabove, cbelow);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * sc);
            const uint32_t slot = (uint32_t)(it & 1) * Cfg::U1_SLOT;
            if (do_rows) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    stu4<DI>(hua[h] + slot, v[h].a, v[h].b, rok[h] && xv3, rok[h] && xv0, rok[h] && xv1, rok[h] && xv2, rok[h] && xv3);
            }
            if (cok) sts1(cu + slot, cv);
            __syncwarp();
            if (lane == 0) mbar_arrive(u1full0 + 8 * (it & 1));
            if (do_rows) { below[0] = cur[0]; below[1] = cur[1]; cur[0] = above[0]; cur[1] = above[1]; }
            cbelow = ccur; ccur = cabove;
            sc = sn;
            if (++sn == NS) { sn = 0; phn ^= 1u; }
        }
    } else {
        // ------------------------------ two-row warps ------------------------------
        const int r0 = 2 * warp;
        const long long jr = j0 + r0;
        const long long ia = i0 + 2 * lane, ib = ia + 64;
        RowK2 k;
        k.ok[0] = jr < n1; k.ok[1] = jr + 1 < n1;
        k.xv0 = ia < n0; k.xv1 = ia + 1 < n0; k.xv2 = ib < n0; k.xv3 = ib + 1 < n0;
        k.pa = u0_0 + (uint32_t)(((r0 + 2) * BX + 2 + 2 * lane) * 8);
        k.ra = rhs_0 + (uint32_t)(((r0 + 1) * BX + 2 + 2 * lane) * 8);
        k.ua = u1_0 + (uint32_t)((r0 + 1) * BX * 8) + u1_pos<DI>(2 + 2 * lane);
        k.full0 = full0; k.empty0 = empty0; k.u1full0 = u1full0;
        k.c = p.c;
        k.pitch = p.g.pitch; k.plane_sz = p.g.plane;
        k.xy_off = p.g.xoff + (ia + p.g.ng) + k.pitch * (jr + p.g.ng);
        k.lane = lane;
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fy0 = make_face1(p.faces.f[2], true); k.fy1 = make_face1(p.faces.f[3], true);
        k.fz0 = make_face1(p.faces.f[4], true); k.fz1 = make_face1(p.faces.f[5], true);
        k.do_xlo = k.fx0.on && k.xv0 && ia == 0;
        const long long last = n0 - 1;
        k.xhi_sel = !k.fx1.on ? -1 : (k.xv0 && ia == last) ? 0 : (k.xv1 && ia + 1 == last) ? 1 : (k.xv2 && ib == last) ? 2 : (k.xv3 && ib + 1 == last) ? 3 : -1;
        k.xhi_off = DI ? (uint32_t)((k.xhi_sel >= 2 ? 256 : 0) + ((k.xhi_sel & 1) ? 8 : U1_ODD))
                       : (uint32_t)((k.xhi_sel >= 2 ? 512 : 0) + ((k.xhi_sel & 1) + 1) * 8);
        k.xlo_off = DI ? U1_ODD - 8 : (uint32_t)-8;
        k.any_x = k.do_xlo || k.xhi_sel >= 0;
        k.do_ylo = k.fy0.on && jr == 0;
        k.yhi_row = !k.fy1.on ? -1 : (jr == n1 - 1 ? 0 : (jr + 1 == n1 - 1 ? 1 : -1));
        k.it_zlo = k0 == 0 ? 1 : -1;
        k.it_u1lo = a_sync_lo ? 0 : -1;
        k.it_u1hi = a_sync_hi ? nL1 - 1 : -1;

        RowSt2 st;
        const double2 zero2 = make_double2(0.0, 0.0);
        const C4 zero4 = { zero2, zero2 };
        mbar_wait(full0, 0);
        st.below[0] = ld4(k.pa); st.below[1] = ld4(k.pa + RB);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        st.cur[0] = ld4(k.pa + Cfg::U0_STAGE); st.cur[1] = ld4(k.pa + Cfg::U0_STAGE + RB);
        st.u1lo[0] = zero4; st.u1lo[1] = zero4; st.u1c[0] = zero4; st.u1c[1] = zero4; st.rp[0] = zero4; st.rp[1] = zero4;
        st.oa = p.out + p.g.at(ia, jr, k0, 0);
        st.sc = 1; st.sn = 2; st.phn = 0;

        int it = 0;
        // warp-uniform: both rows are whole rows of the frame away from the x / y faces, and the chunk has no slab role
        const bool interior = SPEC && k.ok[1] && i0 + TX <= n0 && !(k.fx0.on && i0 == 0) && !(k.fx1.on && i0 + TX == n0) &&
                              !k.do_ylo && k.yhi_row < 0 && !a_sync_lo && !a_sync_hi && !is_b;
        const bool xface_only = SPEC && !interior && k.ok[1] && i0 + TX <= n0 && !k.do_ylo && k.yhi_row < 0 && !a_sync_lo && !a_sync_hi && !is_b;
        if (SPEC && interior) {
            for (; it < it_l2_first; ++it) row2_iter<TY, NS, true, false, 0, DI>(p, k, st, it);
            // the first two-level iteration may meet the z-lo face: general code, once
            if (it < nL1) { row2_iter<TY, NS, true, true, 2, DI>(p, k, st, it); ++it; }
            for (; it < nL1; it += 3) {
                row2_iter<TY, NS, true, true, 0, DI>(p, k, st, it);
                if (it + 1 >= nL1) break;
                row2_iter<TY, NS, true, true, 0, DI>(p, k, st, it + 1);
                if (it + 2 >= nL1) break;
                row2_iter<TY, NS, true, true, 0, DI>(p, k, st, it + 2);
            }
        } else if (SPEC && xface_only) {
            for (; it < it_l2_first; ++it) row2_iter<TY, NS, true, false, 1, DI>(p, k, st, it);
            if (it < nL1) { row2_iter<TY, NS, true, true, 2, DI>(p, k, st, it); ++it; }
            for (; it < nL1; it += 3) {
                row2_iter<TY, NS, true, true, 1, DI>(p, k, st, it);
                if (it + 1 >= nL1) break;
                row2_iter<TY, NS, true, true, 1, DI>(p, k, st, it + 1);
                if (it + 2 >= nL1) break;
                row2_iter<TY, NS, true, true, 1, DI>(p, k, st, it + 2);
            }
        } else {
        for (; it < it_l2_first; ++it) row2_iter<TY, NS, true, false, 2, DI>(p, k, st, it);
        for (; it < nL1; it += 3) {
            row2_iter<TY, NS, true, true, 2, DI>(p, k, st, it);
            if (it + 1 >= nL1) break;
            row2_iter<TY, NS, true, true, 2, DI>(p, k, st, it + 1);
            if (it + 2 >= nL1) break;
            row2_iter<TY, NS, true, true, 2, DI>(p, k, st, it + 2);
        }
        }
        if (top) row2_iter<TY, NS, false, true, 2, DI>(p, k, st, nL1);
    }
    if (a_sync_lo || a_sync_hi || is_b) {
        if (p.ps.enabled) {
            __syncthreads();
            if (threadIdx.x == 0) {
                if (a_sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.half_sig_lo, p.ps.epoch + 1);
                if (a_sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.half_sig_hi, p.ps.epoch + 1);
                if (b_is_lo) peer_publish(p.done_b + 0, p.target_b_lo, p.ps.sig_lo, p.ps.epoch + 2);
                if (b_is_hi) peer_publish(p.done_b + 1, p.target_b_hi, p.ps.sig_hi, p.ps.epoch + 2);
            }
        }
    }
}

template <int TY, int NS, bool SPEC = false, bool QUAD = false>
struct Tb2Launcher {
    using Cfg = Tb2Cfg<TY, NS>;
    // chunk length: whole waves of one CTA per SM, each CTA paying chunk+3 plane iterations
    static void plan(const Tb2Args& t, dim3& grid, int& chunk, int& n_a)
    {
        const FrameGeom& g = t.a.g;
        const long long span = t.a.k_end - t.a.k_begin;
        const long long tiles = ((g.n[0] + 127) / 128) * ((g.n[1] + TY - 1) / TY);
        long long best_c = 1, best_cost = -1;
        static const int cand[] = { 2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256 };
        for (int c : cand) {
            const long long cc = c > span ? span : c;
            if (cc < 1) break;
            const long long chunks = (span + cc - 1) / cc;
            if (chunks <= 60000) {
                const long long waves = (tiles * chunks + 147) / 148;
                const long long cost = waves * (cc + 3);
                if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_c = cc; }
            }
            if (c >= span) break;
        }
        if (t.a.chunk_override > 0) best_c = t.a.chunk_override > span ? span : t.a.chunk_override;
        if (best_c < 1) best_c = 1;
        chunk = (int)best_c;
        n_a = (int)((span + best_c - 1) / best_c);
        grid = dim3((unsigned)((g.n[0] + 127) / 128), (unsigned)((g.n[1] + TY - 1) / TY), (unsigned)(n_a + (t.b_lo ? 1 : 0) + (t.b_hi ? 1 : 0)));
    }
    static int launch(const Tb2Args& t, cudaStream_t s)
    {
        const SweepArgs& a = t.a;
        dim3 grid;
        int chunk, n_a;
        plan(t, grid, chunk, n_a);
        CUtensorMap tm_u0, tm_rhs;
        if (tma3d_encode_map(&tm_u0, a.g, a.in, Cfg::BX, Cfg::U0_ROWS)) return -1;
        if (tma3d_encode_map(&tm_rhs, a.g, a.rhs, Cfg::BX, Cfg::R_ROWS)) return -1;
        Tb2Params p;
        p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces;
        p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk; p.n_a = n_a;
        p.b_lo = t.b_lo; p.b_hi = t.b_hi;
        p.peer_u1_lo = t.peer_u1_lo; p.peer_u1_hi = t.peer_u1_hi;
        p.u1_in_lo = t.u1_in_lo; p.u1_in_hi = t.u1_in_hi;
        p.peer_lo_plane = a.peer_lo_plane; p.peer_hi_plane = a.peer_hi_plane;
        p.ps = a.ps;
        p.half_wait_lo = t.half_wait_lo; p.half_wait_hi = t.half_wait_hi;
        p.half_sig_lo = t.half_sig_lo; p.half_sig_hi = t.half_sig_hi;
        p.done_b = t.done_b; p.target_b_lo = t.target_b_lo; p.target_b_hi = t.target_b_hi;
#ifndef RNMF_EMULATE
        static PerDeviceOnce attr_set;
        if (!attr_set.done()) {
            if (cudaFuncSetAttribute(jacobi3d_tb2_kernel<TY, NS, SPEC, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
            attr_set.mark();
        }
#endif
        RNMF_KERNEL_LAUNCH((jacobi3d_tb2_kernel<TY, NS, SPEC, QUAD>), grid, Cfg::THREADS, Cfg::SMEM, s, tm_u0, tm_rhs, p);
        return 1;
    }
};

// two rows per row warp: same tiling / planner as Tb2Launcher<TY, NS>, its own kernel and thread count
template <int TY, int NS, bool MERGE, bool SPEC = false, bool DI = false>
struct Tb2xLauncher {
    using Cfg = Tb2xCfg<TY, NS, MERGE>;
    static int launch(const Tb2Args& t, cudaStream_t s)
    {
        const SweepArgs& a = t.a;
        dim3 grid;
        int chunk, n_a;
        Tb2Launcher<TY, NS>::plan(t, grid, chunk, n_a);
        CUtensorMap tm_u0, tm_rhs;
        if (tma3d_encode_map(&tm_u0, a.g, a.in, Cfg::BX, Cfg::U0_ROWS)) return -1;
        if (tma3d_encode_map(&tm_rhs, a.g, a.rhs, Cfg::BX, Cfg::R_ROWS)) return -1;
        Tb2Params p;
        p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces;
        p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk; p.n_a = n_a;
        p.b_lo = t.b_lo; p.b_hi = t.b_hi;
        p.peer_u1_lo = t.peer_u1_lo; p.peer_u1_hi = t.peer_u1_hi;
        p.u1_in_lo = t.u1_in_lo; p.u1_in_hi = t.u1_in_hi;
        p.peer_lo_plane = a.peer_lo_plane; p.peer_hi_plane = a.peer_hi_plane;
        p.ps = a.ps;
        p.half_wait_lo = t.half_wait_lo; p.half_wait_hi = t.half_wait_hi;
        p.half_sig_lo = t.half_sig_lo; p.half_sig_hi = t.half_sig_hi;
        p.done_b = t.done_b; p.target_b_lo = t.target_b_lo; p.target_b_hi = t.target_b_hi;
#ifndef RNMF_EMULATE
        static PerDeviceOnce attr_set;
        if (!attr_set.done()) {
            if (cudaFuncSetAttribute(jacobi3d_tb2x_kernel<TY, NS, MERGE, SPEC, DI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
            attr_set.mark();
        }
#endif
        RNMF_KERNEL_LAUNCH((jacobi3d_tb2x_kernel<TY, NS, MERGE, SPEC, DI>), grid, Cfg::THREADS, Cfg::SMEM, s, tm_u0, tm_rhs, p);
        return 1;
    }
};

}  // namespace

// variants: 20 (and auto) = 128x12 tile, one row per row warp; 21 = 128x8  (16 / 12 warps: registers are allocated per
// 4 warps, so 128x14 or 128x16 tiles would leave only 96 registers per thread and spill);
// 24 = 128x12 tile, two rows per row warp, one helper warp (8 warps); 25 = 128x16 tile, two rows per row warp (11 warps)
int launch_jacobi_double_sweep(const Tb2Args& t, cudaStream_t s)
{
    if (t.a.k_end <= t.a.k_begin && !t.b_lo && !t.b_hi) return 0;
    switch (t.a.variant) {
    case 21: return Tb2Launcher<8, 6>::launch(t, s);
    case 22: return Tb2Launcher<12, 6, true, true>::launch(t, s);    // quad mapping (2 rows x 64 columns per warp) + interior-specialised loop
    case 23: return Tb2Launcher<12, 6, false, true>::launch(t, s);   // quad mapping, general loop only
    case 24: return Tb2xLauncher<12, 6, true>::launch(t, s);
    case 25: return Tb2xLauncher<16, 4, false>::launch(t, s);
    case 26: return Tb2Launcher<12, 6, true>::launch(t, s);          // interior-specialised plane loop (opt-in)
    case 27: return Tb2xLauncher<12, 6, true, true>::launch(t, s);   // two rows per warp + interior-specialised loop (opt-in)
    case 28: return Tb2xLauncher<12, 6, true, true, true>::launch(t, s);   // 27 + de-interleaved u1 ring rows (opt-in)
    default: return Tb2Launcher<12, 6>::launch(t, s);
    }
}

int double_sweep_boundary_ctas(const Tb2Args& t)
{
    dim3 grid;
    int chunk, n_a;
    switch (t.a.variant) {
    case 21: Tb2Launcher<8, 6>::plan(t, grid, chunk, n_a); break;
    case 25: Tb2Launcher<16, 4>::plan(t, grid, chunk, n_a); break;
    default: Tb2Launcher<12, 6>::plan(t, grid, chunk, n_a); break;
    }
    return (int)(grid.x * grid.y);
}
