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
//   * TY "quad" warps own TWO adjacent tile rows x 64 columns each: lane l holds the 2 x 2 patch
//     {rows jA, jA+1} x {columns ia, ia+1}, ia = i0 + 64*xh + 2l.  The N neighbour of row jA is the lane's own
//     row jA+1 and vice versa, at both levels, so per warp and plane the N/S loads cost 8 + 8 instead of
//     16 + 16 shared-memory wavefronts (72 instead of 88 in total with one 128-cell row per warp: the
//     shared-memory pipe bounded that mapping; measured 380 -> 444 GLUP/s at 768^3 together with the
//     interior-specialised loop below).  Two more warps own the halo rows j0-1 and j0+TY (level 1 only,
//     one 128-cell row each), one warp owns the two halo columns i0-1 and i0+128 (level 1 only).
//   * Level 1 is register z-marched over u0 exactly like the single-sweep kernel; its result goes to a
//     2-slot shared-memory ring of u1 planes, together with the ghost values fill_bc would give u1 on
//     the physical x/y faces.  Level 2 is register z-marched over the lane's own u1 values (the z
//     ghosts of u1 are formed in registers) and reads the remaining neighbours of u1 from the ring.
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
//     levels rotate by renaming.  Quad warps whose 2 x 64 patch is whole and away from the y faces run an
//     interior-specialised copy of the plane loop without ghost code / store masks (EDGE 0; EDGE 1 keeps
//     the x-ghost code for the patches next to an x face); the few plane iterations that meet a z face or
//     a slab neighbour run the general code (EDGE 2).
//   * u2 is stored with the fused ghost fill of the output frame, as in the single-sweep kernel.
//
// Slabs (one GPU per z-slab), DEEP HALO: every eligible frame buffer carries one extra plane in front of
// its z-lo ghost plane and one behind its z-hi ghost plane.  With them a slab face towards a neighbour is
// nothing but a chunk boundary: level 1 also forms u1 on the ghost-plane position (from u0 of the deep
// plane, the ghost plane and the first valid plane), so u2 of the slab's first plane needs no data that
// arrives DURING the launch.  What the neighbour needs for its next launch -- u2 of my two planes next to
// the face, whole grown rows including the x/y face ghosts -- is stored straight into its ghost + deep
// planes over NVLink by the boundary CTAs, under the SAME epoch handshake as the single-sweep kernel
// (PeerSync, common.cuh: wait flag >= epoch, publish epoch+1; boundary chunks are scheduled last).  The
// redundant work is one plane of level 1 per neighbour and launch; there are no extra chunks, no
// half-step flags and no scratch planes (round 1's A/B-chunk scheme cost three pipeline fills per
// boundary plane, which dominated on thin slabs).
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
    long long m_begin, m_end;    // output planes
    long long hole_begin, hole_end;   // planes [hole_begin, hole_end) are left out (hole_end == hole_begin: none); no chunk straddles it
    int chunk;
    int n_chunks;
    int nb_lo, nb_hi;            // a slab neighbour behind the z-lo / z-hi face (deep halo planes hold its data)
    int zshift;                  // the tensor maps start this many planes before the z-lo ghost plane (the buffers' deep planes)
    // neighbours' planes of THEIR `out` buffer that receive my u2 planes 0, 1 (lo) and n2-1, n2-2 (hi)
    double* peer_lo_ghost;
    double* peer_lo_deep;
    double* peer_hi_ghost;
    double* peer_hi_deep;
    PeerSync ps;
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
    static constexpr int NCONS = TY + 3;                  // TY quad warps + 2 halo-row warps + 1 halo-column warp
    static constexpr int THREADS = 32 * (TY + 4);         // + producer warp
    static_assert(NS >= 3 && SMEM <= 227 * 1024, "shared-memory budget");
    static_assert(TY % 2 == 0 && TY <= 16, "row pairs; the halo-column warp maps 16 rows per side");
};

// first two terms of poisson.rs:83-88 (the z term is added last: same association as upd3)
__device__ __forceinline__ double part2(double c0, double c1, double c2, double r, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, __dadd_rn(e, w))), __dmul_rn(c2, __dadd_rn(n, s)));
}
__device__ __forceinline__ double fin2(double part, double c3, double u, double d) { return __dadd_rn(part, __dmul_rn(c3, __dadd_rn(u, d))); }

// ring position shared by every consumer: stage of plane a+it, of plane a+it+1, parity its full barrier completes with
struct Ring {
    uint32_t sc, sn, phn;
    __device__ __forceinline__ void advance(int ns) { sc = sn; if (++sn == (uint32_t)ns) { sn = 0; phn ^= 1u; } }
};

// ---------------------------------------------------------------------------------------------------
// quad warps
// ---------------------------------------------------------------------------------------------------
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
    int it_zlo;                          // iteration whose level 2 produces plane 0 next to a PHYSICAL z-lo face (-1: none)
    int it_send[4];                      // iterations whose u2 plane also goes to a neighbour: lo ghost, lo deep, hi deep, hi ghost
};

// rotating per-thread state of a quad warp
struct QuadSt {
    double2 below_a, below_b, cur_a, cur_b;      // u0 of planes a+it-1, a+it        (own cells)
    double2 u1lo_a, u1lo_b, u1c_a, u1c_b;        // u1 of planes q-1, q, q = a+it-1  (own cells)
    double2 rp_a, rp_b;                          // rhs of plane q
    double* oa;                                  // where u2 of plane q goes
    Ring r;
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

// u2 of the lane's patch into one grown plane: the valid cells and the x / y face ghosts fill_bc gives them
// (EDGE as in quad_iter).  Used for the frame's own plane and for the copies that go to a slab neighbour.
template <int EDGE>
__device__ __forceinline__ void plane_store(const QuadK& k, double* oa, const double2& na, const double2& nb)
{
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
}

// One plane iteration of a quad warp: level 1 of plane a+it (DO_L1) and level 2 of plane q = a+it-1 (DO_L2).
// All shared-memory loads of both levels are issued behind ONE wait phase; level 2's x/y part is formed
// before level 1's result exists (part2), its z term is added once u1 of plane a+it is known (fin2).
// EDGE: 2 = general; 0 = the warp's patch is whole, touches no x / y face, and the iteration meets neither a z face nor a
// plane that goes to a slab neighbour -- the ghost code and the store masks are compiled out; 1 = as 0, but the patch sits
// next to an x face: only the x-ghost code stays.
// R: it % 6 when it is known at compile time (the steady loop is unrolled by 6 = lcm of the NS = 6 stages, the two u1 slots and
// the three-plane register windows), -1 otherwise.  With R >= 0 every ring index is a constant, so all shared-memory and barrier
// addresses are immediate offsets and the stage arithmetic (about a sixth of the loop's instructions) disappears; the barrier
// parities are st.r.phn (= (it / 6) & 1 throughout a trip, toggled by the caller after it) xor a constant.
template <int TY, int NS, bool DO_L1, bool DO_L2, int EDGE = 2, int R = -1>
__device__ __forceinline__ void quad_iter(const Tb2Params& p, const QuadK& k, QuadSt& st, const int it)
{
    using Cfg = Tb2Cfg<TY, NS>;
    static_assert(R < 0 || NS == 6, "static ring indices assume six stages");
    constexpr uint32_t RB = Cfg::BX * 8;                                  // one tile row in shared memory
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 above_a = zero2, above_b = zero2, u1n_a = zero2, u1n_b = zero2, rc_a = zero2, rc_b = zero2, P_a = zero2, P_b = zero2;
    // ring position: stage of plane a+it (sc), of plane a+it+1 (sn), parity the wait for sn completes with; u1 slots / parities
    const uint32_t sc = R >= 0 ? (uint32_t)((R + 1) % 6) : st.r.sc;
    const uint32_t sn = R >= 0 ? (uint32_t)((R + 2) % 6) : st.r.sn;
    const uint32_t phn = R >= 0 ? st.r.phn ^ (R >= 4 ? 1u : 0u) : st.r.phn;
    const uint32_t slot = R >= 0 ? (uint32_t)(R & 1) : (uint32_t)(it & 1);                      // u1 slot of plane a+it
    const uint32_t pslot = slot ^ 1u;                                                           // ... of plane a+it-1
    // ((it-1) >> 1) & 1 for it = 6t + R:  R = 0: t+1, 1: t, 2: t, 3: t+1, 4: t+1, 5: t   (mod 2)
    const uint32_t pu = R >= 0 ? st.r.phn ^ ((R == 0 || R == 3 || R == 4) ? 1u : 0u) : (uint32_t)((it - 1) >> 1) & 1u;

    // ---- wait phase --------------------------------------------------------------------------------
    if (DO_L1) mbar_wait(k.full0 + 8 * sn, phn);                          // u0 plane a+it+1 (plane a+it was waited for one iteration ago)
    // u1 plane a+it-1 complete: level 2 reads it; and every warp has finished reading plane a+it-2,
    // whose slot level 1 overwrites below (reads precede the arrival in program order)
    if (DO_L2 || it >= 2) mbar_wait(k.u1full0 + 8 * pslot, pu);

    // ---- loads ---------------------------------------------------------------------------------------
    if (DO_L2) {
        const uint32_t up = k.ua + pslot * Cfg::U1_SLOT;
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
        const uint32_t pn = k.pa + sn * Cfg::U0_STAGE, pc = k.pa + sc * Cfg::U0_STAGE, pr = k.ra + sc * Cfg::RHS_STAGE;
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
        if (k.lane == 0) mbar_arrive(k.empty0 + 8 * sc);                 // this warp is done with the stage of plane a+it
        const uint32_t us = k.ua + slot * Cfg::U1_SLOT;
        qsstore(us, RB, u1n_a, u1n_b, EDGE != 2 || k.full, k.va0, k.va1, k.vb0, k.vb1);
        // what fill_bc gives u1 on the physical x / y faces (read by level 2 of the edge cells)
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
        if (k.lane == 0) mbar_arrive(k.u1full0 + 8 * slot);
    }
    if (DO_L2) {
        // ---------------- level 2: u2 of plane q = a+it-1 ----------------
        // z neighbours of u1: registers (a slab neighbour's side included: level 1 ran on the ghost-plane position), or
        // the ghost rule of a physical z face
        double2 dn_a = st.u1lo_a, dn_b = st.u1lo_b, up_a = u1n_a, up_b = u1n_b;
        const bool zlo = EDGE == 2 && it == k.it_zlo;
        if (zlo) { dn_a = ghost1(k.fz0, st.u1c_a); dn_b = ghost1(k.fz0, st.u1c_b); }
        if (!DO_L1) { up_a = ghost1(k.fz1, st.u1c_a); up_b = ghost1(k.fz1, st.u1c_b); }      // the frame's last plane next to a physical z-hi face
        double2 na, nb;
        na.x = fin2(P_a.x, k.c3, up_a.x, dn_a.x);
        na.y = fin2(P_a.y, k.c3, up_a.y, dn_a.y);
        nb.x = fin2(P_b.x, k.c3, up_b.x, dn_b.x);
        nb.y = fin2(P_b.y, k.c3, up_b.y, dn_b.y);

        double* oa = st.oa;
        plane_store<EDGE>(k, oa, na, nb);
        if (EDGE == 2) {
            const bool full = k.full;
            if (zlo) qstore(oa - k.plane_sz, k.pitch, ghost1(k.fz0, na), ghost1(k.fz0, nb), full, k.va0, k.va1, k.vb0, k.vb1);
            if (!DO_L1) qstore(oa + k.plane_sz, k.pitch, ghost1(k.fz1, na), ghost1(k.fz1, nb), full, k.va0, k.va1, k.vb0, k.vb1);
            // planes next to a slab neighbour: the same grown rows into its ghost / deep plane (NVLink stores)
            double* po = it == k.it_send[0] ? p.peer_lo_ghost : it == k.it_send[1] ? p.peer_lo_deep
                       : it == k.it_send[2] ? p.peer_hi_deep : it == k.it_send[3] ? p.peer_hi_ghost : nullptr;
            if (po) plane_store<2>(k, po + k.xy_off, na, nb);
        }
        st.oa = oa + k.plane_sz;
    }
    // ---- rotate --------------------------------------------------------------------------------------
    if (DO_L1) {
        st.below_a = st.cur_a; st.below_b = st.cur_b;
        st.cur_a = above_a; st.cur_b = above_b;
        if (R < 0) st.r.advance(NS);          // static trips leave the ring state alone: after six iterations it is the same, parity aside
    }
    st.u1lo_a = st.u1c_a; st.u1lo_b = st.u1c_b;
    st.u1c_a = u1n_a; st.u1c_b = u1n_b;
    st.rp_a = rc_a; st.rp_b = rc_b;
}

// plane iterations [it, end) with both levels.  The steady part runs in trips of six iterations that start at it % 6 == 0, where
// the ring is at (sc, sn) = (1, 2) and st.r.phn == (it / 6) & 1: there every ring index is a compile-time constant (quad_iter's
// R) and the register windows rotate by renaming.  EDGE == 2 (general code, rare) keeps the run-time ring.
template <int TY, int NS, int EDGE>
__device__ __forceinline__ void quad_run(const Tb2Params& p, const QuadK& k, QuadSt& st, int& it, const int end)
{
#ifdef RNMF_DYNAMIC_RING          // A/B builds only (scripts/build_alt.sh): the run-time ring everywhere
    constexpr bool kStatic = false;
#else
    constexpr bool kStatic = true;
#endif
    if (kStatic && EDGE != 2 && NS == 6) {
        for (; it < end && it % 6 != 0; ++it) quad_iter<TY, NS, true, true, EDGE>(p, k, st, it);
        for (; it + 6 <= end; it += 6) {
            quad_iter<TY, NS, true, true, EDGE, 0>(p, k, st, it);
            quad_iter<TY, NS, true, true, EDGE, 1>(p, k, st, it + 1);
            quad_iter<TY, NS, true, true, EDGE, 2>(p, k, st, it + 2);
            quad_iter<TY, NS, true, true, EDGE, 3>(p, k, st, it + 3);
            quad_iter<TY, NS, true, true, EDGE, 4>(p, k, st, it + 4);
            quad_iter<TY, NS, true, true, EDGE, 5>(p, k, st, it + 5);
            st.r.phn ^= 1u;
        }
    } else {
        for (; it + 3 <= end; it += 3) {
            quad_iter<TY, NS, true, true, EDGE>(p, k, st, it);
            quad_iter<TY, NS, true, true, EDGE>(p, k, st, it + 1);
            quad_iter<TY, NS, true, true, EDGE>(p, k, st, it + 2);
        }
    }
    for (; it < end; ++it) quad_iter<TY, NS, true, true, EDGE>(p, k, st, it);
}

// ---------------------------------------------------------------------------------------------------
// halo-row warps: one 128-cell row (lane l: cell pairs 2l,2l+1 and 64+2l,65+2l), level 1 only
// ---------------------------------------------------------------------------------------------------
struct HaloRowK {
    uint32_t pa, ra, ua;
    uint32_t full0, empty0, u1full0;
    double c0, c1, c2, c3;
    bool full, va0, va1, vb0, vb1;
    int lane;
};

template <int TY, int NS>
__device__ __forceinline__ void halo_row_iter(const HaloRowK& k, double2& below_a, double2& below_b, double2& cur_a, double2& cur_b, Ring& r, const int it)
{
    using Cfg = Tb2Cfg<TY, NS>;
    constexpr int BX = Cfg::BX;
    mbar_wait(k.full0 + 8 * r.sn, r.phn);
    if (it >= 2) mbar_wait(k.u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
    const uint32_t pn = k.pa + r.sn * Cfg::U0_STAGE, pc = k.pa + r.sc * Cfg::U0_STAGE, pr = k.ra + r.sc * Cfg::RHS_STAGE;
    const double2 above_a = lds2(pn), above_b = lds2(pn + 512);
    const double2 n_a = lds2(pc + BX * 8), n_b = lds2(pc + BX * 8 + 512);
    const double2 s_a = lds2(pc - BX * 8), s_b = lds2(pc - BX * 8 + 512);
    const double2 rc_a = lds2(pr), rc_b = lds2(pr + 512);
    const double w_a0 = lds1(pc - 8), e_a1 = lds1(pc + 16);
    const double w_b0 = lds1(pc + 512 - 8), e_b1 = lds1(pc + 512 + 16);
    double2 u1n_a, u1n_b;
    u1n_a.x = upd3(k.c0, k.c1, k.c2, k.c3, rc_a.x, cur_a.y, w_a0, n_a.x, s_a.x, above_a.x, below_a.x);
    u1n_a.y = upd3(k.c0, k.c1, k.c2, k.c3, rc_a.y, e_a1, cur_a.x, n_a.y, s_a.y, above_a.y, below_a.y);
    u1n_b.x = upd3(k.c0, k.c1, k.c2, k.c3, rc_b.x, cur_b.y, w_b0, n_b.x, s_b.x, above_b.x, below_b.x);
    u1n_b.y = upd3(k.c0, k.c1, k.c2, k.c3, rc_b.y, e_b1, cur_b.x, n_b.y, s_b.y, above_b.y, below_b.y);
    __syncwarp();
    if (k.lane == 0) mbar_arrive(k.empty0 + 8 * r.sc);
    sstore4(k.ua + (uint32_t)(it & 1) * Cfg::U1_SLOT, u1n_a, u1n_b, k.full, k.va0, k.va1, k.vb0, k.vb1);
    __syncwarp();
    if (k.lane == 0) mbar_arrive(k.u1full0 + 8 * (it & 1));
    below_a = cur_a; below_b = cur_b;
    cur_a = above_a; cur_b = above_b;
    r.advance(NS);
}

template <int TY, int NS>
__global__ void __launch_bounds__(32 * (TY + 4), 1) jacobi3d_tb2_kernel(const __grid_constant__ CUtensorMap tm_u0,
                                                                         const __grid_constant__ CUtensorMap tm_rhs,
                                                                         const __grid_constant__ Tb2Params p)
{
    using Cfg = Tb2Cfg<TY, NS>;
    constexpr int TX = Cfg::TX, BX = Cfg::BX;
    RNMF_DYNAMIC_SMEM(smem_raw);
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t u0_0 = sb, rhs_0 = sb + Cfg::RHS_OFF, u1_0 = sb + Cfg::U1_OFF;
    const uint32_t full0 = sb + Cfg::BAR_OFF, empty0 = full0 + NS * 8, u1full0 = empty0 + NS * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = p.g.n[0], n1 = p.g.n[1], n2 = p.g.n[2];
    const long long i0 = (long long)blockIdx.x * TX;
    const long long j0 = (long long)blockIdx.y * TY;

    // ---- which planes ----------------------------------------------------------------------------------
    const unsigned zc = p.ps.enabled ? peer_chunk_order(blockIdx.z, (unsigned)p.n_chunks) : blockIdx.z;
    long long k0 = p.m_begin + (long long)zc * p.chunk;
    if (k0 >= p.hole_begin) k0 += p.hole_end - p.hole_begin;             // chunks are laid out over the range without the hole
    const long long part_end = k0 < p.hole_begin ? p.hole_begin : p.m_end;
    const long long k1 = k0 + p.chunk < part_end ? k0 + p.chunk : part_end;
    // level-1 planes a..b: one beyond the output planes on either side, unless a PHYSICAL z face is there (its ghost rule
    // gives u1 on the ghost-plane position); towards a slab neighbour the deep halo plane makes plane -1 / n2 computable
    const long long a = (k0 > 0 || p.nb_lo) ? k0 - 1 : 0;
    const long long b = (k1 < n2 || p.nb_hi) ? k1 : n2 - 1;
    const int nL1 = (int)(b - a) + 1;
    const int n_planes = nL1 + 2;                         // u0 planes a-1..b+1  (plane a-1+idx <-> stage idx % NS)
    const bool top = b < k1;                              // the chunk ends at a physical z-hi face: one level-2-only iteration
    const int it_l2_first = (int)(k0 - a) + 1;            // iteration it: level 1 of plane a+it, level 2 of plane a+it-1

    // slab handshake roles of this CTA: its chunk reads the neighbour's planes and stores into the neighbour's buffer
    const bool sync_lo = p.ps.enabled && p.nb_lo && k0 == 0;
    const bool sync_hi = p.ps.enabled && p.nb_hi && k1 == n2;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, Cfg::NCONS);
        }
        mbar_init(u1full0, Cfg::NCONS);
        mbar_init(u1full0 + 8, Cfg::NCONS);
        mbar_fence_init();
        if (sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
    }
    __syncthreads();

    if (warp == TY + 3) {
        // ------------------------------ producer warp ------------------------------
        if (lane == 0) {
            const int cx = (int)(p.g.xoff + p.g.ng + i0 - 2);
            const int cy_u0 = (int)(p.g.ng + j0 - 2);
            const int cy_rhs = cy_u0 + 1;
            int cz = (int)(p.g.kg + a - 1) + p.zshift;
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
        Ring rg = { 1, 2, 0 };
        for (int it = 0; it < nL1; ++it) {
            mbar_wait(full0 + 8 * rg.sn, rg.phn);
            if (it >= 2) mbar_wait(u1full0 + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
            const uint32_t pc = hp + rg.sc * Cfg::U0_STAGE;
            const double above = lds1(hp + rg.sn * Cfg::U0_STAGE);
            const double nn = lds1(pc + BX * 8), ss = lds1(pc - BX * 8);
            const double ee = lds1(pc + 8), ww = lds1(pc - 8);
            const double rr = lds1(hr + rg.sc * Cfg::RHS_STAGE);
            const double v = upd3(c0, c1, c2, c3, rr, ee, ww, nn, ss, above, below);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * rg.sc);
            if (ok) sts1(hu + (uint32_t)(it & 1) * Cfg::U1_SLOT, v);
            __syncwarp();
            if (lane == 0) mbar_arrive(u1full0 + 8 * (it & 1));
            below = cur;
            cur = above;
            rg.advance(NS);
        }
    } else if (warp >= TY) {
        // ------------------------------ halo-row warps: row j0-1 (warp TY), row j0+TY (warp TY+1) ------------------------------
        const int rr = warp == TY ? -1 : TY;
        const long long j = j0 + rr;
        const bool row_ok = j >= 0 && j < n1;
        const long long ia = i0 + 2 * lane, ib = ia + 64;
        HaloRowK k;
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
        k.lane = lane;
        mbar_wait(full0, 0);
        double2 below_a = lds2(k.pa), below_b = lds2(k.pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait(full0 + 8, 0);
        double2 cur_a = lds2(k.pa + Cfg::U0_STAGE), cur_b = lds2(k.pa + Cfg::U0_STAGE + 512);
        Ring rg = { 1, 2, 0 };
        for (int it = 0; it < nL1; ++it) halo_row_iter<TY, NS>(k, below_a, below_b, cur_a, cur_b, rg, it);
    } else {
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
        // offset of the lane's first cell inside a grown xy plane (the neighbours' ghost / deep planes)
        k.xy_off = p.g.xoff + (ia + p.g.ng) + k.pitch * (jA + p.g.ng);
        k.lane = lane;
        // ghost rules of the six faces (a z face with mode 0 is a slab neighbour)
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
        k.it_zlo = (k0 == 0 && !p.nb_lo) ? 1 : -1;
        // u2 planes that also go to a neighbour: plane P is produced by iteration P - a + 1 of the chunk that holds it
        auto it_of = [&](long long P, bool on) -> int { return on && P >= k0 && P < k1 ? (int)(P - a) + 1 : -1; };
        k.it_send[0] = it_of(0, p.peer_lo_ghost != nullptr);
        k.it_send[1] = it_of(1, p.peer_lo_deep != nullptr);
        k.it_send[2] = it_of(n2 - 2, p.peer_hi_deep != nullptr);
        k.it_send[3] = it_of(n2 - 1, p.peer_hi_ghost != nullptr);

        QuadSt st;
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
        st.r = Ring{ 1, 2, 0 };

        // warp-uniform: both rows and all 64 columns are valid cells and no y face touches the patch -> the plane iterations
        // that meet neither a z face nor a neighbour's plane run without ghost code / store masks
        const long long iw = i0 + 64 * xh;
        const bool whole = okB && iw + 64 <= n0 && !k.do_ylo && k.yhi_row < 0;
        const int edge = !whole ? 2 : ((k.fx0.on && iw == 0) || (k.fx1.on && iw + 64 == n0)) ? 1 : 0;
        // general code: the first level-2 iterations of the frame's first chunk (z-lo face: plane 0; neighbour: planes 0, 1)
        // and the last two of its last chunk when a neighbour sits behind the z-hi face
        const int head = k0 == 0 ? (p.nb_lo ? (p.peer_lo_ghost ? 2 : 0) : 1) : 0;
        const int tail = (k1 == n2 && p.nb_hi && p.peer_hi_ghost) ? 2 : 0;
        int it = 0;
        if (edge == 0) for (; it < it_l2_first; ++it) quad_iter<TY, NS, true, false, 0>(p, k, st, it);
        else if (edge == 1) for (; it < it_l2_first; ++it) quad_iter<TY, NS, true, false, 1>(p, k, st, it);
        else for (; it < it_l2_first; ++it) quad_iter<TY, NS, true, false, 2>(p, k, st, it);
        const int head_end = it_l2_first + head < nL1 ? it_l2_first + head : nL1;
        for (; it < head_end; ++it) quad_iter<TY, NS, true, true, 2>(p, k, st, it);
        const int fast_end = nL1 - tail > it ? nL1 - tail : it;
        if (edge == 0) quad_run<TY, NS, 0>(p, k, st, it, fast_end);
        else if (edge == 1) quad_run<TY, NS, 1>(p, k, st, it, fast_end);
        else quad_run<TY, NS, 2>(p, k, st, it, fast_end);
        for (; it < nL1; ++it) quad_iter<TY, NS, true, true, 2>(p, k, st, it);
        if (top) quad_iter<TY, NS, false, true, 2>(p, k, st, nL1);
    }
    if (sync_lo || sync_hi) {
        __syncthreads();                                       // every peer store of this CTA has been issued
        if (threadIdx.x == 0) {
            if (sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.ps.sig_lo, p.ps.epoch + 1);
            if (sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.ps.sig_hi, p.ps.epoch + 1);
        }
    }
}

template <int TY, int NS>
struct Tb2Launcher {
    using Cfg = Tb2Cfg<TY, NS>;
    // chunk length: whole waves of one CTA per SM, each CTA paying chunk+3 plane iterations
    static void plan(const SweepArgs& a, dim3& grid, int& chunk, int& n_chunks)
    {
        const FrameGeom& g = a.g;
        const long long tiles = ((g.n[0] + 127) / 128) * ((g.n[1] + TY - 1) / TY);
        if (a.hole_end > a.hole_begin) {
            // two parts [k_begin, hole_begin) and [hole_end, k_end): one chunk length for both, no chunk across the hole --
            // the lower part is one chunk (or absent), the upper part is cut into chunks of the same length
            const long long lower = a.hole_begin - a.k_begin, upper = a.k_end - a.hole_end;
            const long long c = lower > 0 ? lower : (upper > 0 ? upper : 1);
            chunk = (int)c;
            n_chunks = (int)((lower > 0 ? 1 : 0) + (upper + c - 1) / c);
            grid = dim3((unsigned)((g.n[0] + 127) / 128), (unsigned)((g.n[1] + TY - 1) / TY), (unsigned)n_chunks);
            return;
        }
        const long long span = a.k_end - a.k_begin;
        long long best_c = 1, best_cost = -1;
        // at most 128 planes: CTAs that march longer drift apart, and the halo rows / columns a CTA shares with its neighbours
        // have left the L2 by the time the neighbour asks for them (768^3, ncu: 8.75 GB of DRAM reads per launch with
        // 256-plane chunks, 8.45 with 128, 8.2 with 64; sustained 435 / 444 / 439 GLUP/s: profiles/r2_chunk_l2.md)
        static const int cand[] = { 2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128 };
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
        if (a.chunk_override > 0) best_c = a.chunk_override > span ? span : a.chunk_override;
        // slabs: the first / last chunk holds both planes that go to the neighbour
        if ((a.nb_lo || a.nb_hi) && best_c < 2) best_c = span < 2 ? span : 2;
        if (best_c < 1) best_c = 1;
        chunk = (int)best_c;
        n_chunks = (int)((span + best_c - 1) / best_c);
        // a one-plane last chunk would split the two planes that go to the hi neighbour over two chunks: lengthen the chunks
        while (a.nb_hi && n_chunks > 1 && span - (long long)(n_chunks - 1) * chunk < 2) {
            chunk += 1;
            n_chunks = (int)((span + chunk - 1) / chunk);
        }
        grid = dim3((unsigned)((g.n[0] + 127) / 128), (unsigned)((g.n[1] + TY - 1) / TY), (unsigned)n_chunks);
    }
    static int launch(const SweepArgs& a, cudaStream_t s)
    {
        dim3 grid;
        int chunk, n_chunks;
        plan(a, grid, chunk, n_chunks);
        CUtensorMap tm_u0, tm_rhs;
        const int deep = a.deep;
        if (tma3d_encode_map_deep(&tm_u0, a.g, a.in, Cfg::BX, Cfg::U0_ROWS, deep)) return -1;
        if (tma3d_encode_map_deep(&tm_rhs, a.g, a.rhs, Cfg::BX, Cfg::R_ROWS, deep)) return -1;
        Tb2Params p;
        p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces;
        p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk; p.n_chunks = n_chunks;
        // no hole: hole_begin = hole_end = m_end, so that k0 >= hole_begin never holds
        p.hole_begin = a.hole_end > a.hole_begin ? a.hole_begin : a.k_end;
        p.hole_end = a.hole_end > a.hole_begin ? a.hole_end : a.k_end;
        p.nb_lo = a.nb_lo; p.nb_hi = a.nb_hi; p.zshift = deep;
        // the neighbours' ghost planes and the deep planes behind / in front of them
        p.peer_lo_ghost = a.peer_lo_plane; p.peer_hi_ghost = a.peer_hi_plane;
        p.peer_lo_deep = a.peer_lo_plane ? a.peer_lo_plane + a.g.plane : nullptr;
        p.peer_hi_deep = a.peer_hi_plane ? a.peer_hi_plane - a.g.plane : nullptr;
        p.ps = a.ps;
#ifndef RNMF_EMULATE
        static PerDeviceOnce attr_set;
        if (!attr_set.done()) {
            if (cudaFuncSetAttribute(jacobi3d_tb2_kernel<TY, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) return -1;
            attr_set.mark();
        }
#endif
        RNMF_KERNEL_LAUNCH((jacobi3d_tb2_kernel<TY, NS>), grid, Cfg::THREADS, Cfg::SMEM, s, tm_u0, tm_rhs, p);
        return 1;
    }
};

}  // namespace

// 128 x 12 tile, 6 stages: 16 warps, 128 registers (registers are allocated per 4 warps, so 128x14 or 128x16 tiles would
// leave only 96 registers per thread and spill)
int launch_jacobi_double_sweep(const SweepArgs& a, cudaStream_t s)
{
    if (a.k_end <= a.k_begin) return 0;
    if (a.hole_end > a.hole_begin && (a.hole_begin < a.k_begin || a.hole_end > a.k_end)) { rnmf_set_error("double sweep: hole outside the plane range"); return -1; }
    if (a.hole_end > a.hole_begin && a.hole_begin == a.k_begin && a.hole_end == a.k_end) return 0;
    if ((a.nb_lo || a.nb_hi) && a.deep < 1) { rnmf_set_error("double sweep across slabs needs deep halo planes"); return -1; }
    return Tb2Launcher<12, 6>::launch(a, s);
}

// CTAs per boundary side (tiles across a z-plane): the first / last chunk of the launch takes part in the handshake
int double_sweep_boundary_ctas(const SweepArgs& a)
{
    dim3 grid;
    int chunk, n_chunks;
    Tb2Launcher<12, 6>::plan(a, grid, chunk, n_chunks);
    return (int)(grid.x * grid.y);
}
