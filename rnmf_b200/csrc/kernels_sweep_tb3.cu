// kernels_sweep_tb3.cu -- K1 for 3D frames, temporally blocked over THREE Jacobi sweeps (each followed by the
// ng=1 ghost fill, poisson.rs:80-89 + dataframe.rs:289-355) per pass over HBM: 24 B per three updates.
//
// The two-sweep kernel (kernels_sweep_tb2.cu) is DRAM-bound in real bytes; the only way past it is fewer bytes per
// update.  Same structure, one level more: a CTA streams psi (u0) and rhs once, forms u1 = sweep(u0) and
// u2 = sweep(u1) in shared memory and writes only u3 = sweep(u2).
//
//   * CTA tile: 128 x 12 cells of u3, a chunk of z-planes.  u2 is needed on the tile grown by one cell in x / y (no
//     corners) and one plane in z, u1 on the tile grown by two (a diamond: Manhattan distance <= 2), u0 by three.
//   * per plane one TMA box of u0 (138 x 18: x from i0-4, y from j0-3) and one of rhs (134 x 16: x from
//     i0-2, y from j0-2) into a 4-stage ring (full / empty mbarriers).  The row lengths are chosen so that the column
//     warp's accesses (one row apart) fall into different banks.
//   * u1 planes live in a two-slot ring, u2 planes in a three-slot ring, both with DE-INTERLEAVED rows: even x first, odd x from ODD bytes on.  A lane
//     owns the pair (x, x+1), x even: its W neighbour is an odd position, its E neighbour an even one, so the E/W loads
//     of a warp are dense 8-byte accesses (2 shared-memory wavefronts instead of the 4 that 8-byte accesses 16 bytes
//     apart cost -- the shared-memory pipe is what bounds three levels per plane).
//   * 12 "quad" warps own two adjacent tile rows x 64 columns each (lane: a 2 x 2 patch, as in the two-sweep kernel)
//     and run all three levels on their cells, each level register z-marched over the lane's own values of the level
//     below: in iteration `it` level 1 forms plane a1+it, level 2 plane a1+it-1, level 3 plane a1+it-2; c0*rhs is rounded
//     once and queued in registers for the two later levels.
//     Two inner-row warps own ONE 128-cell row next to the tile each (j0-1, j0+12): levels 1 and 2, the same 8 cell updates
//     in front of their arrivals as a quad warp (row-PAIR warps with 12 paced every iteration).  One warp owns the columns
//     i0-2, i0-1, i0+128, i0+129 (level 1; level 2 on the inner ones) and the four diamond corners.  The 16th warp forms
//     level 1 on the outer rows j0-2, j0+13 and is also the producer: at the top of iteration `it` its lane 0 waits for the
//     stage that held plane a1+it-1 and loads plane a1+it+3 into it.
//   * levels run in the order 1, 2, 3, each loading what it needs right before it uses it (the kernel lives at the
//     128-register limit; spills would go to an L1 that 223 KB of shared memory leave no room for).  A warp arrives on
//     u1full[it & 1] after its level-1 stores and on u2full[it % 3] after its level-2 stores (dummy arrivals while level 2
//     is not active yet: the phases stay in step with `it`).  One wait at the top -- u2full of the previous iteration: u2
//     plane a1+it-2 is complete, every warp has left iteration it-2 (whose level 3 read the slot level 2 overwrites now:
//     hence THREE u2 slots), and nobody reads the u1 slot level 1 overwrites any more (level 2's loads precede that
//     arrival).  The wait for the u1 plane follows level 1.  All ring positions are functions of `it` (ring_at).
//   * quad warps whose patch is whole and touches no y face run an interior-specialised copy of the steady loop
//     (EDGE 0; EDGE 1 keeps the x-ghost code); the iterations that meet a z face or a plane that travels to a slab
//     neighbour run the general code (EDGE 2).  The steady loop is NOT unrolled: 4.8 KB with 86 register moves per
//     iteration beat every unrolled form (compile-time ring indices included) -- it has to fit the instruction caches
//     (profiles/r2_notes.md, section 9).
//
// Restriction: n0 and n1 even (every 2 x 2 patch is whole or absent; the x-hi ghost is an even position); other frames
// take the two-sweep kernel (capi.cu: plan_sweep_call).
//
// Slabs (one GPU per z-slab), DEEP HALO as in kernels_sweep_tb2.cu with one plane more: the buffers carry TWO extra
// planes in front of the z-lo ghost plane and behind the z-hi ghost plane; with the neighbour's three boundary planes
// there (rhs: two) a slab face is a chunk boundary, and u3 of my three planes next to the face goes into the
// neighbour's ghost + deep planes of its `out` buffer under the same one-epoch-per-launch handshake.
//
// Arithmetic: the same explicitly rounded sequence as every other sweep kernel: u3 is bit-identical to three single sweeps.
#include <stdlib.h>

#include "tma_util.cuh"

namespace {
using namespace tmau;

struct Tb3Params {
    FrameGeom g;
    double* out;
    SweepCoef c;
    FaceSet faces;
    long long m_begin, m_end;    // output planes
    int chunk;
    int n_chunks;
    int nb_lo, nb_hi;            // a slab neighbour behind the z-lo / z-hi face (deep halo planes hold its data)
    int zshift;                  // the tensor maps start this many planes before the z-lo ghost plane
    double* peer_lo[3];          // the lo neighbour's planes (of ITS `out`) that receive my u3 planes 0, 1, 2
    double* peer_hi[3];          // the hi neighbour's planes that receive my u3 planes n2-1, n2-2, n2-3
    PeerSync ps;
};

struct Cfg3 {
    static constexpr int TX = 128, TY = 12, NS = 4;
    static constexpr int BX0 = 138, U0_ROWS = TY + 6;      // u0 box: column c <-> x = i0-4+c, row r <-> y = j0-3+r
    static constexpr int BXR = 134, R_ROWS = TY + 4;       // rhs box: column c <-> x = i0-2+c, row r <-> y = j0-2+r
    static constexpr int HALF = 67;                        // ring row: E[m] <-> x = i0-2+2m, O[m] <-> x = i0-1+2m, m < 67
    static constexpr uint32_t RB0 = BX0 * 8, RBR = BXR * 8, ODD = HALF * 8, RBU = 2 * HALF * 8;
    static constexpr int U0_BYTES = BX0 * U0_ROWS * 8;
    static constexpr int U0_STAGE = (U0_BYTES + 127) / 128 * 128;
    static constexpr int RHS_BYTES = BXR * R_ROWS * 8;
    static constexpr int RHS_STAGE = (RHS_BYTES + 127) / 128 * 128;
    static constexpr int U1_SLOT = ((int)RBU * (TY + 4) + 127) / 128 * 128;     // rows y = j0-2 .. j0+13
    static constexpr int U2_SLOT = ((int)RBU * (TY + 2) + 127) / 128 * 128;     // rows y = j0-1 .. j0+12
    static constexpr int RHS_OFF = NS * U0_STAGE;
    static constexpr int U1_OFF = RHS_OFF + NS * RHS_STAGE;
    static constexpr int U2_OFF = U1_OFF + 2 * U1_SLOT;
    static constexpr int NU2 = 3;                          // u2 slots (three: level 3 reads a plane AFTER this warp's arrival for the next one)
    static constexpr int BAR_OFF = U2_OFF + NU2 * U2_SLOT;
    static constexpr int SMEM = BAR_OFF + (2 * NS + 2 + NU2) * 8 + 128;
    static constexpr int NCONS = TY + 4;                   // 12 quad warps + 2 inner-row warps + the column warp + the outer-rows warp (which is also the producer)
    static constexpr int THREADS = 32 * (TY + 4);
    static_assert(SMEM <= 227 * 1024, "shared-memory budget");
    static_assert(U0_BYTES % 16 == 0 && RHS_BYTES % 16 == 0 && (BX0 * 8) % 16 == 0 && (BXR * 8) % 16 == 0, "TMA box rows are multiples of 16 bytes");
};

// Barrier waits of this kernel: mbarrier.try_wait WITHOUT the suspend-time hint of tma_util.cuh's mbar_wait (the other kernels' consumers
// mostly wait for DRAM; here sixteen warps meet at two barriers per plane and a late wake-up is paid every iteration: 497 against 488
// GLUP/s with the hint, 477 with test_wait polling, 395 with polling + nanosleep; profiles/r2_notes.md, section 9).
__device__ __forceinline__ void mbar_wait3(uint32_t bar, uint32_t parity)
{
#ifdef RNMF_EMULATE
    mbar_wait(bar, parity);
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT3_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE3_%=;\n\t"
        "bra WAIT3_%=;\n\t"
        "DONE3_%=:\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
#endif
}

// first two terms of poisson.rs:83-88 on q = c0*rhs (rounded once per cell and plane and reused by all three levels -- the same
// product, so nothing changes in the result); the z term is added last (fin3): same association as upd3
__device__ __forceinline__ double part3(double q, double c1, double c2, double e, double w, double n, double s)
{
    return __dadd_rn(__dadd_rn(q, __dmul_rn(c1, __dadd_rn(e, w))), __dmul_rn(c2, __dadd_rn(n, s)));
}
__device__ __forceinline__ double2 mulc2(double c, const double2& v) { return make_double2(__dmul_rn(c, v.x), __dmul_rn(c, v.y)); }
__device__ __forceinline__ double fin3(double part, double c3, double u, double d) { return __dadd_rn(part, __dmul_rn(c3, __dadd_rn(u, d))); }
__device__ __forceinline__ double2 fin3(const double2& part, double c3, const double2& u, const double2& d)
{
    return make_double2(fin3(part.x, c3, u.x, d.x), fin3(part.y, c3, u.y, d.y));
}

// a pair (x, x+1), x even, in a de-interleaved ring row: b = byte address of its even position
__device__ __forceinline__ double2 ldp(uint32_t b) { return make_double2(lds1(b), lds1(b + Cfg3::ODD)); }
__device__ __forceinline__ void stp(uint32_t b, const double2& v) { sts1(b, v.x); sts1(b + Cfg3::ODD, v.y); }
// its W neighbour (x-1, an odd position) and its E neighbour (x+2, an even position); also where the x-lo / x-hi ghost of a
// pair next to the face lives
constexpr uint32_t kW = Cfg3::ODD - 8, kE = 8;

// Ring positions of iteration `it`, shared by every consumer warp and derived from the iteration number alone (no loop-carried ring
// state: the kernel lives at the 128-register limit, and what spills goes through an L1 that the 223 KB of shared memory leave
// almost no room for).  u0 / rhs: stage of plane a1+it (sc), of plane a1+it+1 (sn), parity its full barrier completes with (phn).
// u1: slot level 1 writes (u1w = it & 1), the slot written one iteration ago and the parity of its barrier (u1r, u1rp).
// u2: three slots, u2w = it % 3; the slot written one iteration ago and its parity (u2r, u2rp).
struct RingAt {
    uint32_t sc, sn, phn, u1w, u1r, u1rp, u2w, u2r, u2rp;
};
__device__ __forceinline__ RingAt ring_at(const int it)
{
    RingAt r;
    const uint32_t u = (uint32_t)it;
    r.sc = (u + 1u) & 3u;
    r.sn = (u + 2u) & 3u;
    r.phn = ((u + 2u) >> 2) & 1u;
    r.u1w = u & 1u;
    r.u1r = r.u1w ^ 1u;
    r.u1rp = ((u - 1u) >> 1) & 1u;                        // it >= 1 wherever it is used
    const uint32_t q = u / 3u;
    r.u2w = u - 3u * q;
    r.u2r = r.u2w == 0 ? 2u : r.u2w - 1u;
    r.u2rp = (r.u2w == 0 ? q + 1u : q) & 1u;              // (it - 1) / 3
    return r;
}
// barriers behind one base address: full[NS], empty[NS], u1full[2], u2full[3]
constexpr uint32_t kEmpty = 8 * Cfg3::NS, kU1Full = 16 * Cfg3::NS, kU2Full = 16 * Cfg3::NS + 16;

// what every consumer warp of the CTA knows about the chunk (CTA-uniform)
struct ChunkPlan {
    long long k0, k1, a1;
    int nL1;                     // level-1 planes a1 .. a1+nL1-1; iteration it: level 1 of plane a1+it, level 2 of a1+it-1, level 3 of a1+it-2
    int it2_first, it2_last;     // iterations in which level 2 / level 3 is active
    int it3_first, it3_last;
    int it_zlo2, it_zlo3;        // iteration whose level 2 / level 3 forms plane 0 next to a PHYSICAL z-lo face (-1: none)
};

// ---------------------------------------------------------------------------------------------------
// quad warps
// ---------------------------------------------------------------------------------------------------
struct QuadK3 {
    uint32_t pa, ra, u1a, u2a;           // the lane's first cell (row A): u0 stage 0, rhs stage 0, u1 slot 0, u2 slot 0 (even half)
    uint32_t bar0;                       // the CTA's barriers (full[0])
    bool ok;                             // the 2 x 2 patch lies inside the frame
    bool xlo, xhi, ylo, yhi;             // it touches that physical face (and the face has a Dirichlet / Neumann rule)
    Face1 fx0, fx1, fy0, fy1, fz0, fz1;
    long long xy_off;
    int lane;
    int it_send[6];                      // iterations whose u3 plane also goes to a neighbour: lo ghost, lo deep 1, lo deep 2, hi ghost, hi deep 1, hi deep 2
};

struct QuadSt3 {
    double2 below_a, below_b, cur_a, cur_b;      // u0 of planes a1+it-1, a1+it
    double2 u1lo_a, u1lo_b, u1c_a, u1c_b;        // u1 of planes a1+it-2, a1+it-1
    double2 u2lo_a, u2lo_b, u2c_a, u2c_b;        // u2 of planes a1+it-3, a1+it-2
    double2 q1_a, q1_b, q2_a, q2_b;              // c0*rhs of planes a1+it-1, a1+it-2
    double* oa;                                  // where u3 of plane a1+it-2 goes
};

// the lane's patch of one level into a ring slot, with the ghost values fill_bc would give that level on the physical x / y faces
template <int EDGE>
__device__ __forceinline__ void ring_store(const QuadK3& k, uint32_t b, const double2& va, const double2& vb)
{
    constexpr uint32_t RBU = Cfg3::RBU;
    if (EDGE == 2 && !k.ok) return;
    stp(b, va);
    stp(b + RBU, vb);
    if (EDGE != 0) {
        if (k.xlo) { sts1(b + kW, ghost1(k.fx0, va.x)); sts1(b + RBU + kW, ghost1(k.fx0, vb.x)); }
        if (k.xhi) { sts1(b + kE, ghost1(k.fx1, va.y)); sts1(b + RBU + kE, ghost1(k.fx1, vb.y)); }
    }
    if (EDGE == 2) {
        if (k.ylo) stp(b - RBU, ghost1(k.fy0, va));
        if (k.yhi) stp(b + 2 * RBU, ghost1(k.fy1, vb));
    }
}

// u3 of the lane's patch into one grown plane of a frame buffer: the valid cells and the x / y face ghosts
template <int EDGE>
__device__ __forceinline__ void plane_store3(const QuadK3& k, const int pitch, double* oa, const double2& na, const double2& nb)
{
    if (EDGE == 2 && !k.ok) return;
#ifdef RNMF_EMULATE
    if (reinterpret_cast<uintptr_t>(oa) % 16 != 0 || (pitch & 1)) emu::fault("misaligned 16-byte global store", reinterpret_cast<uintptr_t>(oa), 16);
#endif
    *reinterpret_cast<double2*>(oa) = na;
    *reinterpret_cast<double2*>(oa + pitch) = nb;
    if (EDGE != 0) {
        if (k.xlo) { oa[-1] = ghost1(k.fx0, na.x); oa[pitch - 1] = ghost1(k.fx0, nb.x); }
        if (k.xhi) { oa[2] = ghost1(k.fx1, na.y); oa[pitch + 2] = ghost1(k.fx1, nb.y); }
    }
    if (EDGE == 2) {
        if (k.ylo) *reinterpret_cast<double2*>(oa - pitch) = ghost1(k.fy0, na);
        if (k.yhi) *reinterpret_cast<double2*>(oa + 2 * pitch) = ghost1(k.fy1, nb);
    }
}

// One plane iteration of a quad warp.  L1 / L2 / L3: which levels are active.  EDGE: 2 = general; 0 = the warp's patches are
// whole, touch no x / y face, and the iteration meets neither a z face nor a plane that goes to a slab neighbour; 1 = as 0,
// but next to an x face.
// Order: level 1 (plane a1+it), level 2 (plane a1+it-1), level 3 (plane a1+it-2): each level loads what it needs right before it
// uses it (few values in flight: the kernel lives at the 128-register limit), a warp arrives for its u1 / u2 plane as early as
// possible, and the work behind the arrivals (level 3: loads, 44 operations, the global stores) is slack for the slower warps.
// Ring hazards: level 2 loads from the u1 slot that level 1 of the NEXT iteration overwrites -- every warp starts that iteration
// behind the wait for this iteration's u2 arrivals, which follow the loads; level 3 loads from the u2 slot written one iteration
// ago AFTER this warp's u2 arrival, so that slot must survive the next iteration's level 2: three u2 slots.
template <bool L1, bool L2, bool L3, int EDGE>
__device__ __forceinline__ void quad_iter3(const Tb3Params& p, const ChunkPlan& cp, const QuadK3& k, QuadSt3& st, const int it)
{
    using C = Cfg3;
    constexpr uint32_t RB0 = C::RB0, RBR = C::RBR, RBU = C::RBU;
    const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2, c3 = p.c.c3;
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 above_a = zero2, above_b = zero2, qc_a = zero2, qc_b = zero2;
    double2 u1n_a = zero2, u1n_b = zero2, u2n_a = zero2, u2n_b = zero2;
    const RingAt rg = ring_at(it);
    const uint32_t sc = rg.sc, sn = rg.sn;
    const uint32_t slot = rg.u1w, pslot = slot ^ 1u, pu = rg.u1rp;      // u1 ring
    const uint32_t w2 = rg.u2w, r2 = rg.u2r;                                                       // u2 ring

    // ---- wait phase ----------------------------------------------------------------------------------
    if (L1) mbar_wait3(k.bar0 + 8 * sn, rg.phn);                                              // u0 plane a1+it+1
    // every warp is through iteration it-1 up to its u2 arrival (dummy arrivals while level 2 is not active yet): u2 plane a1+it-2 is
    // complete (level 3 reads it), every warp has left iteration it-2 (whose level 3 read the u2 slot level 2 overwrites below), and
    // nobody reads the u1 slot level 1 overwrites below any more.  (The wait for the u1 plane follows level 1: the halo-row warps
    // arrive for it late in their iteration.)
    if (it >= 1) mbar_wait3(k.bar0 + kU2Full + 8 * r2, rg.u2rp);

    // ---- level 1: u1 of plane a1+it (row A, then row B) --------------------------------------------------
    if (L1) {
        const uint32_t pn = k.pa + sn * C::U0_STAGE, pc = k.pa + sc * C::U0_STAGE, pr = k.ra + sc * C::RHS_STAGE;
        {
            qc_a = mulc2(c0, lds2(pr));
            const double2 s_a = lds2(pc - RB0);
            const double w_a = lds1(pc - 8), e_a = lds1(pc + 16);
            above_a = lds2(pn);
            u1n_a.x = fin3(part3(qc_a.x, c1, c2, st.cur_a.y, w_a, st.cur_b.x, s_a.x), c3, above_a.x, st.below_a.x);
            u1n_a.y = fin3(part3(qc_a.y, c1, c2, e_a, st.cur_a.x, st.cur_b.y, s_a.y), c3, above_a.y, st.below_a.y);
        }
        {
            qc_b = mulc2(c0, lds2(pr + RBR));
            const double2 n_b = lds2(pc + 2 * RB0);
            const double w_b = lds1(pc + RB0 - 8), e_b = lds1(pc + RB0 + 16);
            above_b = lds2(pn + RB0);
            u1n_b.x = fin3(part3(qc_b.x, c1, c2, st.cur_b.y, w_b, n_b.x, st.cur_a.x), c3, above_b.x, st.below_b.x);
            u1n_b.y = fin3(part3(qc_b.y, c1, c2, e_b, st.cur_b.x, n_b.y, st.cur_a.y), c3, above_b.y, st.below_b.y);
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.bar0 + kEmpty + 8 * sc);                  // this warp is done with the stage of plane a1+it
        ring_store<EDGE>(k, k.u1a + slot * C::U1_SLOT, u1n_a, u1n_b);
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.bar0 + kU1Full + 8 * slot);
    }
    // ---- level 2: u2 of plane a1+it-1; x/y neighbours from the u1 ring, z neighbours from registers --------------
    if (L2) {
        mbar_wait3(k.bar0 + kU1Full + 8 * pslot, pu);                     // u1 plane a1+it-1 complete (it >= 1 wherever level 2 is active)
        const uint32_t b = k.u1a + pslot * C::U1_SLOT;
        double2 dn_a = st.u1lo_a, dn_b = st.u1lo_b, up_a = u1n_a, up_b = u1n_b;
        if (EDGE == 2 && it == cp.it_zlo2) { dn_a = ghost1(k.fz0, st.u1c_a); dn_b = ghost1(k.fz0, st.u1c_b); }
        if (!L1) { up_a = ghost1(k.fz1, st.u1c_a); up_b = ghost1(k.fz1, st.u1c_b); }         // the frame's last plane next to a physical z-hi face
        {
            const double2 s_a = ldp(b - RBU);
            const double w_a = lds1(b + kW), e_a = lds1(b + kE);
            u2n_a.x = fin3(part3(st.q1_a.x, c1, c2, st.u1c_a.y, w_a, st.u1c_b.x, s_a.x), c3, up_a.x, dn_a.x);
            u2n_a.y = fin3(part3(st.q1_a.y, c1, c2, e_a, st.u1c_a.x, st.u1c_b.y, s_a.y), c3, up_a.y, dn_a.y);
        }
        {
            const double2 n_b = ldp(b + 2 * RBU);
            const double w_b = lds1(b + RBU + kW), e_b = lds1(b + RBU + kE);
            u2n_b.x = fin3(part3(st.q1_b.x, c1, c2, st.u1c_b.y, w_b, n_b.x, st.u1c_a.x), c3, up_b.x, dn_b.x);
            u2n_b.y = fin3(part3(st.q1_b.y, c1, c2, e_b, st.u1c_b.x, n_b.y, st.u1c_a.y), c3, up_b.y, dn_b.y);
        }
        ring_store<EDGE>(k, k.u2a + w2 * C::U2_SLOT, u2n_a, u2n_b);
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.bar0 + kU2Full + 8 * w2);
    } else if (L1) {
        if (k.lane == 0) mbar_arrive(k.bar0 + kU2Full + 8 * w2);                 // level 2 not active yet: keep the barrier's phases in step with `it`
    }
    // ---- level 3: u3 of plane a1+it-2; x/y neighbours from the u2 ring, z neighbours from registers --------------
    if (L3) {
        const uint32_t b = k.u2a + r2 * C::U2_SLOT;
        double2 dn_a = st.u2lo_a, dn_b = st.u2lo_b, up_a = u2n_a, up_b = u2n_b;
        const bool zlo = EDGE == 2 && it == cp.it_zlo3;
        if (zlo) { dn_a = ghost1(k.fz0, st.u2c_a); dn_b = ghost1(k.fz0, st.u2c_b); }
        if (!L2) { up_a = ghost1(k.fz1, st.u2c_a); up_b = ghost1(k.fz1, st.u2c_b); }
        double2 na, nb;
        {
            const double2 s_a = ldp(b - RBU);
            const double w_a = lds1(b + kW), e_a = lds1(b + kE);
            na.x = fin3(part3(st.q2_a.x, c1, c2, st.u2c_a.y, w_a, st.u2c_b.x, s_a.x), c3, up_a.x, dn_a.x);
            na.y = fin3(part3(st.q2_a.y, c1, c2, e_a, st.u2c_a.x, st.u2c_b.y, s_a.y), c3, up_a.y, dn_a.y);
            const double2 n_b = ldp(b + 2 * RBU);
            const double w_b = lds1(b + RBU + kW), e_b = lds1(b + RBU + kE);
            nb.x = fin3(part3(st.q2_b.x, c1, c2, st.u2c_b.y, w_b, n_b.x, st.u2c_a.x), c3, up_b.x, dn_b.x);
            nb.y = fin3(part3(st.q2_b.y, c1, c2, e_b, st.u2c_b.x, n_b.y, st.u2c_a.y), c3, up_b.y, dn_b.y);
        }
        double* oa = st.oa;
        const int pitch = (int)p.g.pitch, plane_sz = (int)p.g.plane;      // kernel parameters (constant bank), not registers
        plane_store3<EDGE>(k, pitch, oa, na, nb);
        if (EDGE == 2) {
            if (k.ok) {
                if (zlo) { *reinterpret_cast<double2*>(oa - plane_sz) = ghost1(k.fz0, na); *reinterpret_cast<double2*>(oa - plane_sz + pitch) = ghost1(k.fz0, nb); }
                if (!L2) { *reinterpret_cast<double2*>(oa + plane_sz) = ghost1(k.fz1, na); *reinterpret_cast<double2*>(oa + plane_sz + pitch) = ghost1(k.fz1, nb); }
            }
            // planes next to a slab neighbour: the same grown rows into its ghost / deep planes (NVLink stores)
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                if (it == k.it_send[e]) plane_store3<2>(k, pitch, p.peer_lo[e] + k.xy_off, na, nb);
                if (it == k.it_send[3 + e]) plane_store3<2>(k, pitch, p.peer_hi[e] + k.xy_off, na, nb);
            }
        }
        st.oa = oa + plane_sz;
    }
    // ---- rotate ------------------------------------------------------------------------------------------
    if (L1) {
        st.below_a = st.cur_a; st.below_b = st.cur_b;
        st.cur_a = above_a; st.cur_b = above_b;
    }
    st.u2lo_a = st.u2c_a; st.u2lo_b = st.u2c_b;
    st.u2c_a = u2n_a; st.u2c_b = u2n_b;
    st.u1lo_a = st.u1c_a; st.u1lo_b = st.u1c_b;
    st.u1c_a = u1n_a; st.u1c_b = u1n_b;
    st.q2_a = st.q1_a; st.q2_b = st.q1_b;
    st.q1_a = qc_a; st.q1_b = qc_b;
}

// an iteration with the general code, whatever levels are active in it
__device__ __forceinline__ void quad_general3(const Tb3Params& p, const ChunkPlan& cp, const QuadK3& k, QuadSt3& st, const int it)
{
    const bool l1 = it < cp.nL1, l2 = it >= cp.it2_first && it <= cp.it2_last, l3 = it >= cp.it3_first && it <= cp.it3_last;
    if (l1 && l2 && l3) quad_iter3<true, true, true, 2>(p, cp, k, st, it);
    else if (l1 && l2) quad_iter3<true, true, false, 2>(p, cp, k, st, it);
    else if (l1) quad_iter3<true, false, false, 2>(p, cp, k, st, it);
    else if (l2 && l3) quad_iter3<false, true, true, 2>(p, cp, k, st, it);
    else if (l2) quad_iter3<false, true, false, 2>(p, cp, k, st, it);
    else if (l3) quad_iter3<false, false, true, 2>(p, cp, k, st, it);
}

// plane iterations [it, end) with all three levels.  NOT unrolled: the 4.8 KB loop with its 86 register moves per iteration beat the
// forms unrolled by 2 / 3 / 12 (the last with compile-time ring indices, 25 % fewer instructions): profiles/r2_notes.md, section 9
template <int EDGE>
__device__ __forceinline__ void quad_run3(const Tb3Params& p, const ChunkPlan& cp, const QuadK3& k, QuadSt3& st, int& it, const int end)
{
#pragma unroll 1
    for (; it < end; ++it) quad_iter3<true, true, true, EDGE>(p, cp, k, st, it);
}

// ---------------------------------------------------------------------------------------------------
// inner-row warps: ONE 128-cell row next to the tile (j0-1 or j0+12; lane l: pairs at x = i0+2l and i0+64+2l), levels 1 and 2.
// The same 8 cell updates per lane and iteration as the first two levels of a quad warp; both warps run the same code (the row
// is a run-time offset), every x/y neighbour comes from shared memory.
// ---------------------------------------------------------------------------------------------------
struct RowK3 {
    uint32_t pa, ra, u1a, u2a;           // pair a of the row: u0 stage 0, rhs stage 0, u1 slot 0, u2 slot 0 (even half)
    uint32_t bar0;
    bool ok_a, ok_b;                     // pair a / pair b lies inside the frame
    bool xlo;                            // pair a starts at x = 0 next to a physical face with a rule
    int xhi;                             // 0 / 1: pair a / b ends at x = n0-1 next to a physical face with a rule; -1: neither
    Face1 fx0, fx1, fz0, fz1;
    int lane;
};

struct RowSt3 {
    double2 below[2], cur[2];            // u0 of planes a1+it-1, a1+it (pairs a, b)
    double2 u1lo[2], u1c[2];             // u1 of planes a1+it-2, a1+it-1
    double2 q1[2];                       // c0*rhs of plane a1+it-1
};

template <bool L1, bool L2>
__device__ __forceinline__ void row_iter3(const Tb3Params& p, const ChunkPlan& cp, const RowK3& k, RowSt3& st, const int it)
{
    using C = Cfg3;
    constexpr uint32_t RB0 = C::RB0, RBU = C::RBU;
    const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2, c3 = p.c.c3;
    const double2 zero2 = make_double2(0.0, 0.0);
    double2 above[2] = { zero2, zero2 }, u1n[2] = { zero2, zero2 }, qc[2] = { zero2, zero2 };
    const RingAt rg = ring_at(it);

    if (L1) mbar_wait3(k.bar0 + 8 * rg.sn, rg.phn);
    if (it >= 1) mbar_wait3(k.bar0 + kU2Full + 8 * rg.u2r, rg.u2rp);      // as in quad_iter3
    if (L1) {
        const uint32_t pn = k.pa + rg.sn * C::U0_STAGE, pc = k.pa + rg.sc * C::U0_STAGE, pr = k.ra + rg.sc * C::RHS_STAGE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t o = 512 * h;
            qc[h] = mulc2(c0, lds2(pr + o));
            const double2 n = lds2(pc + RB0 + o), s = lds2(pc - RB0 + o);
            const double w = lds1(pc + o - 8), e = lds1(pc + o + 16);
            above[h] = lds2(pn + o);
            u1n[h].x = fin3(part3(qc[h].x, c1, c2, st.cur[h].y, w, n.x, s.x), c3, above[h].x, st.below[h].x);
            u1n[h].y = fin3(part3(qc[h].y, c1, c2, e, st.cur[h].x, n.y, s.y), c3, above[h].y, st.below[h].y);
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.bar0 + kEmpty + 8 * rg.sc);
        const uint32_t b = k.u1a + rg.u1w * C::U1_SLOT;
        if (k.ok_a) stp(b, u1n[0]);
        if (k.ok_b) stp(b + 256, u1n[1]);
        // the x-face ghosts of u1 that level 2 of this row reads
        if (k.xlo) sts1(b + kW, ghost1(k.fx0, u1n[0].x));
        if (k.xhi >= 0) sts1(b + 256 * k.xhi + kE, ghost1(k.fx1, k.xhi ? u1n[1].y : u1n[0].y));
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.bar0 + kU1Full + 8 * rg.u1w);
    }
    if (L2) {
        mbar_wait3(k.bar0 + kU1Full + 8 * rg.u1r, rg.u1rp);               // u1 plane a1+it-1 complete
        const uint32_t b = k.u1a + rg.u1r * C::U1_SLOT, b2 = k.u2a + rg.u2w * C::U2_SLOT;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t o = 256 * h;
            const double2 n = ldp(b + RBU + o), s = ldp(b - RBU + o);
            const double w = lds1(b + o + kW), e = lds1(b + o + kE);
            double2 dn = st.u1lo[h], up = u1n[h];
            if (it == cp.it_zlo2) dn = ghost1(k.fz0, st.u1c[h]);
            if (!L1) up = ghost1(k.fz1, st.u1c[h]);
            double2 v;
            v.x = fin3(part3(st.q1[h].x, c1, c2, st.u1c[h].y, w, n.x, s.x), c3, up.x, dn.x);
            v.y = fin3(part3(st.q1[h].y, c1, c2, e, st.u1c[h].x, n.y, s.y), c3, up.y, dn.y);
            if (h == 0 ? k.ok_a : k.ok_b) stp(b2 + o, v);
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(k.bar0 + kU2Full + 8 * rg.u2w);
    } else {
        if (k.lane == 0) mbar_arrive(k.bar0 + kU2Full + 8 * rg.u2w);     // keeps the barrier's phases in step with `it`
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (L1) { st.below[h] = st.cur[h]; st.cur[h] = above[h]; }
        st.u1lo[h] = st.u1c[h];
        st.u1c[h] = u1n[h];
        st.q1[h] = qc[h];
    }
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(Cfg3::THREADS, 1) jacobi3d_tb3_kernel(const __grid_constant__ CUtensorMap tm_u0,
                                                                        const __grid_constant__ CUtensorMap tm_rhs,
                                                                        const __grid_constant__ Tb3Params p)
{
    using C = Cfg3;
    constexpr int TX = C::TX, TY = C::TY, NS = C::NS;
    constexpr uint32_t RB0 = C::RB0, RBU = C::RBU;
    RNMF_DYNAMIC_SMEM(smem_raw);
    const uint32_t sb = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t u0_0 = sb, rhs_0 = sb + C::RHS_OFF, u1_0 = sb + C::U1_OFF, u2_0 = sb + C::U2_OFF;
    const uint32_t full0 = sb + C::BAR_OFF, empty0 = full0 + NS * 8, u1full0 = empty0 + NS * 8, u2full0 = u1full0 + 16;

    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps what derives from it in uniform registers
    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps what derives from it in uniform registers.
    // Roles: warps 0..11 quad warps, 12 / 13 the inner rows j0-1 / j0+12, 14 the columns, 15 the outer rows + producer (one special
    // warp per scheduler; gathering the four on ONE scheduler measured 3.5 % slower)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const long long n0 = p.g.n[0], n1 = p.g.n[1], n2 = p.g.n[2];
    const long long i0 = (long long)blockIdx.x * TX;
    const long long j0 = (long long)blockIdx.y * TY;

    // ---- which planes ----------------------------------------------------------------------------------
    ChunkPlan cp;
    {
        const unsigned zc = p.ps.enabled ? peer_chunk_order(blockIdx.z, (unsigned)p.n_chunks) : blockIdx.z;
        cp.k0 = p.m_begin + (long long)zc * p.chunk;
        cp.k1 = cp.k0 + p.chunk < p.m_end ? cp.k0 + p.chunk : p.m_end;
        // level 1 runs two planes, level 2 one plane beyond the output planes, unless a PHYSICAL z face is in the way (its ghost
        // rule gives that level on the ghost-plane position); towards a slab neighbour the deep halo planes make them computable
        const long long lo1 = cp.k0 - 2, lo2 = cp.k0 - 1, hi1 = cp.k1 + 1, hi2 = cp.k1;
        cp.a1 = p.nb_lo || lo1 > 0 ? lo1 : 0;
        const long long a2 = p.nb_lo || lo2 > 0 ? lo2 : 0;
        const long long b1 = p.nb_hi || hi1 < n2 - 1 ? hi1 : n2 - 1;
        const long long b2 = p.nb_hi || hi2 < n2 - 1 ? hi2 : n2 - 1;
        cp.nL1 = (int)(b1 - cp.a1) + 1;
        cp.it2_first = (int)(a2 - cp.a1) + 1;
        cp.it2_last = (int)(b2 - cp.a1) + 1;
        cp.it3_first = (int)(cp.k0 - cp.a1) + 2;
        cp.it3_last = (int)(cp.k1 - 1 - cp.a1) + 2;
        cp.it_zlo2 = (!p.nb_lo && a2 == 0) ? cp.it2_first : -1;
        cp.it_zlo3 = (!p.nb_lo && cp.k0 == 0) ? cp.it3_first : -1;
    }
    const long long k0 = cp.k0, k1 = cp.k1, a1 = cp.a1;
    const int nL1 = cp.nL1;
    const int n_planes = nL1 + 2;                         // u0 planes a1-1 .. a1+nL1  (plane a1-1+idx <-> stage idx % NS)

    // slab handshake roles of this CTA: its chunk reads the neighbour's planes and stores into the neighbour's buffer
    const bool sync_lo = p.ps.enabled && p.nb_lo && k0 == 0;
    const bool sync_hi = p.ps.enabled && p.nb_hi && k1 == n2;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, C::NCONS);
        }
        mbar_init(u1full0, C::NCONS);
        mbar_init(u1full0 + 8, C::NCONS);
#pragma unroll
        for (int q = 0; q < C::NU2; ++q) mbar_init(u2full0 + 8 * q, C::NCONS);
        mbar_fence_init();
        if (sync_lo) peer_spin_wait(p.ps.wait_lo, p.ps.epoch, p.ps.status);
        if (sync_hi) peer_spin_wait(p.ps.wait_hi, p.ps.epoch, p.ps.status);
    }
    __syncthreads();

    if (warp == TY + 3) {
        // ------------------------------ outer-rows warp, also the producer ------------------------------
        // Level 1 on the rows j0-2 and j0+13 (lane l: pairs at x = i0+2l and i0+64+2l of either row; every neighbour from shared
        // memory).  Lane 0 also feeds the ring: at the top of iteration `it` it waits for the stage that held plane a1+it-1 (every
        // warp released it during iteration it-1) and loads plane a1+it+3 into it -- two iterations before it is read.
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2, c3 = p.c.c3;
        const int cx_u0 = (int)(p.g.xoff + p.g.ng + i0 - 4), cx_rhs = cx_u0 + 2;
        const int cy_u0 = (int)(p.g.ng + j0 - 3), cy_rhs = cy_u0 + 1;
        const int cz0 = (int)(p.g.kg + a1 - 1) + p.zshift;                    // z coordinate of plane index 0 (plane a1-1)
        auto load_plane = [&](int idx) {                                      // lane 0: plane index idx -> stage idx % NS
            const uint32_t sg = (uint32_t)idx & (NS - 1);
            const bool with_rhs = idx >= 1 && idx <= nL1;
            mbar_arrive_expect_tx(full0 + 8 * sg, (uint32_t)(C::U0_BYTES + (with_rhs ? C::RHS_BYTES : 0)));
            tma_load_3d(u0_0 + sg * C::U0_STAGE, &tm_u0, cx_u0, cy_u0, cz0 + idx, full0 + 8 * sg);
            if (with_rhs) tma_load_3d(rhs_0 + sg * C::RHS_STAGE, &tm_rhs, cx_rhs, cy_rhs, cz0 + idx, full0 + 8 * sg);
        };
        if (lane == 0)
            for (int idx = 0; idx < NS && idx < n_planes; ++idx) load_plane(idx);
        const bool ok_bot = j0 > 0, ok_top = j0 + TY + 1 < n1;
        const long long xa = i0 + 2 * lane, xb = xa + 64;
        const bool okx[2] = { xa < n0, xb < n0 };
        // rows: q = 0: j0-2 (tile row -2), q = 1: j0+13 (tile row TY+1); pair h of row q lives in v[2q+h]
        const int trow[2] = { -2, TY + 1 };
        uint32_t pa[2], ra[2], ua[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            pa[q] = u0_0 + (uint32_t)(((trow[q] + 3) * C::BX0 + 4 + 2 * lane) * 8);
            ra[q] = rhs_0 + (uint32_t)(((trow[q] + 2) * C::BXR + 2 + 2 * lane) * 8);
            ua[q] = u1_0 + (uint32_t)(trow[q] + 2) * RBU + (uint32_t)(8 * (1 + lane));
        }
        double2 below[4], cur[4];
        mbar_wait3(full0, 0);
#pragma unroll
        for (int v = 0; v < 4; ++v) below[v] = lds2(pa[v >> 1] + 512 * (v & 1));
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait3(full0 + 8, 0);
#pragma unroll
        for (int v = 0; v < 4; ++v) cur[v] = lds2(pa[v >> 1] + C::U0_STAGE + 512 * (v & 1));
#pragma unroll 1
        for (int it = 0; it <= cp.it2_last; ++it) {
            const RingAt rg = ring_at(it);
            const bool l1 = it < nL1;
            if (lane == 0 && it + NS < n_planes) {
                mbar_wait3(empty0 + 8 * (uint32_t)(it & (NS - 1)), (uint32_t)(it >> 2) & 1u);      // plane index `it` released by every warp
                load_plane(it + NS);
            }
            if (l1) mbar_wait3(full0 + 8 * rg.sn, rg.phn);
            if (it >= 1) mbar_wait3(u2full0 + 8 * rg.u2r, rg.u2rp);            // as in quad_iter3 (the u1 slot written below is free)
            if (l1) {
                double2 above[4], u1n[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint32_t o = 512 * (v & 1);
                    const uint32_t pn = pa[v >> 1] + rg.sn * C::U0_STAGE + o, pc = pa[v >> 1] + rg.sc * C::U0_STAGE + o;
                    const double2 q = mulc2(c0, lds2(ra[v >> 1] + rg.sc * C::RHS_STAGE + o));
                    const double2 n = lds2(pc + RB0), s = lds2(pc - RB0);
                    const double w = lds1(pc - 8), e = lds1(pc + 16);
                    above[v] = lds2(pn);
                    u1n[v].x = fin3(part3(q.x, c1, c2, cur[v].y, w, n.x, s.x), c3, above[v].x, below[v].x);
                    u1n[v].y = fin3(part3(q.y, c1, c2, e, cur[v].x, n.y, s.y), c3, above[v].y, below[v].y);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * rg.sc);
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (((v >> 1) ? ok_top : ok_bot) && okx[v & 1]) stp(ua[v >> 1] + rg.u1w * C::U1_SLOT + 256 * (v & 1), u1n[v]);
                __syncwarp();
                if (lane == 0) mbar_arrive(u1full0 + 8 * rg.u1w);
#pragma unroll
                for (int v = 0; v < 4; ++v) { below[v] = cur[v]; cur[v] = above[v]; }
            }
            if (lane == 0) mbar_arrive(u2full0 + 8 * rg.u2w);                 // no level 2 here: keeps the barrier's phases in step with `it`
        }
    } else if (warp == TY + 2) {
        // ------------------------------ column warp ------------------------------
        // lanes 0..15: left of the tile, lanes 16..31: right of it.  r = lane & 15 < 12: row j0+r, an INNER cell (x = i0-1 / i0+128:
        // levels 1 and 2) and an OUTER cell (x = i0-2 / i0+129: level 1); r = 12 / 13: the diamond corner (inner x, j0-1) /
        // (inner x, j0+12), level 1 only; r = 14, 15: idle
        const double c0 = p.c.c0, c1 = p.c.c1, c2 = p.c.c2, c3 = p.c.c3;
        const int side = lane >> 4, r = lane & 15;
        const bool has_out = r < TY, corner = r == TY || r == TY + 1;
        const int row = has_out ? r : (r == TY ? -1 : TY);             // tile row of the lane's cells (idle lanes: that of lane 13)
        const int xin = side ? TX : -1, xout = side ? TX + 1 : -2;     // tile columns
        const long long y = j0 + row;
        const bool side_ok = side ? i0 + TX < n0 : i0 > 0;
        const bool ok = (has_out || corner) && side_ok && y >= 0 && y < n1;
        const bool ok2 = ok && has_out;                                // level 2 on the inner cell
        // byte addresses: u0 cell (c, row) at ((row+3)*BX0 + c+4)*8; rhs at ((row+2)*BXR + c+2)*8; ring: row (row+2) of a u1 slot,
        // row (row+1) of a u2 slot, x = i0+c at the even / odd half by parity, position (c+2)/2
        const uint32_t hpi = u0_0 + (uint32_t)(((row + 3) * C::BX0 + xin + 4) * 8), hpo = u0_0 + (uint32_t)(((row + 3) * C::BX0 + xout + 4) * 8);
        const uint32_t hri = rhs_0 + (uint32_t)(((row + 2) * C::BXR + xin + 2) * 8), hro = rhs_0 + (uint32_t)(((row + 2) * C::BXR + xout + 2) * 8);
        auto ring_pos = [](int c) -> uint32_t { return (uint32_t)(((c + 2) >> 1) * 8 + (((c + 2) & 1) ? Cfg3::ODD : 0)); };
        const uint32_t u1i = u1_0 + (uint32_t)(row + 2) * RBU + ring_pos(xin), u1o = u1_0 + (uint32_t)(row + 2) * RBU + ring_pos(xout);
        const uint32_t u1t = u1_0 + (uint32_t)(row + 2) * RBU + ring_pos(side ? TX - 1 : 0);      // the tile cell next to the inner cell
        const uint32_t u2i = u2_0 + (uint32_t)(row + 1) * RBU + ring_pos(xin);
        const Face1 fy0 = make_face1(p.faces.f[2], true), fy1 = make_face1(p.faces.f[3], true);
        const Face1 fz0 = make_face1(p.faces.f[4], true), fz1 = make_face1(p.faces.f[5], true);
        const bool ylo = ok2 && fy0.on && y == 0, yhi = ok2 && fy1.on && y == n1 - 1;

        mbar_wait3(full0, 0);
        double below_i = lds1(hpi), below_o = lds1(hpo);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait3(full0 + 8, 0);
        double cur_i = lds1(hpi + C::U0_STAGE), cur_o = lds1(hpo + C::U0_STAGE);
        double u1lo_i = 0.0, u1c_i = 0.0, q1_i = 0.0;
        for (int it = 0; it <= cp.it2_last; ++it) {
            const RingAt rg = ring_at(it);
            const bool l1 = it < nL1, l2 = it >= cp.it2_first;
            const uint32_t slot = rg.u1w, pslot = rg.u1r, w2 = rg.u2w;
            if (l1) mbar_wait3(full0 + 8 * rg.sn, rg.phn);
            if (it >= 1) { mbar_wait3(u1full0 + 8 * pslot, rg.u1rp); mbar_wait3(u2full0 + 8 * rg.u2r, rg.u2rp); }      // as in quad_iter3
            double P2 = 0.0, u1n_i = 0.0, above_i = 0.0, above_o = 0.0, qc_i = 0.0;
            if (l2) {
                const uint32_t so = pslot * C::U1_SLOT;
                const double nn = lds1(u1i + so + RBU), ss = lds1(u1i + so - RBU);
                const double tt = lds1(u1t + so), oo = lds1(u1o + so);
                // E / W: left of the tile the tile cell is the E neighbour, right of it the W neighbour
                P2 = part3(q1_i, c1, c2, side ? oo : tt, side ? tt : oo, nn, ss);
            }
            if (l1) {
                const uint32_t so = rg.sc * C::U0_STAGE, sn_ = rg.sn * C::U0_STAGE;
                above_i = lds1(hpi + sn_); above_o = lds1(hpo + sn_);
                const double n_i = lds1(hpi + so + RB0), s_i = lds1(hpi + so - RB0), e_i = lds1(hpi + so + 8), w_i = lds1(hpi + so - 8);
                const double n_o = lds1(hpo + so + RB0), s_o = lds1(hpo + so - RB0), e_o = lds1(hpo + so + 8), w_o = lds1(hpo + so - 8);
                qc_i = __dmul_rn(c0, lds1(hri + rg.sc * C::RHS_STAGE));
                const double qc_o = __dmul_rn(c0, lds1(hro + rg.sc * C::RHS_STAGE));
                u1n_i = fin3(part3(qc_i, c1, c2, e_i, w_i, n_i, s_i), c3, above_i, below_i);
                const double u1n_o = fin3(part3(qc_o, c1, c2, e_o, w_o, n_o, s_o), c3, above_o, below_o);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * rg.sc);
                const uint32_t so1 = slot * C::U1_SLOT;
                if (ok) sts1(u1i + so1, u1n_i);
                if (ok2) sts1(u1o + so1, u1n_o);
                // the y-face ghosts of u1 in the inner column (read by level 2 of this warp)
                if (ylo) sts1(u1i + so1 - RBU, ghost1(fy0, u1n_i));
                if (yhi) sts1(u1i + so1 + RBU, ghost1(fy1, u1n_i));
                __syncwarp();
                if (lane == 0) mbar_arrive(u1full0 + 8 * slot);
            }
            if (l2) {
                double dn = u1lo_i, up = u1n_i;
                if (it == cp.it_zlo2) dn = ghost1(fz0, u1c_i);
                if (!l1) up = ghost1(fz1, u1c_i);
                const double v = fin3(P2, c3, up, dn);
                if (ok2) sts1(u2i + w2 * C::U2_SLOT, v);
                __syncwarp();
                if (lane == 0) mbar_arrive(u2full0 + 8 * w2);
            } else {
                if (lane == 0) mbar_arrive(u2full0 + 8 * w2);             // keeps the barrier's phases in step with `it`
            }
            if (l1) {
                below_i = cur_i; below_o = cur_o;
                cur_i = above_i; cur_o = above_o;
            }
            u1lo_i = u1c_i; u1c_i = u1n_i; q1_i = qc_i;
        }
    } else if (warp >= TY) {
        // ------------------------------ inner-row warps: row j0-1 (warp TY), row j0+12 (warp TY+1) ------------------------------
        const int trow = warp == TY ? -1 : TY;
        const long long y = j0 + trow;
        const bool row_ok = y >= 0 && y < n1;
        const long long xa = i0 + 2 * lane, xb = xa + 64;
        RowK3 k;
        k.ok_a = row_ok && xa < n0;
        k.ok_b = row_ok && xb < n0;
        k.pa = u0_0 + (uint32_t)(((trow + 3) * C::BX0 + 4 + 2 * lane) * 8);
        k.ra = rhs_0 + (uint32_t)(((trow + 2) * C::BXR + 2 + 2 * lane) * 8);
        k.u1a = u1_0 + (uint32_t)(trow + 2) * RBU + (uint32_t)(8 * (1 + lane));
        k.u2a = u2_0 + (uint32_t)(trow + 1) * RBU + (uint32_t)(8 * (1 + lane));
        k.bar0 = full0;
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fz0 = make_face1(p.faces.f[4], true); k.fz1 = make_face1(p.faces.f[5], true);
        k.xlo = k.ok_a && k.fx0.on && xa == 0;
        k.xhi = !k.fx1.on ? -1 : (k.ok_a && xa + 2 == n0) ? 0 : (k.ok_b && xb + 2 == n0) ? 1 : -1;
        k.lane = lane;
        RowSt3 st;
        const double2 zero2 = make_double2(0.0, 0.0);
        mbar_wait3(full0, 0);
        st.below[0] = lds2(k.pa); st.below[1] = lds2(k.pa + 512);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait3(full0 + 8, 0);
        st.cur[0] = lds2(k.pa + C::U0_STAGE); st.cur[1] = lds2(k.pa + C::U0_STAGE + 512);
        for (int h = 0; h < 2; ++h) { st.u1lo[h] = zero2; st.u1c[h] = zero2; st.q1[h] = zero2; }
        int it = 0;
        for (; it < cp.it2_first && it < nL1; ++it) row_iter3<true, false>(p, cp, k, st, it);
#pragma unroll 1
        for (; it < nL1; ++it) row_iter3<true, true>(p, cp, k, st, it);
        for (; it <= cp.it2_last; ++it) row_iter3<false, true>(p, cp, k, st, it);
    } else {
        // ------------------------------ quad warps: rows j0+2*rp, j0+2*rp+1 x columns i0+64*xh .. +63 ------------------------------
        const int xh = warp & 1, rp = warp >> 1;
        const int ra = 2 * rp;                                  // tile row of row A
        const long long jA = j0 + ra, jB = jA + 1;
        const int t = 32 * xh + lane;                           // pair index within the tile row
        const long long ia = i0 + 2 * t;
        QuadK3 k;
        k.ok = jA < n1 && ia < n0;
        k.pa = u0_0 + (uint32_t)(((ra + 3) * C::BX0 + 4 + 2 * t) * 8);
        k.ra = rhs_0 + (uint32_t)(((ra + 2) * C::BXR + 2 + 2 * t) * 8);
        k.u1a = u1_0 + (uint32_t)(ra + 2) * RBU + (uint32_t)(8 * (1 + t));
        k.u2a = u2_0 + (uint32_t)(ra + 1) * RBU + (uint32_t)(8 * (1 + t));
        k.bar0 = full0;
        // offset of the lane's first cell inside a grown xy plane (the neighbours' ghost / deep planes)
        k.xy_off = p.g.xoff + (ia + p.g.ng) + p.g.pitch * (jA + p.g.ng);
        k.lane = lane;
        // ghost rules of the six faces (a z face with mode 0 is a slab neighbour)
        k.fx0 = make_face1(p.faces.f[0], true); k.fx1 = make_face1(p.faces.f[1], true);
        k.fy0 = make_face1(p.faces.f[2], true); k.fy1 = make_face1(p.faces.f[3], true);
        k.fz0 = make_face1(p.faces.f[4], true); k.fz1 = make_face1(p.faces.f[5], true);
        k.xlo = k.ok && k.fx0.on && ia == 0;
        k.xhi = k.ok && k.fx1.on && ia + 2 == n0;
        k.ylo = k.ok && k.fy0.on && jA == 0;
        k.yhi = k.ok && k.fy1.on && jB == n1 - 1;
        // u3 planes that also go to a neighbour: plane P is formed by iteration P - a1 + 2 of the chunk that holds it
        auto it_of = [&](long long P, bool on) -> int { return on && P >= k0 && P < k1 ? (int)(P - a1) + 2 : -1; };
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            k.it_send[e] = it_of(e, p.peer_lo[e] != nullptr);
            k.it_send[3 + e] = it_of(n2 - 1 - e, p.peer_hi[e] != nullptr);
        }

        QuadSt3 st;
        const double2 zero2 = make_double2(0.0, 0.0);
        mbar_wait3(full0, 0);
        st.below_a = lds2(k.pa); st.below_b = lds2(k.pa + RB0);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
        mbar_wait3(full0 + 8, 0);
        st.cur_a = lds2(k.pa + C::U0_STAGE); st.cur_b = lds2(k.pa + C::U0_STAGE + RB0);
        st.u1lo_a = zero2; st.u1lo_b = zero2; st.u1c_a = zero2; st.u1c_b = zero2;
        st.u2lo_a = zero2; st.u2lo_b = zero2; st.u2c_a = zero2; st.u2c_b = zero2;
        st.q1_a = zero2; st.q1_b = zero2; st.q2_a = zero2; st.q2_b = zero2;
        st.oa = p.out + p.g.at(ia, jA, k0, 0);

        // warp-uniform: all 32 patches are whole and no y face touches them -> the steady iterations that meet neither a z face
        // nor a neighbour's plane run without ghost code / store masks
        const long long iw = i0 + 64 * xh;
        const bool whole = jA < n1 && iw + 64 <= n0 && !(k.fy0.on && jA == 0) && !(k.fy1.on && jB == n1 - 1);
        const int edge = !whole ? 2 : ((k.fx0.on && iw == 0) || (k.fx1.on && iw + 64 == n0)) ? 1 : 0;
        // general code: the first level-3 iterations of the frame's first chunk (z-lo face: plane 0; neighbour: planes 0, 1, 2) and
        // the last three of its last chunk when a neighbour sits behind the z-hi face
        const int head = k0 == 0 ? (p.nb_lo ? (p.peer_lo[0] ? 3 : 0) : 1) : 0;
        const int tail = (k1 == n2 && p.nb_hi && p.peer_hi[0]) ? 3 : 0;
        const int it_end = cp.it3_last + 1;
        int it = 0;
        const int fast_begin = cp.it3_first + head < nL1 ? cp.it3_first + head : nL1;
        const int fast_end = nL1 - tail > fast_begin ? nL1 - tail : fast_begin;
        for (; it < fast_begin; ++it) quad_general3(p, cp, k, st, it);
        if (edge == 0) quad_run3<0>(p, cp, k, st, it, fast_end);
        else if (edge == 1) quad_run3<1>(p, cp, k, st, it, fast_end);
        else quad_run3<2>(p, cp, k, st, it, fast_end);
        for (; it < it_end; ++it) quad_general3(p, cp, k, st, it);
    }
    if (sync_lo || sync_hi) {
        __syncthreads();                                       // every peer store of this CTA has been issued
        if (threadIdx.x == 0) {
            if (sync_lo) peer_publish(p.ps.done + 0, p.ps.target_lo, p.ps.sig_lo, p.ps.epoch + 1);
            if (sync_hi) peer_publish(p.ps.done + 1, p.ps.target_hi, p.ps.sig_hi, p.ps.epoch + 1);
        }
    }
}

// chunk length: a CTA pays (its planes + 5) plane iterations, and a launch of W = CTAs / 148 waves takes about W + 1 CTA times (one
// CTA per SM; the CTAs of a wave do not finish together, so the last wave costs a whole one however full it is).  Fitted to thin
// frames, where it matters (1024 x 1024 x 128: 416 GLUP/s with one 128-plane chunk per tile, 455 with 48-plane chunks; the
// ceil(W) model of the two-sweep kernel chose the former): profiles/r2_notes.md, section 9
void plan_tb3(const SweepArgs& a, dim3& grid, int& chunk, int& n_chunks)
{
    const FrameGeom& g = a.g;
    const long long tiles = ((g.n[0] + 127) / 128) * ((g.n[1] + Cfg3::TY - 1) / Cfg3::TY);
    const long long span = a.k_end - a.k_begin;
    long long best_c = 1;
    double best_cost = -1.0;
    static const int cand[] = { 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128 };
    for (int c : cand) {
        const long long cc = c > span ? span : c;
        if (cc < 1) break;
        const long long chunks = (span + cc - 1) / cc;
        if (chunks <= 60000) {
            const double waves = (double)(tiles * chunks) / 148.0;
            const double cost = (waves + 1.0) * ((double)span / (double)chunks + 5.0);
            // that many chunks of (almost) equal length
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = (span + chunks - 1) / chunks; }
        }
        if (c >= span) break;
    }
    if (a.chunk_override > 0) best_c = a.chunk_override > span ? span : a.chunk_override;
    // slabs: the first / last chunk holds all three planes that go to the neighbour
    if ((a.nb_lo || a.nb_hi) && best_c < 3) best_c = span < 3 ? span : 3;
    if (best_c < 1) best_c = 1;
    chunk = (int)best_c;
    n_chunks = (int)((span + best_c - 1) / best_c);
    // a last chunk of fewer than three planes would split the planes that go to the hi neighbour over two chunks: lengthen the chunks
    while (a.nb_hi && n_chunks > 1 && span - (long long)(n_chunks - 1) * chunk < 3) {
        chunk += 1;
        n_chunks = (int)((span + chunk - 1) / chunk);
    }
    grid = dim3((unsigned)((g.n[0] + 127) / 128), (unsigned)((g.n[1] + Cfg3::TY - 1) / Cfg3::TY), (unsigned)n_chunks);
}

}  // namespace

// frames the three-sweep kernel takes (the others run the two-sweep kernel)
bool triple_sweep_applicable(const FrameGeom& g)
{
    return g.dim == 3 && g.ng == 1 && g.ncomp == 1 && g.n[0] % 2 == 0 && g.n[1] % 2 == 0;
}

int launch_jacobi_triple_sweep(const SweepArgs& a, cudaStream_t s)
{
    using C = Cfg3;
    if (a.k_end <= a.k_begin) return 0;
    if (!triple_sweep_applicable(a.g)) { rnmf_set_error("triple sweep: needs a 3D one-component frame with one ghost layer and even n0, n1"); return -1; }
    if (a.hole_end > a.hole_begin) { rnmf_set_error("triple sweep: plane ranges with a hole are not supported"); return -1; }
    if ((a.nb_lo || a.nb_hi) && (a.deep < 2 || a.k_begin != 0 || a.k_end != a.g.n[2] || a.g.n[2] < 6)) {
        rnmf_set_error("triple sweep across slabs needs two deep halo planes, the whole slab and >= 6 planes");
        return -1;
    }
    dim3 grid;
    int chunk, n_chunks;
    plan_tb3(a, grid, chunk, n_chunks);
    CUtensorMap tm_u0, tm_rhs;
    if (tma3d_encode_map_deep(&tm_u0, a.g, a.in, C::BX0, C::U0_ROWS, a.deep)) return -1;
    if (tma3d_encode_map_deep(&tm_rhs, a.g, a.rhs, C::BXR, C::R_ROWS, a.deep)) return -1;
    Tb3Params p;
    p.g = a.g; p.out = a.out; p.c = a.c; p.faces = a.faces;
    p.m_begin = a.k_begin; p.m_end = a.k_end; p.chunk = chunk; p.n_chunks = n_chunks;
    p.nb_lo = a.nb_lo; p.nb_hi = a.nb_hi; p.zshift = a.deep;
    for (int e = 0; e < 3; ++e) {
        // the neighbours' ghost planes and the deep planes behind / in front of them
        p.peer_lo[e] = a.peer_lo_plane ? a.peer_lo_plane + a.g.plane * e : nullptr;
        p.peer_hi[e] = a.peer_hi_plane ? a.peer_hi_plane - a.g.plane * e : nullptr;
    }
    p.ps = a.ps;
#ifndef RNMF_EMULATE
    static PerDeviceOnce attr_set;
    if (!attr_set.done()) {
        if (cudaFuncSetAttribute(jacobi3d_tb3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) return -1;
        attr_set.mark();
    }
#endif
    RNMF_KERNEL_LAUNCH(jacobi3d_tb3_kernel, grid, C::THREADS, C::SMEM, s, tm_u0, tm_rhs, p);
    return 1;
}

// CTAs per boundary side (tiles across a z-plane): the first / last chunk of the launch takes part in the handshake
int triple_sweep_boundary_ctas(const SweepArgs& a)
{
    dim3 grid;
    int chunk, n_chunks;
    plan_tb3(a, grid, chunk, n_chunks);
    return (int)(grid.x * grid.y);
}
