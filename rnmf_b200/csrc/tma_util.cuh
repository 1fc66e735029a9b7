// tma_util.cuh -- device helpers shared by the 3D TMA sweep kernels (kernels_sweep_tma.cu,
// kernels_sweep_tb2.cu): mbarrier / cp.async.bulk.tensor wrappers, shared-memory accessors, the
// explicitly rounded 7-point update and the branch-free ng=1 ghost rule.
#pragma once
#ifdef RNMF_EMULATE
#include "emu_cuda.hpp"   // tests/emu (test builds only: g++ -DRNMF_EMULATE -Itests/emu); never part of librnmf_b200.so
#endif
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "common.cuh"

namespace tmau {

#ifdef RNMF_EMULATE
// ---- host emulation of the PTX wrappers (tests/emu/emu_cuda.hpp; test builds only) --------------------------
inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::g_cta->smem); }
inline void mbar_init(uint32_t bar, uint32_t count) { emu::mbar_init(bar, count); }
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { emu::mbar_arrive(bar, bytes); }
inline void mbar_arrive(uint32_t bar) { emu::mbar_arrive(bar); }
inline void mbar_wait(uint32_t bar, uint32_t parity) { emu::mbar_wait(bar, parity); }
inline void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) { emu::tma_load(dst, tm, c0, c1, c2, bar); }
inline void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) { emu::tma_load(dst, tm, c0, c1, 0, bar); }
inline uint64_t policy_evict_first() { return 0; }
inline void tma_load_3d_hint(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar, uint64_t) { emu::tma_load(dst, tm, c0, c1, c2, bar); }
inline void st2_hint(double* p, const double2& v, uint64_t) { p[0] = v.x; p[1] = v.y; }
__attribute__((always_inline)) inline double2 lds2(uint32_t a) { emu::smem_access(a, 16, false); double2 v; std::memcpy(&v, emu::g_cta->smem + a, 16); return v; }
__attribute__((always_inline)) inline double lds1(uint32_t a) { emu::smem_access(a, 8, false); double v; std::memcpy(&v, emu::g_cta->smem + a, 8); return v; }
__attribute__((always_inline)) inline void sts2(uint32_t a, const double2& v) { emu::smem_access(a, 16, true); std::memcpy(emu::g_cta->smem + a, &v, 16); }
__attribute__((always_inline)) inline void sts1(uint32_t a, double v) { emu::smem_access(a, 8, true); std::memcpy(emu::g_cta->smem + a, &v, 8); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// The suspend-time hint lets a waiting warp sleep in the barrier unit instead of spinning through try_wait / branch / yield:
// the consumers of a DRAM-bound pipeline wait most of the time, and under the 1 kW cap every issued instruction costs clock.
#ifndef RNMF_MBAR_HINT_NS
#define RNMF_MBAR_HINT_NS 20000
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
#if RNMF_MBAR_HINT_NS > 0
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(RNMF_MBAR_HINT_NS)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// L2 eviction-priority hints: rhs and psi' are touched once per sweep, psi tiles are re-read by the
// neighbouring CTAs (xy halo) a few microseconds later; at 6.6 TB/s the 126 MB L2 turns over in
// ~19 us, so streaming the one-touch traffic with evict_first keeps the psi lines alive longer.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void st2_hint(double* p, const double2& v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
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
__device__ __forceinline__ void sts2(uint32_t a, const double2& v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void sts1(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

#endif  // RNMF_EMULATE

// poisson.rs:83-88 with the reference's association, every operation explicitly rounded
__device__ __forceinline__ double upd3(double c0, double c1, double c2, double c3, double r, double e, double w,
                                       double n, double s, double u, double d)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c0, r), __dmul_rn(c1, __dadd_rn(e, w))),
                               __dmul_rn(c2, __dadd_rn(n, s))),
                     __dmul_rn(c3, __dadd_rn(u, d)));
}

// One face of the ng=1 ghost fill in branch-free form:  ghost = (neg ? -m : m) + add
//   Dirichlet : 2v - m  == (-m) + 2v        Neumann lo: m - s == m + (-s)        Neumann hi: m + s
// (IEEE: a - b == a + (-b) bit for bit, addition commutes), dataframe.rs:289-355.
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
__device__ __forceinline__ double2 ghost1(const Face1& f, const double2& m) { return make_double2(ghost1(f, m.x), ghost1(f, m.y)); }

// store the four cells of a lane (two pairs, 64 cells apart) with the tile-edge validity masks
__device__ __forceinline__ void store4(double* a, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
#ifdef RNMF_EMULATE
    if ((full || a1 || b1) && reinterpret_cast<uintptr_t>(a) % 16 != 0) emu::fault("misaligned 16-byte global store", reinterpret_cast<uintptr_t>(a), 16);
#endif
    if (full) {
        *reinterpret_cast<double2*>(a) = va;
        *reinterpret_cast<double2*>(a + 64) = vb;
    } else {
        if (a1) *reinterpret_cast<double2*>(a) = va; else if (a0) a[0] = va.x;
        if (b1) *reinterpret_cast<double2*>(a + 64) = vb; else if (b0) a[64] = vb.x;
    }
}
// same, into shared memory (byte address)
__device__ __forceinline__ void sstore4(uint32_t a, const double2& va, const double2& vb, bool full, bool a0, bool a1, bool b0, bool b1)
{
    if (full) {
        sts2(a, va);
        sts2(a + 512, vb);
    } else {
        if (a1) sts2(a, va); else if (a0) sts1(a, va.x);
        if (b1) sts2(a + 512, vb); else if (b0) sts1(a + 512, vb.x);
    }
}

}  // namespace tmau

// dynamic shared memory of the CTA (the emulator hands each CTA a host buffer)
#ifdef RNMF_EMULATE
#define RNMF_DYNAMIC_SMEM(name) unsigned char* const name = emu::g_cta->smem
#else
#define RNMF_DYNAMIC_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

// kernel launch: <<<>>> on the device, the thread-per-lane emulator in -DRNMF_EMULATE test builds
#ifdef RNMF_EMULATE
#define RNMF_KERNEL_LAUNCH(kernel, grid, threads, smem, stream, ...) emu_launch_checked(kernel, grid, threads, smem, __VA_ARGS__)
#else
#define RNMF_KERNEL_LAUNCH(kernel, grid, threads, smem, stream, ...) kernel<<<grid, threads, smem, stream>>>(__VA_ARGS__)
#endif

// tensor map of a frame buffer with a (bx, by, 1) box (kernels_sweep_tma.cu; cached per buffer/geometry/box)
int tma3d_encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by);
// same; deep > 0: the tensor starts `deep` planes BEFORE `base` and has G2 + 2*deep planes (the deep halo planes of the buffer),
// so a box at z coordinate c covers grown plane c - deep
int tma3d_encode_map_deep(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by, int deep);
// tensor map of a 2D frame buffer with a (bx, 1) box (kernels_sweep2d_tma.cu)
int tma2d_encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx);
// same; deep > 0: the tensor starts `deep` rows BEFORE `base` and has G1 + 2*deep rows, so a box at y coordinate c covers grown row c - deep
int tma2d_encode_map_deep(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int deep);
