// emu_cuda.hpp -- TEST INFRASTRUCTURE, never part of librnmf_b200.so.
// A functional SIMT emulator for this repository's own kernels: with -DRNMF_EMULATE the kernel
// sources (kernels_sweep_tb2.cu, kernels_sweep2d_tb2.cu) compile with g++ and run on host threads,
// one std::thread per CUDA thread, one CTA at a time per launch:
//   * __syncthreads / __syncwarp            -> counting barriers
//   * mbarrier (init / arrive / expect_tx / complete_tx / try_wait.parity) -> a table of phase counters
//   * cp.async.bulk.tensor (TMA)            -> a box copy with out-of-bounds zero fill + complete_tx
//   * ld/st.shared                          -> accesses to the CTA's shared-memory array by byte offset
//   * __dadd_rn / __dmul_rn / __dsub_rn     -> plain IEEE operations (build with -ffp-contract=off)
//   * the peer handshake (acquire/release flags, completion counters) -> std atomics
// It checks what the GPU box is too expensive to iterate on: tile / halo / ghost index logic, ring and
// barrier protocols (a wrong parity or a missing arrival deadlocks here as it would on the device: the
// launcher gives up after a timeout), and -- with several launches running concurrently on "devices"
// that share the process's memory -- the slab handshake.  It does not model timing, memory-model
// reordering or register limits.  It does check alignment and bounds of shared-memory accesses, TMA destinations and
// boxes (device faults), and an optional static bank model counts shared-memory wavefronts (below).  tests/test_emulated_kernels.py drives it on the CPU box.
#pragma once
#ifndef RNMF_EMULATE
#error "emu_cuda.hpp is for -DRNMF_EMULATE builds (tests) only"
#endif

#define __launch_bounds__(...)
#define __grid_constant__

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace emu {

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int count = 0, waiting = 0;
    unsigned long long gen = 0;
    explicit Barrier(int n) : count(n) {}
    // abort: set when some wait of the CTA timed out; every waiter then falls through so the CTA can end
    void arrive_and_wait(const std::atomic<bool>& abort)
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g = gen;
        if (++waiting == count) {
            waiting = 0;
            ++gen;
            cv.notify_all();
        } else {
            while (gen == g && !abort.load()) cv.wait_for(lk, std::chrono::milliseconds(200));
        }
    }
};

struct MBar {
    int init = 0, pending = 0;
    long long tx = 0;
    unsigned phase = 0;          // completed phases; the phase in progress has parity phase & 1
};

struct Cta {
    std::mutex mu;
    std::condition_variable cv;
    std::map<uint32_t, MBar> bars;
    std::unique_ptr<Barrier> cta_bar;
    std::vector<std::unique_ptr<Barrier>> warp_bars;
    unsigned char* smem = nullptr;
    size_t smem_bytes = 0;
    std::atomic<bool> timed_out{false};
};

// ---- access checks the device would turn into "misaligned address" / "illegal address" faults ----
// every ld/st.shared must be naturally aligned and inside the CTA's dynamic shared memory; a TMA destination must be
// 128-byte aligned (no swizzle), an mbarrier 8-byte aligned.  Violations are counted (the launch then reports failure).
inline std::atomic<int> g_faults{0};
inline void fault(const char* what, unsigned long long a, unsigned long long n)
{
    if (g_faults.fetch_add(1) < 5) std::fprintf(stderr, "emu: %s (address %llu, %llu bytes)\n", what, a, n);
}

inline thread_local Cta* g_cta = nullptr;
inline thread_local int g_warp = 0;
inline thread_local int g_lane = 0;
inline thread_local unsigned g_seq = 0;           // shared-memory accesses this thread has made in its CTA (trace key)

// ---- optional shared-memory wavefront accounting (a STATIC bank model, not a timing model) ---------------
// With tracing on, every ld/st.shared of a warp is keyed by its call site and how often the lane has executed it, so
// one key = one warp instruction (of the lanes that take part in it).  Its cost is the classic bank model: 32 banks of 4 bytes, a
// warp instruction is served in passes of 32 / 16 / 8 lanes for 4- / 8- / 16-byte accesses, and a pass needs as many
// wavefronts as the largest number of distinct 128-byte rows any bank is asked for.
struct Access { uint32_t addr; unsigned char lane, bytes, store; };
inline std::atomic<bool> g_trace{false};
inline std::mutex g_trace_mu;
struct TraceKey {
    int warp; const void* site; unsigned nth;
    bool operator<(const TraceKey& o) const { return warp != o.warp ? warp < o.warp : site != o.site ? site < o.site : nth < o.nth; }
};
inline std::map<TraceKey, std::vector<Access>> g_trace_cta;      // (warp, call site, n-th execution by the lane) -> accesses, current CTA
inline thread_local std::map<const void*, unsigned> g_site_count;
inline std::atomic<long long> g_wf_load{0}, g_wf_store{0}, g_instr_load{0}, g_instr_store{0};
// one warp instruction = the same call site (the ld/st wrappers are always_inline, so the return address identifies the
// source-level access) executed for the n-th time by each lane: robust against lanes that make extra accesses elsewhere
__attribute__((noinline)) inline void smem_access(uint32_t addr, int bytes, bool store)
{
    if (addr % (uint32_t)bytes != 0) fault("misaligned ld/st.shared", addr, (unsigned long long)bytes);
    if ((size_t)addr + (size_t)bytes > g_cta->smem_bytes) fault("ld/st.shared outside the CTA's shared memory", addr, (unsigned long long)bytes);
    if (!g_trace.load(std::memory_order_relaxed)) return;
    const void* site = __builtin_return_address(0);
    const unsigned nth = g_site_count[site]++;
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_trace_cta[TraceKey{ g_warp, site, nth }].push_back(Access{ addr, (unsigned char)g_lane, (unsigned char)bytes, (unsigned char)store });
}
inline int wavefronts_of(const std::vector<Access>& v)
{
    const int bytes = v[0].bytes, lpp = bytes == 16 ? 8 : bytes == 8 ? 16 : 32;
    int total = 0;
    for (int pass = 0; pass < 32 / lpp; ++pass) {
        std::map<int, std::vector<uint32_t>> rows;     // bank -> distinct rows
        bool any = false;
        for (const Access& a : v) {
            if (a.lane / lpp != pass) continue;
            any = true;
            for (int w = 0; w < bytes / 4; ++w) {
                const uint32_t word = a.addr / 4 + w;
                auto& r = rows[(int)(word % 32)];
                bool seen = false;
                for (uint32_t x : r) seen = seen || x == word / 32;
                if (!seen) r.push_back(word / 32);
            }
        }
        if (!any) continue;
        size_t worst = 1;
        for (auto& kv : rows) worst = kv.second.size() > worst ? kv.second.size() : worst;
        total += (int)worst;
    }
    return total;
}
inline void trace_flush_cta()
{
    std::lock_guard<std::mutex> lk(g_trace_mu);
    for (auto& kv : g_trace_cta) {
        const int wf = wavefronts_of(kv.second);
        if (kv.second[0].store) { g_wf_store += wf; ++g_instr_store; } else { g_wf_load += wf; ++g_instr_load; }
    }
    g_trace_cta.clear();
}
inline std::atomic<int> g_wait_timeout_s{20};     // a barrier wait longer than this = deadlock

inline void mbar_check(MBar& b, Cta& c)
{
    if (b.pending == 0 && b.tx == 0) {
        ++b.phase;
        b.pending = b.init;
        c.cv.notify_all();
    }
}
inline void mbar_init(uint32_t bar, uint32_t count)
{
    if (bar % 8 != 0 || (size_t)bar + 8 > g_cta->smem_bytes) fault("mbarrier misaligned / outside shared memory", bar, 8);
    Cta& c = *g_cta;
    std::lock_guard<std::mutex> lk(c.mu);
    MBar b;
    b.init = b.pending = (int)count;
    c.bars[bar] = b;
}
inline void mbar_arrive(uint32_t bar, long long expect_bytes = 0)
{
    Cta& c = *g_cta;
    std::lock_guard<std::mutex> lk(c.mu);
    MBar& b = c.bars.at(bar);
    b.tx += expect_bytes;
    --b.pending;
    mbar_check(b, c);
}
inline void mbar_complete_tx(uint32_t bar, long long bytes)
{
    Cta& c = *g_cta;
    std::lock_guard<std::mutex> lk(c.mu);
    MBar& b = c.bars.at(bar);
    b.tx -= bytes;
    mbar_check(b, c);
}
inline void mbar_wait(uint32_t bar, uint32_t parity)
{
    Cta& c = *g_cta;
    std::unique_lock<std::mutex> lk(c.mu);
    MBar& b = c.bars.at(bar);
    const bool ok = c.cv.wait_for(lk, std::chrono::seconds(g_wait_timeout_s.load()), [&] { return (b.phase & 1u) != (parity & 1u) || c.timed_out.load(); });
    if (!ok) {
        c.timed_out = true;
        c.cv.notify_all();
    }
}

// ---- tensor maps ---------------------------------------------------------------------------------
struct Map {
    const double* base;
    int rank;
    unsigned long long dims[3];
    unsigned long long stride[3];     // elements
    unsigned box[3];
};
inline void encode(CUtensorMap* tm, const Map& m)
{
    static_assert(sizeof(Map) <= sizeof(CUtensorMap), "Map must fit the opaque tensor map");
    std::memset(tm, 0, sizeof(*tm));
    std::memcpy(tm, &m, sizeof(m));
}
inline void tma_load(uint32_t dst, const CUtensorMap* tm, long long c0, long long c1, long long c2, uint32_t bar)
{
    Map m;
    std::memcpy(&m, tm, sizeof(m));
    double* out = reinterpret_cast<double*>(g_cta->smem + dst);
    const unsigned bz = m.rank == 3 ? m.box[2] : 1;
    const size_t box_bytes = (size_t)m.box[0] * m.box[1] * bz * 8;
    if (dst % 128 != 0) fault("TMA destination not 128-byte aligned", dst, box_bytes);
    if ((size_t)dst + box_bytes > g_cta->smem_bytes) fault("TMA box outside the CTA's shared memory", dst, box_bytes);
    if ((m.box[0] * 8) % 16 != 0 || m.box[0] > 256 || m.box[1] > 256) fault("TMA box shape not encodable", m.box[0], m.box[1]);
    if ((reinterpret_cast<uintptr_t>(m.base) % 16) != 0 || (m.stride[1] * 8) % 16 != 0) fault("TMA global base / stride not 16-byte aligned", 0, 0);
    long long bytes = 0;
    for (unsigned z = 0; z < bz; ++z)
        for (unsigned y = 0; y < m.box[1]; ++y)
            for (unsigned x = 0; x < m.box[0]; ++x) {
                const long long X = c0 + x, Y = c1 + y, Z = c2 + z;
                const bool in = X >= 0 && Y >= 0 && Z >= 0 && (unsigned long long)X < m.dims[0] && (unsigned long long)Y < m.dims[1] &&
                                (m.rank < 3 || (unsigned long long)Z < m.dims[2]);
                *out++ = in ? m.base[X + (long long)m.stride[1] * Y + (m.rank == 3 ? (long long)m.stride[2] * Z : 0)] : 0.0;
                bytes += 8;
            }
    mbar_complete_tx(bar, bytes);
}

// ---- launch ------------------------------------------------------------------------------------------
// One CTA at a time in blockIdx order (x fastest): the most adversarial schedule for anything that
// assumes co-residency.  Returns false if a barrier wait timed out (deadlock) in some CTA.
template <class Kernel, class... Args>
bool launch(Kernel kernel, dim3 grid, int threads, unsigned char* smem, size_t smem_bytes, Args... args);

}  // namespace emu

// CUDA built-ins
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 gridDim, blockDim;
inline void __syncthreads() { emu::g_cta->cta_bar->arrive_and_wait(emu::g_cta->timed_out); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_cta->warp_bars[emu::g_warp]->arrive_and_wait(emu::g_cta->timed_out); }
// broadcast of a value that every lane of the warp holds anyway (the kernels use it to tell the compiler "warp-uniform")
inline int __shfl_sync(unsigned, int v, int) { return v; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }

namespace emu {
template <class Kernel, class... Args>
bool launch(Kernel kernel, dim3 grid, int threads, unsigned char* smem, size_t smem_bytes, Args... args)
{
    bool ok = true;
    const int faults_before = g_faults.load();
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Cta cta;
                cta.smem = smem;
                cta.smem_bytes = smem_bytes;
                std::memset(smem, 0xA5, smem_bytes);                 // uninitialised shared memory is garbage, not zero
                cta.cta_bar.reset(new Barrier(threads));
                for (int w = 0; w < (threads + 31) / 32; ++w) {
                    const int lanes = threads - 32 * w < 32 ? threads - 32 * w : 32;
                    cta.warp_bars.emplace_back(new Barrier(lanes));
                }
                std::vector<std::thread> pool;
                pool.reserve(threads);
                for (int t = 0; t < threads; ++t)
                    pool.emplace_back([&, t] {
                        g_cta = &cta;
                        g_warp = t / 32;
                        g_lane = t % 32;
                        g_seq = 0;
                        g_site_count.clear();
                        threadIdx = uint3{ (unsigned)t, 0, 0 };
                        blockIdx = uint3{ bx, by, bz };
                        gridDim = grid;
                        blockDim = dim3(threads, 1, 1);
                        kernel(args...);
                    });
                for (auto& th : pool) th.join();
                if (g_trace.load()) trace_flush_cta();
                if (cta.timed_out) ok = false;
                if (g_faults.load() > faults_before) ok = false;
            }
    return ok;
}
}  // namespace emu
