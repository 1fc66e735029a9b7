// emu_driver.cpp -- TEST INFRASTRUCTURE: the double-sweep kernel SOURCES (kernels_sweep_tb2.cu,
// kernels_sweep2d_tb2.cu, and the K-sweep kernels_sweep2d_tbk.cu) compiled with g++ -DRNMF_EMULATE and run on host threads by the SIMT emulator
// (tests/emu/emu_cuda.hpp).  Exposes a small C interface for tests/test_emulated_kernels.py:
// dense reference-layout frames in, dense frames out, through the very launchers (planner, tensor maps,
// A/B chunk roles, handshake flags) the library uses.  Slabs: every "rank" is a host thread with its own
// buffers; peer pointers are plain pointers into the neighbours' buffers.
#define RNMF_EMULATE 1
#include "emu_cuda.hpp"
#include "../../rnmf_b200/csrc/tma_util.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdlib>

namespace {
std::atomic<int> g_deadlocks{0};

template <class Kernel, class... Args>
void emu_launch_checked(Kernel kernel, dim3 grid, int threads, size_t smem_bytes, Args... args)
{
    std::vector<unsigned char> raw(smem_bytes + 2048);                 // one CTA at a time per launch
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw.data()) + 1023) & ~uintptr_t(1023));
    if (!emu::launch(kernel, grid, threads, smem, smem_bytes, args...)) ++g_deadlocks;
}
}  // namespace

void rnmf_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}

// what cuTensorMapEncodeTiled describes, for the emulated TMA (dims / strides as kernels_sweep_tma.cu encodes them)
int tma3d_encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by)
{
    emu::Map m;
    m.base = base; m.rank = 3;
    m.dims[0] = (unsigned long long)g.pitch; m.dims[1] = (unsigned long long)g.G[1]; m.dims[2] = (unsigned long long)(g.G[2] * g.ncomp);
    m.stride[0] = 1; m.stride[1] = (unsigned long long)g.pitch; m.stride[2] = (unsigned long long)g.plane;
    m.box[0] = (unsigned)bx; m.box[1] = (unsigned)by; m.box[2] = 1;
    emu::encode(tm, m);
    return 0;
}
int tma3d_encode_map_deep(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int by, int deep)
{
    emu::Map m;
    m.base = base - (long long)deep * g.plane; m.rank = 3;
    m.dims[0] = (unsigned long long)g.pitch; m.dims[1] = (unsigned long long)g.G[1]; m.dims[2] = (unsigned long long)(g.G[2] * g.ncomp + 2 * deep);
    m.stride[0] = 1; m.stride[1] = (unsigned long long)g.pitch; m.stride[2] = (unsigned long long)g.plane;
    m.box[0] = (unsigned)bx; m.box[1] = (unsigned)by; m.box[2] = 1;
    emu::encode(tm, m);
    return 0;
}
int tma2d_encode_map(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx)
{
    emu::Map m;
    m.base = base; m.rank = 2;
    m.dims[0] = (unsigned long long)g.pitch; m.dims[1] = (unsigned long long)(g.G[1] * g.ncomp); m.dims[2] = 1;
    m.stride[0] = 1; m.stride[1] = (unsigned long long)g.pitch; m.stride[2] = 0;
    m.box[0] = (unsigned)bx; m.box[1] = 1; m.box[2] = 1;
    emu::encode(tm, m);
    return 0;
}

int tma2d_encode_map_deep(CUtensorMap* tm, const FrameGeom& g, const double* base, int bx, int deep)
{
    emu::Map m;
    m.base = base - (long long)deep * g.pitch; m.rank = 2;
    m.dims[0] = (unsigned long long)g.pitch; m.dims[1] = (unsigned long long)(g.G[1] * g.ncomp + 2 * deep); m.dims[2] = 1;
    m.stride[0] = 1; m.stride[1] = (unsigned long long)g.pitch; m.stride[2] = 0;
    m.box[0] = (unsigned)bx; m.box[1] = 1; m.box[2] = 1;
    emu::encode(tm, m);
    return 0;
}

#include "../../rnmf_b200/csrc/kernels_sweep_tb2.cu"
#include "../../rnmf_b200/csrc/kernels_sweep_tb3.cu"
#include "../../rnmf_b200/csrc/kernels_sweep2d_tb2.cu"
#include "../../rnmf_b200/csrc/kernels_sweep2d_tbk.cu"

namespace {

constexpr long long kTail = 64 + 256;       // tail pad + guard band of the library's buffers (capi.cu)

// a frame as the library allocates it (capi.cu: alloc_buffer): [deep plane][grown frame][deep plane][tail pad + guard band];
// 3D: two deep planes either side, 2D: three deep rows.  d(b) = the frame pointer (lo ghost slice) of ping-pong buffer b.
struct Frame {
    FrameGeom g;
    long long deep = 0;             // doubles
    int deep_slices = 0;
    std::vector<double> buf[2];
    std::vector<double> rhs_buf;
    int cur = 0;
    double* d(int b) { return buf[b].data() + deep; }
    double* rhs() { return rhs_buf.data() + deep; }
};

void unpack(const FrameGeom& g, const double* dense, double* pitched)
{
    const long long rows = g.G[1] * g.G[2];
    for (long long r = 0; r < rows; ++r)
        for (long long i = 0; i < g.G[0]; ++i) pitched[g.xoff + i + g.pitch * r] = dense[i + g.G[0] * r];
}
void pack(const FrameGeom& g, const double* pitched, double* dense)
{
    const long long rows = g.G[1] * g.G[2];
    for (long long r = 0; r < rows; ++r)
        for (long long i = 0; i < g.G[0]; ++i) dense[i + g.G[0] * r] = pitched[g.xoff + i + g.pitch * r];
}
// one dense grown slice (3D: a plane of G1 rows, 2D: one row) -> a pitched slice
void unpack_slice(const FrameGeom& g, const double* dense_slice, double* pitched_slice)
{
    const long long rows = g.dim == 3 ? g.G[1] : 1;
    for (long long r = 0; r < rows; ++r)
        for (long long i = 0; i < g.G[0]; ++i) pitched_slice[g.xoff + i + g.pitch * r] = dense_slice[i + g.G[0] * r];
}

const double kPoison[3] = { -7.0e300, 9.0e300, 5.0e300 };

void init_frame(Frame& f, int dim, const long long* n, const double* dx, const double* psi_dense, const double* rhs_dense)
{
    const int64_t n64[3] = { n[0], n[1], dim == 3 ? n[2] : 1 };
    const double lo[3] = { 0, 0, 0 };
    make_geom(f.g, dim, n64, lo, dx, 1, 1);
    f.deep_slices = dim == 3 ? 2 : 3;
    f.deep = dim == 3 ? 2 * f.g.plane : 3 * f.g.pitch;
    const size_t len = (size_t)(f.deep + f.g.total + f.deep + kTail);
    // poison what is not data: padding columns, the deep planes, the tail
    f.buf[0].assign(len, kPoison[0]); f.buf[1].assign(len, kPoison[1]); f.rhs_buf.assign(len, kPoison[2]);
    unpack(f.g, psi_dense, f.d(0));
    unpack(f.g, psi_dense, f.d(1));       // the library copies the ghost shell into the second buffer
    unpack(f.g, rhs_dense, f.rhs());
}

// cells that are not data (padding columns of every pitched row, the tail pad / guard band, and -- unless a slab neighbour owns
// them -- the deep planes) must still hold the poison init_frame put there: a kernel that stores outside the grown rows of its
// frame is caught here (on the device the canary band of rnmf_frame_check_guards only sees the end of the buffer)
long long clobbered_padding(Frame& f, bool deep_lo_written = false, bool deep_hi_written = false)
{
    long long bad = 0;
    std::vector<double>* bufs[3] = { &f.buf[0], &f.buf[1], &f.rhs_buf };
    const long long rows = f.g.G[1] * f.g.G[2];
    for (int b = 0; b < 3; ++b) {
        const std::vector<double>& v = *bufs[b];
        const double* d = v.data() + f.deep;
        for (long long r = 0; r < rows; ++r)
            for (long long i = 0; i < f.g.pitch; ++i)
                if ((i < f.g.xoff || i >= f.g.xoff + f.g.G[0]) && d[f.g.pitch * r + i] != kPoison[b]) ++bad;
        const bool lo_w = deep_lo_written, hi_w = deep_hi_written;     // psi: the neighbours' stores; rhs: the driver's "entry exchange"
        for (long long q = 0; q < f.deep; ++q) {
            const long long col = q % f.g.pitch;
            const bool pad = col < f.g.xoff || col >= f.g.xoff + f.g.G[0];
            if (v[(size_t)q] != kPoison[b] && (pad || !lo_w)) ++bad;
            if (d[f.g.total + q] != kPoison[b] && (pad || !hi_w)) ++bad;
        }
        for (size_t q = (size_t)(f.deep + f.g.total + f.deep); q < v.size(); ++q)
            if (v[q] != kPoison[b]) ++bad;
    }
    return bad;
}

FaceSet make_faces(int dim, const int* kind, const double* value, const double* dx)
{
    FaceSet fs;
    for (int i = 0; i < 6; ++i) { fs.f[i].mode = 0; fs.f[i].value = 0; fs.f[i].s1 = 0; }
    for (int d = 0; d < dim; ++d)
        for (int side = 0; side < 2; ++side) {
            rnmf_bc_face r;
            r.kind = kind[d * 2 + side]; r._pad = 0; r.value = value[d * 2 + side]; r.prescribed = nullptr; r.n_prescribed = 0;
            fs.f[d * 2 + side] = make_face(r, side, dx[d]);
        }
    return fs;
}

}  // namespace

namespace {
// the checker must be able to fail: a CTA whose second warp waits for a phase nobody completes
void selftest_deadlock_kernel(int dummy)
{
    (void)dummy;
    RNMF_DYNAMIC_SMEM(smem);
    const uint32_t bar = tmau::smem_u32(smem);
    if (threadIdx.x == 0) tmau::mbar_init(bar, 2);
    __syncthreads();
    if (threadIdx.x == 32) tmau::mbar_arrive(bar);             // one of the two arrivals is missing
    if (threadIdx.x >= 32) tmau::mbar_wait(bar, 0);
}
// ... and to pass: both arrivals present, a transaction count completed by an emulated TMA load
void selftest_ok_kernel(CUtensorMap tm)
{
    RNMF_DYNAMIC_SMEM(smem);
    const uint32_t bar = tmau::smem_u32(smem), dst = bar + 128;
    if (threadIdx.x == 0) tmau::mbar_init(bar, 2);
    __syncthreads();
    if (threadIdx.x == 0) { tmau::mbar_arrive_expect_tx(bar, 4 * 8); tmau::tma_load_2d(dst, &tm, -1, 0, bar); }
    if (threadIdx.x == 32) tmau::mbar_arrive(bar);
    tmau::mbar_wait(bar, 0);
    // box (-1..2, row 0) of a 3-wide row {5,6,7}: out-of-bounds column -1 reads as zero
    if (tmau::lds1(dst) != 0.0 || tmau::lds1(dst + 8) != 5.0 || tmau::lds1(dst + 24) != 7.0) g_deadlocks += 100;
}
void selftest_misaligned_kernel(int dummy)
{
    (void)dummy;
    if (threadIdx.x == 0) (void)tmau::lds2(8);             // 16-byte load from an 8-byte aligned address
    if (threadIdx.x == 1) (void)tmau::lds1(256);           // one past the 256 bytes of the CTA
}
}  // namespace

extern "C" {

void emu_set_timeout(int seconds) { emu::g_wait_timeout_s = seconds; }
int emu_faults() { return emu::g_faults.load(); }          // access faults (misaligned / out-of-bounds shared memory, bad TMA boxes) so far

// shared-memory wavefront accounting (emu_cuda.hpp): switch it on, run launches on ONE host thread, read the totals
void emu_trace(int on)
{
    emu::g_trace = on != 0;
    if (on) { emu::g_wf_load = 0; emu::g_wf_store = 0; emu::g_instr_load = 0; emu::g_instr_store = 0; }
}
// out[0..3] = load wavefronts, store wavefronts, load instructions, store instructions (warp level) since emu_trace(1)
void emu_trace_read(long long* out)
{
    out[0] = emu::g_wf_load; out[1] = emu::g_wf_store; out[2] = emu::g_instr_load; out[3] = emu::g_instr_store;
}

// returns (deadlock detected ? 1 : 0) + 2 * (well-formed CTA completed and read the right TMA data ? 1 : 0)
// + 4 * (misaligned / out-of-bounds shared-memory accesses reported ? 1 : 0): 7 = all as expected
int emu_selftest()
{
    g_deadlocks = 0;
    const int saved = emu::g_wait_timeout_s.load();
    emu::g_wait_timeout_s = 1;
    emu_launch_checked(selftest_deadlock_kernel, dim3(1, 1, 1), 64, 256, 0);
    emu::g_wait_timeout_s = saved;
    const int dead = g_deadlocks.load() == 1;
    g_deadlocks = 0;
    alignas(16) static const double row[4] = { 5.0, 6.0, 7.0, -1.0 };       // a 3-wide row with a 16-byte aligned pitch, as TMA demands
    emu::Map m;
    m.base = row; m.rank = 2; m.dims[0] = 3; m.dims[1] = 1; m.dims[2] = 1; m.stride[0] = 1; m.stride[1] = 4; m.stride[2] = 0;
    m.box[0] = 4; m.box[1] = 1; m.box[2] = 1;
    CUtensorMap tm;
    emu::encode(&tm, m);
    emu_launch_checked(selftest_ok_kernel, dim3(2, 1, 1), 64, 512, tm);
    const int ok = g_deadlocks.load() == 0;
    // ... and a misaligned 16-byte shared-memory load / an access past the CTA's shared memory must be reported as a failed launch
    g_deadlocks = 0;
    emu_launch_checked(selftest_misaligned_kernel, dim3(1, 1, 1), 32, 256, 0);
    const int caught = g_deadlocks.load() == 1;
    g_deadlocks = 0;
    return dead + 2 * ok + 4 * caught;
}

// n_pairs double sweeps of one frame on one "device".  face kinds: RNMF_BC_*; coef4 from the oracle.
// Returns 0, or the number of launches that hit a barrier timeout (deadlock).
int emu_double_sweeps(int dim, const long long* n, const double* dx, const int* face_kind, const double* face_value, const double* coef4,
                      const double* psi_dense, const double* rhs_dense, double* out_dense, int variant, int chunk, int n_pairs)
{
    g_deadlocks = 0;
    Frame f;
    init_frame(f, dim, n, dx, psi_dense, rhs_dense);
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.g = f.g;
    a.c = SweepCoef{ coef4[0], coef4[1], coef4[2], coef4[3] };
    a.faces = make_faces(dim, face_kind, face_value, dx);
    a.fuse_fill = 1;
    a.variant = variant;
    a.chunk_override = chunk;
    a.rhs = f.rhs();
    a.k_begin = 0;
    a.k_end = f.g.n[dim - 1];
    a.deep = f.deep_slices;
    for (int it = 0; it < n_pairs; ++it) {
        a.in = f.d(f.cur);
        a.out = f.d(f.cur ^ 1);
        if (dim == 3 && variant == 3) {
            if (launch_jacobi_triple_sweep(a, nullptr) < 0) return -1;          // n_pairs launches of three sweeps each
        } else if (dim == 3) {
            if (launch_jacobi_double_sweep(a, nullptr) < 0) return -1;
        } else if (variant == 30) {
            if (launch_jacobi_double_sweep_2d(a, nullptr) < 0) return -1;
        } else {
            if (launch_jacobi_multi_sweep_2d(a, nullptr) < 0) return -1;        // n_pairs launches of four sweeps each
        }
        f.cur ^= 1;
    }
    pack(f.g, f.d(f.cur), out_dense);
    const long long bad = clobbered_padding(f);
    if (bad) fprintf(stderr, "emu: %lld padding / tail cells were overwritten\n", bad);
    return g_deadlocks.load() + (bad ? 1000 : 0);
}

// the chunk planner of the three-sweep kernel (kernels_sweep_tb3.cu: plan_tb3) for a frame of n0 x n1 x n2 cells whose launch covers
// all planes: out[0] = planes per chunk, out[1] = chunks, out[2] = CTAs
void emu_plan_triple(long long n0, long long n1, long long n2, int nb_lo, int nb_hi, int chunk_override, int* out)
{
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    const int64_t n64[3] = { n0, n1, n2 };
    const double lo[3] = { 0, 0, 0 }, dx[3] = { 1, 1, 1 };
    make_geom(a.g, 3, n64, lo, dx, 1, 1);
    a.k_begin = 0; a.k_end = n2;
    a.nb_lo = nb_lo; a.nb_hi = nb_hi;
    a.chunk_override = chunk_override;
    dim3 grid;
    int chunk, n_chunks;
    plan_tb3(a, grid, chunk, n_chunks);
    out[0] = chunk; out[1] = n_chunks; out[2] = (int)(grid.x * grid.y * grid.z);
}

// the band schedules themselves (common.cuh), for the abstract dependency check in tests/test_skew_schedule.py
void emu_skew_region(long long n2, int n_chunks, int c, long long s, long long* lo, long long* hi) { skew_region(n2, n_chunks, c, s, lo, hi); }
void emu_skew_tail_region(long long n2, int n_chunks, int c, long long s, long long levels, long long* lo, long long* hi)
{
    skew_tail_region(n2, n_chunks, c, s, levels, lo, hi);
}

// The time-skewed start of rnmf_solve_poisson_host (capi.cu: skewed_upload_and_first_sweeps): the frame "arrives" in
// n_chunks chunks of z-planes (planes that have not arrived yet hold poison); after every chunk the first `pairs` double
// sweeps run on the bands skew_region() declares computable, ping-ponging between the two buffers.  At the end every
// plane must hold u_{2*pairs}.  Returns deadlocks, or -1 on a launch failure.
int emu_skewed_double_sweeps(const long long* n, const double* dx, const int* face_kind, const double* face_value, const double* coef4,
                             const double* psi_dense, const double* rhs_dense, double* out_dense, int variant, int chunk, int n_chunks, int pairs)
{
    g_deadlocks = 0;
    Frame f;
    init_frame(f, 3, n, dx, psi_dense, rhs_dense);
    const FrameGeom& g = f.g;
    // nothing has arrived: poison all three buffers (init_frame filled them from the dense frames)
    const std::vector<double> full_psi(f.d(0), f.d(0) + g.total), full_rhs(f.rhs(), f.rhs() + g.total);
    std::fill(f.buf[0].begin(), f.buf[0].end(), -3.0e300);
    std::fill(f.buf[1].begin(), f.buf[1].end(), 4.0e300);
    std::fill(f.rhs_buf.begin(), f.rhs_buf.end(), 6.0e300);
    SweepArgs t;
    memset(&t, 0, sizeof(t));
    t.g = g;
    t.c = SweepCoef{ coef4[0], coef4[1], coef4[2], coef4[3] };
    t.faces = make_faces(3, face_kind, face_value, dx);
    t.fuse_fill = 1;
    t.variant = variant;
    t.chunk_override = chunk;
    t.rhs = f.rhs();
    t.deep = 2;
    const long long n2 = g.n[2];
    for (int ch = 0; ch < n_chunks; ++ch) {
        // "upload" of chunk ch: grown planes [p0, p1) of psi (+ its ghost shell into the partner buffer) and of rhs
        const long long p0 = ch == 0 ? 0 : n2 * ch / n_chunks + 1, p1 = ch == n_chunks - 1 ? g.G[2] : n2 * (ch + 1) / n_chunks + 1;
        for (long long r = g.G[1] * p0; r < g.G[1] * p1; ++r) {
            const long long jj = r % g.G[1], kk = r / g.G[1];
            const bool shell_row = jj < 1 || jj >= g.n[1] + 1 || kk < 1 || kk >= n2 + 1;
            for (long long i = 0; i < g.G[0]; ++i) {
                const long long q = g.xoff + i + g.pitch * r;
                f.d(0)[q] = full_psi[q];
                f.rhs()[q] = full_rhs[q];
                if (shell_row || i < 1 || i >= g.n[0] + 1) f.d(1)[q] = full_psi[q];
            }
        }
        for (long long sp = 1; sp <= pairs; ++sp) {
            long long lo, hi;
            skew_region(n2, n_chunks, ch, sp, &lo, &hi);
            if (hi <= lo) continue;
            t.in = f.d((sp - 1) & 1);
            t.out = f.d(sp & 1);
            t.k_begin = lo; t.k_end = hi;
            if (launch_jacobi_double_sweep(t, nullptr) < 0) return -1;
        }
    }
    pack(g, f.d(pairs & 1), out_dense);
    return g_deadlocks.load();
}

// The band-wise END of rnmf_solve_poisson_host (capi.cu: skewed_last_sweeps_and_download): the last `levels` double sweeps run
// chunk by chunk from the bottom (skew_tail_region); after chunk c's last band the planes below Z(c+1) are "downloaded" into
// out_dense and everything no later band may read (both buffers below plane Z(c+1)-2) is poisoned.
int emu_skewed_tail_double_sweeps(const long long* n, const double* dx, const int* face_kind, const double* face_value, const double* coef4,
                                  const double* psi_dense, const double* rhs_dense, double* out_dense, int variant, int chunk, int n_chunks, int levels)
{
    g_deadlocks = 0;
    Frame f;
    init_frame(f, 3, n, dx, psi_dense, rhs_dense);
    const FrameGeom& g = f.g;
    SweepArgs t;
    memset(&t, 0, sizeof(t));
    t.g = g;
    t.c = SweepCoef{ coef4[0], coef4[1], coef4[2], coef4[3] };
    t.faces = make_faces(3, face_kind, face_value, dx);
    t.fuse_fill = 1;
    t.variant = variant;
    t.chunk_override = chunk;
    t.rhs = f.rhs();
    t.deep = 2;
    const long long n2 = g.n[2], G0 = g.G[0], G1 = g.G[1];
    std::vector<double> dense((size_t)(G0 * G1 * g.G[2]));
    for (int ch = 0; ch < n_chunks; ++ch) {
        for (long long lv = 1; lv <= levels; ++lv) {
            long long lo, hi;
            skew_tail_region(n2, n_chunks, ch, lv, levels, &lo, &hi);
            if (hi <= lo) continue;
            t.in = f.d((lv - 1) & 1);
            t.out = f.d(lv & 1);
            t.k_begin = lo; t.k_end = hi;
            if (launch_jacobi_double_sweep(t, nullptr) < 0) return -1;
        }
        // "download" of chunk ch from the final buffer
        const long long p0 = ch == 0 ? 0 : n2 * ch / n_chunks + 1, p1 = ch == n_chunks - 1 ? g.G[2] : n2 * (ch + 1) / n_chunks + 1;
        pack(g, f.d(levels & 1), dense.data());
        memcpy(out_dense + G0 * G1 * p0, dense.data() + G0 * G1 * p0, (size_t)(G0 * G1 * (p1 - p0)) * sizeof(double));
        // nothing below valid plane Z(ch+1)-2 may be read again
        const long long dead = n2 * (ch + 1) / n_chunks - 2;           // valid planes < dead, i.e. grown planes <= dead
        if (ch < n_chunks - 1)
            for (long long kk = 0; kk <= dead && kk < g.G[2]; ++kk)
                for (long long q = g.plane * kk; q < g.plane * (kk + 1); ++q) { f.d(0)[q] = -8.0e300; f.d(1)[q] = 8.5e300; }
    }
    return g_deadlocks.load();
}

// The same over slabs of the slowest direction (3D: z-slabs, two sweeps per launch; 2D: y-slabs, four sweeps per launch):
// `nranks` host threads, each launching its slab's kernel.  Deep halo: a rank starts with its neighbours' `depth` boundary slices
// in its ghost + deep slices (what the entry exchange of rnmf_jacobi_sweeps delivers; rhs: depth-1 slices) and stores u_depth of
// its own `depth` boundary slices into the neighbours' ghost + deep slices, ordered by the full-step flags -- the argument
// wiring mirrors rnmf_jacobi_sweeps (capi.cu).  Global dense frames in / out (ng = 1).
int emu_sweeps_slabs(int dim, const long long* n, const double* dx, const int* face_kind, const double* face_value, const double* coef4,
                     const double* psi_dense, const double* rhs_dense, double* out_dense, int variant, int chunk, int n_launches, int nranks)
{
    g_deadlocks = 0;
    const int D = dim - 1;
    const bool triple = dim == 3 && variant == 3;
    const int depth = dim == 3 ? (triple ? 3 : 2) : 4;
    const long long G0 = n[0] + 2, slice_dense = dim == 3 ? G0 * (n[1] + 2) : G0;
    std::vector<Frame> fr(nranks);
    std::vector<long long> start(nranks), count(nranks);
    struct Flags { unsigned long long w[16]; };
    std::vector<Flags> flags(nranks);
    for (auto& fl : flags) memset(&fl, 0, sizeof(fl));
    for (int r = 0; r < nranks; ++r) {
        const long long base = n[D] / nranks;                     // Domain::new_rectangular: remainder to the last block
        start[r] = base * r;
        count[r] = r == nranks - 1 ? n[D] - start[r] : base;
        long long nl[3] = { n[0], n[1], dim == 3 ? n[2] : 1 };
        nl[D] = count[r];
        // the slab's dense frame = global grown slices start .. start+count+1 (its ghost slices hold the neighbours' slices)
        Frame& f = fr[r];
        init_frame(f, dim, nl, dx, psi_dense + slice_dense * start[r], rhs_dense + slice_dense * start[r]);
        const long long slice = dim == 3 ? f.g.plane : f.g.pitch;
        // deep slices: the neighbours' slices further from the face (psi: depth-1 of them; rhs: depth-2)
        for (int e = 1; e < depth; ++e) {
            if (r > 0) {
                for (int b = 0; b < 2; ++b) unpack_slice(f.g, psi_dense + slice_dense * (start[r] - e), f.d(b) - slice * e);
                if (e < depth - 1) unpack_slice(f.g, rhs_dense + slice_dense * (start[r] - e), f.rhs() - slice * e);
            }
            if (r < nranks - 1) {
                for (int b = 0; b < 2; ++b) unpack_slice(f.g, psi_dense + slice_dense * (start[r] + count[r] + 1 + e), f.d(b) + slice * (count[r] + 1 + e));
                if (e < depth - 1) unpack_slice(f.g, rhs_dense + slice_dense * (start[r] + count[r] + 1 + e), f.rhs() + slice * (count[r] + 1 + e));
            }
        }
    }
    std::vector<std::thread> ranks;
    for (int r = 0; r < nranks; ++r)
        ranks.emplace_back([&, r] {
            Frame& f = fr[r];
            const bool has_lo = r > 0, has_hi = r < nranks - 1;
            const long long m = count[r];
            const long long slice = dim == 3 ? f.g.plane : f.g.pitch;
            int kind[6];
            for (int i = 0; i < 6; ++i) kind[i] = face_kind[i];
            if (has_lo) kind[2 * D] = RNMF_BC_HALO;
            if (has_hi) kind[2 * D + 1] = RNMF_BC_HALO;
            unsigned long long epoch = 0, done[2] = { 0, 0 };
            for (int it = 0; it < n_launches; ++it) {
                SweepArgs a;
                memset(&a, 0, sizeof(a));
                a.g = f.g;
                a.c = SweepCoef{ coef4[0], coef4[1], coef4[2], coef4[3] };
                a.faces = make_faces(dim, kind, face_value, dx);
                a.fuse_fill = 1;
                a.chunk_override = chunk;
                a.rhs = f.rhs();
                a.in = f.d(f.cur);
                a.out = f.d(f.cur ^ 1);
                a.k_begin = 0;
                a.k_end = m;
                a.deep = f.deep_slices;
                a.nb_lo = has_lo; a.nb_hi = has_hi;
                const int out_phys = f.cur ^ 1;
                if (has_lo) {
                    Frame& lo = fr[r - 1];
                    a.peer_lo_plane = lo.d(out_phys) + slice * (count[r - 1] + 1);                    // its hi ghost slice
                    a.ps.wait_lo = &flags[r].w[0];
                    a.ps.sig_lo = &flags[r - 1].w[1];
                }
                if (has_hi) {
                    Frame& hi = fr[r + 1];
                    a.peer_hi_plane = hi.d(out_phys);                                                  // its lo ghost slice
                    a.ps.wait_hi = &flags[r].w[1];
                    a.ps.sig_hi = &flags[r + 1].w[0];
                }
                const unsigned long long tiles = (unsigned long long)(triple ? triple_sweep_boundary_ctas(a) : dim == 3 ? double_sweep_boundary_ctas(a) : multi_sweep_2d_boundary_ctas(a));
                a.ps.enabled = 1;
                a.ps.done = &flags[r].w[3];
                if (has_lo) done[0] += tiles;
                if (has_hi) done[1] += tiles;
                a.ps.target_lo = done[0]; a.ps.target_hi = done[1];
                a.ps.epoch = epoch;
                a.ps.status = &flags[r].w[2];
                epoch += 1;
                if (triple) launch_jacobi_triple_sweep(a, nullptr);
                else if (dim == 3) launch_jacobi_double_sweep(a, nullptr);
                else launch_jacobi_multi_sweep_2d(a, nullptr);
                f.cur ^= 1;
            }
            // the closing wait of rnmf_jacobi_sweeps: the neighbours' last stores into my ghost slices have landed
            if (has_lo) peer_spin_wait(&flags[r].w[0], epoch, &flags[r].w[2]);
            if (has_hi) peer_spin_wait(&flags[r].w[1], epoch, &flags[r].w[2]);
        });
    for (auto& th : ranks) th.join();
    int bad = g_deadlocks.load();
    for (int r = 0; r < nranks; ++r) {
        if (flags[r].w[2]) ++bad;                                   // a handshake wait timed out
        const long long clob = clobbered_padding(fr[r], r > 0, r < nranks - 1);
        if (clob) { fprintf(stderr, "emu: rank %d: %lld padding / tail / deep cells were overwritten\n", r, clob); bad += 1000; }
        // valid slices of the slab -> the global dense frame; ghost slices of the outer ranks too
        std::vector<double> dense((size_t)(slice_dense * (count[r] + 2)));
        pack(fr[r].g, fr[r].d(fr[r].cur), dense.data());
        const long long p0 = r == 0 ? 0 : 1, p1 = r == nranks - 1 ? count[r] + 2 : count[r] + 1;
        memcpy(out_dense + slice_dense * (start[r] + p0), dense.data() + slice_dense * p0, (size_t)(slice_dense * (p1 - p0)) * sizeof(double));
    }
    return bad;
}

// The time-skewed start ACROSS SLABS (capi.cu: skewed_upload_and_first_sweeps with neighbours): every rank's slab "arrives" in
// n_chunks chunks; double sweep s runs on the bands skew_region() makes computable, clipped to [2s, m - 2s) next to a
// neighbour (no halo needed); then all ranks meet ("entry exchange" of the level-0 halos: ghost + deep plane, rhs ghost plane)
// and the wedges catch up level by level, each level ONE launch over both wedges (a hole in the plane range) with the slab
// handshake.  Unarrived planes hold poison.  At the end every plane must hold u_{2*pairs}.
int emu_skewed_double_sweeps_slabs(const long long* n, const double* dx, const int* face_kind, const double* face_value, const double* coef4,
                                   const double* psi_dense, const double* rhs_dense, double* out_dense, int chunk, int n_chunks, int pairs, int nranks)
{
    g_deadlocks = 0;
    const long long G0 = n[0] + 2, plane_dense = G0 * (n[1] + 2);
    std::vector<Frame> fr(nranks);
    std::vector<std::vector<double>> full_psi(nranks), full_rhs(nranks);
    std::vector<long long> start(nranks), count(nranks);
    struct Flags { unsigned long long w[16]; };
    std::vector<Flags> flags(nranks);
    for (auto& fl : flags) memset(&fl, 0, sizeof(fl));
    for (int r = 0; r < nranks; ++r) {
        const long long base = n[2] / nranks;
        start[r] = base * r;
        count[r] = r == nranks - 1 ? n[2] - start[r] : base;
        const long long nl[3] = { n[0], n[1], count[r] };
        Frame& f = fr[r];
        // the host frame of a rank: its slab with whatever ghost planes the caller has (here: the neighbours' planes, which the
        // schedule must NOT rely on before the exchange -- they are poisoned below together with everything else)
        init_frame(f, 3, nl, dx, psi_dense + plane_dense * start[r], rhs_dense + plane_dense * start[r]);
        full_psi[r].assign(f.d(0), f.d(0) + f.g.total);
        full_rhs[r].assign(f.rhs(), f.rhs() + f.g.total);
        std::fill(f.buf[0].begin(), f.buf[0].end(), -3.0e300);
        std::fill(f.buf[1].begin(), f.buf[1].end(), 4.0e300);
        std::fill(f.rhs_buf.begin(), f.rhs_buf.end(), 6.0e300);
    }
    std::atomic<int> arrived{0};
    std::vector<std::thread> ranks;
    for (int r = 0; r < nranks; ++r)
        ranks.emplace_back([&, r] {
            Frame& f = fr[r];
            const FrameGeom& g = f.g;
            const bool has_lo = r > 0, has_hi = r < nranks - 1;
            const long long n2 = count[r];
            int kind[6];
            for (int i = 0; i < 6; ++i) kind[i] = face_kind[i];
            if (has_lo) kind[4] = RNMF_BC_HALO;
            if (has_hi) kind[5] = RNMF_BC_HALO;
            SweepArgs t;
            memset(&t, 0, sizeof(t));
            t.g = g;
            t.c = SweepCoef{ coef4[0], coef4[1], coef4[2], coef4[3] };
            t.faces = make_faces(3, kind, face_value, dx);
            t.fuse_fill = 1;
            t.chunk_override = chunk;
            t.rhs = f.rhs();
            t.deep = 2;
            // ---- phase 1: chunks arrive, bands run (clipped next to the neighbours) ----
            for (int ch = 0; ch < n_chunks; ++ch) {
                const long long p0 = ch == 0 ? 0 : n2 * ch / n_chunks + 1, p1 = ch == n_chunks - 1 ? g.G[2] : n2 * (ch + 1) / n_chunks + 1;
                for (long long rr = g.G[1] * p0; rr < g.G[1] * p1; ++rr) {
                    const long long jj = rr % g.G[1], kk = rr / g.G[1];
                    const bool shell_row = jj < 1 || jj >= g.n[1] + 1 || kk < 1 || kk >= n2 + 1;
                    // a ghost plane towards a neighbour is not data the caller can provide: leave it poisoned until the exchange
                    if ((kk == 0 && has_lo) || (kk == n2 + 1 && has_hi)) continue;
                    for (long long i = 0; i < g.G[0]; ++i) {
                        const long long q = g.xoff + i + g.pitch * rr;
                        f.d(0)[q] = full_psi[r][q];
                        f.rhs()[q] = full_rhs[r][q];
                        if (shell_row || i < 1 || i >= g.n[0] + 1) f.d(1)[q] = full_psi[r][q];
                    }
                }
                for (long long sp = 1; sp <= pairs; ++sp) {
                    long long lo, hi;
                    skew_region(n2, n_chunks, ch, sp, &lo, &hi);
                    if (has_lo && lo < 2 * sp) lo = 2 * sp;
                    if (has_hi && hi > n2 - 2 * sp) hi = n2 - 2 * sp;
                    if (hi <= lo) continue;
                    t.in = f.d((sp - 1) & 1);
                    t.out = f.d(sp & 1);
                    t.k_begin = lo; t.k_end = hi;
                    launch_jacobi_double_sweep(t, nullptr);
                }
            }
            // ---- all ranks have everything: the level-0 halo exchange (ghost + deep plane of psi, ghost plane of rhs) ----
            ++arrived;
            while (arrived.load() < nranks) std::this_thread::yield();
            if (has_lo) {
                Frame& lo = fr[r - 1];
                const long long m_lo = count[r - 1];
                memcpy(f.d(0) - g.plane, lo.d(0) + g.plane * (m_lo - 1), (size_t)(2 * g.plane) * sizeof(double));      // its planes m-2, m-1 -> my deep, ghost
                memcpy(f.rhs(), lo.rhs() + g.plane * m_lo, (size_t)g.plane * sizeof(double));
            }
            if (has_hi) {
                Frame& hi = fr[r + 1];
                memcpy(f.d(0) + g.plane * (n2 + 1), hi.d(0) + g.plane, (size_t)(2 * g.plane) * sizeof(double));         // its planes 0, 1 -> my ghost, deep
                memcpy(f.rhs() + g.plane * (n2 + 1), hi.rhs() + g.plane, (size_t)g.plane * sizeof(double));
            }
            ++arrived;
            while (arrived.load() < 2 * nranks) std::this_thread::yield();
            // ---- phase 2: the wedges catch up ----
            unsigned long long epoch = 0, done[2] = { 0, 0 };
            for (long long sp = 1; sp <= pairs; ++sp) {
                SweepArgs w = t;
                w.in = f.d((sp - 1) & 1);
                w.out = f.d(sp & 1);
                w.k_begin = 0; w.k_end = n2;
                w.hole_begin = has_lo ? 2 * sp : 0;
                w.hole_end = has_hi ? n2 - 2 * sp : n2;
                w.nb_lo = has_lo; w.nb_hi = has_hi;
                const int out_phys = (int)(sp & 1);
                if (has_lo) {
                    Frame& lo = fr[r - 1];
                    w.peer_lo_plane = lo.d(out_phys) + g.plane * (count[r - 1] + 1);
                    w.ps.wait_lo = &flags[r].w[0];
                    w.ps.sig_lo = &flags[r - 1].w[1];
                }
                if (has_hi) {
                    Frame& hi = fr[r + 1];
                    w.peer_hi_plane = hi.d(out_phys);
                    w.ps.wait_hi = &flags[r].w[1];
                    w.ps.sig_hi = &flags[r + 1].w[0];
                }
                const unsigned long long tiles = (unsigned long long)double_sweep_boundary_ctas(w);
                w.ps.enabled = 1;
                w.ps.done = &flags[r].w[3];
                if (has_lo) done[0] += tiles;
                if (has_hi) done[1] += tiles;
                w.ps.target_lo = done[0]; w.ps.target_hi = done[1];
                w.ps.epoch = epoch;
                w.ps.status = &flags[r].w[2];
                epoch += 1;
                launch_jacobi_double_sweep(w, nullptr);
            }
            if (has_lo) peer_spin_wait(&flags[r].w[0], epoch, &flags[r].w[2]);
            if (has_hi) peer_spin_wait(&flags[r].w[1], epoch, &flags[r].w[2]);
        });
    for (auto& th : ranks) th.join();
    int bad = g_deadlocks.load();
    for (int r = 0; r < nranks; ++r) {
        if (flags[r].w[2]) ++bad;
        std::vector<double> dense((size_t)(plane_dense * (count[r] + 2)));
        pack(fr[r].g, fr[r].d(pairs & 1), dense.data());
        const long long p0 = r == 0 ? 0 : 1, p1 = r == nranks - 1 ? count[r] + 2 : count[r] + 1;
        memcpy(out_dense + plane_dense * (start[r] + p0), dense.data() + plane_dense * p0, (size_t)(plane_dense * (p1 - p0)) * sizeof(double));
    }
    return bad;
}

}  // extern "C"
