// capi.cu -- the C ABI of librnmf_b200 (include/rnmf_b200.h): contexts, device-resident data
// frames, and the host-side sequencing of the sweep / ghost-fill / norm kernels, including the
// slab halo exchange (NCCL, loaded with dlopen so single-GPU users never need it).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local unsigned g_err_seq = 0;     // bumped by every rnmf_set_error (lets callers add context only when needed)

void rnmf_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    ++g_err_seq;
}

extern "C" const char* rnmf_last_error(void) { return g_err; }
extern "C" int32_t rnmf_abi_version(void) { return RNMF_ABI_VERSION; }
extern "C" const char* rnmf_build_info(void)
{
    return "librnmf_b200 abi=1 arch=sm_100a fmad=false kernels="
           "jacobi3d_tma{128x8x4},jacobi3d_tb3{128x12 quad, 3 sweeps},jacobi3d_tb2{128x12 quad},jacobi2d_tma{512x8},jacobi2d_tb2{512},jacobi2d_tbk{4x512 di},"
           "jacobi_resident,fill_bc_faces,fill_prescribed,gs_wavefront_2d,varcoef_2d,abs_sum_valid,ew{scale,add,mul,axpby}";
}

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (types from the system header; the library is whichever libnccl.so.2 the
// process already has -- torch's bundled one under Python -- or the system one)
// ------------------------------------------------------------------------------------------
#include <nccl.h>

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.handle) return RNMF_OK;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        rnmf_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return RNMF_ERR_NCCL;
    }
#define LOAD(field, sym)                                                          \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                                     \
    if (!g_nccl.field) { rnmf_set_error("libnccl lacks %s", sym); return RNMF_ERR_NCCL; }
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.handle = h;
    return RNMF_OK;
}
}  // namespace

// a kernel launcher returns the number of kernels it launched, or < 0 (message already set)
#define RNMF_LAUNCH(ctx, expr)                                                               \
    do {                                                                                     \
        const unsigned _seq = g_err_seq;                                                     \
        int _k = (expr);                                                                     \
        if (_k < 0) {                                                                        \
            if (_seq == g_err_seq) rnmf_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(cudaGetLastError())); \
            return RNMF_ERR_CUDA;                                                            \
        }                                                                                    \
        (ctx)->launches += _k;                                                               \
    } while (0)

#define RNMF_NCCL_TRY(expr)                                                                  \
    do {                                                                                     \
        ncclResult_t _r = (expr);                                                            \
        if (_r != ncclSuccess) {                                                             \
            rnmf_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return RNMF_ERR_NCCL;                                                            \
        }                                                                                    \
    } while (0)

// ------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------
struct rnmf_frame;

struct rnmf_ctx {
    int device = 0;
    int rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;        // compute
    cudaStream_t stream_halo = nullptr;   // halo exchange (high priority)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_boundary = nullptr, ev_halo = nullptr;
    long long launches = 0;
    double* d_scalar = nullptr;           // device scalars (norms)
    double* h_scalar = nullptr;           // pinned
    double* d_partials = nullptr;
    long long partial_cap = 0;
    int sweep_variant = 0;
    long long double_sweeps = 0;                  // launches of the two-sweeps-per-pass kernel (rnmf_double_sweep_launches)
    long long multi_sweeps = 0;                   // launches of the K-sweeps-per-pass 2D kernel (opt-in variants 32..40)
    long long multi_sweep_sweeps = 0;             // sweeps those launches covered
    int sweep_chunk = 0;
    int overlap_halo = 1;
    ncclComm_t comm = nullptr;
    // fused peer-memory halo: [0] written by the lo neighbour, [1] by the hi neighbour, [2] status
    unsigned long long* d_flags = nullptr;
    unsigned long long* peer_flag_lo = nullptr;   // lo neighbour's flags[1] (peer-mapped)
    unsigned long long* peer_flag_hi = nullptr;   // hi neighbour's flags[0] (peer-mapped)
    void* peer_flag_base_lo = nullptr;
    void* peer_flag_base_hi = nullptr;
    unsigned long long epoch = 0;                 // sweeps completed in peer mode (same on all ranks)
    unsigned long long done_target[2] = { 0, 0 }; // cumulative boundary-CTA counts (in-kernel handshake), lo / hi
    int peer_inkernel = 1;                        // option "peer_inkernel": handshake inside the sweep kernel
    int peer_halo = 1;                            // option "peer_halo": use the fused path when attached
    int peer_used = 0;                            // a fused-halo sweep has been launched (check_peer_status has something to check)
    double* d_staging = nullptr;                  // dense staging for large uploads (1-D PCIe copy + device repack)
    long long staging_cap = 0;                    // doubles
    int resident = 1;                             // option "resident": multi-sweep cooperative kernel for L2-resident frames
    long long resident_max_bytes = 48LL << 20;    // psi + psi' + rhs must fit comfortably in the 126 MB L2
    rnmf_frame* hp_psi = nullptr;         // cached frames of rnmf_solve_poisson_host
    rnmf_frame* hp_rhs = nullptr;
    int e2e_skew = 2;                     // option "e2e_skew": 1 = time-skewed start of rnmf_solve_poisson_host, 2 = and end (default; measured 294 -> 347 GLUP/s end to end at 768^3), 0 = plain
    int e2e_skew_chunks = 8, e2e_skew_pairs = 56; // upload chunks, double sweeps run band-wise during the upload
    int e2e_skew_tail = 32;                       // e2e_skew >= 2: last launches run band-wise so the download overlaps them
    std::vector<cudaEvent_t> ev_chunks;   // one per upload chunk
    long long skew_launches = 0;          // double-sweep launches issued by the skewed start (a test reads it back)
};

struct rnmf_frame {
    rnmf_ctx* ctx = nullptr;
    FrameGeom g;
    double* d = nullptr;                  // current buffer
    double* d2 = nullptr;                 // ping-pong partner (lazily allocated)
    int parity = 0;                       // physical id of `d` (0 = the buffer the frame was created with)
    int peer_attached = 0;
    double* peer_lo[2] = { nullptr, nullptr };   // lo neighbour's physical buffers 0/1 (peer-mapped bases)
    double* peer_hi[2] = { nullptr, nullptr };
    int distributed = 0;
    int bc_set = 0;
    int has_reflect = 0;
    std::vector<rnmf_bc_face> raw;        // as given (global BCs), prescribed pointers cleared
    std::vector<FaceSet> faces;           // per component, HALO applied
    std::vector<double*> prescribed_dev;  // per raw face (nullptr if not Prescribed)
    std::vector<long long> prescribed_n;
};

namespace {

constexpr long long kTailPad = 64;   // doubles; lets edge threads read a little past the last row
constexpr long long kGuard = 256;    // doubles after the tail pad holding a canary pattern (rnmf_frame_check_guards)
constexpr unsigned long long kCanary = 0x7FF8C0DEC0DE0000ULL;   // a quiet-NaN payload

int ensure_partials(rnmf_ctx* c, long long n)
{
    if (n <= c->partial_cap) return RNMF_OK;
    if (c->d_partials) RNMF_CUDA_TRY(cudaFree(c->d_partials));
    c->d_partials = nullptr;
    c->partial_cap = 0;
    long long cap = n < 4096 ? 4096 : n;
    RNMF_CUDA_TRY(cudaMalloc(&c->d_partials, (size_t)cap * sizeof(double)));
    c->partial_cap = cap;
    return RNMF_OK;
}

int bind(rnmf_ctx* c)
{
    RNMF_CUDA_TRY(cudaSetDevice(c->device));
    return RNMF_OK;
}

// Fused peer-memory halo: a handshake wait that ran out of patience leaves the awaited epoch in d_flags[2] and lets the
// kernel go on (never a hang), which means stale halo slices.  Every synchronising entry point (rnmf_sync, the downloads,
// norm-carrying sweep calls) ends with this check, so that no result computed after a time-out is handed to the caller as
// good.  The word is cleared once reported.  The stream must be idle.
int check_peer_status(rnmf_ctx* c)
{
    if (c->nranks == 1 || !c->d_flags || !c->peer_used) return RNMF_OK;
    unsigned long long st = 0;
    RNMF_CUDA_TRY(cudaMemcpy(&st, c->d_flags + 2, 8, cudaMemcpyDeviceToHost));
    if (!st) return RNMF_OK;
    RNMF_CUDA_TRY(cudaMemset(c->d_flags + 2, 0, 8));
    rnmf_set_error("peer halo handshake timed out waiting for epoch %llu: halo slices are stale, results since then are invalid", st);
    return RNMF_ERR_NCCL;
}

// DEEP HALO: 3D one-component ng=1 frame buffers carry one extra z-plane in front of the z-lo ghost plane and one behind
// the z-hi ghost plane (kernels_sweep_tb2.cu): with the neighbour slab's two boundary planes in ghost + deep plane, two sweeps
// per launch need no data that arrives during the launch (K sweeps per launch: K slices, i.e. K - 1 deep ones).  A frame pointer (`d`, `d2`) addresses the z-lo ghost plane as
// before; the allocation starts deep_elems() doubles earlier and the tail pad / canary band follow the rear deep plane.
// 2D frames: three extra ROWS in front of the y-lo ghost row and behind the y-hi ghost row (four sweeps per launch across y-slabs).
int deep_slices(const FrameGeom& g) { return g.ng == 1 && g.ncomp == 1 ? (g.dim == 3 ? 2 : 3) : 0; }
long long deep_elems(const FrameGeom& g) { return deep_slices(g) * (g.dim == 3 ? g.plane : g.pitch); }
long long tail_off(const FrameGeom& g) { return g.total + deep_elems(g); }       // first double behind the (rear deep plane of the) frame

int write_guard(double* buf, const FrameGeom& g, cudaStream_t s)
{
    static unsigned long long pattern[kGuard];
    static bool init = false;
    if (!init) { for (long long i = 0; i < kGuard; ++i) pattern[i] = kCanary + (unsigned long long)i; init = true; }
    RNMF_CUDA_TRY(cudaMemcpyAsync(buf + tail_off(g) + kTailPad, pattern, sizeof(pattern), cudaMemcpyHostToDevice, s));
    return RNMF_OK;
}

int alloc_buffer(const FrameGeom& g, double** out, cudaStream_t s)
{
    const long long deep = deep_elems(g);
    const size_t bytes = (size_t)(deep + g.total + deep + kTailPad + kGuard) * sizeof(double);
    double* raw = nullptr;
    RNMF_CUDA_TRY(cudaMalloc(&raw, bytes));
    RNMF_CUDA_TRY(cudaMemsetAsync(raw, 0, bytes, s));   // new_from zero-initialises (dataframe.rs:155)
    *out = raw + deep;
    return write_guard(*out, g, s);
}

void free_buffer(const FrameGeom& g, double* buf) { if (buf) cudaFree(buf - deep_elems(g)); }

bool same_shape(const FrameGeom& a, const FrameGeom& b)
{
    return a.dim == b.dim && a.ng == b.ng && a.ncomp == b.ncomp && a.n[0] == b.n[0] && a.n[1] == b.n[1] &&
           a.n[2] == b.n[2];
}

int slow_dim(const FrameGeom& g) { return g.dim - 1; }

}  // namespace

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
static int ctx_common_init(rnmf_ctx* c)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        rnmf_set_error("no CUDA device available (%s): librnmf_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return RNMF_ERR_CUDA;
    }
    RNMF_REQUIRE(c->device >= 0 && c->device < count, "device %d out of range (count %d)", c->device, count);
    RNMF_CUDA_TRY(cudaSetDevice(c->device));
    int lo_p = 0, hi_p = 0;
    RNMF_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    RNMF_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, lo_p));
    RNMF_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream_halo, cudaStreamNonBlocking, hi_p));
    RNMF_CUDA_TRY(cudaEventCreate(&c->ev_t0));
    RNMF_CUDA_TRY(cudaEventCreate(&c->ev_t1));
    RNMF_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
    RNMF_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    RNMF_CUDA_TRY(cudaMalloc(&c->d_scalar, 8 * sizeof(double)));
    RNMF_CUDA_TRY(cudaMallocHost(&c->h_scalar, 8 * sizeof(double)));
    // words: [0],[1] full-step flags (written by the lo / hi neighbour), [2] handshake status, [3],[4] boundary-CTA
    // counters, [7] attach tag, [8],[9] half-step flags of the double-sweep kernel, [10],[11] its B-chunk counters
    RNMF_CUDA_TRY(cudaMalloc(&c->d_flags, 128));
    RNMF_CUDA_TRY(cudaMemset(c->d_flags, 0, 128));
    const char* v = getenv("RNMF_SWEEP_VARIANT");
    if (v) c->sweep_variant = atoi(v);
    const char* o = getenv("RNMF_HALO_OVERLAP");
    if (o) c->overlap_halo = atoi(o);
    const char* k = getenv("RNMF_E2E_SKEW");
    if (k) c->e2e_skew = atoi(k);
    return RNMF_OK;
}

extern "C" int32_t rnmf_ctx_create(int32_t device, rnmf_ctx** out)
{
    RNMF_REQUIRE(out, "out is null");
    rnmf_ctx* c = new rnmf_ctx();
    c->device = device;
    int rc = ctx_common_init(c);
    if (rc) { rnmf_ctx_destroy(c); return rc; }      // releases whatever the partial initialisation created
    *out = c;
    return RNMF_OK;
}

extern "C" int32_t rnmf_nccl_unique_id(void* id128)
{
    RNMF_REQUIRE(id128, "id128 is null");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    RNMF_NCCL_TRY(g_nccl.GetUniqueId((ncclUniqueId*)id128));
    return RNMF_OK;
}

extern "C" int32_t rnmf_ctx_create_dist(int32_t device, int32_t rank, int32_t nranks, const void* nccl_id, rnmf_ctx** out)
{
    RNMF_REQUIRE(out && nccl_id, "null argument");
    RNMF_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / nranks %d", rank, nranks);
    rnmf_ctx* c = new rnmf_ctx();
    c->device = device; c->rank = rank; c->nranks = nranks;
    int rc = ctx_common_init(c);
    if (rc) { rnmf_ctx_destroy(c); return rc; }
    if (nranks > 1) {
        rc = nccl_load();
        if (rc) { rnmf_ctx_destroy(c); return rc; }
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            rnmf_set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r));
            c->comm = nullptr;
            rnmf_ctx_destroy(c);
            return RNMF_ERR_NCCL;
        }
    }
    *out = c;
    return RNMF_OK;
}

extern "C" int32_t rnmf_ctx_destroy(rnmf_ctx* c)
{
    if (!c) return RNMF_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->hp_psi) rnmf_frame_destroy(c->hp_psi);
    if (c->hp_rhs) rnmf_frame_destroy(c->hp_rhs);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->peer_flag_base_lo) cudaIpcCloseMemHandle(c->peer_flag_base_lo);
    if (c->peer_flag_base_hi) cudaIpcCloseMemHandle(c->peer_flag_base_hi);
    if (c->d_flags) cudaFree(c->d_flags);
    if (c->d_partials) cudaFree(c->d_partials);
    if (c->d_staging) cudaFree(c->d_staging);
    if (c->d_scalar) cudaFree(c->d_scalar);
    if (c->h_scalar) cudaFreeHost(c->h_scalar);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->ev_boundary) cudaEventDestroy(c->ev_boundary);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    for (cudaEvent_t e : c->ev_chunks) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream_halo) cudaStreamDestroy(c->stream_halo);
    delete c;
    return RNMF_OK;
}

extern "C" int32_t rnmf_sync(rnmf_ctx* c)
{
    RNMF_REQUIRE(c, "ctx is null");
    if (bind(c)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream_halo));
    RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_peer_status(c);
}

extern "C" int32_t rnmf_ctx_stream(rnmf_ctx* c, void** s)
{
    RNMF_REQUIRE(c && s, "null argument");
    *s = (void*)c->stream;
    return RNMF_OK;
}

extern "C" int32_t rnmf_ctx_rank(rnmf_ctx* c, int32_t* rank, int32_t* nranks)
{
    RNMF_REQUIRE(c, "ctx is null");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return RNMF_OK;
}

extern "C" int64_t rnmf_kernel_launches(rnmf_ctx* c) { return c ? c->launches : 0; }
extern "C" int64_t rnmf_double_sweep_launches(rnmf_ctx* c) { return c ? c->double_sweeps : 0; }
extern "C" int32_t rnmf_multi_sweep_stats(rnmf_ctx* c, int64_t* launches, int64_t* sweeps)
{
    RNMF_REQUIRE(c, "ctx is null");
    if (launches) *launches = c->double_sweeps + c->multi_sweeps;
    if (sweeps) *sweeps = 2 * c->double_sweeps + c->multi_sweep_sweeps;
    return RNMF_OK;
}

extern "C" int32_t rnmf_timer_begin(rnmf_ctx* c)
{
    RNMF_REQUIRE(c, "ctx is null");
    if (bind(c)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaEventRecord(c->ev_t0, c->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_timer_end(rnmf_ctx* c, double* ms)
{
    RNMF_REQUIRE(c && ms, "null argument");
    if (bind(c)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaEventRecord(c->ev_t1, c->stream));
    RNMF_CUDA_TRY(cudaEventSynchronize(c->ev_t1));
    float f = 0.f;
    RNMF_CUDA_TRY(cudaEventElapsedTime(&f, c->ev_t0, c->ev_t1));
    *ms = (double)f;
    return RNMF_OK;
}

extern "C" int32_t rnmf_set_option(rnmf_ctx* c, const char* key, const char* value)
{
    RNMF_REQUIRE(c && key && value, "null argument");
    if (!strcmp(key, "sweep_variant")) { c->sweep_variant = !strcmp(value, "auto") ? 0 : atoi(value); return RNMF_OK; }
    if (!strcmp(key, "sweep_chunk")) { c->sweep_chunk = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "peer_halo")) { c->peer_halo = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "peer_inkernel")) { c->peer_inkernel = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "resident")) { c->resident = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "halo_overlap")) { c->overlap_halo = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "e2e_skew")) { c->e2e_skew = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "e2e_skew_chunks")) { c->e2e_skew_chunks = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "e2e_skew_pairs")) { c->e2e_skew_pairs = atoi(value); return RNMF_OK; }
    if (!strcmp(key, "e2e_skew_tail")) { c->e2e_skew_tail = atoi(value); return RNMF_OK; }
    rnmf_set_error("unknown option '%s'", key);
    return RNMF_ERR_INVALID;
}

// Domain::new_rectangular (cartesian2d/mod.rs:64-73,109-116): floor(n/P) cells per block, the
// last block takes the remainder.
extern "C" int32_t rnmf_slab_partition(int64_t n, int32_t nranks, int32_t rank, int64_t* start, int64_t* count)
{
    RNMF_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && n >= 0, "bad partition arguments");
    const int64_t per = n / nranks;
    const int64_t rem = n % nranks;
    if (start) *start = per * rank;
    if (count) *count = per + (rank == nranks - 1 ? rem : 0);
    return RNMF_OK;
}

// ------------------------------------------------------------------------------------------
// frames
// ------------------------------------------------------------------------------------------
static int frame_create_impl(rnmf_ctx* c, int dim, const int64_t* n_local, const int64_t* n_global,
                             const int64_t* offset, const double* lo, const double* dx, int ng, int ncomp,
                             int distributed, rnmf_frame** out)
{
    RNMF_REQUIRE(c && n_local && lo && dx && out, "null argument");
    RNMF_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3 (got %d)", dim);
    RNMF_REQUIRE(ng >= 1 && ng <= RNMF_ROW_ALIGN, "n_ghost must be in [1,16] (got %d)", ng);
    RNMF_REQUIRE(ncomp >= 1, "n_comp must be >= 1");
    for (int d = 0; d < dim; ++d) RNMF_REQUIRE(n_local[d] >= 1, "n[%d] must be >= 1", d);
    if (bind(c)) return RNMF_ERR_CUDA;
    rnmf_frame* f = new rnmf_frame();
    f->ctx = c;
    make_geom(f->g, dim, n_local, lo, dx, ng, ncomp);
    for (int d = 0; d < dim; ++d) {
        f->g.nglob[d] = n_global[d];
        f->g.off[d] = offset[d];
    }
    f->distributed = distributed;
    int rc = alloc_buffer(f->g, &f->d, c->stream);
    if (rc) { delete f; return rc; }
    f->faces.assign(ncomp, FaceSet());
    for (auto& fs : f->faces)
        for (int i = 0; i < 6; ++i) { fs.f[i].mode = 0; fs.f[i].value = 0.0; fs.f[i].s1 = 0.0; }
    *out = f;
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_create(rnmf_ctx* c, int32_t dim, const int64_t* n, const double* lo, const double* dx,
                                     int32_t ng, int32_t ncomp, rnmf_frame** out)
{
    RNMF_REQUIRE(n, "n is null");
    const int64_t zero[3] = { 0, 0, 0 };
    return frame_create_impl(c, dim, n, n, zero, lo, dx, ng, ncomp, 0, out);
}

extern "C" int32_t rnmf_frame_create_slab(rnmf_ctx* c, int32_t dim, const int64_t* ng_n, const double* lo,
                                          const double* dx, int32_t ng, int32_t ncomp, rnmf_frame** out)
{
    RNMF_REQUIRE(c && ng_n && lo && dx, "null argument");
    RNMF_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    const int D = dim - 1;
    RNMF_REQUIRE(ng_n[D] >= c->nranks, "global n[%d]=%lld smaller than nranks=%d", D, (long long)ng_n[D], c->nranks);
    int64_t start = 0, count = 0;
    rnmf_slab_partition(ng_n[D], c->nranks, c->rank, &start, &count);
    int64_t nl[3] = { 1, 1, 1 }, off[3] = { 0, 0, 0 };
    double lo_l[3] = { 0, 0, 0 };
    for (int d = 0; d < dim; ++d) { nl[d] = ng_n[d]; lo_l[d] = lo[d]; }
    nl[D] = count;
    off[D] = start;
    lo_l[D] = lo[D] + (double)start * dx[D];
    RNMF_REQUIRE(count >= ng, "slab of %lld cells thinner than n_ghost=%d", (long long)count, ng);
    return frame_create_impl(c, dim, nl, ng_n, off, lo_l, dx, ng, ncomp, c->nranks > 1 ? 1 : 0, out);
}

extern "C" int32_t rnmf_frame_destroy(rnmf_frame* f)
{
    if (!f) return RNMF_OK;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    cudaStreamSynchronize(f->ctx->stream_halo);
    for (int i = 0; i < 2; ++i) {
        if (f->peer_lo[i]) cudaIpcCloseMemHandle(f->peer_lo[i] - deep_elems(f->g));     // the mapped allocation bases
        if (f->peer_hi[i]) cudaIpcCloseMemHandle(f->peer_hi[i] - deep_elems(f->g));
    }
    free_buffer(f->g, f->d);
    free_buffer(f->g, f->d2);
    for (double* p : f->prescribed_dev) if (p) cudaFree(p);
    delete f;
    return RNMF_OK;
}

// Debug aid (compute-sanitizer is not available on every pool): every buffer ends in a canary band;
// a kernel that stores past the end of a frame destroys it.  ok = 1 when both buffers are intact.
extern "C" int32_t rnmf_frame_check_guards(rnmf_frame* f, int32_t* ok)
{
    RNMF_REQUIRE(f && ok, "null argument");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaStreamSynchronize(f->ctx->stream));
    *ok = 1;
    unsigned long long host[kGuard];
    for (double* buf : { f->d, f->d2 }) {
        if (!buf) continue;
        RNMF_CUDA_TRY(cudaMemcpy(host, buf + tail_off(f->g) + kTailPad, sizeof(host), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < kGuard; ++i)
            if (host[i] != kCanary + (unsigned long long)i) { *ok = 0; break; }
    }
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_info_get(const rnmf_frame* f, rnmf_frame_info* o)
{
    RNMF_REQUIRE(f && o, "null argument");
    memset(o, 0, sizeof(*o));
    o->dim = f->g.dim; o->n_ghost = f->g.ng; o->n_comp = f->g.ncomp; o->distributed = f->distributed;
    for (int d = 0; d < 3; ++d) {
        o->n[d] = f->g.n[d]; o->n_global[d] = f->g.nglob[d]; o->offset[d] = f->g.off[d];
        o->dx[d] = f->g.dx[d]; o->lo[d] = f->g.lo[d];
    }
    o->pitch = f->g.pitch; o->x_offset = f->g.xoff;
    o->n_rows = f->g.G[1] * f->g.G[2] * f->g.ncomp;
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_set_bc(rnmf_frame* f, const rnmf_bc_face* faces, int32_t n_faces)
{
    RNMF_REQUIRE(f && faces, "null argument");
    const FrameGeom& g = f->g;
    RNMF_REQUIRE(n_faces == g.ncomp * g.dim * 2, "expected %d faces (n_comp*dim*2), got %d", g.ncomp * g.dim * 2, n_faces);
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    for (int i = 0; i < n_faces; ++i)
        RNMF_REQUIRE(faces[i].kind >= RNMF_BC_DIRICHLET && faces[i].kind <= RNMF_BC_HALO, "face %d: unknown kind %d", i, faces[i].kind);
    for (double* p : f->prescribed_dev) if (p) RNMF_CUDA_TRY(cudaFree(p));
    f->prescribed_dev.assign(n_faces, nullptr);
    f->prescribed_n.assign(n_faces, 0);
    f->raw.assign(faces, faces + n_faces);
    f->has_reflect = 0;
    const int D = slow_dim(g);
    for (int c = 0; c < g.ncomp; ++c)
        for (int d = 0; d < g.dim; ++d)
            for (int side = 0; side < 2; ++side) {
                const int idx = (c * g.dim + d) * 2 + side;
                rnmf_bc_face& r = f->raw[idx];
                // internal faces of a slab are halos whatever the global BC says
                if (f->distributed && d == D) {
                    const bool internal = side == 0 ? f->ctx->rank > 0 : f->ctx->rank < f->ctx->nranks - 1;
                    if (internal) r.kind = RNMF_BC_HALO;
                }
                if (r.kind == RNMF_BC_REFLECT) f->has_reflect = 1;
                if (r.kind == RNMF_BC_PRESCRIBED && r.n_prescribed > 0) {
                    RNMF_REQUIRE(r.prescribed, "face %d: Prescribed with null values", idx);
                    RNMF_CUDA_TRY(cudaMalloc(&f->prescribed_dev[idx], (size_t)r.n_prescribed * sizeof(double)));
                    RNMF_CUDA_TRY(cudaMemcpyAsync(f->prescribed_dev[idx], r.prescribed, (size_t)r.n_prescribed * sizeof(double),
                                                  cudaMemcpyHostToDevice, f->ctx->stream));
                    RNMF_CUDA_TRY(cudaStreamSynchronize(f->ctx->stream));   // caller may free its Vec right away
                    f->prescribed_n[idx] = r.n_prescribed;
                }
                r.prescribed = nullptr;
                f->faces[c].f[d * 2 + side] = make_face(r, side, g.dx[d]);
            }
    f->bc_set = 1;
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_upload(rnmf_frame* f, const double* host)
{
    RNMF_REQUIRE(f && host, "null argument");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    const FrameGeom& g = f->g;
    const size_t rows = (size_t)(g.G[1] * g.G[2] * g.ncomp);
    rnmf_ctx* c = f->ctx;
    const long long dense = g.G[0] * (long long)rows;
    if (dense * 8 >= (64LL << 20)) {
        // Large frames: a pitched 2-D PCIe copy runs at ~37 GB/s on this platform, a 1-D one at ~55
        // (scripts/h2d_probe.py), so stage the dense host frame 1-D and repack on the device
        // (HBM speed, ~1 ms for 768^3).
        if (c->staging_cap < dense) {
            if (c->d_staging) { RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream)); RNMF_CUDA_TRY(cudaFree(c->d_staging)); c->d_staging = nullptr; c->staging_cap = 0; }
            RNMF_CUDA_TRY(cudaMalloc(&c->d_staging, (size_t)dense * 8));
            c->staging_cap = dense;
        }
        RNMF_CUDA_TRY(cudaMemcpyAsync(c->d_staging, host, (size_t)dense * 8, cudaMemcpyHostToDevice, c->stream));
        RNMF_LAUNCH(c, launch_unpack_dense(g, c->d_staging, f->d, c->stream));
        RNMF_CUDA_TRY(cudaGetLastError());
        return RNMF_OK;
    }
    RNMF_CUDA_TRY(cudaMemcpy2DAsync(f->d + g.xoff, (size_t)g.pitch * 8, host, (size_t)g.G[0] * 8, (size_t)g.G[0] * 8, rows,
                                    cudaMemcpyHostToDevice, c->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_download(rnmf_frame* f, double* host)
{
    RNMF_REQUIRE(f && host, "null argument");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    const FrameGeom& g = f->g;
    const size_t rows = (size_t)(g.G[1] * g.G[2] * g.ncomp);
    rnmf_ctx* c = f->ctx;
    const long long dense = g.G[0] * (long long)rows;
    if (dense * 8 >= (64LL << 20)) {
        // Large frames, the mirror image of the upload: pack on the device (HBM speed) and send the dense frame home with ONE
        // 1-D copy (~55 GB/s here; the pitched 2-D copy reaches ~37: scripts/h2d_probe.py)
        if (c->staging_cap < dense) {
            if (c->d_staging) { RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream)); RNMF_CUDA_TRY(cudaFree(c->d_staging)); c->d_staging = nullptr; c->staging_cap = 0; }
            RNMF_CUDA_TRY(cudaMalloc(&c->d_staging, (size_t)dense * 8));
            c->staging_cap = dense;
        }
        RNMF_LAUNCH(c, launch_pack_dense_rows(g, f->d, c->d_staging, 0, (long long)rows, c->stream));
        RNMF_CUDA_TRY(cudaMemcpyAsync(host, c->d_staging, (size_t)dense * 8, cudaMemcpyDeviceToHost, c->stream));
    } else {
        RNMF_CUDA_TRY(cudaMemcpy2DAsync(host, (size_t)g.G[0] * 8, f->d + g.xoff, (size_t)g.pitch * 8, (size_t)g.G[0] * 8, rows,
                                        cudaMemcpyDeviceToHost, c->stream));
    }
    RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_peer_status(c);
}

static int copy_valid(rnmf_frame* f, double* host, bool to_host)
{
    const FrameGeom& g = f->g;
    const size_t per_comp = (size_t)(g.n[0] * g.n[1] * g.n[2]);
    for (int c = 0; c < g.ncomp; ++c) {
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        cudaPitchedPtr dev = make_cudaPitchedPtr(f->d + g.comp * c, (size_t)g.pitch * 8, (size_t)g.pitch * 8, (size_t)g.G[1]);
        cudaPitchedPtr hst = make_cudaPitchedPtr(host + per_comp * c, (size_t)g.n[0] * 8, (size_t)g.n[0] * 8, (size_t)g.n[1]);
        cudaPos dpos = make_cudaPos((size_t)(g.xoff + g.ng) * 8, (size_t)g.ng, (size_t)g.kg);
        cudaPos hpos = make_cudaPos(0, 0, 0);
        p.extent = make_cudaExtent((size_t)g.n[0] * 8, (size_t)g.n[1], (size_t)g.n[2]);
        if (to_host) { p.srcPtr = dev; p.srcPos = dpos; p.dstPtr = hst; p.dstPos = hpos; p.kind = cudaMemcpyDeviceToHost; }
        else { p.srcPtr = hst; p.srcPos = hpos; p.dstPtr = dev; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
        RNMF_CUDA_TRY(cudaMemcpy3DAsync(&p, f->ctx->stream));
    }
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_download_valid(rnmf_frame* f, double* host)
{
    RNMF_REQUIRE(f && host, "null argument");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    int rc = copy_valid(f, host, true);
    if (rc) return rc;
    RNMF_CUDA_TRY(cudaStreamSynchronize(f->ctx->stream));
    return check_peer_status(f->ctx);
}

extern "C" int32_t rnmf_frame_upload_valid(rnmf_frame* f, const double* host)
{
    RNMF_REQUIRE(f && host, "null argument");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    return copy_valid(f, const_cast<double*>(host), false);
}

static int cell_offset(const rnmf_frame* f, int64_t i, int64_t j, int64_t k, int32_t comp, long long* off)
{
    const FrameGeom& g = f->g;
    const long long ng = g.ng;
    if (g.dim == 2) k = 0;
    RNMF_REQUIRE(comp >= 0 && comp < g.ncomp, "component %d out of range", comp);
    RNMF_REQUIRE(i >= -ng && i < g.n[0] + ng && j >= -ng && j < g.n[1] + ng && k >= -(long long)g.kg && k < g.n[2] + g.kg,
                 "index (%lld, %lld, %lld) outside the grown frame", (long long)i, (long long)j, (long long)k);
    *off = g.at(i, j, k, comp);
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_get_cell(rnmf_frame* f, int64_t i, int64_t j, int64_t k, int32_t comp, double* value)
{
    RNMF_REQUIRE(f && value, "null argument");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    long long off = 0;
    int rc = cell_offset(f, i, j, k, comp, &off);
    if (rc) return rc;
    RNMF_CUDA_TRY(cudaMemcpyAsync(value, f->d + off, 8, cudaMemcpyDeviceToHost, f->ctx->stream));
    RNMF_CUDA_TRY(cudaStreamSynchronize(f->ctx->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_set_cell(rnmf_frame* f, int64_t i, int64_t j, int64_t k, int32_t comp, double value)
{
    RNMF_REQUIRE(f, "frame is null");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    long long off = 0;
    int rc = cell_offset(f, i, j, k, comp, &off);
    if (rc) return rc;
    RNMF_CUDA_TRY(cudaMemcpyAsync(f->d + off, &value, 8, cudaMemcpyHostToDevice, f->ctx->stream));
    RNMF_CUDA_TRY(cudaStreamSynchronize(f->ctx->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_collect_side(rnmf_frame* f, int32_t dir, int32_t side, int32_t which, int32_t comp,
                                           double* host_out, int64_t* n_out)
{
    RNMF_REQUIRE(f && n_out, "null argument");
    const FrameGeom& g = f->g;
    RNMF_REQUIRE(dir >= 0 && dir < g.dim && (side == 0 || side == 1) && (which == 0 || which == 1), "bad dir/side/which");
    RNMF_REQUIRE(comp < g.ncomp, "component %d out of range", comp);
    rnmf_ctx* c = f->ctx;
    if (bind(c)) return RNMF_ERR_CUDA;
    const int comp0 = comp < 0 ? 0 : comp, ncomp = comp < 0 ? g.ncomp : 1;
    const long long total = collect_side_count(g, dir, which, ncomp);
    *n_out = total;
    if (!host_out || total == 0) return RNMF_OK;
    int rc = ensure_partials(c, total);              // reuse the scratch buffer
    if (rc) return rc;
    RNMF_LAUNCH(c, launch_collect_side(g, f->d, dir, side, which, comp0, ncomp, c->d_partials, c->stream));
    RNMF_CUDA_TRY(cudaMemcpyAsync(host_out, c->d_partials, (size_t)total * 8, cudaMemcpyDeviceToHost, c->stream));
    RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_fill_zero(rnmf_frame* f)
{
    RNMF_REQUIRE(f, "frame is null");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaMemsetAsync(f->d, 0, (size_t)f->g.total * 8, f->ctx->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_copy(rnmf_frame* dst, const rnmf_frame* src)
{
    RNMF_REQUIRE(dst && src, "null argument");
    RNMF_REQUIRE(dst->ctx == src->ctx && same_shape(dst->g, src->g), "frame shapes differ");
    if (bind(dst->ctx)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaMemcpyAsync(dst->d, src->d, (size_t)src->g.total * 8, cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_fill_synthetic(rnmf_frame* f, uint64_t seed)
{
    RNMF_REQUIRE(f, "frame is null");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    RNMF_CUDA_TRY(cudaMemsetAsync(f->d, 0, (size_t)f->g.total * 8, f->ctx->stream));
    RNMF_LAUNCH(f->ctx, launch_fill_synthetic(f->g, f->d, seed, f->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_device_ptr(rnmf_frame* f, void** p)
{
    RNMF_REQUIRE(f && p, "null argument");
    *p = f->d;
    return RNMF_OK;
}

// ------------------------------------------------------------------------------------------
// fill_bc
// ------------------------------------------------------------------------------------------
static int fill_bc_on(rnmf_frame* f, double* buf, cudaStream_t s)
{
    RNMF_REQUIRE(f->bc_set, "fill_bc: frame has no boundary conditions (BCs::empty() would panic in the reference)");
    if (f->has_reflect) {
        rnmf_set_error("Reflect boundary condition not supported yet (dataframe.rs:375-377)");
        return RNMF_ERR_UNIMPLEMENTED;
    }
    const FrameGeom& g = f->g;
    int k = launch_fill_bc_faces(g, buf, f->faces.data(), s);
    f->ctx->launches += k;
    // Prescribed sides overlap each other at the corners: keep the reference's order
    // comp -> dim -> lo, hi (dataframe.rs:362-430)
    for (int c = 0; c < g.ncomp; ++c)
        for (int d = 0; d < g.dim; ++d)
            for (int side = 0; side < 2; ++side) {
                const int idx = (c * g.dim + d) * 2 + side;
                if (f->raw[idx].kind == RNMF_BC_PRESCRIBED && f->prescribed_n[idx] > 0)
                    RNMF_LAUNCH(f->ctx, launch_fill_prescribed(g, buf, c, d, side, f->prescribed_dev[idx], f->prescribed_n[idx], s));
            }
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_fill_bc(rnmf_frame* f)
{
    RNMF_REQUIRE(f, "frame is null");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    return fill_bc_on(f, f->d, f->ctx->stream);
}

// poisson.rs:72-76; 3D by the extrapolation rule.  Host arithmetic, compiled with
// -ffp-contract=off: the same IEEE operations as the reference.
extern "C" int32_t rnmf_poisson_coefs(int32_t dim, const double* dx, double* coef)
{
    RNMF_REQUIRE(dx && coef && (dim == 2 || dim == 3), "bad arguments");
    if (dim == 2) {
        const double dx2 = dx[0] * dx[0], dy2 = dx[1] * dx[1];
        coef[0] = dx2 * dy2 / (2.0 * (dx2 + dy2));
        coef[1] = dy2 / (2.0 * (dx2 + dy2));
        coef[2] = dx2 / (2.0 * (dx2 + dy2));
        coef[3] = 0.0;
    } else {
        const double dx2 = dx[0] * dx[0], dy2 = dx[1] * dx[1], dz2 = dx[2] * dx[2];
        const double a = 1.0 / (2.0 * (1.0 / dx2 + 1.0 / dy2 + 1.0 / dz2));
        coef[0] = a; coef[1] = a / dx2; coef[2] = a / dy2; coef[3] = a / dz2;
    }
    return RNMF_OK;
}

// ------------------------------------------------------------------------------------------
// halo exchange: the slab's first/last valid slices go to the neighbours' ghost slices.
// Slices of the slowest direction are contiguous in memory (pitch*G1 [*1] doubles per z-plane,
// pitch doubles per y-row in 2D), so no pack kernel is needed: pack order == collect_typed_cells
// order (dataframe.rs:273-278, pinned by :1148-1201), unpack == fill_prescribed_bc (:280-284).
// ------------------------------------------------------------------------------------------
// depth = ng: the ghost slices (the reference's halo).  depth = ng + 1 (deep halo, 3D one-component ng=1 frames only): the
// slab's TWO slices next to the face go into the neighbour's ghost + deep plane, which are adjacent in memory on both sides.
static int halo_exchange_on(rnmf_frame* f, double* buf, cudaStream_t s, int extra_depth = 0)
{
    rnmf_ctx* c = f->ctx;
    if (!f->distributed || c->nranks == 1) return RNMF_OK;
    const FrameGeom& g = f->g;
    const int D = slow_dim(g);
    const long long slice = D == 2 ? g.plane : g.pitch;          // doubles per slowest-direction slice
    const long long ng = g.ng;
    const long long depth = ng + extra_depth;
    RNMF_REQUIRE(extra_depth == 0 || (deep_elems(g) >= slice * extra_depth && g.n[D] >= depth), "deep halo exchange needs deep planes and a slab of >= %lld slices", depth);
    const long long cnt = slice * depth;
    for (int comp = 0; comp < g.ncomp; ++comp) {
        double* base = buf + g.comp * comp;                       // grown slice 0 of this component
        double* lo_ghost = base + slice * (ng - depth);           // slices [ng-depth, ng)  (deep slices sit in front of the buffer)
        double* lo_valid = base + slice * ng;                     // slices [ng, ng+depth)
        double* hi_valid = base + slice * (g.n[D] + ng - depth);  // slices [n+ng-depth, n+ng)   (grown index)
        double* hi_ghost = base + slice * (g.n[D] + ng);          // slices [n+ng, n+ng+depth)
        RNMF_NCCL_TRY(g_nccl.GroupStart());
        if (c->rank > 0) {
            RNMF_NCCL_TRY(g_nccl.Send(lo_valid, (size_t)cnt, ncclDouble, c->rank - 1, c->comm, s));
            RNMF_NCCL_TRY(g_nccl.Recv(lo_ghost, (size_t)cnt, ncclDouble, c->rank - 1, c->comm, s));
        }
        if (c->rank < c->nranks - 1) {
            RNMF_NCCL_TRY(g_nccl.Send(hi_valid, (size_t)cnt, ncclDouble, c->rank + 1, c->comm, s));
            RNMF_NCCL_TRY(g_nccl.Recv(hi_ghost, (size_t)cnt, ncclDouble, c->rank + 1, c->comm, s));
        }
        RNMF_NCCL_TRY(g_nccl.GroupEnd());
    }
    return RNMF_OK;
}

extern "C" int32_t rnmf_halo_exchange(rnmf_frame* f)
{
    RNMF_REQUIRE(f, "frame is null");
    if (bind(f->ctx)) return RNMF_ERR_CUDA;
    return halo_exchange_on(f, f->d, f->ctx->stream);
}

extern "C" int32_t rnmf_allreduce_sum(rnmf_ctx* c, double* inout)
{
    RNMF_REQUIRE(c && inout, "null argument");
    if (c->nranks == 1) return RNMF_OK;
    if (bind(c)) return RNMF_ERR_CUDA;
    c->h_scalar[0] = *inout;
    RNMF_CUDA_TRY(cudaMemcpyAsync(c->d_scalar, c->h_scalar, 8, cudaMemcpyHostToDevice, c->stream));
    RNMF_NCCL_TRY(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclDouble, ncclSum, c->comm, c->stream));
    RNMF_CUDA_TRY(cudaMemcpyAsync(c->h_scalar, c->d_scalar, 8, cudaMemcpyDeviceToHost, c->stream));
    RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *inout = c->h_scalar[0];
    return RNMF_OK;
}

// ------------------------------------------------------------------------------------------
// sweeps
// ------------------------------------------------------------------------------------------
static int check_sweep_args(rnmf_frame* psi, const rnmf_frame* rhs, const double* coef, int64_t n_sweeps)
{
    RNMF_REQUIRE(psi && rhs && coef, "null argument");
    RNMF_REQUIRE(psi->ctx == rhs->ctx, "frames belong to different contexts");
    RNMF_REQUIRE(same_shape(psi->g, rhs->g), "psi and rhs shapes differ");
    RNMF_REQUIRE(psi->g.ncomp == 1, "sweeps need n_comp == 1 (poisson.rs uses component 0 of a 1-component frame)");
    RNMF_REQUIRE(n_sweeps >= 0, "n_sweeps < 0");
    RNMF_REQUIRE(psi->bc_set, "psi has no boundary conditions");
    if (psi->has_reflect) {
        rnmf_set_error("Reflect boundary condition not supported yet (dataframe.rs:375-377)");
        return RNMF_ERR_UNIMPLEMENTED;
    }
    return RNMF_OK;
}

static bool has_prescribed(const rnmf_frame* f)
{
    for (auto& r : f->raw) if (r.kind == RNMF_BC_PRESCRIBED) return true;
    return false;
}

// A temporally blocked launch across slabs: u_depth of my `depth` slices next to each neighbour goes into the ghost + deep
// slices of ITS buffer with physical id `out_phys` (0 = the buffer the frame was created with), under the same
// one-epoch-per-launch handshake as a single sweep.  boundary_ctas = CTAs per side that take part in the handshake.
static void setup_slab_launch(rnmf_frame* psi, int out_phys, int boundary_ctas, SweepArgs& t)
{
    rnmf_ctx* c = psi->ctx;
    const FrameGeom& g = psi->g;
    const int D = slow_dim(g);
    const long long slice_sz = D == 2 ? g.plane : g.pitch;
    const bool has_lo = c->rank > 0, has_hi = c->rank < c->nranks - 1;
    int64_t n_lo_slab = 0;
    if (has_lo) rnmf_slab_partition(g.nglob[D], c->nranks, c->rank - 1, nullptr, &n_lo_slab);
    t.nb_lo = has_lo; t.nb_hi = has_hi;
    t.peer_lo_plane = has_lo ? psi->peer_lo[out_phys] + slice_sz * (n_lo_slab + g.ng) : nullptr;    // its hi ghost slice
    t.peer_hi_plane = has_hi ? psi->peer_hi[out_phys] + slice_sz * (g.ng - 1) : nullptr;             // its lo ghost slice
    t.ps.enabled = 1;
    t.ps.wait_lo = has_lo ? c->d_flags + 0 : nullptr;
    t.ps.wait_hi = has_hi ? c->d_flags + 1 : nullptr;
    t.ps.sig_lo = has_lo ? c->peer_flag_lo : nullptr;
    t.ps.sig_hi = has_hi ? c->peer_flag_hi : nullptr;
    t.ps.done = c->d_flags + 3;
    if (has_lo) c->done_target[0] += (unsigned long long)boundary_ctas;
    if (has_hi) c->done_target[1] += (unsigned long long)boundary_ctas;
    t.ps.target_lo = c->done_target[0]; t.ps.target_hi = c->done_target[1];
    t.ps.epoch = c->epoch;
    t.ps.status = c->d_flags + 2;
    c->epoch += 1;
    c->peer_used = 1;
}

namespace {

// The decisions of one rnmf_jacobi_sweeps call (the same on every rank of a slab job: they depend on the frame's GLOBAL shape,
// the options and the table of sweep variants only).
struct SweepCall {
    bool dist;            // slab frame on a multi-rank context
    bool fused;           // the kernels write the face ghosts themselves (ng == 1, no Prescribed side)
    bool peer;            // halo exchange by peer-memory stores from inside the sweep kernels
    bool has_lo, has_hi;  // slab neighbours
    bool tb;              // groups of `depth` sweeps run temporally blocked
    bool tb_peer;         // ... across slabs (deep halo)
    bool pairs;           // leftovers of the groups run as pairs (2D, single GPU)
    int depth;            // sweeps per temporally blocked launch
    int D;                // slowest direction
    long long m;          // slices of this rank along D
};

int plan_sweep_call(const rnmf_frame* psi, const SweepArgs& a, SweepCall* out)
{
    const rnmf_ctx* c = psi->ctx;
    const FrameGeom& g = psi->g;
    SweepCall p;
    p.dist = psi->distributed && c->nranks > 1;
    p.D = slow_dim(g);
    p.m = g.n[p.D];
    p.fused = a.fuse_fill != 0;
    // fused compute + halo exchange over NVLink peer memory: one launch per sweep (or per group of sweeps), the boundary
    // slices are stored into the neighbours' ghost slices by the sweep kernel itself
    p.peer = p.dist && psi->peer_attached && c->peer_halo && p.fused && sweep_supports_peer_halo(a);
    p.has_lo = p.dist && c->rank > 0;
    p.has_hi = p.dist && c->rank < c->nranks - 1;
    // Temporal blocking (several sweeps per pass over HBM) for every group of sweeps that does not carry the norm.
    // A face without a Dirichlet/Neumann rule must be a slab neighbour.
    const SweepVariantInfo* vi = sweep_variant_info(g.dim, c->sweep_variant);
    RNMF_REQUIRE(vi, "unknown sweep_variant %d for %dD frames", c->sweep_variant, g.dim);
    bool all_faces = true, side_faces = true;                       // side faces: those across the marching direction
    for (int fi = 0; fi < 2 * g.dim; ++fi) all_faces = all_faces && a.faces.f[fi].mode != 0;
    for (int fi = 0; fi < 2 * (g.dim - 1); ++fi) side_faces = side_faces && a.faces.f[fi].mode != 0;
    const bool slab_faces = side_faces && (p.has_lo || a.faces.f[2 * p.D].mode != 0) && (p.has_hi || a.faces.f[2 * p.D + 1].mode != 0);
    // three sweeps per launch (3D) need even n0 / n1 (kernels_sweep_tb3.cu) and, across slabs, six planes per slab: frames that do
    // not qualify run two sweeps per launch (the same decision on every rank: global shape and options only)
    int depth = vi->depth;
    if (g.dim == 3 && depth == 3 && !triple_sweep_applicable(g)) depth = 2;
    // slabs: the fused peer path with the in-kernel handshake and the deep halo; the thinnest slab (floor(n/P),
    // cartesian2d/mod.rs:64-73) has to hold the 2*depth slices that travel
    // (the 2D pair kernel, kernels_sweep2d_tb2.cu, has no slab form: variant 30 runs single sweeps across slabs)
    auto slab_ok = [&](int d) {
        return d > 1 && (g.dim == 3 || d > 2) && p.peer && c->peer_inkernel && slab_faces && deep_slices(g) >= d - 1 && g.nglob[p.D] / c->nranks >= 2 * d;
    };
    if (p.dist && g.dim == 3 && depth == 3 && !slab_ok(3)) depth = 2;
    p.depth = depth;
    p.tb_peer = p.dist && slab_ok(depth);
    p.tb = p.tb_peer || (depth > 1 && !p.dist && p.fused && all_faces);
    // sweeps left over by the groups run as pairs: 2D on a single GPU, 3D also across slabs (the two-sweep kernel has a slab form)
    p.pairs = p.tb && depth > 2 && vi->pairs && (g.dim == 3 || !p.dist);
    *out = p;
    return RNMF_OK;
}

void swap_buffers(rnmf_frame* psi)
{
    double* t = psi->d; psi->d = psi->d2; psi->d2 = t;
    psi->parity ^= 1;
}

// ONE temporally blocked launch of `sweeps` sweeps (3D: 3, or 2 for the pairs / frames the three-sweep kernel does not take; 2D: 4, or 2 for the pairs).  Returns RNMF_OK / an error code.
int launch_blocked_sweeps(rnmf_frame* psi, const SweepArgs& a, const SweepCall& p, int sweeps)
{
    rnmf_ctx* c = psi->ctx;
    SweepArgs t = a;
    t.in = psi->d; t.out = psi->d2; t.partials = nullptr;
    t.k_begin = 0; t.k_end = p.m;
    const bool d3 = a.g.dim == 3;
    if (p.tb_peer) {
        t.nb_lo = p.has_lo; t.nb_hi = p.has_hi;
        const int ctas = !d3 ? multi_sweep_2d_boundary_ctas(t) : sweeps == 3 ? triple_sweep_boundary_ctas(t) : double_sweep_boundary_ctas(t);
        setup_slab_launch(psi, psi->parity ^ 1, ctas, t);
    }
    if (d3 && sweeps == 3) RNMF_LAUNCH(c, launch_jacobi_triple_sweep(t, c->stream));   // kernels_sweep_tb3.cu
    else if (d3) RNMF_LAUNCH(c, launch_jacobi_double_sweep(t, c->stream));             // kernels_sweep_tb2.cu
    else if (sweeps > 2) RNMF_LAUNCH(c, launch_jacobi_multi_sweep_2d(t, c->stream));   // kernels_sweep2d_tbk.cu
    else RNMF_LAUNCH(c, launch_jacobi_double_sweep_2d(t, c->stream));                  // kernels_sweep2d_tb2.cu
    if (sweeps == 2) c->double_sweeps += 1;
    else { c->multi_sweeps += 1; c->multi_sweep_sweeps += sweeps; }
    swap_buffers(psi);
    return RNMF_OK;
}

// ONE sweep (+ fill_bc, + halo exchange on slabs); want_norm: per-CTA partial sums of |new - old| go to c->d_partials, *n_part of them.
int launch_single_sweep(rnmf_frame* psi, SweepArgs a, const SweepCall& p, bool want_norm, long long* n_part)
{
    rnmf_ctx* c = psi->ctx;
    const FrameGeom& g = psi->g;
    cudaStream_t s = c->stream;
    int rc;
    a.in = psi->d;
    a.out = psi->d2;
    a.partials = nullptr;
    a.k_begin = 0; a.k_end = p.m;
    *n_part = 0;
    auto with_partials = [&](SweepArgs& q, long long offset) -> int {
        const long long n = sweep_num_partials(q);
        int r = ensure_partials(c, offset + n);
        if (r) return r;
        q.partials = c->d_partials + offset;
        *n_part = offset + n;
        return RNMF_OK;
    };
    if (p.peer) {
        unsigned long long* my_flag_lo = p.has_lo ? c->d_flags + 0 : nullptr;
        unsigned long long* my_flag_hi = p.has_hi ? c->d_flags + 1 : nullptr;
        if (want_norm) { rc = with_partials(a, 0); if (rc) return rc; }
        if (c->peer_inkernel) {
            // handshake inside the sweep kernel: boundary CTAs (scheduled last) wait / publish
            const int tiles = sweep_boundary_ctas(a);           // (peer pointers and flags do not change the plan)
            setup_slab_launch(psi, psi->parity ^ 1, tiles, a);
            a.nb_lo = a.nb_hi = 0;                              // single sweeps need no deep halo
            RNMF_LAUNCH(c, launch_jacobi_sweep(a, s));
        } else {
            // the same protocol with 1-thread flag kernels around the sweep (option peer_inkernel = 0)
            RNMF_LAUNCH(c, launch_peer_wait(my_flag_lo, my_flag_hi, c->epoch, c->d_flags + 2, s));
            int64_t n_lo_slab = 0;
            if (p.has_lo) rnmf_slab_partition(g.nglob[p.D], c->nranks, c->rank - 1, nullptr, &n_lo_slab);
            const long long slice = p.D == 2 ? g.plane : g.pitch;
            const int out_phys = psi->parity ^ 1;
            a.peer_lo_plane = p.has_lo ? psi->peer_lo[out_phys] + slice * (n_lo_slab + g.ng) : nullptr;
            a.peer_hi_plane = p.has_hi ? psi->peer_hi[out_phys] + slice * (g.ng - 1) : nullptr;
            RNMF_LAUNCH(c, launch_jacobi_sweep(a, s));
            RNMF_LAUNCH(c, launch_peer_signal(p.has_lo ? c->peer_flag_lo : nullptr, p.has_hi ? c->peer_flag_hi : nullptr, c->epoch + 1, s));
            c->epoch += 1;
            c->peer_used = 1;
        }
    } else if (!p.dist || p.m < 3 || !c->overlap_halo || !p.fused) {
        if (want_norm) { rc = with_partials(a, 0); if (rc) return rc; }
        RNMF_LAUNCH(c, launch_jacobi_sweep(a, s));
        if (!p.fused) { rc = fill_bc_on(psi, psi->d2, s); if (rc) return rc; }
        if (p.dist) { rc = halo_exchange_on(psi, psi->d2, s); if (rc) return rc; }
    } else {
        // NCCL path: boundary slices first, then exchange them on the halo stream while the interior runs
        SweepArgs lo = a, hi = a, mid = a;
        lo.k_begin = 0; lo.k_end = 1;
        hi.k_begin = p.m - 1; hi.k_end = p.m;
        mid.k_begin = 1; mid.k_end = p.m - 1;
        if (want_norm) {
            rc = with_partials(lo, 0); if (rc) return rc;
            rc = with_partials(hi, *n_part); if (rc) return rc;
            rc = with_partials(mid, *n_part); if (rc) return rc;
            // ensure_partials may have moved the buffer: re-base the earlier segments
            const long long p_lo = sweep_num_partials(lo), p_hi = sweep_num_partials(hi);
            lo.partials = c->d_partials; hi.partials = c->d_partials + p_lo; mid.partials = c->d_partials + p_lo + p_hi;
        }
        RNMF_LAUNCH(c, launch_jacobi_sweep(lo, s));
        RNMF_LAUNCH(c, launch_jacobi_sweep(hi, s));
        RNMF_CUDA_TRY(cudaEventRecord(c->ev_boundary, s));
        RNMF_CUDA_TRY(cudaStreamWaitEvent(c->stream_halo, c->ev_boundary, 0));
        rc = halo_exchange_on(psi, psi->d2, c->stream_halo);
        if (rc) return rc;
        RNMF_CUDA_TRY(cudaEventRecord(c->ev_halo, c->stream_halo));
        RNMF_LAUNCH(c, launch_jacobi_sweep(mid, s));
        RNMF_CUDA_TRY(cudaStreamWaitEvent(s, c->ev_halo, 0));
    }
    swap_buffers(psi);
    return RNMF_OK;
}

}  // namespace

extern "C" int32_t rnmf_jacobi_sweeps(rnmf_frame* psi, const rnmf_frame* rhs, const double* coef, int64_t n_sweeps,
                                      double* l1_update)
{
    int rc = check_sweep_args(psi, rhs, coef, n_sweeps);
    if (rc) return rc;
    rnmf_ctx* c = psi->ctx;
    if (bind(c)) return RNMF_ERR_CUDA;
    const FrameGeom& g = psi->g;
    cudaStream_t s = c->stream;
    if (l1_update) *l1_update = 0.0;
    if (n_sweeps == 0) return RNMF_OK;

    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.c = SweepCoef{ coef[0], coef[1], coef[2], coef[3] };
    a.faces = psi->faces[0];
    a.fuse_fill = g.ng == 1 && !has_prescribed(psi) ? 1 : 0;
    a.variant = c->sweep_variant;
    a.chunk_override = c->sweep_chunk;
    a.rhs = rhs->d;
    a.deep = deep_slices(g);
    SweepCall p;
    rc = plan_sweep_call(psi, a, &p);
    if (rc) return rc;

    if (!psi->d2) {
        rc = alloc_buffer(g, &psi->d2, s);
        if (rc) return rc;
    }
    // cells no sweep/fill ever writes (edges, corners; all ghosts for ng>1 until the fill) must
    // agree in both buffers: the reference's in-place frame keeps them untouched.
    RNMF_LAUNCH(c, launch_copy_ghost_shell(g, psi->d, psi->d2, s));

    if (p.dist) {
        // Entry exchange of the current buffer's halos (deep halo: the `depth` slices next to each face; rhs: depth-1 of
        // them, the lower levels on the ghost-slice positions read them).  It is also what orders the shell copy above
        // (which rewrites d2's ghost slices) before the neighbours' first peer stores into d2 of this call: a neighbour
        // cannot leave its own entry exchange before this rank's send/recv, which is stream-ordered after the copy.
        rc = halo_exchange_on(psi, psi->d, s, p.tb_peer ? p.depth - 1 : 0);
        if (rc) return rc;
        if (p.tb_peer) { rc = halo_exchange_on(const_cast<rnmf_frame*>(rhs), rhs->d, s, p.depth - 2); if (rc) return rc; }
    }

    // every sweep but a norm-carrying last one may be grouped
    const int64_t n_groupable = n_sweeps - (l1_update ? 1 : 0);
    int64_t it = 0;
    // small (L2-resident) frames: all of them in ONE cooperative launch
    if (!p.dist && p.fused && c->resident && g.total * 8 * 3 <= c->resident_max_bytes && n_groupable >= 4) {
        int k = launch_jacobi_resident(g, psi->d, psi->d2, rhs->d, a.c, a.faces, n_groupable, s);
        if (k < 0) return RNMF_ERR_CUDA;
        c->launches += k;
        if (n_groupable & 1) swap_buffers(psi);
        it = n_groupable;
    }
    long long n_part = 0;
    while (it < n_sweeps) {
        if (p.tb && it + p.depth <= n_groupable) {                       // a full group: `depth` sweeps per pass over HBM
            rc = launch_blocked_sweeps(psi, a, p, p.depth);
            if (rc) return rc;
            it += p.depth;
        } else if (p.pairs && p.depth > 2 && it + 2 <= n_groupable) {    // leftovers as a pair
            rc = launch_blocked_sweeps(psi, a, p, 2);
            if (rc) return rc;
            it += 2;
        } else {
            const bool want_norm = l1_update && it == n_sweeps - 1;
            rc = launch_single_sweep(psi, a, p, want_norm, &n_part);
            if (rc) return rc;
            if (want_norm) {
                RNMF_LAUNCH(c, launch_reduce_partials(c->d_partials, (int)n_part, c->d_scalar, s));
                if (p.dist) RNMF_NCCL_TRY(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclDouble, ncclSum, c->comm, s));
                RNMF_CUDA_TRY(cudaMemcpyAsync(c->h_scalar, c->d_scalar, 8, cudaMemcpyDeviceToHost, s));
            }
            it += 1;
        }
    }
    // neighbours' last sweep must have landed in this frame's ghost slices before anything else in this stream touches the frame
    if (p.peer) {
        RNMF_LAUNCH(c, launch_peer_wait(p.has_lo ? c->d_flags + 0 : nullptr, p.has_hi ? c->d_flags + 1 : nullptr, c->epoch, c->d_flags + 2, s));
        c->peer_used = 1;
    }
    RNMF_CUDA_TRY(cudaGetLastError());
    if (l1_update) {
        RNMF_CUDA_TRY(cudaStreamSynchronize(s));
        *l1_update = c->h_scalar[0];
        return check_peer_status(c);
    }
    // no norm wanted: the call stays asynchronous; a handshake time-out surfaces at the next synchronising entry point
    // (rnmf_sync, rnmf_frame_download*, rnmf_abs_sum_valid, a norm-carrying sweep call)
    return RNMF_OK;
}

// ------------------------------------------------------------------------------------------
// peer-memory halo: CUDA IPC export / attach (one process per GPU on one NVSwitch box)
// ------------------------------------------------------------------------------------------
static unsigned long long peer_tag(int rank, int which)
{
    return 0x524E4D4600000000ULL ^ ((unsigned long long)(rank + 1) << 8) ^ (unsigned long long)which;
}

static int verify_peer_tag(const void* mapped, int rank, int which)
{
    unsigned long long v = 0;
    RNMF_CUDA_TRY(cudaMemcpy(&v, mapped, 8, cudaMemcpyDeviceToHost));
    if (v != peer_tag(rank, which)) {
        rnmf_set_error("peer attach: mapped pointer is not the base of rank %d's buffer %d (tag %llx)", rank, which, v);
        return RNMF_ERR_CUDA;
    }
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_peer_export(rnmf_frame* f, void* handles192)
{
    RNMF_REQUIRE(f && handles192, "null argument");
    rnmf_ctx* c = f->ctx;
    if (bind(c)) return RNMF_ERR_CUDA;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!f->d2) {
        int rc = alloc_buffer(f->g, &f->d2, c->stream);
        if (rc) return rc;
        RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles192;
    double* phys0 = f->parity == 0 ? f->d : f->d2;
    double* phys1 = f->parity == 0 ? f->d2 : f->d;
    RNMF_CUDA_TRY(cudaIpcGetMemHandle(&h[0], phys0 - deep_elems(f->g)));     // handles name the allocation bases
    RNMF_CUDA_TRY(cudaIpcGetMemHandle(&h[1], phys1 - deep_elems(f->g)));
    RNMF_CUDA_TRY(cudaIpcGetMemHandle(&h[2], c->d_flags));
    // tags in the tail padding / spare flag word: attach verifies that the pointer it maps really
    // is the base of the neighbour's buffer (guards against allocator sub-allocation offsets)
    const unsigned long long t0 = peer_tag(c->rank, 0), t1 = peer_tag(c->rank, 1), t2 = peer_tag(c->rank, 2);
    RNMF_CUDA_TRY(cudaMemcpy(phys0 + tail_off(f->g) + 8, &t0, 8, cudaMemcpyHostToDevice));
    RNMF_CUDA_TRY(cudaMemcpy(phys1 + tail_off(f->g) + 8, &t1, 8, cudaMemcpyHostToDevice));
    RNMF_CUDA_TRY(cudaMemcpy(c->d_flags + 7, &t2, 8, cudaMemcpyHostToDevice));
    return RNMF_OK;
}

extern "C" int32_t rnmf_frame_peer_attach(rnmf_frame* f, const void* lo192, const void* hi192)
{
    RNMF_REQUIRE(f, "frame is null");
    rnmf_ctx* c = f->ctx;
    RNMF_REQUIRE(f->distributed && c->nranks > 1, "peer halo needs a slab frame on a distributed context");
    RNMF_REQUIRE((c->rank == 0) == (lo192 == nullptr), "lo handles must be given exactly on ranks > 0");
    RNMF_REQUIRE((c->rank == c->nranks - 1) == (hi192 == nullptr), "hi handles must be given exactly on ranks < nranks-1");
    // doubles in a neighbour's buffer holding n_peer slices of the slowest direction
    auto peer_total = [&](int64_t n_peer) -> long long {
        const FrameGeom& g = f->g;
        const long long slice = g.dim == 3 ? g.plane : g.pitch;
        return slice * (n_peer + 2 * g.ng) * g.ncomp;
    };
    if (bind(c)) return RNMF_ERR_CUDA;
    auto open = [&](const cudaIpcMemHandle_t& h, void** out) -> int {
        RNMF_CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
        return RNMF_OK;
    };
    // a neighbour's buffer: the handle maps its allocation base; the frame pointer (its z-lo ghost plane) sits deep_elems()
    // further (same x / y extents, hence the same plane size as mine), the tag in its tail pad
    const long long deep = deep_elems(f->g);
    auto open_buffer = [&](const cudaIpcMemHandle_t& h, double** out, int64_t n_peer, int peer_rank, int which) -> int {
        void* mapped = nullptr;
        int rc = open(h, &mapped);
        if (rc) return rc;
        double* frame = (double*)mapped + deep;
        rc = verify_peer_tag(frame + peer_total(n_peer) + deep + 8, peer_rank, which);
        if (rc) { cudaIpcCloseMemHandle(mapped); return rc; }
        *out = frame;
        return RNMF_OK;
    };
    if (lo192) {
        const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)lo192;
        for (int i = 0; i < 2; ++i) {
            int rc = open_buffer(h[i], &f->peer_lo[i], f->g.nglob[f->g.dim - 1] / c->nranks, c->rank - 1, i);
            if (rc) return rc;
        }
        if (!c->peer_flag_base_lo) {
            int rc = open(h[2], &c->peer_flag_base_lo);
            if (!rc) rc = verify_peer_tag((unsigned long long*)c->peer_flag_base_lo + 7, c->rank - 1, 2);
            if (rc) return rc;
            c->peer_flag_lo = (unsigned long long*)c->peer_flag_base_lo + 1;    // its hi-side flag
        }
    }
    if (hi192) {
        const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)hi192;
        int64_t n_hi = 0;
        rnmf_slab_partition(f->g.nglob[f->g.dim - 1], c->nranks, c->rank + 1, nullptr, &n_hi);
        for (int i = 0; i < 2; ++i) {
            int rc = open_buffer(h[i], &f->peer_hi[i], n_hi, c->rank + 1, i);
            if (rc) return rc;
        }
        if (!c->peer_flag_base_hi) {
            int rc = open(h[2], &c->peer_flag_base_hi);
            if (!rc) rc = verify_peer_tag((unsigned long long*)c->peer_flag_base_hi + 7, c->rank + 1, 2);
            if (rc) return rc;
            c->peer_flag_hi = (unsigned long long*)c->peer_flag_base_hi + 0;    // its lo-side flag
        }
    }
    f->peer_attached = 1;
    return RNMF_OK;
}

extern "C" int32_t rnmf_gs_sweeps(rnmf_frame* psi, const rnmf_frame* rhs, const double* coef, int64_t n_sweeps)
{
    int rc = check_sweep_args(psi, rhs, coef, n_sweeps);
    if (rc) return rc;
    rnmf_ctx* c = psi->ctx;
    const FrameGeom& g = psi->g;
    RNMF_REQUIRE(g.dim == 2 && g.ng == 1, "reference-order GS: 2D frames with n_ghost=1 only");
    RNMF_REQUIRE(!(psi->distributed && c->nranks > 1), "reference-order GS is single-GPU (replicas only)");
    for (int i = 0; i < 4; ++i)
        RNMF_REQUIRE(psi->faces[0].f[i].mode != 0, "reference-order GS needs Dirichlet/Neumann on all four faces");
    if (bind(c)) return RNMF_ERR_CUDA;
    if (n_sweeps == 0) return RNMF_OK;
    int k = launch_gs_wavefront_2d(g, psi->d, rhs->d, SweepCoef{ coef[0], coef[1], coef[2], coef[3] }, psi->faces[0],
                                   n_sweeps, c->stream);
    if (k < 0) return RNMF_ERR_CUDA;
    c->launches += k;
    return fill_bc_on(psi, psi->d, c->stream);     // the fill after the last sweep (poisson.rs:88)
}

// ------------------------------------------------------------------------------------------
// operators, norm, variable-mu operator
// ------------------------------------------------------------------------------------------
static int check_same(const rnmf_frame* a, const rnmf_frame* b)
{
    RNMF_REQUIRE(a && b, "null argument");
    RNMF_REQUIRE(a->ctx == b->ctx && same_shape(a->g, b->g), "frame shapes differ");
    return RNMF_OK;
}

extern "C" int32_t rnmf_scale(rnmf_frame* out, double a, const rnmf_frame* x)
{
    int rc = check_same(out, x);
    if (rc) return rc;
    if (bind(out->ctx)) return RNMF_ERR_CUDA;
    RNMF_LAUNCH(out->ctx, launch_scale(x->g.total, a, x->d, out->d, out->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_add(rnmf_frame* out, const rnmf_frame* x, const rnmf_frame* y)
{
    int rc = check_same(out, x);
    if (!rc) rc = check_same(out, y);
    if (rc) return rc;
    if (bind(out->ctx)) return RNMF_ERR_CUDA;
    RNMF_LAUNCH(out->ctx, launch_add(x->g.total, x->d, y->d, out->d, out->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_mul(rnmf_frame* out, const rnmf_frame* x, const rnmf_frame* y)
{
    int rc = check_same(out, x);
    if (!rc) rc = check_same(out, y);
    if (rc) return rc;
    if (bind(out->ctx)) return RNMF_ERR_CUDA;
    RNMF_LAUNCH(out->ctx, launch_mul(x->g.total, x->d, y->d, out->d, out->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_axpby(rnmf_frame* out, double a, const rnmf_frame* x, double b, const rnmf_frame* y)
{
    int rc = check_same(out, x);
    if (!rc) rc = check_same(out, y);
    if (rc) return rc;
    if (bind(out->ctx)) return RNMF_ERR_CUDA;
    RNMF_LAUNCH(out->ctx, launch_axpby(x->g.total, a, x->d, b, y->d, out->d, out->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_abs_sum_valid(rnmf_frame* f, double* out)
{
    RNMF_REQUIRE(f && out, "null argument");
    rnmf_ctx* c = f->ctx;
    if (bind(c)) return RNMF_ERR_CUDA;
    const int nb = abs_sum_num_partials(f->g);
    int rc = ensure_partials(c, nb);
    if (rc) return rc;
    int k = launch_abs_sum_valid(f->g, f->d, c->d_partials, (int)c->partial_cap, c->d_scalar + 1, c->stream);
    if (k < 0) { rnmf_set_error("abs_sum_valid: partial buffer too small"); return RNMF_ERR_INVALID; }
    c->launches += k;
    RNMF_CUDA_TRY(cudaMemcpyAsync(c->h_scalar + 1, c->d_scalar + 1, 8, cudaMemcpyDeviceToHost, c->stream));
    RNMF_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = c->h_scalar[1];
    return check_peer_status(c);
}

static int varcoef_impl(const rnmf_frame* phi, const rnmf_mu_params* mu, int with_source, double relax, rnmf_frame* out)
{
    int rc = check_same(out, phi);
    if (rc) return rc;
    RNMF_REQUIRE(mu, "mu params null");
    RNMF_REQUIRE(phi->g.dim == 2 && phi->g.ncomp == 1, "variable-mu operator: 2D, n_comp=1 (poisson.rs:13-44)");
    RNMF_REQUIRE(out != phi, "out must not alias phi");
    // mu(i, j) is evaluated from the LOCAL index and the Dirichlet(0) mirror is written on all four sides: a slab frame
    // on rank > 0 would get a misplaced droplet and mirrored values in its halo rows.  The outer loop of the example is
    // "replicas only" (reference-order GS does not decompose), so slabs are refused rather than silently wrong.
    RNMF_REQUIRE(!(phi->distributed && phi->ctx->nranks > 1) && !(out->distributed && out->ctx->nranks > 1),
                 "variable-mu operator: slab (distributed) frames are not supported; use replicated frames");
    if (bind(out->ctx)) return RNMF_ERR_CUDA;
    RNMF_LAUNCH(out->ctx, launch_varcoef_2d(phi->g, phi->d, *mu, with_source, relax, out->d, out->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

extern "C" int32_t rnmf_varcoef_laplace_2d(const rnmf_frame* phi, const rnmf_mu_params* mu, rnmf_frame* out)
{
    return varcoef_impl(phi, mu, 0, 1.0, out);
}

extern "C" int32_t rnmf_poisson_source_2d(const rnmf_frame* phi, const rnmf_mu_params* mu, double relax, rnmf_frame* out)
{
    return varcoef_impl(phi, mu, 1, relax, out);
}

extern "C" int32_t rnmf_gradient_h_2d(const rnmf_frame* psi, rnmf_frame* h)
{
    RNMF_REQUIRE(psi && h, "null argument");
    RNMF_REQUIRE(psi->ctx == h->ctx, "different contexts");
    RNMF_REQUIRE(psi->g.dim == 2 && psi->g.ncomp == 1 && h->g.dim == 2 && h->g.ncomp == 2, "get_h: 2D psi (1 comp) -> h (2 comps)");
    RNMF_REQUIRE(psi->g.n[0] == h->g.n[0] && psi->g.n[1] == h->g.n[1] && psi->g.ng == h->g.ng, "shapes differ");
    if (bind(h->ctx)) return RNMF_ERR_CUDA;
    RNMF_LAUNCH(h->ctx, launch_gradient_h_2d(psi->g, psi->d, h->g, h->d, h->ctx->stream));
    RNMF_CUDA_TRY(cudaGetLastError());
    return RNMF_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer entry point (the reference-facing call: solve_poisson with host frames)
// ------------------------------------------------------------------------------------------
// Time-skewed start of rnmf_solve_poisson_host (opt-in, option "e2e_skew"): psi and rhs go up in chunks of z-planes on
// the halo stream (1-D PCIe copy into the staging buffer + device repack, which also seeds the ghost shell of the
// ping-pong buffer), and after every chunk the compute stream runs the first `pairs` double sweeps on the bands that
// chunk makes computable (skew_region, common.cuh), so the sweeps overlap most of the upload.  Afterwards every plane
// holds u_{2*pairs} and the frame continues with ordinary whole-frame launches.  Same kernels, same operations: the
// result is bit-identical to the plain path.  Returns the number of sweeps done (0: not applicable, nothing touched).
static bool skew_applicable(const rnmf_ctx* c, const rnmf_frame* psi)
{
    const FrameGeom& g = psi->g;
    const bool dist = psi->distributed && c->nranks > 1;
    const bool has_lo = dist && c->rank > 0, has_hi = dist && c->rank < c->nranks - 1;
    bool faces = psi->bc_set && !psi->has_reflect;
    for (int fi = 0; fi < 6 && faces; ++fi) faces = psi->faces[0].f[fi].mode != 0 || (fi == 4 && has_lo) || (fi == 5 && has_hi);
    const int C = c->e2e_skew_chunks;
    const SweepVariantInfo* vi = sweep_variant_info(3, c->sweep_variant);
    // slabs: only through the fused deep-halo path (every rank takes the same decision: the options are per job)
    const bool slab_ok = !dist || (psi->peer_attached && c->peer_halo && c->peer_inkernel);
    // every rank of a slab job must take the same decision (the wedge catch-up is collective): size the tests by the thinnest slab
    // (floor(n/P), cartesian2d/mod.rs:64-73), which is the same number on every rank
    const long long n2_min = dist ? g.nglob[2] / c->nranks : g.n[2];
    return c->e2e_skew && g.dim == 3 && g.ng == 1 && g.ncomp == 1 && (c->nranks == 1 || dist) && slab_ok && faces && C >= 2 && n2_min / C >= 16 &&
           g.G[0] * g.G[1] * (n2_min + 2) * 8 >= (64LL << 20) && vi && vi->depth >= 2;
}

static int64_t skewed_upload_and_first_sweeps(rnmf_ctx* c, rnmf_frame* psi, rnmf_frame* rhs, const double* psi_in, const double* rhs_in,
                                              const double* coef, int64_t pairable, int* rc_out)
{
    *rc_out = RNMF_OK;
    const FrameGeom& g = psi->g;
    const int C = c->e2e_skew_chunks;
    const long long n2 = g.n[2];
    const long long dense = g.G[0] * g.G[1] * g.G[2];
    int64_t pairs = c->e2e_skew_pairs < pairable / 2 ? c->e2e_skew_pairs : pairable / 2;
    // Slabs: while the frames are still arriving a rank cannot exchange halos (the neighbour's boundary planes come with ITS
    // last / first chunk), so double sweep s leaves a wedge of 2s planes next to each neighbour untouched (it needs no halo:
    // level s on planes >= 2s reads level s-1 on planes >= 2s-2).  Once everything is up the wedges catch up, level by level,
    // each level ONE launch over both wedges with the usual slab handshake (the catch-up costs about S^2/n2 of a sweep per
    // level, S = pairs).  Wedges of both sides must not meet: 4S <= the thinnest slab (the same on every rank).
    const bool dist = psi->distributed && c->nranks > 1;
    const bool has_lo = dist && c->rank > 0, has_hi = dist && c->rank < c->nranks - 1;
    if (dist && pairs > g.nglob[2] / c->nranks / 4) pairs = g.nglob[2] / c->nranks / 4;
    if (!skew_applicable(c, psi) || pairs < 4) return 0;
    cudaStream_t s = c->stream, cs = c->stream_halo;
#define SKEW_TRY(expr) do { int _rc = (expr); if (_rc) { *rc_out = _rc; return 0; } } while (0)
#define SKEW_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rnmf_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); *rc_out = RNMF_ERR_CUDA; return 0; } } while (0)
    if (!psi->d2) SKEW_TRY(alloc_buffer(g, &psi->d2, s));
    if (c->staging_cap < dense) {
        SKEW_CUDA(cudaStreamSynchronize(s));
        if (c->d_staging) { SKEW_CUDA(cudaFree(c->d_staging)); c->d_staging = nullptr; c->staging_cap = 0; }
        SKEW_CUDA(cudaMalloc(&c->d_staging, (size_t)dense * 8));
        c->staging_cap = dense;
    }
    while ((int)c->ev_chunks.size() < C) {
        cudaEvent_t e;
        SKEW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_chunks.push_back(e);
    }
    // the copy stream starts after everything queued on the compute stream (earlier readers of the staging buffer / frames)
    SKEW_CUDA(cudaEventRecord(c->ev_boundary, s));
    SKEW_CUDA(cudaStreamWaitEvent(cs, c->ev_boundary, 0));
    const long long plane_dense = g.G[0] * g.G[1];
    for (int ch = 0; ch < C; ++ch) {
        // grown planes of chunk ch: valid planes [Z(ch), Z(ch+1)) plus the bottom / top ghost plane in the first / last chunk
        const long long p0 = ch == 0 ? 0 : n2 * ch / C + 1, p1 = ch == C - 1 ? g.G[2] : n2 * (ch + 1) / C + 1;
        const long long off = plane_dense * p0, cnt = plane_dense * (p1 - p0);
        SKEW_CUDA(cudaMemcpyAsync(c->d_staging + off, psi_in + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, cs));
        { int k = launch_unpack_dense_rows(g, c->d_staging, psi->d, psi->d2, g.G[1] * p0, g.G[1] * p1, cs); if (k < 0) { *rc_out = RNMF_ERR_CUDA; return 0; } c->launches += k; }
        SKEW_CUDA(cudaMemcpyAsync(c->d_staging + off, rhs_in + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, cs));
        { int k = launch_unpack_dense_rows(g, c->d_staging, rhs->d, nullptr, g.G[1] * p0, g.G[1] * p1, cs); if (k < 0) { *rc_out = RNMF_ERR_CUDA; return 0; } c->launches += k; }
        SKEW_CUDA(cudaEventRecord(c->ev_chunks[ch], cs));
    }
    SweepArgs t;
    memset(&t, 0, sizeof(t));
    t.g = g;
    t.c = SweepCoef{ coef[0], coef[1], coef[2], coef[3] };
    t.faces = psi->faces[0];
    t.fuse_fill = 1;
    t.variant = c->sweep_variant;
    t.chunk_override = c->sweep_chunk;
    t.rhs = rhs->d;
    t.deep = deep_slices(g);
    double* buf[2] = { psi->d, psi->d2 };
    for (int ch = 0; ch < C; ++ch) {
        SKEW_CUDA(cudaStreamWaitEvent(s, c->ev_chunks[ch], 0));
        for (int64_t sp = 1; sp <= pairs; ++sp) {
            long long lo, hi;
            skew_region(n2, C, ch, sp, &lo, &hi);
            if (has_lo && lo < 2 * sp) lo = 2 * sp;                 // the wedges next to the neighbours wait for the catch-up below
            if (has_hi && hi > n2 - 2 * sp) hi = n2 - 2 * sp;
            if (hi <= lo) continue;
            t.in = buf[(sp - 1) & 1];
            t.out = buf[sp & 1];
            t.k_begin = lo; t.k_end = hi;
            int k = launch_jacobi_double_sweep(t, s);
            if (k < 0) { *rc_out = RNMF_ERR_CUDA; return 0; }
            c->launches += k;
            c->skew_launches += k;
        }
    }
    if (dist) {
        // every plane of every rank has arrived: level-0 halos (ghost + deep plane; rhs: ghost plane), then the wedges
        SKEW_TRY(halo_exchange_on(psi, buf[0], s, 1));
        SKEW_TRY(halo_exchange_on(rhs, rhs->d, s, 0));
        for (int64_t sp = 1; sp <= pairs; ++sp) {
            SweepArgs w = t;
            w.in = buf[(sp - 1) & 1];
            w.out = buf[sp & 1];
            w.k_begin = 0; w.k_end = n2;
            w.hole_begin = has_lo ? 2 * sp : 0;
            w.hole_end = has_hi ? n2 - 2 * sp : n2;
            w.nb_lo = has_lo; w.nb_hi = has_hi;
            setup_slab_launch(psi, psi->parity ^ (int)(sp & 1), double_sweep_boundary_ctas(w), w);
            int k = launch_jacobi_double_sweep(w, s);
            if (k < 0) { *rc_out = RNMF_ERR_CUDA; return 0; }
            c->launches += k;
            c->skew_launches += k;
        }
    }
    SKEW_CUDA(cudaGetLastError());
#undef SKEW_TRY
#undef SKEW_CUDA
    if (pairs & 1) { double* tmp = psi->d; psi->d = psi->d2; psi->d2 = tmp; psi->parity ^= 1; }
    return 2 * pairs;
}

// The end of the same call (option e2e_skew >= 2): the last `levels` launches -- double sweeps, the very last one the single
// norm-carrying sweep when the norm is wanted -- run band-wise from the bottom of the frame (skew_tail_region, common.cuh),
// and every chunk of z-planes goes back to the host (device pack + 1-D PCIe copy on the halo stream) as soon as its last
// band is done, while the bands of the chunks above are still being computed.  Bit-identical field; the norm is the same
// deterministic two-stage tree over per-CTA partials, only grouped by band (so it agrees to rounding, not to the bit).
static int skewed_last_sweeps_and_download(rnmf_ctx* c, rnmf_frame* psi, rnmf_frame* rhs, const double* coef, int64_t levels,
                                           bool last_is_single, double* psi_out, double* l1_update)
{
    const FrameGeom& g = psi->g;
    const int C = c->e2e_skew_chunks;
    const long long n2 = g.n[2];
    const long long dense = g.G[0] * g.G[1] * g.G[2], plane_dense = g.G[0] * g.G[1];
    cudaStream_t s = c->stream, cs = c->stream_halo;
    if (!psi->d2) { int rc = alloc_buffer(g, &psi->d2, s); if (rc) return rc; }
    if (c->staging_cap < dense) {
        RNMF_CUDA_TRY(cudaStreamSynchronize(s));
        if (c->d_staging) { RNMF_CUDA_TRY(cudaFree(c->d_staging)); c->d_staging = nullptr; c->staging_cap = 0; }
        RNMF_CUDA_TRY(cudaMalloc(&c->d_staging, (size_t)dense * 8));
        c->staging_cap = dense;
    }
    while ((int)c->ev_chunks.size() < C) {
        cudaEvent_t e;
        RNMF_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_chunks.push_back(e);
    }
    RNMF_LAUNCH(c, launch_copy_ghost_shell(g, psi->d, psi->d2, s));      // edges / corners agree in both buffers (as rnmf_jacobi_sweeps does)
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.c = SweepCoef{ coef[0], coef[1], coef[2], coef[3] };
    a.faces = psi->faces[0];
    a.fuse_fill = 1;
    a.variant = c->sweep_variant;
    a.chunk_override = c->sweep_chunk;
    a.rhs = rhs->d;
    a.deep = deep_slices(g);
    // partial sums of the norm-carrying level, one segment per band
    std::vector<long long> part_off(C + 1, 0);
    if (last_is_single && l1_update) {
        for (int ch = 0; ch < C; ++ch) {
            long long lo, hi;
            skew_tail_region(n2, C, ch, levels, levels, &lo, &hi);
            SweepArgs q = a;
            q.k_begin = lo; q.k_end = hi;
            part_off[ch + 1] = part_off[ch] + (hi > lo ? sweep_num_partials(q) : 0);
        }
        int rc = ensure_partials(c, part_off[C]);
        if (rc) return rc;
    }
    double* buf[2] = { psi->d, psi->d2 };
    double* fin = buf[levels & 1];
    for (int ch = 0; ch < C; ++ch) {
        for (int64_t lv = 1; lv <= levels; ++lv) {
            long long lo, hi;
            skew_tail_region(n2, C, ch, lv, levels, &lo, &hi);
            if (hi <= lo) continue;
            if (lv == levels && last_is_single) {
                SweepArgs q = a;
                q.in = buf[(lv - 1) & 1]; q.out = buf[lv & 1];
                q.k_begin = lo; q.k_end = hi;
                q.partials = l1_update ? c->d_partials + part_off[ch] : nullptr;
                RNMF_LAUNCH(c, launch_jacobi_sweep(q, s));
            } else {
                SweepArgs t = a;
                t.in = buf[(lv - 1) & 1]; t.out = buf[lv & 1];
                t.k_begin = lo; t.k_end = hi; t.partials = nullptr;
                RNMF_LAUNCH(c, launch_jacobi_double_sweep(t, s));
                c->skew_launches += 1;
            }
        }
        // planes below Z(ch+1) are final: pack them and send them home while the next chunk's bands run
        RNMF_CUDA_TRY(cudaEventRecord(c->ev_chunks[ch], s));
        RNMF_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_chunks[ch], 0));
        const long long p0 = ch == 0 ? 0 : n2 * ch / C + 1, p1 = ch == C - 1 ? g.G[2] : n2 * (ch + 1) / C + 1;
        const long long off = plane_dense * p0, cnt = plane_dense * (p1 - p0);
        RNMF_LAUNCH(c, launch_pack_dense_rows(g, fin, c->d_staging, g.G[1] * p0, g.G[1] * p1, cs));
        RNMF_CUDA_TRY(cudaMemcpyAsync(psi_out + off, c->d_staging + off, (size_t)cnt * 8, cudaMemcpyDeviceToHost, cs));
    }
    if (last_is_single && l1_update) {
        RNMF_LAUNCH(c, launch_reduce_partials(c->d_partials, (int)part_off[C], c->d_scalar, s));
        RNMF_CUDA_TRY(cudaMemcpyAsync(c->h_scalar, c->d_scalar, 8, cudaMemcpyDeviceToHost, s));
    }
    RNMF_CUDA_TRY(cudaGetLastError());
    if (levels & 1) { double* tmp = psi->d; psi->d = psi->d2; psi->d2 = tmp; psi->parity ^= 1; }
    RNMF_CUDA_TRY(cudaStreamSynchronize(s));
    RNMF_CUDA_TRY(cudaStreamSynchronize(cs));
    if (l1_update) *l1_update = last_is_single ? c->h_scalar[0] : 0.0;
    return RNMF_OK;
}

// Host frames in, `n_sweeps` Jacobi sweeps (+ fill_bc each), host frame out, on EXISTING device frames (single-GPU or slab):
// psi_in / rhs_in / psi_out are this rank's frames in the reference's dense grown layout (cartesian2d/mod.rs:31-37).
// With option e2e_skew (default) the upload is chunked and overlapped with the first double sweeps -- also across slabs --
// and, on a single GPU, the download with the last ones.  Bit-identical to upload + rnmf_jacobi_sweeps + download.
extern "C" int32_t rnmf_jacobi_sweeps_host(rnmf_frame* psi, rnmf_frame* rhs, const double* psi_in, const double* rhs_in,
                                           const double* coef, int64_t n_sweeps, double* psi_out, double* l1_update)
{
    int rc = check_sweep_args(psi, rhs, coef, n_sweeps);
    if (rc) return rc;
    RNMF_REQUIRE(psi_in && rhs_in && psi_out, "null host frame");
    RNMF_REQUIRE(psi != rhs, "psi and rhs must be different frames");
    rnmf_ctx* c = psi->ctx;
    if (bind(c)) return RNMF_ERR_CUDA;
    int64_t done = 0;
    if (c->e2e_skew) {
        done = skewed_upload_and_first_sweeps(c, psi, rhs, psi_in, rhs_in, coef, n_sweeps - (l1_update ? 1 : 0), &rc);
        if (rc) return rc;
    }
    if (done == 0) {
        rc = rnmf_frame_upload(psi, psi_in);
        if (rc) return rc;
        rc = rnmf_frame_upload(rhs, rhs_in);
        if (rc) return rc;
    }
    // band-wise end of the call (single GPU): `levels` launches = 2*levels sweeps, one fewer when the last launch is the single
    // norm-carrying sweep
    const int64_t rest = n_sweeps - done;
    int64_t levels = c->e2e_skew_tail;
    const bool last_single = l1_update != nullptr;
    while (levels >= 4 && 2 * levels - (last_single ? 1 : 0) > rest) --levels;
    if (c->e2e_skew >= 2 && levels >= 4 && c->nranks == 1 && skew_applicable(c, psi)) {
        rc = rnmf_jacobi_sweeps(psi, rhs, coef, rest - (2 * levels - (last_single ? 1 : 0)), nullptr);
        if (rc) return rc;
        return skewed_last_sweeps_and_download(c, psi, rhs, coef, levels, last_single, psi_out, l1_update);
    }
    rc = rnmf_jacobi_sweeps(psi, rhs, coef, rest, l1_update);
    if (rc) return rc;
    return rnmf_frame_download(psi, psi_out);
}

extern "C" int32_t rnmf_solve_poisson_host(rnmf_ctx* c, int32_t dim, const int64_t* n, const double* lo, const double* dx,
                                           int32_t ng, const rnmf_bc_face* faces, const double* psi_in, const double* rhs_in,
                                           int64_t n_sub_iter, int32_t mode, double* psi_out, double* l1_update)
{
    RNMF_REQUIRE(c && n && lo && dx && faces && psi_in && rhs_in && psi_out, "null argument");
    RNMF_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (Jacobi) or 1 (reference-order GS)");
    FrameGeom want;
    RNMF_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    make_geom(want, dim, n, lo, dx, ng, 1);
    if (c->hp_psi && !same_shape(c->hp_psi->g, want)) {
        rnmf_frame_destroy(c->hp_psi); c->hp_psi = nullptr;
        rnmf_frame_destroy(c->hp_rhs); c->hp_rhs = nullptr;
    }
    int rc;
    if (!c->hp_psi) {
        rc = rnmf_frame_create(c, dim, n, lo, dx, ng, 1, &c->hp_psi);
        if (rc) return rc;
        rc = rnmf_frame_create(c, dim, n, lo, dx, ng, 1, &c->hp_rhs);
        if (rc) return rc;
    }
    for (int d = 0; d < dim; ++d) { c->hp_psi->g.dx[d] = dx[d]; c->hp_psi->g.lo[d] = lo[d]; c->hp_rhs->g.dx[d] = dx[d]; c->hp_rhs->g.lo[d] = lo[d]; }
    rc = rnmf_frame_set_bc(c->hp_psi, faces, dim * 2);
    if (rc) return rc;
    double coef[4];
    rnmf_poisson_coefs(dim, dx, coef);
    if (mode == 0) return rnmf_jacobi_sweeps_host(c->hp_psi, c->hp_rhs, psi_in, rhs_in, coef, n_sub_iter, psi_out, l1_update);
    rc = rnmf_frame_upload(c->hp_psi, psi_in);
    if (rc) return rc;
    rc = rnmf_frame_upload(c->hp_rhs, rhs_in);
    if (rc) return rc;
    rc = rnmf_gs_sweeps(c->hp_psi, c->hp_rhs, coef, n_sub_iter);
    if (l1_update) *l1_update = 0.0;
    if (rc) return rc;
    return rnmf_frame_download(c->hp_psi, psi_out);
}

extern "C" int32_t rnmf_host_alloc(int64_t bytes, void** out)
{
    RNMF_REQUIRE(out && bytes >= 0, "bad arguments");
    RNMF_CUDA_TRY(cudaMallocHost(out, (size_t)bytes));
    return RNMF_OK;
}

extern "C" int32_t rnmf_host_free(void* p)
{
    if (p) RNMF_CUDA_TRY(cudaFreeHost(p));
    return RNMF_OK;
}
