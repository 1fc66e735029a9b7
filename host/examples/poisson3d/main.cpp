// poisson3d -- BASELINE config "3D 7-point f64 Poisson, n^3 with mixed Dirichlet/Neumann ghost
// fill" driven straight through the C ABI from C++ (the reference has no working 3D frame; the 3D
// semantics are the extrapolation rule of SURVEY 8a).  Synthetic psi/rhs generated on the device.
//
//   poisson3d [n=512] [sweeps=200] [norm_every=10]
//
// Prints one JSON line: GLUP/s (CUDA-event timed), the last L1 update norm and sum|psi|.
#include <cstdio>
#include <cstdlib>

#include "../../../include/rnmf_b200.h"

#define CHECK(expr)                                                                  \
    do {                                                                             \
        if ((expr) != RNMF_OK) {                                                     \
            std::fprintf(stderr, "panic: %s: %s\n", #expr, rnmf_last_error());       \
            return 101;                                                              \
        }                                                                            \
    } while (0)

int main(int argc, char** argv)
{
    const int64_t n1 = argc > 1 ? std::atoll(argv[1]) : 512;
    const int64_t sweeps = argc > 2 ? std::atoll(argv[2]) : 200;
    const int64_t every = argc > 3 ? std::atoll(argv[3]) : 10;
    const int64_t n[3] = { n1, n1, n1 };
    const double lo[3] = { 0, 0, 0 }, dx[3] = { 1, 1, 1 };
    rnmf_ctx* ctx = nullptr;
    CHECK(rnmf_ctx_create(0, &ctx));
    rnmf_frame *psi = nullptr, *rhs = nullptr;
    CHECK(rnmf_frame_create(ctx, 3, n, lo, dx, 1, 1, &psi));
    CHECK(rnmf_frame_create(ctx, 3, n, lo, dx, 1, 1, &rhs));
    // x Dirichlet(0)/Dirichlet(1), y Neumann(0)/Neumann(-0.5), z Dirichlet(0.25)/Neumann(0.5)
    const rnmf_bc_face faces[6] = {
        { RNMF_BC_DIRICHLET, 0, 0.0, nullptr, 0 }, { RNMF_BC_DIRICHLET, 0, 1.0, nullptr, 0 },
        { RNMF_BC_NEUMANN, 0, 0.0, nullptr, 0 },   { RNMF_BC_NEUMANN, 0, -0.5, nullptr, 0 },
        { RNMF_BC_DIRICHLET, 0, 0.25, nullptr, 0 }, { RNMF_BC_NEUMANN, 0, 0.5, nullptr, 0 },
    };
    CHECK(rnmf_frame_set_bc(psi, faces, 6));
    CHECK(rnmf_frame_set_bc(rhs, faces, 6));
    CHECK(rnmf_frame_fill_synthetic(psi, 0x5EED0001ULL));
    CHECK(rnmf_frame_fill_synthetic(rhs, 0x5EED0002ULL));
    CHECK(rnmf_fill_bc(psi));
    double coef[4];
    CHECK(rnmf_poisson_coefs(3, dx, coef));
    double l1 = 0.0;
    CHECK(rnmf_jacobi_sweeps(psi, rhs, coef, every, &l1));          // warm-up
    CHECK(rnmf_sync(ctx));
    CHECK(rnmf_timer_begin(ctx));
    for (int64_t done = 0; done < sweeps; done += every) CHECK(rnmf_jacobi_sweeps(psi, rhs, coef, every, &l1));
    double ms = 0.0;
    CHECK(rnmf_timer_end(ctx, &ms));
    double total = 0.0;
    CHECK(rnmf_abs_sum_valid(psi, &total));
    const double lups = (double)n1 * n1 * n1 * (double)((sweeps + every - 1) / every * every);
    std::printf("{\"n\": %lld, \"sweeps\": %lld, \"ms\": %.3f, \"glups\": %.2f, \"l1_update\": %.17g, \"abs_sum\": %.17g, \"launches\": %lld}\n",
                (long long)n1, (long long)((sweeps + every - 1) / every * every), ms, lups / (ms * 1e-3) / 1e9, l1, total,
                (long long)rnmf_kernel_launches(ctx));
    CHECK(rnmf_frame_destroy(psi));
    CHECK(rnmf_frame_destroy(rhs));
    CHECK(rnmf_ctx_destroy(ctx));
    return 0;
}
