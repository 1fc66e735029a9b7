// poisson.hpp -- mirror of examples/magnetostatics/poisson.rs: same function names and argument
// meaning, bodies call the B200 library through the C ABI.
#pragma once

#include <rnmf/rnmf.hpp>

#include "model.hpp"

namespace magnetostatics {

using rnmf::CartesianDataFrame2D;
using rnmf::boundary_conditions::BCs;
using rnmf::boundary_conditions::BcType;
using rnmf::boundary_conditions::ComponentBCs;

inline rnmf_mu_params mu_params(const Config& config)
{
    rnmf_mu_params p;
    p.bubble_centre[0] = config.model.bubble_centre[0];
    p.bubble_centre[1] = config.model.bubble_centre[1];
    p.bubble_radius = config.model.bubble_radius;
    p.mu[0] = config.model.mu[0];
    p.mu[1] = config.model.mu[1];
    return p;
}

// poisson.rs:13-36
inline CartesianDataFrame2D get_laplace(const CartesianDataFrame2D& phi, const Config& config)
{
    const BcType z = BcType::Dirichlet(0.0);
    CartesianDataFrame2D lap = phi.like(BCs({ ComponentBCs({ z, z }, { z, z }) }));
    const rnmf_mu_params p = mu_params(config);
    rnmf::check(rnmf_varcoef_laplace_2d(phi.handle(), &p, lap.handle()), "get_laplace");
    return lap;
}

// poisson.rs:38-44 (one fused kernel: source * get_laplace over all data)
inline CartesianDataFrame2D get_source(const CartesianDataFrame2D& phi, const Config& config)
{
    CartesianDataFrame2D src = phi.like(phi.bc);
    const rnmf_mu_params p = mu_params(config);
    rnmf::check(rnmf_poisson_source_2d(phi.handle(), &p, config.model.relax, src.handle()), "get_source");
    return src;
}

enum class Order { Reference, Jacobi };

// poisson.rs:67-92.  Order::Reference keeps the reference's lexicographic Gauss-Seidel update order
// (bit-identical results); Order::Jacobi is the same expression out of place (the throughput path).
inline CartesianDataFrame2D solve_poisson(const CartesianDataFrame2D& psi_init, const CartesianDataFrame2D& rhs,
                                          const Config& config, Order order = Order::Reference)
{
    CartesianDataFrame2D psi = psi_init.clone();                         // poisson.rs:71
    const auto& m = *psi.underlying_mesh;
    const double dx[3] = { m.dx[0], m.dx[1], 1.0 };
    double coef[4];
    rnmf::check(rnmf_poisson_coefs(2, dx, coef), "poisson coefs");      // poisson.rs:72-76
    const int64_t n = (int64_t)config.model.n_sub_iter;
    if (order == Order::Reference) rnmf::check(rnmf_gs_sweeps(psi.handle(), rhs.handle(), coef, n), "solve_poisson");
    else rnmf::check(rnmf_jacobi_sweeps(psi.handle(), rhs.handle(), coef, n, nullptr), "solve_poisson");
    psi.invalidate_mirror();
    return psi;
}

// poisson.rs:94-100
inline rnmf::Real sum(const CartesianDataFrame2D& psi)
{
    double v = 0.0;
    rnmf::check(rnmf_abs_sum_valid(psi.handle(), &v), "sum");
    return v;
}

}  // namespace magnetostatics
