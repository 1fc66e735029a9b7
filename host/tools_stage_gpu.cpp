// stage_gpu: the solver API (Stage / Step / GlobalStop, solver.hpp) driving the GPU Jacobi sweep:
// a block with frames "psi" and "rhs", sweeps until the L1 update norm drops below `tol`, checked
// every `check_every` iterations.  Prints one JSON line.  Used by the GPU tests.
//   stage_gpu n max_iters check_every tol [batched]     (batched: one device call per check, solver.hpp)
#include <cstdlib>
#include <iostream>
#include <string>

#include <rnmf/solver.hpp>

struct M { static M from_lua(const rnmf::lua::Table&) { return M(); } };

int main(int argc, char** argv)
{
    using namespace rnmf;
    using namespace rnmf::boundary_conditions;
    if (argc < 5) { std::cerr << "usage: stage_gpu n max_iters check_every tol" << std::endl; return 2; }
    const std::size_t n = std::strtoul(argv[1], nullptr, 10), max_iters = std::strtoul(argv[2], nullptr, 10),
                      check_every = std::strtoul(argv[3], nullptr, 10);
    const double tol = std::strtod(argv[4], nullptr);
    try {
        config::Config<M> conf{ M() };
        BCs bc({ ComponentBCs({ BcType::Dirichlet(0.0), BcType::Dirichlet(0.0) }, { BcType::Dirichlet(1.0), BcType::Dirichlet(1.0) }) });   // src/main.rs:20-28
        CartesianBlock blk("b", { 0.0, 0.0 }, { 1.0, 1.0 }, { n, n }, { "psi", "rhs" }, { 1, 1 }, { 1, 1 }, 0, { bc, bc });
        blk["rhs"].fill_ic([](Real x, Real y, std::size_t) { return 8.0 * x * (1.0 - y); });
        blk["psi"].fill_bc();
        const bool batched = argc > 5 && std::string(argv[5]) == "batched";
        auto stage = batched ? solver::make_jacobi_stage_batched<M>("psi", "rhs", max_iters, check_every, tol)
                             : solver::make_jacobi_stage<M>("psi", "rhs", max_iters, check_every, tol);
        solver::Step<M> top(stage);
        top.execute(conf, blk);
        const std::size_t sweeps = batched ? stage->iterations_run * (check_every ? check_every : 1) : stage->iterations_run;
        std::cout.precision(17);
        std::cout << "{\"iterations\": " << sweeps << ", \"norm\": " << stage->stop.global_stop[0].global.value
                  << ", \"psi_mid\": " << blk["psi"]((std::ptrdiff_t)n / 2, (std::ptrdiff_t)n / 2, 0) << "}" << std::endl;
        return 0;
    } catch (const Panic& e) {
        std::cerr << "panic: " << e.what() << std::endl;
        return 101;
    }
}
