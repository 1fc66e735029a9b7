// main.cpp -- mirror of examples/magnetostatics/main.rs: the under-relaxed outer loop of the
// ferrofluid-droplet Poisson problem, driven by a .lua case, running on the B200 through the
// data-frame API (rnmf.hpp) whose bodies are librnmf_b200.
//
//   magnetostatics <case.lua> [--order reference|jacobi] [--no-vtk] [--out prefix] [--json]
//
// --order reference (default) keeps the reference's update order => its exact numbers;
// --order jacobi runs the throughput path.
#include <chrono>
#include <cstring>
#include <iostream>

#include <rnmf/io.hpp>

#include "poisson.hpp"

using namespace magnetostatics;
using rnmf::Real;

// model.rs:60-88 output callbacks
static rnmf::io::VtkData get_psi(const CartesianDataFrame2D& data) { return rnmf::io::vtk_from_frame(data); }
static rnmf::io::VtkData get_h(const CartesianDataFrame2D& data)
{
    CartesianDataFrame2D h = CartesianDataFrame2D::new_from(data.underlying_mesh, BCs::empty(), 2, data.n_ghost);
    rnmf::check(rnmf_gradient_h_2d(data.handle(), h.handle()), "get_h");
    return rnmf::io::vtk_from_frame(h);
}

int main(int argc, char** argv)
{
    try {
        std::vector<std::string> args(argv, argv + argc);
        Order order = Order::Reference;
        bool vtk = true, json = false;
        std::string out = "./ferro_droplet";
        std::vector<std::string> pos{ args[0] };
        for (std::size_t i = 1; i < args.size(); ++i) {
            if (args[i] == "--order" && i + 1 < args.size()) order = args[++i] == "jacobi" ? Order::Jacobi : Order::Reference;
            else if (args[i] == "--no-vtk") vtk = false;
            else if (args[i] == "--json") json = true;
            else if (args[i] == "--out" && i + 1 < args.size()) out = args[++i];
            else pos.push_back(args[i]);
        }
        // configure the simulation (main.rs:23-26)
        Config conf = rnmf::config::init(pos, UserModel());

        // mesh, boundary conditions, data frame (main.rs:29-54)
        auto mesh = rnmf::mesh::CartesianMesh2D::create({ 0.0, 0.0 }, { conf.geom.length[0], conf.geom.length[1] },
                                                        { conf.geom.n_cells[0], conf.geom.n_cells[1] });
        const Real gx = -conf.model.h_far[0] / mesh->dx[0], gy = -conf.model.h_far[1] / mesh->dx[1];
        BCs bc({ ComponentBCs({ BcType::Neumann(gx), BcType::Neumann(gy) }, { BcType::Neumann(gx), BcType::Neumann(gy) }) });
        CartesianDataFrame2D psi = CartesianDataFrame2D::new_from(mesh, bc, 1, 1);
        psi.fill_ic([&](Real x, Real y, std::size_t) { return -conf.model.h_far[0] / mesh->dx[0] * x - conf.model.h_far[1] / mesh->dx[1] * y; });
        psi.fill_bc();

        rnmf::io::OutputCallBackHashMap out_cb{ { "psi", get_psi }, { "h", get_h } };
        std::cout << "Calculating solution:" << std::endl;
        if (vtk) rnmf::io::write_vtk(out, "ferro_droplet", 0, psi, out_cb);

        // begin solving (main.rs:63-74)
        const auto t0 = std::chrono::steady_clock::now();
        Real sum_diff = 1.0;
        std::size_t iter = 0;
        std::vector<Real> trace;
        while (iter < conf.model.n_iter && sum_diff > conf.model.tol) {
            CartesianDataFrame2D source = get_source(psi, conf);
            CartesianDataFrame2D psi_star = solve_poisson(psi, source, conf, order);
            CartesianDataFrame2D diff = psi_star + (-1.0 * psi);
            sum_diff = sum(diff);
            psi = psi + (conf.model.relax * diff);
            psi.fill_bc();
            trace.push_back(sum_diff);
            ++iter;
        }
        rnmf::Context::get().sync();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        if (vtk) rnmf::io::write_vtk(out, "ferro_droplet", conf.model.n_iter, psi, out_cb);
        std::cout.precision(17);
        std::cout << "sum diff = " << sum_diff << std::endl;
        if (json) {
            std::cout << "{\"iters\": " << iter << ", \"seconds\": " << secs << ", \"order\": \"" << (order == Order::Jacobi ? "jacobi" : "reference")
                      << "\", \"psi_centre\": " << psi((std::ptrdiff_t)mesh->n[0] / 2, (std::ptrdiff_t)mesh->n[1] / 2, 0) << ", \"sum_diff\": [";
            for (std::size_t i = 0; i < trace.size(); ++i) std::cout << (i ? ", " : "") << trace[i];
            std::cout << "]}" << std::endl;
        }
        std::cout << "Done." << std::endl;
        return 0;
    } catch (const rnmf::Panic& e) {
        std::cerr << "panic: " << e.what() << std::endl;      // the reference panics on every error of this path
        return 101;                                            // Rust's panic exit code
    }
}
