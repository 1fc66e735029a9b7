// config_dump: reads a .lua case with the Lua-subset reader and prints the Config as JSON.
// Host-only (no GPU): used by the CPU tests of the reader.
#include <iostream>

#include "examples/magnetostatics/model.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) { std::cerr << "usage: config_dump case.lua" << std::endl; return 2; }
    try {
        magnetostatics::Config c = rnmf::config::read_lua(argv[1], magnetostatics::UserModel());
        std::cout.precision(17);
        std::cout << "{\"dim\": " << c.geom.dim << ", \"length\": [" << c.geom.length[0] << ", " << c.geom.length[1] << "], \"n_cells\": ["
                  << c.geom.n_cells[0] << ", " << c.geom.n_cells[1] << "], \"residual_iters\": " << c.residual_iters
                  << ", \"model\": {\"H_far\": [" << c.model.h_far[0] << ", " << c.model.h_far[1] << "], \"bubble_centre\": ["
                  << c.model.bubble_centre[0] << ", " << c.model.bubble_centre[1] << "], \"bubble_radius\": " << c.model.bubble_radius
                  << ", \"mu\": [" << c.model.mu[0] << ", " << c.model.mu[1] << "], \"n_iter\": " << c.model.n_iter << ", \"n_sub_iter\": "
                  << c.model.n_sub_iter << ", \"relax\": " << c.model.relax << ", \"tol\": " << c.model.tol << "}}" << std::endl;
        return 0;
    } catch (const rnmf::Panic& e) {
        std::cerr << "panic: " << e.what() << std::endl;
        return 101;
    }
}
