// solver_selftest: exercises the Stage / Step / Stop / GlobalStop semantics of solver.hpp with
// host-only callbacks (no device work, no frames).  Prints one JSON line; used by the CPU tests.
#include <iostream>

#include <rnmf/solver.hpp>

struct M { static M from_lua(const rnmf::lua::Table&) { return M(); } };

// a block is needed by the signatures only; callbacks below never touch it
int main()
{
    using namespace rnmf;
    using namespace rnmf::solver;
    config::Config<M> conf{ M() };
    CartesianBlock* blk = nullptr;      // never dereferenced
    int calls = 0, checks = 0;
    double value = 100.0;
    auto stage = std::make_shared<Stage<M>>();
    stage->stop.max_iters = 50;
    stage->iteration.emplace_back(Step<M>::Callback([&](const config::Config<M>&, CartesianBlock&) { ++calls; value *= 0.5; }));
    GlobalStop gs;
    gs.global.get_value = [&](const CartesianBlock&) { ++checks; return value; };
    gs.condition = { global::GlobalCompare::Less, 1.0 };
    gs.check_every = 3;
    stage->stop.global_stop.push_back(gs);
    // nested: an outer stage that runs the inner stage twice
    auto outer = std::make_shared<Stage<M>>();
    outer->stop.max_iters = 2;
    outer->iteration.emplace_back(Step<M>(stage));
    Step<M> top(outer);
    top.execute(conf, *blk);
    // 100 * 0.5^k < 1 first at k = 7; checks happen at k = 3, 6, 9 -> the first inner run stops after 9
    // iterations; the second run starts below the threshold and stops at its first check (3 iterations)
    std::cout << "{\"calls\": " << calls << ", \"checks\": " << checks << ", \"inner_last_run\": " << stage->iterations_run
              << ", \"outer_run\": " << outer->iterations_run << "}" << std::endl;
    // a stage without global stops runs max_iters
    auto plain = std::make_shared<Stage<M>>();
    plain->stop.max_iters = 7;
    int n = 0;
    plain->iteration.emplace_back(Step<M>::Callback([&](const config::Config<M>&, CartesianBlock&) { ++n; }));
    Step<M>(plain).execute(conf, *blk);
    std::cout << "{\"plain\": " << n << "}" << std::endl;
    return 0;
}
