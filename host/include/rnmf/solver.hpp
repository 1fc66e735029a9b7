// solver.hpp -- compilable mirror of src/solver.rs + src/global.rs + the block of
// src/mesh/cartesian2d/block.rs.  In the reference both modules are commented out of lib.rs
// (lib.rs:23-27) and would not compile; they define the API SHAPE the GPU iteration loop sits
// behind: Actions > Action > Stage{iteration: Vec<Step>, stop: Stop}, Step::{Callback, Stage},
// GlobalStop{global, condition, check_every}.  Two defects of the sketch are fixed, not copied
// (SURVEY 3.3): `check_every % iter_num` (division by zero at iter 0, operands swapped) and a
// `break` that only leaves the inner `for stop` loop.  Intended semantics implemented here:
// after every `check_every`-th iteration compute the global value and end the stage if the
// comparison holds.
#pragma once

#include <functional>
#include <map>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "config.hpp"
#include "rnmf.hpp"

namespace rnmf {

// block.rs:24-86: the data frames of one block behind human-readable names
class CartesianBlock {
public:
    std::size_t id = 0;
    std::string name;
    std::shared_ptr<mesh::CartesianMesh2D> mesh;
    std::array<std::ptrdiff_t, 2> low_connect{ -1, -1 }, high_connect{ -1, -1 };   // block.rs:37-41 (None = -1)

    CartesianBlock(std::string nm, RealVec2 low, RealVec2 high, UIntVec2 n_cells, const std::vector<std::string>& data_names,
                   const std::vector<std::size_t>& num_comps, const std::vector<std::size_t>& num_ghost, std::size_t id_,
                   const std::vector<boundary_conditions::BCs>& bcs)
        : id(id_), name(std::move(nm)), mesh(mesh::CartesianMesh2D::create(low, high, n_cells))
    {
        for (std::size_t i = 0; i < data_names.size(); ++i)
            dfs_.emplace(data_names[i], CartesianDataFrame2D::new_from(mesh, bcs.at(i), num_comps.at(i), num_ghost.at(i)));
    }
    // Index<String>, block.rs:45-56 (panics when the name is unknown)
    CartesianDataFrame2D& operator[](const std::string& nm)
    {
        auto it = dfs_.find(nm);
        if (it == dfs_.end()) throw Panic(nm + " not in CartesianBlock with id: " + std::to_string(id) + ", name: " + name);
        return it->second;
    }
    const CartesianDataFrame2D& operator[](const std::string& nm) const { return const_cast<CartesianBlock*>(this)->operator[](nm); }

private:
    std::map<std::string, CartesianDataFrame2D> dfs_;
};

namespace global {

// global.rs:5-11
struct GlobalCompare {
    enum Op { Greater, Less, GreaterEq, LessEq, Eq } op;
    Real threshold;
};

// global.rs:14-23
struct GlobalVal {
    Real value = 0.0;
    std::function<Real(const CartesianBlock&)> get_value;
    void calc_value(const CartesianBlock& dfs) { value = get_value(dfs); }
};

}  // namespace global

namespace solver {

// solver.rs:24-41
struct GlobalStop {
    global::GlobalVal global;
    global::GlobalCompare condition;
    std::size_t check_every = 1;
    bool check_stop_condition() const
    {
        const Real v = global.value, t = condition.threshold;
        switch (condition.op) {
        case global::GlobalCompare::Greater: return v > t;
        case global::GlobalCompare::Less: return v < t;
        case global::GlobalCompare::GreaterEq: return v >= t;
        case global::GlobalCompare::LessEq: return v <= t;
        case global::GlobalCompare::Eq: return v == t;
        }
        return false;
    }
};

// solver.rs:9-22
struct Stop {
    std::size_t max_iters = 0;
    std::vector<GlobalStop> global_stop;     // Option<&mut [GlobalStop]>: empty = None
};

template <class T>
struct Stage;

// solver.rs:161-203
template <class T>
struct Step {
    using Callback = std::function<void(const config::Config<T>&, CartesianBlock&)>;
    std::variant<Callback, std::shared_ptr<Stage<T>>> what;
    Step(Callback cb) : what(std::move(cb)) {}
    Step(std::shared_ptr<Stage<T>> st) : what(std::move(st)) {}
    void execute(const config::Config<T>& config, CartesianBlock& data_buf);
};

// solver.rs:124-159
template <class T>
struct Iteration : std::vector<Step<T>> {
    void execute(const config::Config<T>& config, CartesianBlock& data_buf)
    {
        std::size_t iter = 0;
        for (auto& step : *this) {
            ++iter;
            step.execute(config, data_buf);
            if (config.residual_iters != 0 && iter % config.residual_iters == 0) {
                // "compute all the globals and residuals" (solver.rs:154-156, empty in the sketch):
                // residuals are produced by the steps themselves (fused norm of rnmf_jacobi_sweeps)
            }
        }
    }
};

// solver.rs:98-122
template <class T>
struct Stage {
    Iteration<T> iteration;
    Stop stop;
    std::size_t iterations_run = 0;          // bookkeeping (not in the sketch)
    void execute(const config::Config<T>& config, CartesianBlock& data_buf) { iteration.execute(config, data_buf); }
};

template <class T>
void Step<T>::execute(const config::Config<T>& config, CartesianBlock& data_buf)
{
    if (auto* cb = std::get_if<Callback>(&what)) {
        (*cb)(config, data_buf);                                   // solver.rs:177
        return;
    }
    Stage<T>& stage = *std::get<std::shared_ptr<Stage<T>>>(what);
    stage.iterations_run = 0;
    for (std::size_t iter_num = 0; iter_num < stage.stop.max_iters; ++iter_num) {     // solver.rs:181
        stage.execute(config, data_buf);
        ++stage.iterations_run;
        bool done = false;
        for (auto& stop : stage.stop.global_stop) {
            const std::size_t every = stop.check_every ? stop.check_every : 1;
            if ((iter_num + 1) % every == 0) {                      // fixed form of solver.rs:189
                stop.global.calc_value(data_buf);
                if (stop.check_stop_condition()) { done = true; break; }
            }
        }
        if (done) break;                                            // leaves the STAGE loop (solver.rs:191-193 did not)
    }
}

// solver.rs:66-96, 43-63
template <class T>
struct Action {
    Stage<T> action;
};
template <class T>
struct Actions : std::vector<Action<T>> {
    void execute(const config::Config<T>& config, CartesianBlock& data_buf)
    {
        for (auto& a : *this) {
            Step<T> s(std::make_shared<Stage<T>>(a.action));
            s.execute(config, data_buf);
        }
    }
};

// The GPU iteration loop behind this API: a Stage whose single step is one Jacobi sweep of
// `psi` against `rhs` on the device (fused ghost fill), and whose GlobalStop is the L1 norm of the
// update, produced by the sweep kernel itself on exactly the iterations that are checked, so the
// host only synchronises every `check_every` sweeps.
template <class T>
std::shared_ptr<Stage<T>> make_jacobi_stage(const std::string& psi, const std::string& rhs, std::size_t max_iters,
                                            std::size_t check_every, Real tol)
{
    struct State { std::size_t iter = 0; Real last_norm = HUGE_VAL; };
    auto st = std::make_shared<State>();
    auto stage = std::make_shared<Stage<T>>();
    stage->stop.max_iters = max_iters;
    typename Step<T>::Callback cb = [st, psi, rhs, check_every](const config::Config<T>&, CartesianBlock& b) {
        CartesianDataFrame2D& f = b[psi];
        const auto& m = *f.underlying_mesh;
        const double dx[3] = { m.dx[0], m.dx[1], 1.0 };
        double coef[4];
        check(rnmf_poisson_coefs(2, dx, coef), "rnmf_poisson_coefs");
        const bool want = (st->iter + 1) % (check_every ? check_every : 1) == 0;
        double l1 = 0.0;
        check(rnmf_jacobi_sweeps(f.handle(), b[rhs].handle(), coef, 1, want ? &l1 : nullptr), "rnmf_jacobi_sweeps");
        f.invalidate_mirror();
        if (want) st->last_norm = l1;
        ++st->iter;
    };
    stage->iteration.emplace_back(cb);
    GlobalStop gs;
    gs.global.get_value = [st](const CartesianBlock&) { return st->last_norm; };
    gs.condition = { global::GlobalCompare::Less, tol };
    gs.check_every = check_every;
    stage->stop.global_stop.push_back(gs);
    return stage;
}

// The same loop with the sweeps between two checks handed to the device in ONE call: a step is
// `check_every` sweeps (the update norm comes from the last of them), the stage stops after
// ceil(max_sweeps / check_every) steps or when the norm drops below `tol`.  The stop test therefore
// happens at exactly the sweep counts of make_jacobi_stage and the field is bit-identical, but
// rnmf_jacobi_sweeps can pair the sweeps up (two per pass over HBM) or keep a small frame resident in
// L2, and the host launches once per check instead of once per sweep.  sweeps_run() = steps x check_every.
template <class T>
std::shared_ptr<Stage<T>> make_jacobi_stage_batched(const std::string& psi, const std::string& rhs, std::size_t max_sweeps,
                                                    std::size_t check_every, Real tol)
{
    const std::size_t every = check_every ? check_every : 1;
    struct State { Real last_norm = HUGE_VAL; };
    auto st = std::make_shared<State>();
    auto stage = std::make_shared<Stage<T>>();
    stage->stop.max_iters = (max_sweeps + every - 1) / every;
    typename Step<T>::Callback cb = [st, psi, rhs, every](const config::Config<T>&, CartesianBlock& b) {
        CartesianDataFrame2D& f = b[psi];
        const auto& m = *f.underlying_mesh;
        const double dx[3] = { m.dx[0], m.dx[1], 1.0 };
        double coef[4];
        check(rnmf_poisson_coefs(2, dx, coef), "rnmf_poisson_coefs");
        double l1 = 0.0;
        check(rnmf_jacobi_sweeps(f.handle(), b[rhs].handle(), coef, (int64_t)every, &l1), "rnmf_jacobi_sweeps");
        f.invalidate_mirror();
        st->last_norm = l1;
    };
    stage->iteration.emplace_back(cb);
    GlobalStop gs;
    gs.global.get_value = [st](const CartesianBlock&) { return st->last_norm; };
    gs.condition = { global::GlobalCompare::Less, tol };
    gs.check_every = 1;
    stage->stop.global_stop.push_back(gs);
    return stage;
}

}  // namespace solver
}  // namespace rnmf
