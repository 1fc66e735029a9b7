// model.hpp -- mirror of examples/magnetostatics/model.rs: the user model read from the Lua file.
#pragma once

#include <rnmf/config.hpp>
#include <rnmf/rnmf.hpp>

namespace magnetostatics {

using rnmf::Real;
using rnmf::RealVec2;

// model.rs:11-20; defaults of UserConfig::new, model.rs:24-35
struct UserModel {
    RealVec2 h_far{ 0.0, 0.0 };
    Real relax = 0.0;
    std::size_t n_iter = 0;
    std::size_t n_sub_iter = 0;
    RealVec2 bubble_centre{ 0.0, 0.0 };
    Real bubble_radius = 0.0;
    RealVec2 mu{ 0.0, 0.0 };
    Real tol = 0.4;

    // UserConfig::lua_constructor, model.rs:37-57 (same keys, same failure messages)
    static UserModel from_lua(const rnmf::lua::Table& t)
    {
        UserModel m;
        m.h_far = rnmf::lua::get_realvec2(t, "H_far");
        m.bubble_centre = rnmf::lua::get_realvec2(t, "bubble_centre");
        m.mu = rnmf::lua::get_realvec2(t, "mu");
        m.n_iter = rnmf::lua::get_usize(t, "n_iter");
        m.n_sub_iter = rnmf::lua::get_usize(t, "n_sub_iter");
        m.relax = rnmf::lua::get_real(t, "relax");
        m.bubble_radius = rnmf::lua::get_real(t, "bubble_radius");
        m.tol = rnmf::lua::get_real(t, "tol");
        return m;
    }
};

using Config = rnmf::config::Config<UserModel>;

}  // namespace magnetostatics
