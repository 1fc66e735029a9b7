// rnmf.hpp -- C++17 host-side mirror of wattrg/rnmf's mesh / boundary-condition / data-frame API
// over the C ABI of librnmf_b200 (include/rnmf_b200.h).
//
// The reference's host code is Rust; this image has no Rust toolchain, so the host layer that a
// Rust maintainer would write as `extern "C"` + safe wrappers is written here in
// C++ with the same names, argument meaning and error behaviour:
//   reference                                   here
//   CartesianMesh2D::new(lo,hi,n) -> Rc<..>      mesh::CartesianMesh2D::create(lo,hi,n) -> shared_ptr   (mesh.rs:36)
//   BcType::{Dirichlet,Neumann,Reflect,Prescribed}  boundary_conditions::BcType::...                    (boundary_conditions.rs:11)
//   CartesianDataFrame2D::new_from(&m,bc,nc,ng)  CartesianDataFrame2D::new_from(m,bc,nc,ng)             (dataframe.rs:71)
//   fill_ic / fill_bc / get_valid_cells          same                                                    (:166, :361, :181)
//   df[(i,j,n)]                                  df(i,j,n)  (host mirror, synced on demand)              (:208-233)
//   Real * &df, &a + &b, &a * &b                 a * df, a + b, a * b                                    (:628-691)
// The reference panics on every error of this path; here every failure throws rnmf::Panic.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/rnmf_b200.h"

namespace rnmf {

using Real = double;                         // lib.rs:30-31
using RealVec2 = std::array<Real, 2>;        // lib.rs:40-60
using UIntVec2 = std::array<std::size_t, 2>;
using IntVec2 = std::array<std::ptrdiff_t, 2>;

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int32_t rc, const char* what)
{
    if (rc != RNMF_OK) throw Panic(std::string(what) + ": " + rnmf_last_error());
}

// One CUDA device per process (the reference is single-threaded; frames are !Send).
class Context {
public:
    static Context& get(int device = 0)
    {
        static Context c(device);
        return c;
    }
    rnmf_ctx* handle() const { return h_; }
    void sync() const { check(rnmf_sync(h_), "rnmf_sync"); }
    ~Context() { rnmf_ctx_destroy(h_); }

private:
    explicit Context(int device) { check(rnmf_ctx_create(device, &h_), "rnmf_ctx_create"); }
    rnmf_ctx* h_ = nullptr;
};

namespace boundary_conditions {

// boundary_conditions.rs:11-24
struct BcType {
    enum Kind { DirichletK = RNMF_BC_DIRICHLET, NeumannK = RNMF_BC_NEUMANN, ReflectK = RNMF_BC_REFLECT, PrescribedK = RNMF_BC_PRESCRIBED };
    Kind kind = DirichletK;
    Real value = 0.0;
    std::vector<Real> values;
    static BcType Dirichlet(Real v) { return { DirichletK, v, {} }; }
    static BcType Neumann(Real g) { return { NeumannK, g, {} }; }
    static BcType Reflect() { return { ReflectK, 0.0, {} }; }
    static BcType Prescribed(std::vector<Real> v) { return { PrescribedK, 0.0, std::move(v) }; }
};

// boundary_conditions.rs:41-54
struct ComponentBCs {
    std::vector<BcType> lo, hi;
    ComponentBCs(std::vector<BcType> l, std::vector<BcType> h) : lo(std::move(l)), hi(std::move(h)) {}
};

// boundary_conditions.rs:57-70
struct BCs {
    std::vector<ComponentBCs> bcs;
    BCs() = default;
    explicit BCs(std::vector<ComponentBCs> b) : bcs(std::move(b)) {}
    static BCs empty() { return BCs(); }
};

}  // namespace boundary_conditions

namespace mesh {

// mesh.rs:8-76
struct CartesianMesh2D {
    RealVec2 lo, hi;
    UIntVec2 n;
    RealVec2 dx;
    std::size_t dim = 2;

    static std::shared_ptr<CartesianMesh2D> create(RealVec2 lo, RealVec2 hi, UIntVec2 n)
    {
        auto m = std::make_shared<CartesianMesh2D>();
        m->lo = lo; m->hi = hi; m->n = n;
        for (int d = 0; d < 2; ++d) m->dx[d] = (hi[d] - lo[d]) / static_cast<Real>(n[d]);   // mesh.rs:40
        return m;
    }
    // cell-centre coordinate, mesh.rs:69-70
    Real centre(int d, std::ptrdiff_t i) const { return lo[d] + (static_cast<Real>(i) + 0.5) * dx[d]; }
    // node_pos table (SoA: all x then all y), mesh.rs:65-74
    std::vector<Real> node_pos() const
    {
        std::vector<Real> p(n[0] * n[1] * 2);
        for (std::size_t j = 0; j < n[1]; ++j)
            for (std::size_t i = 0; i < n[0]; ++i) {
                p[i + n[0] * j] = centre(0, (std::ptrdiff_t)i);
                p[i + n[0] * j + n[0] * n[1]] = centre(1, (std::ptrdiff_t)j);
            }
        return p;
    }
};

}  // namespace mesh

// dataframe.rs:35-60 -- storage lives in HBM; `data` is an on-demand host mirror.
class CartesianDataFrame2D {
public:
    using BCs = boundary_conditions::BCs;
    std::shared_ptr<mesh::CartesianMesh2D> underlying_mesh;
    BCs bc;
    std::size_t n_comp = 1, n_ghost = 1;
    UIntVec2 n_grown{};
    std::size_t n_nodes = 0;

    // new_from(&mesh, bc, n_comp, n_ghost), dataframe.rs:71-163 (data zero-initialised, :155)
    static CartesianDataFrame2D new_from(const std::shared_ptr<mesh::CartesianMesh2D>& m, BCs bc, std::size_t n_comp,
                                         std::size_t n_ghost)
    {
        CartesianDataFrame2D f;
        f.underlying_mesh = m;
        f.n_comp = n_comp; f.n_ghost = n_ghost;
        f.n_grown = { m->n[0] + 2 * n_ghost, m->n[1] + 2 * n_ghost };
        f.n_nodes = f.n_grown[0] * f.n_grown[1] * n_comp;
        const int64_t n[3] = { (int64_t)m->n[0], (int64_t)m->n[1], 1 };
        const double lo[3] = { m->lo[0], m->lo[1], 0.0 }, dx[3] = { m->dx[0], m->dx[1], 1.0 };
        rnmf_frame* h = nullptr;
        check(rnmf_frame_create(Context::get().handle(), 2, n, lo, dx, (int32_t)n_ghost, (int32_t)n_comp, &h), "rnmf_frame_create");
        f.h_.reset(h, [](rnmf_frame* p) { rnmf_frame_destroy(p); });
        f.set_bc(std::move(bc));
        return f;
    }

    void set_bc(BCs b)
    {
        bc = std::move(b);
        if (bc.bcs.empty()) return;
        if (bc.bcs.size() != n_comp) throw Panic("BCs: one ComponentBCs per component expected");
        std::vector<rnmf_bc_face> faces(n_comp * 2 * 2);
        for (std::size_t c = 0; c < n_comp; ++c)
            for (int d = 0; d < 2; ++d)
                for (int side = 0; side < 2; ++side) {
                    const auto& t = side == 0 ? bc.bcs[c].lo.at(d) : bc.bcs[c].hi.at(d);
                    rnmf_bc_face& fc = faces[(c * 2 + d) * 2 + side];
                    fc.kind = t.kind; fc._pad = 0; fc.value = t.value;
                    fc.prescribed = t.values.empty() ? nullptr : t.values.data();
                    fc.n_prescribed = (int64_t)t.values.size();
                }
        check(rnmf_frame_set_bc(h_.get(), faces.data(), (int32_t)faces.size()), "rnmf_frame_set_bc");
    }

    // Clone (a deep copy on the device), poisson.rs:71
    CartesianDataFrame2D clone() const
    {
        CartesianDataFrame2D f = new_from(underlying_mesh, bc, n_comp, n_ghost);
        check(rnmf_frame_copy(f.h_.get(), h_.get()), "rnmf_frame_copy");
        return f;
    }
    CartesianDataFrame2D like(const BCs& b) const { return new_from(underlying_mesh, b, n_comp, n_ghost); }

    // fill_ic(|x,y,n| ..), dataframe.rs:166-173: valid cells from the cell-centre table
    void fill_ic(const std::function<Real(Real, Real, std::size_t)>& ic)
    {
        const auto& m = *underlying_mesh;
        std::vector<Real> v(m.n[0] * m.n[1] * n_comp);
        for (std::size_t c = 0; c < n_comp; ++c)
            for (std::size_t j = 0; j < m.n[1]; ++j)
                for (std::size_t i = 0; i < m.n[0]; ++i)
                    v[i + m.n[0] * (j + m.n[1] * c)] = ic(m.centre(0, (std::ptrdiff_t)i), m.centre(1, (std::ptrdiff_t)j), c);
        check(rnmf_frame_upload_valid(h_.get(), v.data()), "rnmf_frame_upload_valid");
        Context::get().sync();
        mirror_valid_ = false;
    }

    // BoundaryCondition::fill_bc, boundary_conditions.rs:4-7 / dataframe.rs:361-433
    void fill_bc()
    {
        check(rnmf_fill_bc(h_.get()), "fill_bc");
        mirror_valid_ = false;
    }

    // get_valid_cells, dataframe.rs:181-189
    std::vector<Real> get_valid_cells() const
    {
        const auto& m = *underlying_mesh;
        std::vector<Real> v(m.n[0] * m.n[1] * n_comp);
        check(rnmf_frame_download_valid(h_.get(), v.data()), "rnmf_frame_download_valid");
        return v;
    }

    // the reference's `data: Vec<S>` (dense grown layout)
    const std::vector<Real>& data() const
    {
        if (!mirror_valid_) {
            mirror_.resize(n_nodes);
            check(rnmf_frame_download(h_.get(), mirror_.data()), "rnmf_frame_download");
            mirror_valid_ = true;
        }
        return mirror_;
    }
    void upload(const std::vector<Real>& host)
    {
        if (host.size() != n_nodes) throw Panic("upload: wrong length");
        check(rnmf_frame_upload(h_.get(), host.data()), "rnmf_frame_upload");
        Context::get().sync();
        mirror_valid_ = false;
    }

    // Index<(isize,isize,usize)>, dataframe.rs:221-233 (valid cells from 0, low ghosts negative)
    Real operator()(std::ptrdiff_t i, std::ptrdiff_t j, std::size_t n) const
    {
        const std::ptrdiff_t ng = (std::ptrdiff_t)n_ghost;
        const std::size_t p = (std::size_t)(i + ng) + n_grown[0] * (std::size_t)(j + ng) + n_grown[0] * n_grown[1] * n;   // mod.rs:31-37
        return data().at(p);
    }

    // IndexMut<(isize,isize,usize)>, dataframe.rs:208-218 (one cell; not for loops)
    void set(std::ptrdiff_t i, std::ptrdiff_t j, std::size_t n, Real value)
    {
        check(rnmf_frame_set_cell(h_.get(), i, j, 0, (int32_t)n, value), "rnmf_frame_set_cell");
        mirror_valid_ = false;
    }
    // collect_typed_cells / the side iterators (dataframe.rs:273-278, 709-817): valid strip next to a
    // face (ghost=false) or the ghost cells of that side (ghost=true), flat order, corners included
    std::vector<Real> collect_side(int dir, int side, bool ghost = false, int comp = -1) const
    {
        int64_t n = 0;
        check(rnmf_frame_collect_side(h_.get(), dir, side, ghost ? 1 : 0, comp, nullptr, &n), "rnmf_frame_collect_side");
        std::vector<Real> v((std::size_t)n);
        if (n) check(rnmf_frame_collect_side(h_.get(), dir, side, ghost ? 1 : 0, comp, v.data(), &n), "rnmf_frame_collect_side");
        return v;
    }

    rnmf_frame* handle() const { return h_.get(); }
    void invalidate_mirror() { mirror_valid_ = false; }

private:
    std::shared_ptr<rnmf_frame> h_;
    mutable std::vector<Real> mirror_;
    mutable bool mirror_valid_ = false;
};

// ---- whole-frame operators, ghosts included (dataframe.rs:628-691); result takes rhs's BCs ----
inline CartesianDataFrame2D operator*(Real a, const CartesianDataFrame2D& rhs)
{
    CartesianDataFrame2D out = rhs.like(rhs.bc);
    check(rnmf_scale(out.handle(), a, rhs.handle()), "rnmf_scale");
    return out;
}
inline CartesianDataFrame2D operator+(const CartesianDataFrame2D& lhs, const CartesianDataFrame2D& rhs)
{
    CartesianDataFrame2D out = rhs.like(rhs.bc);
    check(rnmf_add(out.handle(), lhs.handle(), rhs.handle()), "rnmf_add");
    return out;
}
inline CartesianDataFrame2D operator*(const CartesianDataFrame2D& lhs, const CartesianDataFrame2D& rhs)
{
    CartesianDataFrame2D out = rhs.like(rhs.bc);
    check(rnmf_mul(out.handle(), lhs.handle(), rhs.handle()), "rnmf_mul");
    return out;
}

}  // namespace rnmf
