"""Known-answer / self-consistency tests of the CPU oracle for the rows the reference itself
does not test (sweep, variable-mu operator, operators, norm, outer loop).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as ora

# Trajectory of `sum_diff` for magnetostatics.lua (301x301, 10 x 550 GS sweeps).  An independent
# survey-time C++ restatement of poisson.rs/main.rs produced exactly these values;
# this oracle reproduces them bit-for-bit.  NOT the output of the Rust binary (cannot be built here).
MAGNETOSTATICS_TRACE = [
    55753.058823675652, 24129.691845741065, 18839.571821757232, 15375.960245292415,
    13839.82427758238, 12512.715327638489, 11576.420030824474, 10795.923399300678,
    10165.74476166129, 9627.9718351461779,
]
MAGNETOSTATICS_PSI_150_150 = -150.5093868527895


def _rand_frame(g, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(g.n_nodes)


MIXED_2D = [([("neumann", 0.3), ("dirichlet", -1.0)], [("dirichlet", 2.0), ("neumann", -0.7)])]
MIXED_3D = [([("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 0.25)],
             [("dirichlet", 1.0), ("neumann", -0.5), ("neumann", 0.5)])]


def test_magnetostatics_trajectory_full_case():
    g = ora.geom([301, 301], hi=[301.0, 301.0])
    m = ora.model()
    psi = ora.zeros(g)
    tr = np.zeros(10)
    it = C.c_int64()
    assert ora.lib().ora_magnetostatics_run(C.byref(g), psi, C.byref(m), 0, tr, C.byref(it)) == 0
    assert it.value == 10                      # never drops below tol=0.1
    assert tr.tolist() == MAGNETOSTATICS_TRACE
    assert ora.as_grid(g, psi)[0, 151, 151] == MAGNETOSTATICS_PSI_150_150
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "magnetostatics_trace.json")) as fh:
        gold = json.load(fh)
    assert gold["sum_diff"] == MAGNETOSTATICS_TRACE and gold["psi_150_150"] == MAGNETOSTATICS_PSI_150_150


def test_linear_field_is_fixed_point():
    # 10x10 test.lua geometry; the bubble contains no cell centre -> uniform mu;
    # the linear IC is reproduced exactly by every stage and sum_diff == 0.0 after one iteration.
    g = ora.geom([10, 10], hi=[10.0, 10.0])
    m = ora.model(n_iter=3, n_sub_iter=20, bubble_centre=(5.0, 5.0), bubble_radius=0.25)
    psi = ora.zeros(g)
    tr = np.full(3, -1.0)
    it = C.c_int64()
    assert ora.lib().ora_magnetostatics_run(C.byref(g), psi, C.byref(m), 0, tr, C.byref(it)) == 0
    assert it.value == 1 and tr[0] == 0.0
    ref = ora.zeros(g)
    ora.lib().ora_fill_ic_magnetostatics(C.byref(g), ref, C.byref(m))
    np.testing.assert_array_equal(ora.valid_view(g, psi), ora.valid_view(g, ref))


@pytest.mark.parametrize("n,hi", [([37, 23], [3.0, 5.0]), ([16, 16], [16.0, 16.0]), ([1, 9], [1.0, 2.0])])
def test_wavefront_gs_equals_lexicographic_gs(n, hi):
    g = ora.geom(n, hi=hi)
    bc = ora.bcs(2, MIXED_2D)
    psi = _rand_frame(g, 1)
    rhs = _rand_frame(g, 2)          # ghosts deliberately NOT pre-filled: sweep 0 reads stored ghosts
    a, b = psi.copy(), psi.copy()
    ora.gs(g, a, rhs, bc, 13)
    ora.gs(g, b, rhs, bc, 13, wavefront=True)
    np.testing.assert_array_equal(a, b)


def test_jacobi_matches_numpy_expression_2d_and_3d():
    for dims, spec in (([19, 11], MIXED_2D), ([9, 7, 5], MIXED_3D)):
        g = ora.geom(dims, hi=[1.0 * (d + 1) for d in range(len(dims))])
        bc = ora.bcs(g.dim, spec)
        psi = _rand_frame(g, 3)
        rhs = _rand_frame(g, 4)
        ora.fill_bc(g, psi, bc)
        c = ora.coefs(g)
        u = ora.as_grid(g, psi)[0].copy()
        r = ora.as_grid(g, rhs)[0]
        if g.dim == 2:
            new = c[0] * r[1:-1, 1:-1] + c[1] * (u[1:-1, 2:] + u[1:-1, :-2]) + c[2] * (u[2:, 1:-1] + u[:-2, 1:-1])
        else:
            new = (c[0] * r[1:-1, 1:-1, 1:-1] + c[1] * (u[1:-1, 1:-1, 2:] + u[1:-1, 1:-1, :-2])
                   + c[2] * (u[1:-1, 2:, 1:-1] + u[1:-1, :-2, 1:-1]) + c[3] * (u[2:, 1:-1, 1:-1] + u[:-2, 1:-1, 1:-1]))
        old_valid = ora.valid_view(g, psi)[0].copy()
        l1 = ora.jacobi(g, psi, rhs, bc, 1)
        np.testing.assert_array_equal(ora.valid_view(g, psi)[0], new)
        assert abs(l1 - np.abs(new - old_valid).sum()) <= 1e-12 * max(l1, 1.0)
        # the ghost faces are those of a stand-alone fill; edges/corners keep their old values
        chk = psi.copy()
        ora.fill_bc(g, chk, bc)
        np.testing.assert_array_equal(chk, psi)


def test_jacobi_omp_variant_is_bit_identical():
    g = ora.geom([33, 17, 9], hi=[1.0, 2.0, 3.0])
    bc = ora.bcs(3, MIXED_3D)
    psi, rhs = _rand_frame(g, 5), _rand_frame(g, 6)
    a, b = psi.copy(), psi.copy()
    l1a = ora.jacobi(g, a, rhs, bc, 7)
    l1b = ora.jacobi(g, b, rhs, bc, 7, omp=True)
    np.testing.assert_array_equal(a, b)
    assert l1a == l1b


def test_poisson_coefs_and_ulp_insensitivity():
    # poisson.rs:72-76 with powf(2.0) restated as x*x.  A 1-ulp change of dx2 moves the
    # coefficients by O(1 ulp) only (pow restatement is not a sensitive spot).
    g = ora.geom([301, 301], hi=[301.0, 301.0])
    np.testing.assert_array_equal(ora.coefs(g), [0.25, 0.25, 0.25, 0.0])
    g2 = ora.geom([10, 20], hi=[3.0, 7.0])
    dx2, dy2 = g2.dx[0] * g2.dx[0], g2.dx[1] * g2.dx[1]
    c = ora.coefs(g2)
    assert c[0] == dx2 * dy2 / (2.0 * (dx2 + dy2)) and c[1] == dy2 / (2.0 * (dx2 + dy2))
    dx2p = np.nextafter(dx2, np.inf)
    assert abs(dx2p * dy2 / (2.0 * (dx2p + dy2)) - c[0]) <= 4 * np.spacing(c[0])
    g3 = ora.geom([4, 4, 4], hi=[4.0, 4.0, 4.0])
    np.testing.assert_allclose(ora.coefs(g3), [1 / 6, 1 / 6, 1 / 6, 1 / 6], rtol=1e-15)


def test_manufactured_solution_second_order():
    # -lap(u) = rhs with u = sin(pi x) sin(pi y), Dirichlet 0; the converged
    # discrete solution approaches u with O(h^2).
    errs = []
    for n in (8, 16):
        g = ora.geom([n, n], hi=[1.0, 1.0])
        bc = ora.bcs(2, [([("dirichlet", 0.0)] * 2, [("dirichlet", 0.0)] * 2)])
        exact, rhs, psi = ora.zeros(g), ora.zeros(g), ora.zeros(g)
        ora.fill_ic(g, exact, lambda x, y, c: np.sin(np.pi * x) * np.sin(np.pi * y))
        ora.fill_ic(g, rhs, lambda x, y, c: 2 * np.pi ** 2 * np.sin(np.pi * x) * np.sin(np.pi * y))
        ora.jacobi(g, psi, rhs, bc, 4000)
        errs.append(np.abs(ora.valid_view(g, psi) - ora.valid_view(g, exact)).max())
    assert errs[1] < errs[0] / 3.0


def test_operators_and_sum():
    g = ora.geom([5, 4], hi=[5.0, 4.0], ng=2)
    x, y = _rand_frame(g, 7), _rand_frame(g, 8)
    out = ora.zeros(g)
    L = ora.lib()
    L.ora_scale(g.n_nodes, -1.5, x, out); np.testing.assert_array_equal(out, -1.5 * x)
    L.ora_add(g.n_nodes, x, y, out); np.testing.assert_array_equal(out, x + y)
    L.ora_mul(g.n_nodes, x, y, out); np.testing.assert_array_equal(out, x * y)
    v = ora.valid_view(g, x)
    s = L.ora_sum_abs_valid(C.byref(g), x)
    acc = 0.0
    for val in v.reshape(-1):
        acc += abs(val)
    assert s == acc                                   # naive serial order, poisson.rs:94-100
    assert abs(L.ora_sum_abs_valid_ld(C.byref(g), x) - s) <= 1e-13 * s


def test_get_laplace_uniform_mu_is_5pt_laplacian():
    # with mu == 1 everywhere the operator (poisson.rs:26-31) is E+W+N+S-4C (no 1/dx^2 factor)
    g = ora.geom([12, 9], hi=[12.0, 9.0])
    m = ora.model(mu=(1.0, 1.0))
    phi = _rand_frame(g, 9)
    lap = ora.zeros(g)
    ora.lib().ora_get_laplace_2d(C.byref(g), phi, C.byref(m), lap)
    u = ora.as_grid(g, phi)[0]
    want = u[1:-1, 2:] + u[1:-1, :-2] + u[2:, 1:-1] + u[:-2, 1:-1] - 4.0 * u[1:-1, 1:-1]
    np.testing.assert_allclose(ora.valid_view(g, lap)[0], want, rtol=0, atol=1e-13)
    # ghost faces: Dirichlet(0) => ghost = -mirror
    L = ora.as_grid(g, lap)[0]
    np.testing.assert_array_equal(L[1:-1, 0], -L[1:-1, 1])
    np.testing.assert_array_equal(L[-1, 1:-1], -L[-2, 1:-1])


def test_3d_fill_bc_extrapolation_rule():
    # 3D fill = the 2D formulas per direction with dx[d], face interiors only.
    g = ora.geom([4, 3, 2], hi=[4.0, 6.0, 6.0], ng=2)
    bc = ora.bcs(3, MIXED_3D)
    a = ora.zeros(g)
    ora.fill_ic(g, a, lambda x, y, z, c: x + 10 * y + 100 * z)
    before = a.copy()
    assert ora.fill_bc(g, a, bc) == 0
    A, B = ora.as_grid(g, a)[0], ora.as_grid(g, before)[0]
    ng = 2
    # x low Dirichlet(0): u[-g-1] = 2*0 - u[g]
    np.testing.assert_array_equal(A[ng:-ng, ng:-ng, 1], -B[ng:-ng, ng:-ng, 2])
    np.testing.assert_array_equal(A[ng:-ng, ng:-ng, 0], -B[ng:-ng, ng:-ng, 3])
    # z high Neumann(0.5), dx_z = 3: u[n+g] = u[n-1-g] + 0.5*(2(g+1)-1)*3
    np.testing.assert_array_equal(A[ng + 2, ng:-ng, ng:-ng], B[ng + 1, ng:-ng, ng:-ng] + 0.5 * 1.0 * 3.0)
    np.testing.assert_array_equal(A[ng + 3, ng:-ng, ng:-ng], B[ng + 0, ng:-ng, ng:-ng] + 0.5 * 3.0 * 3.0)
    # edges and corners untouched
    mask = np.zeros_like(A, dtype=bool)
    out_of = [np.r_[0:ng, (s - ng):s] for s in A.shape]
    for ax in range(3):
        for bx in range(ax + 1, 3):
            idx = [slice(None)] * 3
            sel = np.zeros_like(mask)
            ia = np.zeros(A.shape[ax], bool); ia[out_of[ax]] = True
            ib = np.zeros(A.shape[bx], bool); ib[out_of[bx]] = True
            shp_a = [1, 1, 1]; shp_a[ax] = -1
            shp_b = [1, 1, 1]; shp_b[bx] = -1
            sel |= ia.reshape(shp_a) & ib.reshape(shp_b)
            mask |= sel
    np.testing.assert_array_equal(A[mask], B[mask])


def test_gs_3d_extrapolation_matches_plain_python_loop():
    # the 3D lexicographic sweep used by bench.py's reference arm: i outer, j, k inner, in place
    # (the continuation of poisson.rs:80-89's loop nest), then fill_bc
    n = [5, 4, 3]
    g = ora.geom(n, hi=[5.0, 8.0, 9.0])
    bc = ora.bcs(3, MIXED_3D)
    psi, rhs = _rand_frame(g, 11), _rand_frame(g, 12)
    ora.fill_bc(g, psi, bc)
    want = psi.copy()
    c = ora.coefs(g)
    u = ora.as_grid(g, want)[0]
    r = ora.as_grid(g, rhs)[0]
    for _ in range(2):
        for i in range(1, n[0] + 1):
            for j in range(1, n[1] + 1):
                for k in range(1, n[2] + 1):
                    u[k, j, i] = (c[0] * r[k, j, i] + c[1] * (u[k, j, i + 1] + u[k, j, i - 1])
                                  + c[2] * (u[k, j + 1, i] + u[k, j - 1, i]) + c[3] * (u[k + 1, j, i] + u[k - 1, j, i]))
        ora.fill_bc(g, want, bc)
    ora.gs(g, psi, rhs, bc, 2)
    np.testing.assert_array_equal(psi, want)


def test_outer_loop_against_an_independent_pure_python_restatement():
    """Second, independent restatement (plain Python floats, written from poisson.rs / main.rs /
    dataframe.rs line by line, sharing no code with the C oracle) of the whole magnetostatics loop on
    a small droplet case with a bubble that covers several cells: mu, get_laplace, get_source,
    lexicographic GS sweeps + Neumann fill, the operator chain, sum, the under-relaxed update."""
    n, L = 9, 9.0
    h_far, relax, n_iter, n_sub, centre, radius, mus, tol = (0.25, 1.0), 0.2, 3, 4, (4.3, 4.6), 2.2, (3.2, 1.0), 0.1
    dx = (L / n, L / n)                                      # mesh.rs:40
    G = n + 2

    def at(i, j):                                            # mod.rs:31-37, ng = 1
        return (i + 1) + G * (j + 1)

    def mu(i, j):                                            # poisson.rs:47-64
        x, y = (i + 0.5) * dx[0], (j + 0.5) * dx[1]
        inside = (x - centre[0]) * (x - centre[0]) + (y - centre[1]) * (y - centre[1]) <= radius * radius
        return mus[0] if inside else mus[1]

    def fill_neumann(u, grad):                               # dataframe.rs:289-322, ng = 1
        for d in range(2):
            for t in range(n):
                lo_g, lo_m = (at(-1, t), at(0, t)) if d == 0 else (at(t, -1), at(t, 0))
                hi_g, hi_m = (at(n, t), at(n - 1, t)) if d == 0 else (at(t, n), at(t, n - 1))
                u[lo_g] = u[lo_m] - grad[d] * (2.0 * 1 - 1.0) * dx[d]
                u[hi_g] = u[hi_m] + grad[d] * (2.0 * (0 + 1) - 1.0) * dx[d]

    grad = (-h_far[0] / dx[0], -h_far[1] / dx[1])            # main.rs:40-44
    psi = [0.0] * (G * G)
    for j in range(n):
        for i in range(n):
            x, y = 0.0 + (i + 0.5) * dx[0], 0.0 + (j + 0.5) * dx[1]           # mesh.rs:69-70
            psi[at(i, j)] = -h_far[0] / dx[0] * x - h_far[1] / dx[1] * y      # main.rs:50-53
    fill_neumann(psi, grad)
    dx2, dy2 = dx[0] * dx[0], dx[1] * dx[1]
    alpha, beta, gamma = dx2 * dy2 / (2.0 * (dx2 + dy2)), dy2 / (2.0 * (dx2 + dy2)), dx2 / (2.0 * (dx2 + dy2))
    trace, sum_diff, it = [], 1.0, 0
    while it < n_iter and sum_diff > tol:                    # main.rs:65
        lap = [0.0] * (G * G)                                # poisson.rs:13-36
        for j in range(n):
            for i in range(n):
                lap[at(i, j)] = (psi[at(i + 1, j)] * (mu(i + 1, j) + mu(i, j)) / 2.0
                                 + psi[at(i - 1, j)] * (mu(i - 1, j) + mu(i, j)) / 2.0
                                 + psi[at(i, j + 1)] * (mu(i, j + 1) + mu(i, j)) / 2.0
                                 + psi[at(i, j - 1)] * (mu(i, j - 1) + mu(i, j)) / 2.0
                                 - (psi[at(i, j)] * (mu(i + 1, j) + mu(i - 1, j) + mu(i, j + 1) + mu(i, j - 1) + 4.0 * mu(i, j))) / 2.0)
        for t in range(n):                                   # Dirichlet(0) fill of lap, dataframe.rs:324-355
            lap[at(-1, t)] = 2.0 * 0.0 - lap[at(0, t)]
            lap[at(n, t)] = 2.0 * 0.0 - lap[at(n - 1, t)]
        for t in range(n):
            lap[at(t, -1)] = 2.0 * 0.0 - lap[at(t, 0)]
            lap[at(t, n)] = 2.0 * 0.0 - lap[at(t, n - 1)]
        src = [0.0] * (G * G)                                # poisson.rs:38-44
        for j in range(n):
            for i in range(n):
                src[at(i, j)] = -(1.0 - 1.0 / relax) / mu(i, j)
        rhs = [a * b for a, b in zip(src, lap)]
        star = list(psi)                                     # poisson.rs:71
        for _ in range(n_sub):                               # poisson.rs:80-89
            for i in range(n):
                for j in range(n):
                    star[at(i, j)] = (alpha * rhs[at(i, j)] + beta * (star[at(i + 1, j)] + star[at(i - 1, j)])
                                      + gamma * (star[at(i, j + 1)] + star[at(i, j - 1)]))
            fill_neumann(star, grad)
        diff = [a + (-1.0 * b) for a, b in zip(star, psi)]   # main.rs:68
        sum_diff = 0.0
        for j in range(n):                                   # poisson.rs:94-100, flat order
            for i in range(n):
                sum_diff += abs(diff[at(i, j)])
        psi = [a + relax * b for a, b in zip(psi, diff)]     # main.rs:70
        fill_neumann(psi, grad)                              # main.rs:71
        trace.append(sum_diff)
        it += 1

    g = ora.geom([n, n], hi=[L, L])
    m = ora.model(h_far=h_far, relax=relax, n_iter=n_iter, n_sub_iter=n_sub, bubble_centre=centre, bubble_radius=radius, mu=mus, tol=tol)
    got, tr, its = ora.zeros(g), np.zeros(n_iter), C.c_int64()
    assert ora.lib().ora_magnetostatics_run(C.byref(g), got, C.byref(m), 0, tr, C.byref(its)) == 0
    assert its.value == it and tr[:it].tolist() == trace
    np.testing.assert_array_equal(got, np.array(psi))
    assert len({mu(i, j) for i in range(n) for j in range(n)}) == 2      # the droplet really is inside the mesh
