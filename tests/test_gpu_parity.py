"""GPU parity tests: every kernel behind the C ABI against the CPU oracle on the same inputs.
Bar: bit-exact for fill_bc / operators / Jacobi sweeps / variable-mu operator / reference-order GS
(the kernels use explicit round-to-nearest mul/add, no FMA, same association as the reference);
norms within 1e-12 relative of a long-double oracle sum (tree order differs from the serial sum).
All tests call through librnmf_b200.so via ctypes."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as ora

pytestmark = pytest.mark.gpu

NORM_RTOL = 1e-12      # stated tolerance for reductions (summation order differs)


@pytest.fixture(scope="module")
def R():
    import rnmf_b200
    return rnmf_b200


@pytest.fixture(scope="module")
def ctx(R):
    c = R.Context(0)
    yield c
    c.close()


def _bc_pair(R, dim, spec):
    """spec: (lo, hi) lists of (kind, value) -> (oracle faces, rnmf BCs) for ONE component"""
    mk = {"dirichlet": R.BcType.Dirichlet, "neumann": R.BcType.Neumann,
          "prescribed": R.BcType.Prescribed, "reflect": lambda v: R.BcType.Reflect()}
    obc = ora.bcs(dim, [spec])
    bc = R.BCs([R.ComponentBCs([mk[k](v) for k, v in spec[0]], [mk[k](v) for k, v in spec[1]])])
    return obc, bc


def _frames(R, ctx, n, hi, spec, ng=1, seed=0):
    dim = len(n)
    g = ora.geom(n, hi=hi, ng=ng)
    obc, bc = _bc_pair(R, dim, spec)
    mesh = R.CartesianMesh.new([0.0] * dim, hi, n)
    rng = np.random.default_rng(seed)
    psi0 = rng.standard_normal(g.n_nodes)
    rhs0 = rng.standard_normal(g.n_nodes)
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, ng, ctx)
    rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, ng, ctx)
    psi.data = psi0
    rhs.data = rhs0
    return g, obc, psi0, rhs0, psi, rhs


MIXED_2D = ([("neumann", 0.3), ("dirichlet", -1.0)], [("dirichlet", 2.0), ("neumann", -0.7)])
DIRICHLET_2D = ([("dirichlet", 0.0), ("dirichlet", 0.0)], [("dirichlet", 1.0), ("dirichlet", 1.0)])   # src/main.rs:20-28
MIXED_3D = ([("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 0.25)],
            [("dirichlet", 1.0), ("neumann", -0.5), ("neumann", 0.5)])


# ---- ghost fill: the reference's own golden vectors, through the GPU ---------------------------
def test_fill_bc_golden_vectors_on_gpu(R, ctx, golden):
    for case in golden["bc_fill"]:
        pres = case.get("prescribed_values")
        fix = lambda f: (f[0], pres if f[0] == "prescribed" else f[1])
        spec = ([fix(f) for f in case["lo"]], [fix(f) for f in case["hi_bc"]])
        _, bc = _bc_pair(R, 2, spec)
        mesh = R.CartesianMesh.new([0.0, 0.0], case["hi"], case["n"])
        f = R.CartesianDataFrame.new_from(mesh, bc, 1, case["ng"], ctx)
        if case["ic"] == "x+1":
            f.fill_ic(lambda x, y, c: x + 1.0)
        f.fill_bc()
        np.testing.assert_array_equal(f.data, np.array(case["data"]), err_msg=case["name"])


def test_valid_cells_order_on_gpu(R, ctx, golden):
    case = golden["data_frame_iterator"]
    mesh = R.CartesianMesh.new([0.0, 0.0], case["hi"], case["n"])
    f = R.CartesianDataFrame.new_from(mesh, R.BCs.empty(), 1, case["ng"], ctx)
    f.fill_ic(lambda x, y, c: x + y)
    assert f.get_valid_cells().tolist() == case["valid_order"]
    assert f[(0, 0, 0)] == 2.0 and f[(2, 2, 0)] == 10.0 and f[(-1, -1, 0)] == 0.0


def test_reflect_is_unimplemented_on_gpu(R, ctx):
    _, bc = _bc_pair(R, 2, ([("reflect", None), ("dirichlet", 0.0)], [("dirichlet", 0.0)] * 2))
    f = R.CartesianDataFrame.new_from(R.CartesianMesh.new([0, 0], [6, 6], [3, 3]), bc, 1, 1, ctx)
    with pytest.raises(R.RnmfError) as e:
        f.fill_bc()
    assert e.value.code == -4


# Quirk Q1 (declared in oracle/rnmf_oracle.c and DESIGN.md section 2): the reference's fillers loop the TANGENTIAL index over
# 0..n[i_dim] instead of the other dimension's extent (dataframe.rs:290,308,325,341), which is wrong on non-square meshes.  The
# oracle and the library fill the INTENDED tangential extent.  On the square / cubic meshes of every BASELINE config and of the
# reference's own golden vectors the two agree; the non-square shapes below ([7,5], [5,4,3], [33,1,2], ...) therefore pin the
# intended behaviour, not the reference's bug.
@pytest.mark.parametrize("n,hi,ng,spec", [
    ([7, 5], [7.0, 10.0], 1, MIXED_2D),
    ([16, 16], [4.0, 4.0], 3, MIXED_2D),
    ([1, 1], [1.0, 1.0], 1, MIXED_2D),
    ([5, 4, 3], [5.0, 8.0, 9.0], 1, MIXED_3D),
    ([6, 7, 8], [1.0, 2.0, 3.0], 2, MIXED_3D),
    ([33, 1, 2], [1.0, 2.0, 3.0], 1, MIXED_3D),
])
def test_fill_bc_random_frames(R, ctx, n, hi, ng, spec):
    g, obc, psi0, _, psi, _ = _frames(R, ctx, n, hi, spec, ng=ng, seed=11)
    want = psi0.copy()
    assert ora.fill_bc(g, want, obc) == 0
    psi.fill_bc()
    np.testing.assert_array_equal(psi.data, want)


def test_fill_bc_multicomponent_and_prescribed_3d(R, ctx):
    n, hi, ng = [4, 3, 5], [4.0, 3.0, 5.0], 2
    g = ora.geom(n, hi=hi, ng=ng, ncomp=2)
    vals = np.arange(1.0, 400.0)
    comp0 = MIXED_3D
    comp1 = ([("prescribed", vals), ("neumann", 1.5), ("prescribed", vals[::-1])],
             [("dirichlet", -2.0), ("prescribed", vals * 0.5), ("prescribed", vals[:7])])
    obc = ora.bcs(3, [comp0, comp1])
    mk = {"dirichlet": R.BcType.Dirichlet, "neumann": R.BcType.Neumann, "prescribed": R.BcType.Prescribed}
    bc = R.BCs([R.ComponentBCs([mk[k](v) for k, v in c[0]], [mk[k](v) for k, v in c[1]]) for c in (comp0, comp1)])
    f = R.CartesianDataFrame.new_from(R.CartesianMesh.new([0.0] * 3, hi, n), bc, 2, ng, ctx)
    host = np.random.default_rng(5).standard_normal(g.n_nodes)
    f.data = host
    want = host.copy()
    assert ora.fill_bc(g, want, obc) == 0
    f.fill_bc()
    np.testing.assert_array_equal(f.data, want)


# ---- Jacobi sweeps (K1) --------------------------------------------------------------------------
# (non-square cases: quirk Q1 above applies to the fused ghost fill of the sweeps as well)
SWEEP_CASES_2D = [
    ([301, 301], [301.0, 301.0], MIXED_2D, 4),
    ([256, 256], [256.0, 256.0], DIRICHLET_2D, 3),
    ([255, 130], [3.0, 5.0], MIXED_2D, 3),       # odd width, anisotropic dx
    ([2, 3], [2.0, 3.0], MIXED_2D, 5),
    ([1, 1], [1.0, 1.0], MIXED_2D, 3),
    ([1031, 67], [1.0, 1.0], DIRICHLET_2D, 2),
    ([1500, 300], [15.0, 3.0], MIXED_2D, 3),     # several 512-column tiles and row chunks of the TMA kernel
]


@pytest.mark.parametrize("n,hi,spec,sweeps", SWEEP_CASES_2D)
@pytest.mark.parametrize("variant", [0, 20, 30])      # kernels_sweep.cu: the table of sweep variants
def test_jacobi_2d_bit_exact(R, ctx, n, hi, spec, sweeps, variant):
    from rnmf_b200 import poisson
    ctx.set_option("sweep_variant", variant)
    ctx.set_option("resident", 0)          # exercise the streaming kernels themselves
    try:
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=21)
        ora.fill_bc(g, psi0, obc)
        psi.fill_bc()
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=True)
        want = psi0.copy()
        l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
        np.testing.assert_array_equal(psi.data, want)
        assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
        assert psi.guards_intact() and rhs.guards_intact()
    finally:
        ctx.set_option("sweep_variant", "auto")
        ctx.set_option("resident", 1)


SWEEP_CASES_3D = [
    ([64, 64, 64], [64.0, 128.0, 192.0], MIXED_3D, 3),    # anisotropic hi=(n,2n,3n)
    ([65, 33, 17], [1.0, 1.0, 1.0], MIXED_3D, 4),         # odd sizes
    ([130, 9, 70], [130.0, 9.0, 70.0], MIXED_3D, 2),
    ([2, 2, 2], [2.0, 2.0, 2.0], MIXED_3D, 5),
    ([1, 1, 1], [1.0, 1.0, 1.0], MIXED_3D, 3),
    ([31, 40, 3], [1.0, 2.0, 3.0], MIXED_3D, 3),
    ([300, 70, 80], [3.0, 2.0, 1.0], MIXED_3D, 3),       # several x tiles of the TMA variant, ragged edges
]


@pytest.mark.parametrize("n,hi,spec,sweeps", SWEEP_CASES_3D)
@pytest.mark.parametrize("variant", [0, 31, 6])      # three (even n0, n1; otherwise two) / two / one sweep per launch
def test_jacobi_3d_bit_exact(R, ctx, n, hi, spec, sweeps, variant):
    from rnmf_b200 import poisson
    ctx.set_option("sweep_variant", variant)
    ctx.set_option("resident", 0)          # exercise the streaming kernels themselves
    try:
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=22)
        ora.fill_bc(g, psi0, obc)
        psi.fill_bc()
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=True)
        want = psi0.copy()
        l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
        np.testing.assert_array_equal(psi.data, want)
        assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
        assert psi.guards_intact() and rhs.guards_intact()       # no store past the end of a buffer
    finally:
        ctx.set_option("sweep_variant", "auto")
        ctx.set_option("resident", 1)


@pytest.mark.parametrize("n,hi,spec,sweeps,norm", [
    ([301, 301], [301.0, 301.0], MIXED_2D, 6, False),        # three double sweeps
    ([1500, 300], [15.0, 3.0], MIXED_2D, 7, False),          # three 512-column strips, several row chunks; + one single
    ([1031, 67], [1.0, 1.0], DIRICHLET_2D, 8, True),         # 3 double + 2 single; last valid cell is the first of a pair
    ([512, 40], [512.0, 40.0], MIXED_2D, 4, False),          # exactly one strip
    ([513, 9], [1.0, 2.0], MIXED_2D, 5, True),               # one cell past a strip edge
    ([255, 1], [3.0, 5.0], MIXED_2D, 4, False),              # a single row: both y ghosts in registers
    ([2, 3], [2.0, 3.0], MIXED_2D, 9, True),
    ([1, 1], [1.0, 1.0], MIXED_2D, 6, False),
    ([2048, 700], [1.0, 1.0], DIRICHLET_2D, 6, False),
])
def test_jacobi_2d_double_sweep_bit_exact(R, ctx, n, hi, spec, sweeps, norm):
    """kernels_sweep2d_tb2.cu: two 2D sweeps (+ ghost fills) per pass over HBM == two single sweeps, bit for bit."""
    from rnmf_b200 import poisson
    ctx.set_option("sweep_variant", 30)
    ctx.set_option("resident", 0)
    try:
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=31)
        ora.fill_bc(g, psi0, obc)
        psi.fill_bc()
        d0 = ctx.double_sweep_launches
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=norm)
        ctx.sync()
        assert ctx.double_sweep_launches - d0 == (sweeps - (1 if norm else 0)) // 2
        want = psi0.copy()
        l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
        np.testing.assert_array_equal(psi.data, want)
        if norm:
            assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
        assert psi.guards_intact() and rhs.guards_intact()
    finally:
        ctx.set_option("sweep_variant", "auto")
        ctx.set_option("resident", 1)


# four sweeps per pass (kernels_sweep2d_tbk.cu): the 2D default for groups of four sweeps; the rest runs as pairs / singles
@pytest.mark.parametrize("n,hi,spec,sweeps,norm", [
    ([301, 301], [301.0, 301.0], MIXED_2D, 12, False),
    ([1500, 300], [15.0, 3.0], MIXED_2D, 13, False),         # several strips and row chunks; leftover sweeps as pairs / singles
    ([1031, 67], [1.0, 1.0], DIRICHLET_2D, 9, True),         # last valid cell is the first of a pair; the last sweep carries the norm
    ([515, 40], [515.0, 40.0], MIXED_2D, 8, False),          # a 3-column second strip: x-hi ghosts formed by the halo lanes
    ([255, 1], [3.0, 5.0], MIXED_2D, 8, False),              # a single row: every level meets both y faces
    ([2, 3], [2.0, 3.0], MIXED_2D, 9, True),
    ([2048, 700], [1.0, 1.0], DIRICHLET_2D, 12, False),
])
def test_jacobi_2d_multi_sweep_bit_exact(R, ctx, n, hi, spec, sweeps, norm):
    """kernels_sweep2d_tbk.cu: four 2D sweeps (+ ghost fills) per pass over HBM == four single sweeps, bit for bit."""
    from rnmf_b200 import poisson
    K = 4
    ctx.set_option("sweep_variant", "auto")
    ctx.set_option("resident", 0)
    try:
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=33)
        ora.fill_bc(g, psi0, obc)
        psi.fill_bc()
        k0, d0, m0 = ctx.kernel_launches, ctx.double_sweep_launches, ctx.multi_sweep_stats
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=norm)
        ctx.sync()
        pairable = sweeps - (1 if norm else 0)
        groups, rest = divmod(pairable, K)
        assert ctx.double_sweep_launches - d0 == rest // 2             # leftovers run as pairs, then singles
        m1 = ctx.multi_sweep_stats
        assert (m1[0] - m0[0], m1[1] - m0[1]) == (groups + rest // 2, groups * K + 2 * (rest // 2))
        # ghost-shell copy + K-sweep launches + pairs + singles (+ the norm's reduction): the K-sweep kernel really ran
        assert ctx.kernel_launches - k0 == 1 + groups + rest // 2 + rest % 2 + (2 if norm else 0)
        want = psi0.copy()
        l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
        np.testing.assert_array_equal(psi.data, want)
        if norm:
            assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
        assert psi.guards_intact() and rhs.guards_intact()
    finally:
        ctx.set_option("sweep_variant", "auto")
        ctx.set_option("resident", 1)


ALL_NEUMANN_3D = ([("neumann", 0.3), ("neumann", -0.2), ("neumann", 0.1)], [("neumann", 0.4), ("neumann", 0.6), ("neumann", -0.9)])
@pytest.mark.parametrize("n,hi,spec,sweeps,norm", [
    ([64, 64, 64], [64.0, 128.0, 192.0], MIXED_3D, 6, False),     # three double sweeps
    ([65, 33, 17], [1.0, 1.0, 1.0], MIXED_3D, 7, False),          # three double sweeps + one single
    ([65, 33, 17], [1.0, 1.0, 1.0], ALL_NEUMANN_3D, 8, True),     # three double + two single (the last carries the norm)
    ([300, 70, 80], [3.0, 2.0, 1.0], MIXED_3D, 5, True),          # x/y/z tiling with ragged edges
    ([129, 13, 5], [1.0, 2.0, 3.0], ALL_NEUMANN_3D, 4, False),    # one cell / one row past a tile edge
    ([128, 12, 4], [1.0, 2.0, 3.0], MIXED_3D, 2, False),          # exactly one tile
    ([127, 25, 1], [1.0, 2.0, 3.0], MIXED_3D, 4, False),          # a single plane: both z ghosts in registers
    ([1, 1, 1], [1.0, 1.0, 1.0], MIXED_3D, 6, False),
    ([260, 30, 2], [1.0, 1.0, 1.0], ALL_NEUMANN_3D, 9, True),
    ([520, 40, 70], [5.0, 2.0, 1.0], MIXED_3D, 6, False),         # five x tiles: warps with no x / y face run the interior-specialised loop
    ([200, 37, 150], [2.0, 1.0, 3.0], MIXED_3D, 4, False),        # two 96-plane chunks; odd n1
])
def test_jacobi_3d_double_sweep_bit_exact(R, ctx, n, hi, spec, sweeps, norm):
    """kernels_sweep_tb2.cu: two sweeps (+ the ghost fill after each) per pass over HBM must equal two
    single sweeps bit for bit, for every tiling edge case and both parities of the sweep count."""
    from rnmf_b200 import poisson
    ctx.set_option("sweep_variant", 31)
    ctx.set_option("resident", 0)
    try:
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=29)
        ora.fill_bc(g, psi0, obc)
        psi.fill_bc()
        before = ctx.kernel_launches
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=norm)
        ctx.sync()
        pairs = (sweeps - (1 if norm else 0)) // 2
        assert ctx.kernel_launches - before <= 1 + pairs + (sweeps - 2 * pairs) + 1      # shell copy, doubles, singles, reduce
        want = psi0.copy()
        l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
        np.testing.assert_array_equal(psi.data, want)
        if norm:
            assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
        assert psi.guards_intact() and rhs.guards_intact()
    finally:
        ctx.set_option("sweep_variant", "auto")
        ctx.set_option("resident", 1)


@pytest.mark.parametrize("n,hi,spec,sweeps,norm", [
    ([64, 64, 64], [64.0, 128.0, 192.0], MIXED_3D, 6, False),     # two triple sweeps
    ([66, 34, 17], [1.0, 1.0, 1.0], MIXED_3D, 7, False),          # two triples + one single
    ([66, 34, 17], [1.0, 1.0, 1.0], ALL_NEUMANN_3D, 9, True),     # two triples + a pair + the norm-carrying single
    ([300, 70, 80], [3.0, 2.0, 1.0], MIXED_3D, 4, True),          # x/y/z tiling with ragged edges
    ([130, 14, 5], [1.0, 2.0, 3.0], ALL_NEUMANN_3D, 6, False),    # one pair / one row pair past a tile edge
    ([128, 12, 4], [1.0, 2.0, 3.0], MIXED_3D, 3, False),          # exactly one tile
    ([126, 26, 1], [1.0, 2.0, 3.0], MIXED_3D, 6, False),          # a single plane: every level meets both z faces
    ([2, 2, 1], [1.0, 1.0, 1.0], MIXED_3D, 6, False),
    ([260, 30, 2], [1.0, 1.0, 1.0], ALL_NEUMANN_3D, 10, True),
    ([520, 40, 70], [5.0, 2.0, 1.0], MIXED_3D, 6, False),         # five x tiles: warps with no x / y face run the interior-specialised loop
    ([200, 38, 300], [2.0, 1.0, 3.0], MIXED_3D, 3, False),        # several chunks per launch
    ([768, 24, 40], [7.0, 1.0, 3.0], MIXED_3D, 6, False),         # the benchmark's row length
])
def test_jacobi_3d_triple_sweep_bit_exact(R, ctx, n, hi, spec, sweeps, norm):
    """kernels_sweep_tb3.cu (the 3D default for frames with even n0, n1): three sweeps (+ the ghost fill after each) per pass over
    HBM must equal three single sweeps bit for bit; leftovers run as a pair (kernels_sweep_tb2.cu), then singles."""
    from rnmf_b200 import poisson
    ctx.set_option("sweep_variant", "auto")
    ctx.set_option("resident", 0)
    try:
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=33)
        ora.fill_bc(g, psi0, obc)
        psi.fill_bc()
        d0, m0 = ctx.double_sweep_launches, ctx.multi_sweep_stats
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=norm)
        ctx.sync()
        groupable = sweeps - (1 if norm else 0)
        m1, pairs = ctx.multi_sweep_stats, ctx.double_sweep_launches - d0        # the stats count all temporally blocked launches
        assert pairs == (groupable % 3) // 2
        assert (m1[0] - m0[0] - pairs, m1[1] - m0[1] - 2 * pairs) == (groupable // 3, 3 * (groupable // 3)), "the three-sweep kernel did not run"
        want = psi0.copy()
        l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
        np.testing.assert_array_equal(psi.data, want)
        if norm:
            assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
        assert psi.guards_intact() and rhs.guards_intact()
    finally:
        ctx.set_option("sweep_variant", "auto")
        ctx.set_option("resident", 1)


@pytest.mark.parametrize("n,hi,spec,sweeps", [
    ([301, 301], [301.0, 301.0], MIXED_2D, 57), ([301, 301], [301.0, 301.0], MIXED_2D, 24),
    ([7, 5], [7.0, 5.0], MIXED_2D, 9), ([1, 1], [1.0, 1.0], MIXED_2D, 6),
    ([65, 33, 17], [1.0, 2.0, 3.0], MIXED_3D, 21), ([64, 64, 64], [64.0, 64.0, 64.0], MIXED_3D, 12),
])
def test_jacobi_resident_multi_sweep_kernel(R, ctx, n, hi, spec, sweeps):
    """Small (L2-resident) frames run all sweeps in one cooperative launch (kernels_resident.cu);
    with and without the norm-carrying last sweep, odd and even sweep counts."""
    from rnmf_b200 import poisson
    g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, seed=24)
    ora.fill_bc(g, psi0, obc)
    psi.fill_bc()
    before = ctx.kernel_launches
    l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=True)
    ctx.sync()
    assert ctx.kernel_launches - before <= 6                 # shell copy + resident + last sweep + reduce
    want = psi0.copy()
    l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps)
    np.testing.assert_array_equal(psi.data, want)
    assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref)
    poisson.jacobi_sweeps(psi, rhs, sweeps + 1)               # no norm: every sweep in the resident kernel
    ora.jacobi(g, want, rhs0, obc, sweeps + 1)
    np.testing.assert_array_equal(psi.data, want)
    assert psi.guards_intact()


def test_jacobi_two_ghost_layers_and_unfilled_ghosts(R, ctx):
    # ng=2 takes the stand-alone fill path; ghosts deliberately NOT pre-filled (sweep 0 reads stored ghosts)
    from rnmf_b200 import poisson
    for n, hi, spec in (([40, 24], [40.0, 24.0], MIXED_2D), ([20, 12, 9], [20.0, 12.0, 9.0], MIXED_3D)):
        g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, spec, ng=2, seed=23)
        poisson.jacobi_sweeps(psi, rhs, 3)
        want = psi0.copy()
        ora.jacobi(g, want, rhs0, obc, 3)
        np.testing.assert_array_equal(psi.data, want)


def test_jacobi_linearity_and_fixed_point_at_size(R, ctx):
    """Size-independent properties on a grid too big for the oracle to be quick:
    (a) a field linear in x,y,z with matching Neumann/Dirichlet data is a fixed point of the sweep;
    (b) sweep(u+v; rhs_u+rhs_v) == sweep(u)+sweep(v) up to rounding for homogeneous BCs."""
    from rnmf_b200 import poisson
    n = [192, 160, 128]
    hi = [float(v) for v in n]
    # (a) u = 2x - 3y + 0.5z: Neumann gradients per face, rhs = 0
    spec = ([("neumann", 2.0), ("neumann", -3.0), ("neumann", 0.5)], [("neumann", 2.0), ("neumann", -3.0), ("neumann", 0.5)])
    _, bc = _bc_pair(R, 3, spec)
    mesh = R.CartesianMesh.new([0.0] * 3, hi, n)
    u = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    z = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    u.fill_ic(lambda x, y, zz, c: 2.0 * x - 3.0 * y + 0.5 * zz)
    u.fill_bc()
    before = u.get_valid_cells()
    l1 = poisson.jacobi_sweeps(u, z, 6, want_norm=True)
    after = u.get_valid_cells()
    assert np.abs(after - before).max() <= 1e-9 and l1 <= 1e-9 * before.size
    # (b) linearity with homogeneous Dirichlet
    spec0 = ([("dirichlet", 0.0)] * 3, [("dirichlet", 0.0)] * 3)
    _, bc0 = _bc_pair(R, 3, spec0)
    fr = [R.CartesianDataFrame.new_from(mesh, bc0, 1, 1, ctx) for _ in range(6)]
    a, ra, b, rb, s, rs = fr
    a.fill_synthetic(1); ra.fill_synthetic(2); b.fill_synthetic(3); rb.fill_synthetic(4)
    s.axpby_(1.0, a, 1.0, b); rs.axpby_(1.0, ra, 1.0, rb)
    for f in (a, b, s):
        f.fill_bc()
    for f, r in ((a, ra), (b, rb), (s, rs)):
        poisson.jacobi_sweeps(f, r, 4)
    lhs = s.get_valid_cells()
    rhs_ = a.get_valid_cells() + b.get_valid_cells()
    assert np.abs(lhs - rhs_).max() <= 1e-13 * max(1.0, np.abs(rhs_).max()) * 8


# ---- reference-order Gauss-Seidel (K6) ------------------------------------------------------------
@pytest.mark.parametrize("n,hi,sweeps", [([37, 23], [3.0, 5.0], 13), ([64, 64], [64.0, 64.0], 40),
                                         ([1, 9], [1.0, 2.0], 4), ([301, 301], [301.0, 301.0], 5)])
def test_gs_reference_order_bit_exact(R, ctx, n, hi, sweeps):
    from rnmf_b200 import _capi, poisson
    g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, MIXED_2D, seed=31)
    coef = (C.c_double * 4)(*poisson.poisson_coefs(2, psi.underlying_mesh.dx))
    _capi.check(_capi.lib().rnmf_gs_sweeps(psi._h, rhs._h, coef, sweeps))
    want = psi0.copy()
    ora.gs(g, want, rhs0, obc, sweeps)          # the serial lexicographic loop of poisson.rs:80-89
    np.testing.assert_array_equal(psi.data, want)


# ---- operators, norm, variable-mu operator, get_h -----------------------------------------------
def test_operators_bit_exact(R, ctx):
    g, _, x0, y0, x, y = _frames(R, ctx, [33, 20], [33.0, 20.0], MIXED_2D, ng=2, seed=41)
    np.testing.assert_array_equal((-1.5 * x).data, -1.5 * x0)
    np.testing.assert_array_equal((x + y).data, x0 + y0)
    np.testing.assert_array_equal((x * y).data, x0 * y0)
    out = x.clone()
    out.axpby_(1.0, x, 0.2, y)
    np.testing.assert_array_equal(out.data, x0 + 0.2 * y0)
    out.axpby_(1.0, x, -1.0, y)
    np.testing.assert_array_equal(out.data, x0 + (-1.0 * y0))


@pytest.mark.parametrize("n,ng,ncomp", [([301, 301], 1, 1), ([17, 5], 2, 1), ([9, 8, 7], 1, 1), ([64, 48, 40], 2, 1)])
def test_abs_sum_valid(R, ctx, n, ng, ncomp):
    dim = len(n)
    g = ora.geom(n, ng=ng)
    mesh = R.CartesianMesh.new([0.0] * dim, [float(v) for v in n], n)
    f = R.CartesianDataFrame.new_from(mesh, R.BCs.empty(), 1, ng, ctx)
    host = np.random.default_rng(51).standard_normal(g.n_nodes) * 1e3
    f.data = host
    want = ora.lib().ora_sum_abs_valid_ld(C.byref(g), host)
    got = f.abs_sum()
    assert abs(got - want) <= NORM_RTOL * want
    naive = ora.lib().ora_sum_abs_valid(C.byref(g), host)      # the reference's serial order: within N*eps
    assert abs(got - naive) <= g.n_nodes * np.finfo(float).eps * want
    assert f.abs_sum() == got                                    # deterministic


def test_varcoef_laplace_and_source_bit_exact(R, ctx):
    from rnmf_b200 import poisson
    for n, hi, centre, radius in (([301, 301], [301.0, 301.0], (150.5, 150.5), 15.05), ([40, 24], [10.0, 6.0], (4.0, 3.0), 2.0)):
        g = ora.geom(n, hi=hi)
        m = ora.model(bubble_centre=centre, bubble_radius=radius)
        cfg = poisson.Config(poisson.GeomConfig(2, tuple(hi), tuple(n)),
                             poisson.UserModel(relax=0.2, bubble_centre=centre, bubble_radius=radius, mu=(3.2, 1.0)))
        _, bc = _bc_pair(R, 2, MIXED_2D)
        phi = R.CartesianDataFrame.new_from(R.CartesianMesh.new([0.0, 0.0], hi, n), bc, 1, 1, ctx)
        host = np.random.default_rng(61).standard_normal(g.n_nodes)
        phi.data = host
        lap, src = ora.zeros(g), ora.zeros(g)
        ora.lib().ora_get_laplace_2d(C.byref(g), host, C.byref(m), lap)
        ora.lib().ora_get_source_2d(C.byref(g), host, C.byref(m), src)
        np.testing.assert_array_equal(poisson.get_laplace(phi, cfg).data, lap)
        np.testing.assert_array_equal(poisson.get_source(phi, cfg).data, src)


def test_get_h_bit_exact(R, ctx):
    from rnmf_b200 import _capi
    n, hi = [30, 20], [15.0, 40.0]
    g = ora.geom(n, hi=hi)
    g2 = ora.geom(n, hi=hi, ncomp=2)
    _, bc = _bc_pair(R, 2, MIXED_2D)
    mesh = R.CartesianMesh.new([0.0, 0.0], hi, n)
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    h = R.CartesianDataFrame.new_from(mesh, R.BCs.empty(), 2, 1, ctx)
    host = np.random.default_rng(71).standard_normal(g.n_nodes)
    psi.data = host
    want = np.zeros(g2.n_nodes)
    ora.lib().ora_get_h_2d(C.byref(g), host, want)
    _capi.check(_capi.lib().rnmf_gradient_h_2d(psi._h, h._h))
    np.testing.assert_array_equal(h.data, want)


# ---- the shipped case end to end -----------------------------------------------------------------
def test_magnetostatics_small_case_reference_order(R, ctx):
    """BASELINE configs[0]: the 10x10 test.lua geometry with the magnetostatics model block, plus
    a 33x33 droplet; reference update order => bit-exact field, norms within 1e-12."""
    from rnmf_b200 import poisson
    for n, centre, radius, n_iter, sub in ((10, (5.0, 5.0), 0.25, 3, 20), (33, (16.5, 16.5), 4.0, 4, 25)):
        g = ora.geom([n, n], hi=[float(n)] * 2)
        m = ora.model(n_iter=n_iter, n_sub_iter=sub, bubble_centre=centre, bubble_radius=radius)
        ref, tr, it = ora.zeros(g), np.zeros(n_iter), C.c_int64()
        ora.lib().ora_magnetostatics_run(C.byref(g), ref, C.byref(m), 0, tr, C.byref(it))
        cfg = poisson.Config(poisson.GeomConfig(2, (float(n),) * 2, (n, n)),
                             poisson.UserModel((0.0, 1.0), 0.2, n_iter, sub, centre, radius, (3.2, 1.0), 0.1))
        psi, trace, iters = poisson.magnetostatics(cfg, order="reference", ctx=ctx)
        assert iters == it.value
        np.testing.assert_array_equal(psi.data, ref)
        for a, b in zip(trace, tr[:iters]):
            assert abs(a - b) <= NORM_RTOL * abs(b) if b else a == 0.0


def test_magnetostatics_lua_case_full(R, ctx):
    """examples/magnetostatics/magnetostatics.lua verbatim (301x301, 10 x 550 sweeps) in the
    reference's update order: the sum_diff trajectory of the oracle (tests/golden/magnetostatics_trace.json)
    to 1e-12, and psi(150,150) bit-exact."""
    from rnmf_b200 import poisson
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "magnetostatics_trace.json")) as fh:
        gold = json.load(fh)
    MAGNETOSTATICS_TRACE, MAGNETOSTATICS_PSI_150_150 = gold["sum_diff"], gold["psi_150_150"]
    from rnmf_b200 import config
    # the .lua case itself, unedited, through the Lua-subset reader (config.rs:90-158 mirror)
    cfg = config.read_lua(os.path.join(os.path.dirname(os.path.dirname(__file__)), "cases", "ferro_droplet.lua"))
    assert cfg == poisson.Config(poisson.GeomConfig(2, (301.0, 301.0), (301, 301)),
                                 poisson.UserModel((0.0, 1.0), 0.2, 10, 550, (301.0 / 2, 301.0 / 2), 301.0 / 20, (3.2, 1.0), 0.1))
    psi, trace, iters = poisson.magnetostatics(cfg, order="reference", ctx=ctx)
    assert iters == 10
    for a, b in zip(trace, MAGNETOSTATICS_TRACE):
        assert abs(a - b) <= NORM_RTOL * b
    assert psi[(150, 150, 0)] == MAGNETOSTATICS_PSI_150_150


def test_magnetostatics_jacobi_vs_reference_converged_field(R, ctx):
    """Where the update order cannot be kept (Jacobi throughput path) the CONVERGED field is
    compared on a gauge-free quantity: all-Neumann => psi is defined up to a constant, so compare
    H = -grad(psi) (model.rs:72-88).  Stated tolerance: 1e-6 relative to max|H| after both
    orders have converged on a 24x24 droplet."""
    from rnmf_b200 import _capi, poisson
    n = 24
    cfg = poisson.Config(poisson.GeomConfig(2, (float(n),) * 2, (n, n)),
                         poisson.UserModel((0.0, 1.0), 0.2, 60, 4000, (12.0, 12.0), 3.0, (3.2, 1.0), 1e-9))
    fields = []
    for order in ("reference", "jacobi"):
        psi, trace, iters = poisson.magnetostatics(cfg, order=order, ctx=ctx)
        h = R.CartesianDataFrame.new_from(psi.underlying_mesh, R.BCs.empty(), 2, 1, ctx)
        _capi.check(_capi.lib().rnmf_gradient_h_2d(psi._h, h._h))
        fields.append(h.get_valid_cells())
    scale = np.abs(fields[0]).max()
    assert np.abs(fields[0] - fields[1]).max() <= 1e-6 * scale


def test_solve_poisson_host_entry_point(R, ctx):
    """The reference-facing call with HOST frames in the reference's dense layout (poisson.rs:67)."""
    from rnmf_b200 import _capi
    n, hi = [48, 40, 36], [48.0, 40.0, 36.0]
    g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, n, hi, MIXED_3D, seed=81)
    ora.fill_bc(g, psi0, obc)
    _, bc = _bc_pair(R, 3, MIXED_3D)
    from rnmf_b200.dataframe import _faces_array
    faces = _faces_array(bc, 3, 1)
    out = np.empty_like(psi0)
    l1 = C.c_double()
    nn = (C.c_int64 * 3)(*n)
    lo = (C.c_double * 3)(0.0, 0.0, 0.0)
    dx = (C.c_double * 3)(*[g.dx[d] for d in range(3)])
    _capi.check(_capi.lib().rnmf_solve_poisson_host(ctx._h, 3, nn, lo, dx, 1, faces, psi0.ctypes.data, rhs0.ctypes.data,
                                                    6, 0, out.ctypes.data, C.byref(l1)))
    want = psi0.copy()
    l1_ref = ora.jacobi(g, want, rhs0, obc, 6)
    np.testing.assert_array_equal(out, want)
    assert abs(l1.value - l1_ref) <= NORM_RTOL * l1_ref


@pytest.mark.parametrize("sweeps,chunks,pairs", [(41, 4, 10), (24, 8, 56), (19, 2, 4)])
def test_solve_poisson_host_time_skewed_start(R, ctx, sweeps, chunks, pairs):
    """Option e2e_skew (default 2; 0 = plain): chunked upload overlapped with the first double sweeps on the bands each chunk makes computable
    (capi.cu: skewed_upload_and_first_sweeps).  Same kernels, same operations: bit-identical to the plain call; the launch
    counter shows the skewed path really ran.  The frame is just over the 64 MB threshold of the path."""
    from rnmf_b200 import _capi
    from rnmf_b200.dataframe import _faces_array
    n, hi = [254, 126, 260], [254.0, 126.0, 260.0]
    _, bc = _bc_pair(R, 3, MIXED_3D)
    faces = _faces_array(bc, 3, 1)
    rng = np.random.default_rng(83)
    G = (n[0] + 2) * (n[1] + 2) * (n[2] + 2)
    psi0, rhs0 = rng.standard_normal(G), rng.standard_normal(G)
    nn = (C.c_int64 * 3)(*n)
    lo = (C.c_double * 3)(0.0, 0.0, 0.0)
    dx = (C.c_double * 3)(1.0, 1.0, 1.0)
    outs, norms, launches = [], [], []
    try:
        for skew in (0, 1, 2):                          # 2: the band-wise end with the overlapped download as well
            ctx.set_option("e2e_skew", skew)
            ctx.set_option("e2e_skew_chunks", chunks)
            ctx.set_option("e2e_skew_pairs", pairs)
            ctx.set_option("e2e_skew_tail", 5)
            out = np.empty_like(psi0)
            l1 = C.c_double()
            k0 = ctx.kernel_launches
            _capi.check(_capi.lib().rnmf_solve_poisson_host(ctx._h, 3, nn, lo, dx, 1, faces, psi0.ctypes.data, rhs0.ctypes.data,
                                                            sweeps, 0, out.ctypes.data, C.byref(l1)))
            launches.append(ctx.kernel_launches - k0)
            outs.append(out)
            norms.append(l1.value)
    finally:
        ctx.set_option("e2e_skew", 2)
        ctx.set_option("e2e_skew_chunks", 8)
        ctx.set_option("e2e_skew_pairs", 56)
        ctx.set_option("e2e_skew_tail", 32)
    np.testing.assert_array_equal(outs[1], outs[0])
    np.testing.assert_array_equal(outs[2], outs[0])
    assert norms[1] == norms[0]
    assert abs(norms[2] - norms[0]) <= NORM_RTOL * abs(norms[0])      # same partial sums, grouped by band
    assert launches[1] > launches[0] + chunks          # one band launch per (chunk, pair) instead of one launch per pair
    assert launches[2] >= launches[1]                  # equal when too few sweeps remain for a band-wise end (levels < 4)


def test_kernel_launch_counter_moves(R, ctx):
    from rnmf_b200 import poisson
    g, obc, psi0, rhs0, psi, rhs = _frames(R, ctx, [32, 32, 32], [32.0] * 3, MIXED_3D, seed=91)
    ctx.set_option("resident", 0)
    before = ctx.kernel_launches
    m0 = ctx.multi_sweep_stats
    poisson.jacobi_sweeps(psi, rhs, 10)
    ctx.sync()
    ctx.set_option("resident", 1)
    m1 = ctx.multi_sweep_stats
    assert (m1[0] - m0[0], m1[1] - m0[1]) == (3, 9)            # 3D default: three sweeps per launch (+ one single sweep)
    assert ctx.kernel_launches - before >= 4 + 1


# ---- size-independent properties at BASELINE.json's full sizes (no host arrays involved) --------
FULL_SIZES = [
    ([8192, 8192], DIRICHLET_2D),                 # configs[1]
    ([512, 512, 512], MIXED_3D),                  # configs[2]
    ([768, 768, 768], MIXED_3D),                  # configs[4], per-GPU shape
]


def _synthetic_grown(g, n, seed):
    """the bench's synthetic field in the reference's dense grown layout, generated on the host by numpy
    in bounded chunks of rows: valid cell q (flat valid index) = splitmix64(seed ^ q) -> (-1, 1); ghosts 0."""
    out = ora.zeros(g)
    v = ora.valid_view(g, out)[0]                    # [nz,] ny, nx view into `out`
    planes = v if len(n) == 3 else v[None]
    rows_per = max(1, (1 << 23) // n[0])
    for k in range(planes.shape[0]):
        for j0 in range(0, n[1], rows_per):
            j1 = min(n[1], j0 + rows_per)
            q0 = (k * n[1] + j0) * n[0]
            planes[k, j0:j1] = ora.splitmix_field((j1 - j0) * n[0], seed, offset=q0).reshape(j1 - j0, n[0])
    return out


@pytest.mark.parametrize("n,spec", FULL_SIZES)
def test_full_size_oracle_parity(R, ctx, n, spec):
    """BASELINE.json's own sizes against the oracle, bit for bit (VERDICT r1 "next" #1): the default path runs
    5 sweeps (two double-sweep launches + the norm-carrying single sweep) of the update expression of
    examples/magnetostatics/poisson.rs:80-89 (3D: SURVEY 8a's extrapolation) with the fused ghost fill
    (dataframe.rs:289-433) on the bench's synthetic fields; the oracle does the same with its OpenMP Jacobi loop
    (out of place => independent of the thread schedule, bit-identical to the serial loop, which
    tests/test_oracle_known_answers.py checks).  The host-side synthetic field is generated by numpy, so the
    device generator is checked at full size as well.  Norm: 1e-12 relative against the long-double serial sum."""
    from rnmf_b200 import poisson
    dim = len(n)
    hi = [float(v) for v in n]
    g = ora.geom(n, hi=hi)
    obc, bc = _bc_pair(R, dim, spec)
    mesh = R.CartesianMesh.new([0.0] * dim, hi, n)
    sweeps = 6 if dim == 3 else 5      # 3D: one triple-sweep launch + one double-sweep launch + the norm-carrying single sweep
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx)
    psi.fill_synthetic(0x5EED0001)
    rhs.fill_synthetic(0x5EED0002)
    psi.fill_bc()
    want = _synthetic_grown(g, n, 0x5EED0001)
    rhs0 = _synthetic_grown(g, n, 0x5EED0002)
    assert ora.fill_bc(g, want, obc) == 0
    got = psi.data
    assert np.array_equal(got, want), "synthetic field + fill_bc differ from the oracle at full size"
    assert np.array_equal(rhs.data, rhs0)
    d0, m0 = ctx.double_sweep_launches, ctx.multi_sweep_stats
    l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=True)
    if dim == 3:
        m1 = ctx.multi_sweep_stats
        assert ctx.double_sweep_launches - d0 == 1, "the leftover pair did not take the double-sweep kernel"
        assert (m1[0] - m0[0], m1[1] - m0[1]) == (2, 5), "the default 3D path did not take the three-sweep kernel"
    l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps, omp=True)
    del rhs0
    del got
    got = psi.data
    same = np.array_equal(got, want)
    if not same:
        bad = np.flatnonzero(got != want)
        pytest.fail(f"{bad.size} of {got.size} grown cells differ from the oracle; first at flat index {bad[0]}: "
                    f"{got[bad[0]]!r} vs {want[bad[0]]!r}")
    assert abs(l1 - l1_ref) <= NORM_RTOL * abs(l1_ref), (l1, l1_ref)
    assert psi.guards_intact()
    psi.close(); rhs.close()


@pytest.mark.parametrize("n,spec", FULL_SIZES)
def test_full_size_properties(R, ctx, n, spec):
    """Extra, size-independent properties at the benchmark sizes (test_full_size_oracle_parity above is the
    parity test), evaluated entirely on the device:
      (1) determinism: the same call twice gives bit-identical norm and checksum;
      (2) fill_bc is idempotent (second fill changes nothing: |diff| sums to exactly 0);
      (3) linearity: sweeps(a+b; ra+rb) == sweeps(a; ra) + sweeps(b; rb) with homogeneous
          Dirichlet faces, to rounding (relative 1e-13 of the field's L1 norm);
      (4) the default path (3D: three sweeps per pass over HBM, 2D: four) agrees bit for bit with the
          single-sweep kernel (one sweep per launch, different tiling)."""
    from rnmf_b200 import poisson
    dim = len(n)
    hi = [float(v) for v in n]
    mesh = R.CartesianMesh.new([0.0] * dim, hi, n)
    _, bc = _bc_pair(R, dim, spec)
    mk = lambda b: R.CartesianDataFrame.new_from(mesh, b, 1, 1, ctx)
    sweeps = 6          # 3D default: a triple-sweep launch, a double-sweep launch and the norm-carrying single sweep

    def run(variant):
        ctx.set_option("sweep_variant", variant)
        psi, rhs = mk(bc), mk(bc)
        psi.fill_synthetic(0x5EED0001)
        rhs.fill_synthetic(0x5EED0002)
        psi.fill_bc()
        l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=True)
        return psi, rhs, l1, psi.abs_sum()

    try:
        psi, rhs, l1, cs = run(0)
        psi2, rhs2, l1b, csb = run(0)
        assert (l1, cs) == (l1b, csb)                                       # (1)
        diff = mk(bc)
        diff.axpby_(1.0, psi, -1.0, psi2)
        assert diff.abs_sum() == 0.0
        psi2.fill_bc()                                                      # (2)
        diff.axpby_(1.0, psi, -1.0, psi2)
        assert diff.abs_sum() == 0.0
        psi2.close(); rhs2.close()
        psi3, rhs3, l1c, csc = run(6 if dim == 3 else 20)                   # (4) one sweep per launch
        diff.axpby_(1.0, psi, -1.0, psi3)
        assert diff.abs_sum() == 0.0 and l1c == pytest.approx(l1, rel=1e-12)
        psi3.close(); rhs3.close(); psi.close(); rhs.close()
        ctx.set_option("sweep_variant", "auto")
        # (3) linearity
        zero = ([("dirichlet", 0.0)] * dim, [("dirichlet", 0.0)] * dim)
        _, bc0 = _bc_pair(R, dim, zero)
        a, ra, b, rb = mk(bc0), mk(bc0), mk(bc0), mk(bc0)
        a.fill_synthetic(11); ra.fill_synthetic(12); b.fill_synthetic(13); rb.fill_synthetic(14)
        s, rs = mk(bc0), mk(bc0)
        s.axpby_(1.0, a, 1.0, b)
        rs.axpby_(1.0, ra, 1.0, rb)
        for f in (a, b, s):
            f.fill_bc()
        for f, r in ((a, ra), (b, rb), (s, rs)):
            poisson.jacobi_sweeps(f, r, sweeps)
        scale = s.abs_sum()
        diff.axpby_(1.0, a, 1.0, b)
        diff.axpby_(1.0, s, -1.0, diff)
        assert diff.abs_sum() <= 1e-13 * scale * 8
        assert all(f.guards_intact() for f in (a, b, s))
    finally:
        ctx.set_option("sweep_variant", "auto")


# ---- error behaviour of the C ABI (the Rust shim turns these into panics) ------------------------
def test_error_paths(R, ctx):
    from rnmf_b200 import _capi, poisson
    L = _capi.lib()
    mesh2 = R.CartesianMesh.new([0.0, 0.0], [8.0, 8.0], [8, 8])
    mesh2b = R.CartesianMesh.new([0.0, 0.0], [8.0, 8.0], [8, 9])
    _, bc = _bc_pair(R, 2, MIXED_2D)
    a = R.CartesianDataFrame.new_from(mesh2, bc, 1, 1, ctx)
    b = R.CartesianDataFrame.new_from(mesh2b, bc, 1, 1, ctx)
    nobc = R.CartesianDataFrame.new_from(mesh2, R.BCs.empty(), 1, 1, ctx)
    two = R.CartesianDataFrame.new_from(mesh2, R.BCs([bc.bcs[0], bc.bcs[0]]), 2, 1, ctx)
    coef = (C.c_double * 4)(0.25, 0.25, 0.25, 0.0)

    def code(fn, *args):
        rc = fn(*args)
        assert rc != 0 and len(L.rnmf_last_error()) > 0
        return rc

    assert code(L.rnmf_jacobi_sweeps, a._h, b._h, coef, 1, None) == _capi.ERR_INVALID        # shape mismatch
    assert code(L.rnmf_jacobi_sweeps, nobc._h, a._h, coef, 1, None) == _capi.ERR_INVALID     # BCs::empty(): the reference would panic
    assert code(L.rnmf_jacobi_sweeps, two._h, two._h, coef, 1, None) == _capi.ERR_INVALID    # sweeps are for n_comp = 1
    assert code(L.rnmf_jacobi_sweeps, a._h, a._h, coef, -1, None) == _capi.ERR_INVALID
    assert code(L.rnmf_add, a._h, a._h, b._h) == _capi.ERR_INVALID
    assert code(L.rnmf_fill_bc, nobc._h) == _capi.ERR_INVALID
    assert code(L.rnmf_frame_set_bc, a._h, None, 4) == _capi.ERR_INVALID
    h = C.c_void_p()
    n = (C.c_int64 * 3)(4, 4, 1)
    lo = (C.c_double * 3)(0, 0, 0)
    dx = (C.c_double * 3)(1, 1, 1)
    assert code(L.rnmf_frame_create, ctx._h, 4, n, lo, dx, 1, 1, C.byref(h)) == _capi.ERR_INVALID     # dim
    assert code(L.rnmf_frame_create, ctx._h, 2, n, lo, dx, 0, 1, C.byref(h)) == _capi.ERR_INVALID     # n_ghost
    assert code(L.rnmf_frame_create, ctx._h, 2, n, lo, dx, 1, 0, C.byref(h)) == _capi.ERR_INVALID     # n_comp
    # reference-order GS: 2D, one ghost layer, Dirichlet/Neumann everywhere
    mesh3 = R.CartesianMesh.new([0.0] * 3, [4.0] * 3, [4, 4, 4])
    _, bc3 = _bc_pair(R, 3, MIXED_3D)
    f3 = R.CartesianDataFrame.new_from(mesh3, bc3, 1, 1, ctx)
    assert code(L.rnmf_gs_sweeps, f3._h, f3._h, coef, 1) == _capi.ERR_INVALID
    _, bcp = _bc_pair(R, 2, ([("prescribed", [1.0, 2.0]), ("dirichlet", 0.0)], [("dirichlet", 0.0)] * 2))
    fp = R.CartesianDataFrame.new_from(mesh2, bcp, 1, 1, ctx)
    assert code(L.rnmf_gs_sweeps, fp._h, a._h, coef, 1) == _capi.ERR_INVALID
    # a sweep variant that is not in the table (kernels_sweep.cu) is an error, not a silent default
    ctx.set_option("sweep_variant", 26)
    try:
        assert code(L.rnmf_jacobi_sweeps, a._h, a.clone()._h, coef, 1, None) == _capi.ERR_INVALID
    finally:
        ctx.set_option("sweep_variant", "auto")
    # after the errors the context still works
    assert poisson.jacobi_sweeps(a, a.clone(), 2, want_norm=True) >= 0.0


def test_large_upload_uses_staged_path_and_roundtrips(R, ctx):
    """Frames >= 64 MB are uploaded as one 1-D copy into a staging buffer and repacked on the device;
    the round trip through the pitched layout must be exact (odd sizes, 2 ghost layers, 2 components)."""
    n, ng, ncomp = [2051, 2101], 2, 2
    g = ora.geom(n, ng=ng, ncomp=ncomp)
    assert g.n_nodes * 8 >= 64 << 20
    mesh = R.CartesianMesh.new([0.0, 0.0], [1.0, 1.0], n)
    f = R.CartesianDataFrame.new_from(mesh, R.BCs.empty(), ncomp, ng, ctx)
    host = np.random.default_rng(3).standard_normal(g.n_nodes)
    f.data = host
    np.testing.assert_array_equal(f.data, host)
    np.testing.assert_array_equal(f.get_valid_cells(), ora.valid_view(g, host).reshape(-1))
    assert f.guards_intact()


# ---- side iterators / collect_typed_cells / IndexMut --------------------
def test_collect_side_golden_vectors_on_gpu(R, ctx, golden):
    """The halo PACK order pinned by dataframe.rs:1148-1201 and the ghost-side iteration pinned by
    :1065-1145, through the GPU gather."""
    case = golden["ghost_collect_two_edge"]
    _, bc = _bc_pair(R, 2, ([tuple(f) for f in case["lo"]], [tuple(f) for f in case["hi_bc"]]))
    f = R.CartesianDataFrame.new_from(R.CartesianMesh.new([0.0, 0.0], case["hi"], case["n"]), bc, 1, case["ng"], ctx)
    f.fill_ic(lambda x, y, c: x + 1.0)
    f.fill_bc()
    want = {tuple(c["sides"])[0]: c["values"] for c in case["collect"]}
    assert f.collect_side(1, 1).tolist() == want["Valid(North)"]
    assert f.collect_side(1, 0).tolist() == want["Valid(South)"]
    assert f.collect_side(0, 1).tolist() == want["Valid(East)"]
    assert f.collect_side(0, 0).tolist() == want["Valid(West)"]
    case = golden["ghost_iter_two_ghost"]
    _, bc = _bc_pair(R, 2, ([tuple(x) for x in case["lo"]], [tuple(x) for x in case["hi_bc"]]))
    f = R.CartesianDataFrame.new_from(R.CartesianMesh.new([0.0, 0.0], case["hi"], case["n"]), bc, 1, case["ng"], ctx)
    f.fill_ic(lambda x, y, c: x + 1.0)
    f.fill_bc()
    ng, n = case["ng"], case["n"][0]
    east = f.collect_side(0, 1, ghost=True)             # Ghost(East|NE|SE): rows 0..G1-1, ng cells each
    want = {it["sides"][0]: it["values"] for it in case["iters"]}
    assert east[ng * ng:-ng * ng].tolist() == want["Ghost(East)"]
    assert east[-ng * ng:].tolist() == want["Ghost(NorthEast)"]
    south = f.collect_side(1, 0, ghost=True).reshape(ng, n + 2 * ng)
    assert south[:, ng:-ng].reshape(-1).tolist() == want["Ghost(South)"]


@pytest.mark.parametrize("n,ng,ncomp", [([7, 5], 2, 2), ([6, 5, 4], 1, 1), ([5, 6, 7], 2, 2)])
def test_collect_side_matches_dense_slices(R, ctx, n, ng, ncomp):
    dim = len(n)
    g = ora.geom(n, ng=ng, ncomp=ncomp)
    mesh = R.CartesianMesh.new([0.0] * dim, [float(v) for v in n], n)
    f = R.CartesianDataFrame.new_from(mesh, R.BCs.empty(), ncomp, ng, ctx)
    host = np.random.default_rng(17).standard_normal(g.n_nodes)
    f.data = host
    grid = ora.as_grid(g, host)                       # (c, [z,] y, x)
    for d in range(dim):
        ax = grid.ndim - 1 - d                        # numpy axis of direction d
        for side in (0, 1):
            # ghost side: full grown extent in the other directions
            sl = [slice(None)] * grid.ndim
            sl[ax] = slice(0, ng) if side == 0 else slice(n[d] + ng, n[d] + 2 * ng)
            np.testing.assert_array_equal(f.collect_side(d, side, ghost=True), grid[tuple(sl)].reshape(-1))
            np.testing.assert_array_equal(f.collect_side(d, side, ghost=True, comp=ncomp - 1), grid[tuple(sl)][ncomp - 1].reshape(-1))
            # valid strip: valid extent in the other directions
            sl = [slice(None)] + [slice(ng, ng + n[dim - 1 - a]) for a in range(dim)]
            sl[ax] = slice(ng, 2 * ng) if side == 0 else slice(n[d], n[d] + ng)
            np.testing.assert_array_equal(f.collect_side(d, side), grid[tuple(sl)].reshape(-1))


def test_index_and_index_mut(R, ctx):
    n, ng = [6, 5, 4], 2
    g = ora.geom(n, ng=ng, ncomp=2)
    f = R.CartesianDataFrame.new_from(R.CartesianMesh.new([0.0] * 3, [6.0, 5.0, 4.0], n), R.BCs.empty(), 2, ng, ctx)
    host = np.random.default_rng(19).standard_normal(g.n_nodes)
    f.data = host
    grid = ora.as_grid(g, host)
    for (i, j, k, c) in ((0, 0, 0, 0), (-2, -1, 3, 1), (5, 4, 3, 1), (7, 6, 5, 0), (-2, -2, -2, 1)):
        assert f[(i, j, k, c)] == grid[c, k + ng, j + ng, i + ng]
    f[(3, -2, 5, 1)] = 42.5
    host2 = f.data
    assert ora.as_grid(g, host2)[1, 5 + ng, -2 + ng, 3 + ng] == 42.5
    host2.reshape(grid.shape)[1, 5 + ng, -2 + ng, 3 + ng] = grid[1, 5 + ng, -2 + ng, 3 + ng]
    np.testing.assert_array_equal(host2, host)
    with pytest.raises(R.RnmfError):
        f[(8, 0, 0, 0)]


def test_context_and_frame_lifecycle_frees_device_memory(R):
    """Contexts and frames are created and destroyed repeatedly (the Rust wrapper frees in Drop);
    device memory must come back and a context destroyed with live Python frames must not crash."""
    import gc
    import torch
    from rnmf_b200 import poisson
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for rep in range(3):
        c = R.Context(0)
        mesh = R.CartesianMesh.new([0.0] * 3, [96.0] * 3, [96, 96, 96])
        _, bc = _bc_pair(R, 3, MIXED_3D)
        frames = [R.CartesianDataFrame.new_from(mesh, bc, 1, 1, c) for _ in range(4)]
        frames[0].fill_synthetic(1)
        frames[1].fill_synthetic(2)
        frames[0].fill_bc()
        assert poisson.jacobi_sweeps(frames[0], frames[1], 6, want_norm=True) > 0.0
        if rep == 0:
            for f in frames:
                f.close()
        c.close()                       # rep > 0: closes the still-open frames first
        assert all(f._h is None for f in frames)
    gc.collect()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20     # nothing but allocator slack left behind
