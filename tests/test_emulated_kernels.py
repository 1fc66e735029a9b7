"""The double-sweep kernel SOURCES (rnmf_b200/csrc/kernels_sweep_tb2.cu, kernels_sweep2d_tb2.cu) run on the CPU box:
compiled with g++ -DRNMF_EMULATE against a thread-per-lane SIMT emulator (tests/emu/emu_cuda.hpp: barriers,
mbarrier phases, TMA box copies with zero fill, shared memory by byte offset, the slab handshake as atomics) and
driven through the library's own launchers (planner, tensor maps, A/B chunk roles).  Checks, bit for bit against
the oracle, what a GPU box is too expensive to iterate on: tile / halo / ghost indexing, the ring and barrier
protocols (a wrong parity or a missing arrival deadlocks and is reported), the half-step / full-step exchange
between slabs with every "rank" a host thread.  Test infrastructure: nothing here is part of librnmf_b200.so, and
the GPU tests remain the parity tests proper."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "emu_driver.cpp")
OUT = os.path.join(ROOT, "tests", "emu", "_build", "libemu_driver.so")
KIND = {"dirichlet": 0, "neumann": 1}

MIXED_3D = ([("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 0.25)], [("dirichlet", 1.0), ("neumann", -0.5), ("neumann", 0.5)])
NEUMANN_3D = ([("neumann", 0.3), ("neumann", -0.2), ("neumann", 0.1)], [("neumann", 0.4), ("neumann", 0.6), ("neumann", -0.9)])
MIXED_2D = ([("neumann", 0.3), ("dirichlet", -1.0)], [("dirichlet", 2.0), ("neumann", -0.7)])
NEUMANN_2D = ([("neumann", 0.3), ("neumann", -0.2)], [("neumann", 0.4), ("neumann", 0.6)])


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(ROOT, "rnmf_b200", "csrc", f) for f in
                    ("kernels_sweep_tb2.cu", "kernels_sweep_tb3.cu", "kernels_sweep2d_tb2.cu", "kernels_sweep2d_tbk.cu", "tma_util.cuh", "common.cuh")] + [os.path.join(ROOT, "tests", "emu", "emu_cuda.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                        "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "emu"), "-Wno-psabi",
                        "-Wl,-Bsymbolic",          # the driver defines the library's own launcher names: bind them inside the driver
                        "-o", OUT, SRC], check=True)
    lib = C.CDLL(OUT)
    lib.emu_set_timeout(20)
    return lib


def test_the_emulator_reports_deadlocks_access_faults_and_models_tma_zero_fill(emu):
    assert emu.emu_selftest() == 7


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _setup(n, hi, spec, sweeps, seed):
    dim = len(n)
    g = ora.geom(n, hi=hi)
    obc = ora.bcs(dim, [spec])
    rng = np.random.default_rng(seed)
    psi, rhs = rng.standard_normal(g.n_nodes), rng.standard_normal(g.n_nodes)
    ora.fill_bc(g, psi, obc)
    want = psi.copy()
    ora.jacobi(g, want, rhs, obc, sweeps)
    kinds, vals = (C.c_int * 6)(*([0] * 6)), (C.c_double * 6)(*([0.0] * 6))
    for d in range(dim):
        for side in range(2):
            kinds[d * 2 + side], vals[d * 2 + side] = KIND[spec[side][d][0]], spec[side][d][1]
    dx = (C.c_double * 3)(*[hi[d] / n[d] for d in range(dim)], *([1.0] * (3 - dim)))
    nn = (C.c_longlong * 3)(*n, *([1] * (3 - dim)))
    return g, psi, rhs, want, kinds, vals, dx, nn, ora.coefs(g)


@pytest.mark.parametrize("n,hi,spec,pairs,chunk", [
    ([20, 9, 5], [1.0, 2.0, 3.0], MIXED_3D, 1, 0),               # odd n1: the last row is a warp's row A (y-hi ghost of u1 in row B's registers)
    ([129, 13, 4], [1.0, 2.0, 3.0], NEUMANN_3D, 2, 0),           # one cell / one row past a tile edge; two launches (ping-pong)
    ([65, 25, 7], [65.0, 50.0, 21.0], MIXED_3D, 1, 2),           # odd n0: x-hi ghost in the pair's second register (x-half 1 warp); 3 y tiles, 4 z chunks
    ([63, 24, 3], [1.0, 2.0, 3.0], MIXED_3D, 1, 0),              # x-hi face in the x-half 0 warp; n1 = two full tiles
    ([30, 12, 1], [1.0, 1.0, 1.0], MIXED_3D, 1, 0),              # a single plane: both z ghosts of u1 formed in registers
    ([1, 1, 1], [1.0, 1.0, 1.0], MIXED_3D, 2, 0),
    ([2, 2, 2], [1.0, 1.0, 1.0], NEUMANN_3D, 2, 0),
    ([40, 10, 6], [1.0, 2.0, 3.0], MIXED_3D, 1, 3),
    # interior-specialised plane loop: warps whose 2 x 64 patch touches no x / y face
    ([400, 30, 7], [4.0, 2.0, 3.0], MIXED_3D, 1, 0),
    ([385, 13, 4], [1.0, 2.0, 3.0], NEUMANN_3D, 2, 2),           # the last tile holds one column; two launches; 2-plane chunks
    ([384, 27, 9], [1.0, 2.0, 3.0], MIXED_3D, 1, 3),             # x-hi face exactly at a tile edge; odd n1
    ([192, 26, 9], [1.0, 2.0, 3.0], MIXED_3D, 1, 3),             # x-hi face at the edge of an x-half 0 warp
    ([200, 30, 20], [1.0, 2.0, 3.0], MIXED_3D, 1, 7),            # several unrolled trips of the fast loop per chunk
    # chunks long enough for the six-iteration trips with compile-time ring indices (they start at it % 6 == 0), two launches
    ([200, 26, 27], [1.0, 2.0, 3.0], MIXED_3D, 2, 0),            # one 27-plane chunk: three static trips, then the drain
    ([260, 14, 40], [1.0, 2.0, 3.0], NEUMANN_3D, 1, 20),         # two 20-plane chunks: the second starts at a chunk boundary (a = k0 - 1)
])
def test_3d_double_sweep_source_on_the_emulator(emu, n, hi, spec, pairs, chunk):
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 2 * pairs, seed=41)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(3, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 0, chunk, pairs)
    assert rc == 0, "a barrier wait timed out (deadlock) or the launch failed"
    np.testing.assert_array_equal(out, want)      # valid cells, face ghosts, and the untouched edges / corners


@pytest.mark.parametrize("n,hi,spec,launches,chunk", [
    ([20, 8, 5], [1.0, 2.0, 3.0], MIXED_3D, 1, 0),               # one tile, one chunk: both z faces in the same chunk
    ([130, 14, 4], [1.0, 2.0, 3.0], NEUMANN_3D, 2, 0),           # one pair / one row pair past a tile edge; two launches (ping-pong)
    ([66, 26, 7], [66.0, 52.0, 21.0], MIXED_3D, 1, 2),           # x-hi face inside the x-half 1 warp; 3 y tiles, 4 z chunks
    ([64, 24, 3], [1.0, 2.0, 3.0], MIXED_3D, 1, 0),              # x-hi face at the edge of the x-half 0 warp; n1 = two full tiles
    ([30, 12, 1], [1.0, 1.0, 1.0], MIXED_3D, 1, 0),              # a single plane: every level meets both z faces
    ([2, 2, 2], [1.0, 1.0, 1.0], NEUMANN_3D, 2, 0),
    ([2, 2, 1], [1.0, 1.0, 1.0], MIXED_3D, 1, 0),
    ([40, 10, 6], [1.0, 2.0, 3.0], MIXED_3D, 1, 1),              # one-plane chunks: every chunk clips some level at a face or starts next to one
    ([40, 10, 7], [1.0, 2.0, 3.0], NEUMANN_3D, 1, 3),
    # interior-specialised steady loop: warps whose 2 x 64 patch touches no x / y face
    ([400, 30, 9], [4.0, 2.0, 3.0], MIXED_3D, 1, 0),
    ([386, 14, 5], [1.0, 2.0, 3.0], NEUMANN_3D, 2, 2),           # the last tile holds one pair; two launches; 2-plane chunks
    ([384, 28, 9], [1.0, 2.0, 3.0], MIXED_3D, 1, 4),             # x-hi face exactly at a tile edge
    ([192, 26, 9], [1.0, 2.0, 3.0], MIXED_3D, 1, 3),             # x-hi face at the edge of an x-half 0 warp
    ([200, 30, 20], [1.0, 2.0, 3.0], MIXED_3D, 1, 7),            # several steady iterations per chunk
    ([260, 14, 40], [1.0, 2.0, 3.0], NEUMANN_3D, 1, 20),         # two 20-plane chunks: the second starts at a chunk boundary (a1 = k0 - 2)
])
def test_3d_triple_sweep_source_on_the_emulator(emu, n, hi, spec, launches, chunk):
    """Three sweeps per pass (kernels_sweep_tb3.cu; frames with even n0, n1): u after `launches` launches == 3*launches oracle sweeps,
    valid cells, face ghosts and the untouched edges / corners, bit for bit."""
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 3 * launches, seed=51)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(3, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 3, chunk, launches)
    assert rc == 0, "a barrier wait timed out (deadlock), an access fault, or the launch failed"
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,hi,spec,pairs,chunk", [
    ([33, 7], [1.0, 2.0], MIXED_2D, 1, 0),
    ([513, 5], [513.0, 5.0], MIXED_2D, 2, 2),                    # one cell past a 512-column strip; 3 row chunks
    ([40, 1], [1.0, 1.0], MIXED_2D, 1, 0),                       # a single row
    ([2, 3], [2.0, 3.0], MIXED_2D, 3, 1),
])
def test_2d_double_sweep_source_on_the_emulator(emu, n, hi, spec, pairs, chunk):
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 2 * pairs, seed=42)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(2, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 30, chunk, pairs)
    assert rc == 0
    np.testing.assert_array_equal(out, want)


DEPTH = 4      # sweeps per launch of the K-sweep 2D kernel (kernels_sweep2d_tbk.cu: the one shape that ships)


@pytest.mark.parametrize("n,hi,spec,launches,chunk", [
    ([33, 7], [1.0, 2.0], MIXED_2D, 1, 0),
    ([513, 5], [513.0, 5.0], MIXED_2D, 2, 2),                    # one cell past a 512-column strip; 3 row chunks; ping-pong
    ([515, 9], [1.0, 1.0], MIXED_2D, 1, 4),                      # the second strip is 3 columns wide: x-hi ghosts formed by halo lanes
    ([769, 6], [1.0, 3.0], NEUMANN_2D, 1, 3),                    # one cell past a 768-column strip
    ([40, 1], [1.0, 1.0], MIXED_2D, 1, 0),                       # a single row: every level meets both y faces
    ([40, 2], [1.0, 1.0], MIXED_2D, 2, 0),
    ([2, 3], [2.0, 3.0], MIXED_2D, 3, 1),                        # one-row chunks
    ([1, 1], [1.0, 1.0], MIXED_2D, 2, 0),
    ([130, 23], [13.0, 2.3], NEUMANN_2D, 1, 5),                  # chunks whose lower levels are clipped at either face
    ([64, 40], [1.0, 1.0], MIXED_2D, 1, 0),                      # planner-chosen chunks, steady (unrolled) iterations
    ([131, 50], [1.0, 1.0], MIXED_2D, 1, 25),                    # odd n0 in the second warp; long chunks: several unrolled trips
])
def test_2d_multi_sweep_source_on_the_emulator(emu, n, hi, spec, launches, chunk):
    """Four sweeps per pass: u_4 after `launches` launches == 4*launches oracle sweeps, bit for bit."""
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, DEPTH * launches, seed=44)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(2, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 0, chunk, launches)
    assert rc == 0, "a barrier wait timed out (deadlock) or the launch failed"
    np.testing.assert_array_equal(out, want)


from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=30, deadline=None, derandomize=True)
@given(st.integers(1, 1100), st.integers(1, 14), st.integers(0, 14), st.integers(0, 3), st.integers(1, 2))
def test_2d_multi_sweep_random_shapes_on_the_emulator(emu, n0, n1, chunk, bc, launches):
    """Random frame shapes / chunk lengths / face rules for the K-sweep kernel: strips that end anywhere, chunks shorter than the
    ramp-up, levels clipped at both y faces at once."""
    spec = [MIXED_2D, NEUMANN_2D, ([("dirichlet", 0.5), ("neumann", 0.1)], [("neumann", -0.3), ("dirichlet", 1.5)]),
            ([("dirichlet", 0.0), ("dirichlet", 1.0)], [("dirichlet", 2.0), ("dirichlet", -1.0)])][bc]
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup([n0, n1], [1.0 + n0 / 7.0, 2.0], spec, DEPTH * launches, seed=n0 * 31 + n1)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(2, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 0, min(chunk, n1), launches)
    assert rc == 0
    np.testing.assert_array_equal(out, want)


@settings(max_examples=16, deadline=None, derandomize=True)
@given(st.integers(1, 400), st.integers(1, 30), st.integers(1, 8), st.integers(0, 8), st.booleans())
def test_3d_double_sweep_random_shapes_on_the_emulator(emu, n0, n1, n2, chunk, neumann):
    """Random frame shapes / chunk lengths for the 3D double-sweep kernel."""
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup([n0, n1, n2], [1.0, 2.0, 3.0], NEUMANN_3D if neumann else MIXED_3D, 2,
                                                          seed=n0 * 131 + n1 * 17 + n2)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(3, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 0, min(chunk, n2), 1)
    assert rc == 0
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,hi,spec,pairs,nranks,chunk", [
    ([20, 9, 12], [1.0, 2.0, 3.0], MIXED_3D, 2, 2, 0),           # two slabs: each rank has one neighbour
    ([130, 13, 17], [1.0, 2.0, 3.0], MIXED_3D, 2, 3, 0),         # uneven split (5, 5, 7); the middle rank has both
    ([20, 14, 16], [1.0, 2.0, 3.0], NEUMANN_3D, 3, 4, 2),        # 4-plane slabs (the minimum), two-plane chunks: first and last chunk differ
    ([20, 14, 16], [1.0, 2.0, 3.0], MIXED_3D, 3, 4, 4),          # 4-plane slabs, one chunk: the same CTA serves both neighbours
    ([390, 15, 13], [1.0, 2.0, 3.0], MIXED_3D, 2, 3, 2),         # interior-specialised loop between the planes that travel; odd n1
    ([130, 13, 17], [1.0, 2.0, 3.0], NEUMANN_3D, 2, 3, 3),       # a chunk length that leaves a one-plane last chunk (the planner lengthens it)
    ([65, 11, 24], [1.0, 2.0, 3.0], MIXED_3D, 3, 2, 5),          # odd n0: the x-hi ghosts of the travelling planes
    ([200, 26, 60], [1.0, 2.0, 3.0], MIXED_3D, 2, 2, 0),         # 30-plane slabs in one chunk: static trips between the planes that travel
])
def test_slab_handshake_of_the_double_sweep_on_the_emulator(emu, n, hi, spec, pairs, nranks, chunk):
    """Every rank is a host thread launching its slab's kernel (CTAs one at a time, in order: the most adversarial
    schedule).  Deep halo: each rank starts with its neighbours' two boundary planes in its ghost + deep plane and stores u2 of its
    own two boundary planes (whole grown rows, x / y face ghosts included) into the neighbours' planes under the full-step flags."""
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 2 * pairs, seed=43)
    out = psi.copy()
    rc = emu.emu_sweeps_slabs(3, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 0, chunk, pairs, nranks)
    assert rc == 0, "deadlock / handshake timeout"
    np.testing.assert_array_equal(out, want)


@settings(max_examples=16, deadline=None, derandomize=True)
@given(st.integers(1, 200), st.integers(1, 15), st.integers(1, 9), st.integers(0, 9), st.booleans())
def test_3d_triple_sweep_random_shapes_on_the_emulator(emu, h0, h1, n2, chunk, neumann):
    """Random (even n0, n1) frame shapes / chunk lengths for the three-sweep kernel."""
    n0, n1 = 2 * h0, 2 * h1
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup([n0, n1, n2], [1.0, 2.0, 3.0], NEUMANN_3D if neumann else MIXED_3D, 3,
                                                          seed=n0 * 131 + n1 * 17 + n2)
    out = np.zeros_like(psi)
    rc = emu.emu_double_sweeps(3, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 3, min(chunk, n2), 1)
    assert rc == 0
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,hi,spec,launches,nranks,chunk", [
    ([20, 8, 12], [1.0, 2.0, 3.0], MIXED_3D, 2, 2, 0),           # two 6-plane slabs (the minimum): one chunk serves the neighbour and the face
    ([130, 14, 20], [1.0, 2.0, 3.0], MIXED_3D, 2, 3, 0),         # uneven split (6, 6, 8); the middle rank has both neighbours
    ([20, 14, 24], [1.0, 2.0, 3.0], NEUMANN_3D, 3, 4, 3),        # 6-plane slabs, three-plane chunks: first and last chunk differ
    ([20, 14, 24], [1.0, 2.0, 3.0], MIXED_3D, 2, 4, 6),          # 6-plane slabs, one chunk: the same CTA serves both neighbours
    ([390, 16, 21], [1.0, 2.0, 3.0], MIXED_3D, 2, 3, 0),         # interior-specialised loop between the planes that travel
    ([130, 14, 22], [1.0, 2.0, 3.0], NEUMANN_3D, 2, 3, 5),       # chunk lengths that leave a short last chunk (the planner lengthens the chunks)
    ([200, 26, 60], [1.0, 2.0, 3.0], MIXED_3D, 2, 2, 0),         # 30-plane slabs in one chunk
])
def test_slab_handshake_of_the_triple_sweep_on_the_emulator(emu, n, hi, spec, launches, nranks, chunk):
    """z-slabs, three sweeps per launch: each rank starts with its neighbours' three boundary planes in its ghost + two deep planes
    (rhs: two) and stores u3 of its own three boundary planes into the neighbours' planes under the one-epoch-per-launch handshake."""
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 3 * launches, seed=53)
    out = psi.copy()
    rc = emu.emu_sweeps_slabs(3, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 3, chunk, launches, nranks)
    assert rc == 0, "deadlock / handshake timeout"
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,hi,spec,launches,nranks,chunk", [
    ([33, 16], [1.0, 2.0], MIXED_2D, 2, 2, 0),                  # 8-row slabs (the minimum: 2K rows), one chunk serves both faces
    ([513, 40], [513.0, 40.0], MIXED_2D, 2, 3, 0),              # two strips; uneven split (13, 13, 14); the middle rank has both neighbours
    ([130, 64], [13.0, 6.4], NEUMANN_2D, 3, 4, 5),              # 16-row slabs in chunks of 5 (the last chunk is lengthened to hold 4 rows)
    ([1031, 24], [1.0, 1.0], MIXED_2D, 1, 2, 4),                # odd n0: x-hi ghosts of the travelling rows; chunks of exactly K rows
    ([64, 90], [1.0, 1.0], MIXED_2D, 2, 2, 16),                 # several chunks per slab: steady iterations between the rows that travel
])
def test_slab_handshake_of_the_2d_multi_sweep_on_the_emulator(emu, n, hi, spec, launches, nranks, chunk):
    """y-slabs, four sweeps per launch: K boundary rows travel per neighbour and launch (ghost + 3 deep rows), lower levels are
    not clipped at a face that has a neighbour behind it."""
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, DEPTH * launches, seed=47)
    out = psi.copy()
    rc = emu.emu_sweeps_slabs(2, nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), 0, chunk, launches, nranks)
    assert rc == 0, "deadlock / handshake timeout / stores outside the frame"
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,spec,n_chunks,pairs,chunk,variant", [
    ([20, 9, 32], MIXED_3D, 2, 5, 0, 0),           # 16-plane chunks, bands shrink to nothing at the bottom (2*5 < 16)
    ([20, 9, 24], NEUMANN_3D, 3, 6, 3, 0),         # 8-plane chunks, more pairs than half a chunk: bands of chunk 0 vanish
    ([130, 13, 20], MIXED_3D, 4, 3, 2, 0),         # two x tiles, two y tiles, several CTAs per band
])
def test_time_skewed_start_schedule_on_the_emulator(emu, n, spec, n_chunks, pairs, chunk, variant):
    """The band schedule of the opt-in time-skewed start of rnmf_solve_poisson_host (skew_region, common.cuh): planes that have
    not "arrived" hold poison, yet after the last chunk every plane equals 2*pairs oracle sweeps bit for bit -- no band reads
    a plane before it exists or after the ping-pong partner overwrote it."""
    hi = [1.0, 2.0, 3.0]
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 2 * pairs, seed=45)
    out = np.zeros_like(psi)
    rc = emu.emu_skewed_double_sweeps(nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), variant, chunk, n_chunks, pairs)
    assert rc == 0, "deadlock or launch failure"
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,spec,n_chunks,pairs,chunk,nranks", [
    ([20, 9, 64], MIXED_3D, 2, 4, 0, 2),            # 32-plane slabs, 16-plane chunks, wedges of up to 8 planes
    ([20, 9, 96], NEUMANN_3D, 3, 5, 3, 3),          # the middle rank has both neighbours: two wedges per catch-up launch (hole)
    ([130, 13, 50], MIXED_3D, 2, 6, 0, 2),          # uneven split (25, 25), 4 * pairs = 24 <= 25: the wedges almost meet
    ([20, 9, 80], MIXED_3D, 4, 3, 2, 4),            # four ranks, four chunks of five planes each
])
def test_time_skewed_start_across_slabs_on_the_emulator(emu, n, spec, n_chunks, pairs, chunk, nranks):
    """The time-skewed start of rnmf_jacobi_sweeps_host on slab frames: bands clipped to [2s, m - 2s) next to a neighbour while the
    frames arrive (the neighbours' ghost planes are poison until the exchange), then the wedge catch-up, one launch per level over
    both wedges with the slab handshake.  Every plane ends as 2*pairs oracle sweeps of the GLOBAL frame, bit for bit."""
    hi = [1.0, 2.0, 3.0]
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 2 * pairs, seed=48)
    out = psi.copy()
    rc = emu.emu_skewed_double_sweeps_slabs(nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), chunk, n_chunks, pairs, nranks)
    assert rc == 0, "deadlock / handshake timeout"
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,spec,n_chunks,levels,chunk,variant", [
    ([20, 9, 32], MIXED_3D, 2, 5, 0, 0),
    ([20, 9, 24], NEUMANN_3D, 3, 6, 3, 0),         # 8-plane chunks, 2*levels > chunk: the upper bands are clipped at the top
    ([130, 13, 20], MIXED_3D, 4, 3, 2, 0),
])
def test_time_skewed_end_schedule_on_the_emulator(emu, n, spec, n_chunks, levels, chunk, variant):
    """The band schedule of the END of the opt-in time-skewed host call (skew_tail_region): every chunk is "downloaded" right after
    its last band and all planes no later band may read are poisoned at once, yet the assembled frame equals 2*levels oracle sweeps."""
    hi = [1.0, 2.0, 3.0]
    g, psi, rhs, want, kinds, vals, dx, nn, coef = _setup(n, hi, spec, 2 * levels, seed=46)
    out = np.zeros_like(psi)
    rc = emu.emu_skewed_tail_double_sweeps(nn, dx, kinds, vals, _p(coef), _p(psi), _p(rhs), _p(out), variant, chunk, n_chunks, levels)
    assert rc == 0, "deadlock or launch failure"
    np.testing.assert_array_equal(out, want)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.integers(1, 1100), st.integers(1, 1100), st.integers(1, 1100), st.booleans(), st.booleans(), st.integers(0, 200))
def test_triple_sweep_chunk_planner_properties(emu, h0, h1, n2, nb_lo, nb_hi, override):
    """plan_tb3 (kernels_sweep_tb3.cu): the chunks cover the plane range, hold at most 128 planes unless overridden, and next to a slab
    neighbour the first and the last chunk hold the three planes that travel (the launcher refuses slabs of fewer than six planes)."""
    if (nb_lo or nb_hi) and n2 < 6:
        n2 += 6
    out = (C.c_int * 3)()
    emu.emu_plan_triple(C.c_longlong(2 * h0), C.c_longlong(2 * h1), C.c_longlong(n2), int(nb_lo), int(nb_hi), override, out)
    chunk, n_chunks, ctas = out[0], out[1], out[2]
    assert chunk >= 1 and n_chunks >= 1 and (n_chunks - 1) * chunk < n2 <= n_chunks * chunk
    assert ctas == n_chunks * ((2 * h0 + 127) // 128) * ((2 * h1 + 11) // 12)
    if override == 0:
        assert chunk <= 128 or (nb_hi and chunk <= 131)
    last = n2 - (n_chunks - 1) * chunk
    if nb_lo:
        assert chunk >= 3
    if nb_hi:
        assert last >= 3


def test_triple_sweep_chunk_planner_choices(emu):
    """The shapes the planner was fitted to (profiles/r2_notes.md, section 9): equal chunks, (waves + 1) x (planes per chunk + 5)."""
    out = (C.c_int * 3)()
    want = {(768, 768, 768): (96, 8), (512, 512, 512): (47, 11), (1024, 1024, 128): (64, 2), (1024, 1024, 1024): (128, 8)}
    for (n0, n1, n2), (chunk, n_chunks) in want.items():
        emu.emu_plan_triple(C.c_longlong(n0), C.c_longlong(n1), C.c_longlong(n2), 0, 0, 0, out)
        assert (out[0], out[1]) == (chunk, n_chunks), ((n0, n1, n2), out[0], out[1])
