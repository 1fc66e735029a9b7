"""The band schedules of the opt-in time-skewed host call (skew_region / skew_tail_region, rnmf_b200/csrc/common.cuh) checked
abstractly at full size: every plane of both ping-pong buffers carries the sweep level it holds (-1 = not uploaded / dead), a band
must find exactly the level it expects on every plane it reads (two beyond its output either side) and leaves its output two
levels higher.  The emulator runs the same schedules with real kernels on small frames (test_emulated_kernels); this covers the
benchmark's shape (768 planes, 8 chunks, 56 / 32 launches) and random ones."""
import ctypes as C

import pytest
from hypothesis import given, settings, strategies as st

from test_emulated_kernels import emu  # noqa: F401  (fixture: builds / loads the emulator driver, which exports the schedules)


def _region(lib, tail, n2, chunks, c, s, levels):
    lo, hi = C.c_longlong(), C.c_longlong()
    if tail:
        lib.emu_skew_tail_region(C.c_longlong(n2), chunks, c, C.c_longlong(s), C.c_longlong(levels), C.byref(lo), C.byref(hi))
    else:
        lib.emu_skew_region(C.c_longlong(n2), chunks, c, C.c_longlong(s), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def _check_start(lib, n2, chunks, pairs):
    buf = [[-1] * n2, [-1] * n2]
    covered = [[0] * n2 for _ in range(pairs + 1)]
    for c in range(chunks):
        for z in range(n2 * c // chunks, n2 * (c + 1) // chunks):       # chunk c arrives (valid planes < Z(c+1))
            buf[0][z] = 0
        for s in range(1, pairs + 1):
            lo, hi = _region(lib, False, n2, chunks, c, s, 0)
            assert 0 <= lo and hi <= n2
            if hi <= lo:
                continue
            src, dst = buf[(s - 1) & 1], buf[s & 1]
            for z in range(max(0, lo - 2), min(n2, hi + 2)):
                assert src[z] == 2 * (s - 1), f"band (chunk {c}, pair {s}) reads plane {z} at level {src[z]}"
            for z in range(lo, hi):
                dst[z] = 2 * s
                covered[s][z] += 1
    assert all(v == 2 * pairs for v in buf[pairs & 1])
    assert all(v == 1 for s in range(1, pairs + 1) for v in covered[s])     # bands of one pair tile the frame exactly once


def _check_end(lib, n2, chunks, levels):
    buf = [[0] * n2, [-1] * n2]
    out = [-1] * n2
    for c in range(chunks):
        for s in range(1, levels + 1):
            lo, hi = _region(lib, True, n2, chunks, c, s, levels)
            assert 0 <= lo and hi <= n2
            if hi <= lo:
                continue
            src, dst = buf[(s - 1) & 1], buf[s & 1]
            for z in range(max(0, lo - 2), min(n2, hi + 2)):
                assert src[z] == 2 * (s - 1), f"band (chunk {c}, level {s}) reads plane {z} at level {src[z]}"
            for z in range(lo, hi):
                dst[z] = 2 * s
        z1 = n2 if c == chunks - 1 else n2 * (c + 1) // chunks
        for z in range(n2 * c // chunks, z1):                            # chunk c goes home
            out[z] = buf[levels & 1][z]
        if c < chunks - 1:
            for z in range(0, max(0, z1 - 2)):                           # nothing below Z(c+1)-2 may be read again
                buf[0][z] = buf[1][z] = -99
    assert all(v == 2 * levels for v in out)


@pytest.mark.parametrize("n2,chunks,pairs", [(768, 8, 56), (768, 8, 275), (512, 8, 56), (260, 4, 10), (260, 8, 11), (130, 2, 4), (32, 2, 40)])
def test_start_schedule_dependencies(emu, n2, chunks, pairs):
    _check_start(emu, n2, chunks, pairs)


@pytest.mark.parametrize("n2,chunks,levels", [(768, 8, 32), (768, 8, 200), (512, 8, 32), (260, 4, 5), (130, 2, 5), (32, 2, 30)])
def test_end_schedule_dependencies(emu, n2, chunks, levels):
    _check_end(emu, n2, chunks, levels)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(st.integers(2, 12), st.integers(1, 40), st.integers(1, 60), st.data())
def test_schedules_random_shapes(emu, chunks, per_chunk, launches, data):
    n2 = chunks * per_chunk + data.draw(st.integers(0, chunks - 1))
    _check_start(emu, n2, chunks, launches)
    _check_end(emu, n2, chunks, launches)
