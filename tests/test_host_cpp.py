"""The C++ host layer (host/): the Lua-subset case reader (config.rs mirror), the solver Stage/Step/
Stop semantics (solver.rs mirror, defects fixed) -- CPU only -- and, on a GPU, the magnetostatics
example binary end to end on .lua cases against the oracle."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "bin")


@pytest.fixture(scope="module")
def host_bins():
    entry.build_library()
    subprocess.run(["make", "-C", os.path.join(ROOT, "host"), "-s"], check=True)
    return BIN


def _dump(host_bins, path):
    r = subprocess.run([os.path.join(host_bins, "config_dump"), path], capture_output=True, text=True)
    return r.returncode, (json.loads(r.stdout) if r.returncode == 0 else r.stderr)


def test_lua_reader_shipped_case(host_bins):
    rc, cfg = _dump(host_bins, os.path.join(ROOT, "cases", "ferro_droplet.lua"))
    assert rc == 0
    # values of examples/magnetostatics/magnetostatics.lua:1-16, arithmetic evaluated as Lua does (IEEE double)
    assert cfg == {"dim": 2, "length": [301.0, 301.0], "n_cells": [301, 301], "residual_iters": 1,
                   "model": {"H_far": [0.0, 1.0], "bubble_centre": [301.0 / 2, 301.0 / 2], "bubble_radius": 301.0 / 20,
                             "mu": [3.2, 1.0], "n_iter": 10, "n_sub_iter": 550, "relax": 0.2, "tol": 0.1}}


def test_lua_reader_syntax_subset(host_bins, tmp_path):
    rc, cfg = _dump(host_bins, os.path.join(ROOT, "cases", "small_droplet.lua"))
    assert rc == 0
    assert cfg["n_cells"] == [33, 33] and cfg["length"] == [33.0, 33.0]
    m = cfg["model"]
    assert m["bubble_centre"] == [16.5, 16.5] and m["bubble_radius"] == 4.0 and m["n_sub_iter"] == 25 and m["relax"] == 1 / 5
    p = tmp_path / "t.lua"
    p.write_text("dim = 2 -- c\nlength = RealVec2(7 % 4 + 1.5, -(-2) ^ 2 * -1)\nn_cells = UIntVec2(7 // 2, (1 + 2) * 3)\n"
                 "t = { a = { b = 5 }, [2] = 9, 11 }\nmodel = Model({ H_far = RealVec2(t.a.b, t[2]), bubble_centre = RealVec2(t[1], math.sqrt(16)),\n"
                 " mu = RealVec2(1, 2), n_iter = 1, n_sub_iter = 2, relax = 0.5, bubble_radius = math.pi, tol = 1e-3 })\n")
    rc, cfg = _dump(host_bins, str(p))
    assert rc == 0
    assert cfg["length"] == [4.5, 4.0] and cfg["n_cells"] == [3, 9]
    assert cfg["model"]["H_far"] == [5.0, 9.0] and cfg["model"]["bubble_centre"] == [11.0, 4.0]
    assert cfg["model"]["bubble_radius"] == 3.141592653589793


def test_lua_reader_error_behaviour(host_bins, tmp_path):
    # the reference panics (config.rs:97,135,153; model.rs:39-54): exit code 101 and the same messages
    rc, err = _dump(host_bins, str(tmp_path / "missing.lua"))
    assert rc == 101 and "Could not read lua file" in err
    p = tmp_path / "nomodel.lua"
    p.write_text("dim = 2\nlength = RealVec2(10.0, 10.0)\nn_cells = UIntVec2(10, 10)\nx = IntVec2(7, 8)\ny = RealVec2(3.0, 4.0)\n")   # test.lua's content
    rc, err = _dump(host_bins, str(p))
    assert rc == 101 and "failed reading model parameters" in err
    p2 = tmp_path / "nokey.lua"
    p2.write_text("dim = 2\nlength = RealVec2(1.0, 1.0)\nn_cells = UIntVec2(4, 4)\nmodel = Model({ H_far = RealVec2(0.0, 1.0) })\n")
    rc, err = _dump(host_bins, str(p2))
    assert rc == 101 and "failed reading 'bubble_centre'" in err
    p3 = tmp_path / "syntax.lua"
    p3.write_text("dim = = 2\n")
    rc, err = _dump(host_bins, str(p3))
    assert rc == 101 and "failed executing lua file" in err


def test_solver_stage_semantics(host_bins):
    out = subprocess.run([os.path.join(host_bins, "solver_selftest")], capture_output=True, text=True, check=True).stdout.splitlines()
    a, b = json.loads(out[0]), json.loads(out[1])
    # inner stage: value halves per iteration from 100, stop when < 1 checked every 3rd iteration:
    # first run 9 iterations / 3 checks, second run 3 iterations / 1 check; outer stage runs twice
    assert a == {"calls": 12, "checks": 4, "inner_last_run": 3, "outer_run": 2}
    assert b == {"plain": 7}


def _run_example(host_bins, case, *extra):
    r = subprocess.run([os.path.join(host_bins, "magnetostatics"), case, "--json", *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line), r.stdout


@pytest.mark.gpu
def test_example_small_case_matches_oracle(host_bins, tmp_path):
    from oracle import oracle as ora
    res, out = _run_example(host_bins, os.path.join(ROOT, "cases", "small_droplet.lua"), "--out", str(tmp_path / "run"))
    g = ora.geom([33, 33], hi=[33.0, 33.0])
    m = ora.model(n_iter=4, n_sub_iter=25, bubble_centre=(16.5, 16.5), bubble_radius=4.0, relax=1 / 5)
    ref, tr, it = ora.zeros(g), np.zeros(4), C.c_int64()
    ora.lib().ora_magnetostatics_run(C.byref(g), ref, C.byref(m), 0, tr, C.byref(it))
    assert res["iters"] == it.value
    for a, b in zip(res["sum_diff"], tr):
        assert abs(a - b) <= 1e-12 * abs(b)
    assert res["psi_centre"] == ora.as_grid(g, ref)[0, 17, 17]          # valid (16,16)
    assert "sum diff = " in out and "Done." in out
    # write_vtk: iter 0 and iter n_iter files with psi and h arrays (io.rs:97-198, main.rs:60,77)
    for it_no in (0, 4):
        txt = open(str(tmp_path / f"run_ferro_droplet_{it_no}.vti")).read()
        assert 'Name="psi"' in txt and 'Name="h" NumberOfComponents="2"' in txt and 'WholeExtent="0 33 0 33 0 0"' in txt


@pytest.mark.gpu
def test_example_shipped_lua_case(host_bins):
    with open(os.path.join(ROOT, "tests", "golden", "magnetostatics_trace.json")) as fh:
        gold = json.load(fh)
    res, out = _run_example(host_bins, os.path.join(ROOT, "cases", "ferro_droplet.lua"), "--no-vtk")
    assert res["iters"] == gold["iters"]
    for a, b in zip(res["sum_diff"], gold["sum_diff"]):
        assert abs(a - b) <= 1e-12 * b
    assert res["psi_centre"] == gold["psi_150_150"]
    # the throughput (Jacobi) order runs the same case; different update order => different numbers, same scale
    res_j, _ = _run_example(host_bins, os.path.join(ROOT, "cases", "ferro_droplet.lua"), "--no-vtk", "--order", "jacobi")
    assert res_j["iters"] == 10 and 0.2 < res_j["sum_diff"][-1] / gold["sum_diff"][-1] < 5.0


@pytest.mark.gpu
def test_solver_stage_drives_gpu_sweeps(host_bins):
    """solver.rs's Stage/Step/GlobalStop shape around the device sweep: stop when the L1 update norm
    (fused in the sweep kernel, read back only every check_every iterations) drops below tol.  The
    oracle iterated the same way must stop at the same iteration with the same field."""
    from oracle import oracle as ora
    n, max_iters, every, tol = 48, 4000, 25, 1.0
    r = subprocess.run([os.path.join(host_bins, "stage_gpu"), str(n), str(max_iters), str(every), str(tol)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    # the batched stage (one device call per check: sweeps paired up / kept L2-resident) stops at the same sweep
    # with the same field and norm
    rb = subprocess.run([os.path.join(host_bins, "stage_gpu"), str(n), str(max_iters), str(every), str(tol), "batched"],
                        capture_output=True, text=True, timeout=300)
    assert rb.returncode == 0, rb.stdout + rb.stderr
    resb = json.loads(rb.stdout.strip().splitlines()[-1])
    assert resb["iterations"] == res["iterations"] and resb["psi_mid"] == res["psi_mid"]
    assert abs(resb["norm"] - res["norm"]) <= 1e-12 * res["norm"]
    g = ora.geom([n, n], hi=[1.0, 1.0])
    bc = ora.bcs(2, [([("dirichlet", 0.0), ("dirichlet", 0.0)], [("dirichlet", 1.0), ("dirichlet", 1.0)])])
    psi, rhs = ora.zeros(g), ora.zeros(g)
    ora.fill_ic(g, rhs, lambda x, y, c: 8.0 * x * (1.0 - y))
    ora.fill_bc(g, psi, bc)
    it, norm = 0, np.inf
    while it < max_iters:
        norm = ora.jacobi(g, psi, rhs, bc, every)      # norm of the last of `every` sweeps
        it += every
        if norm < tol:
            break
    assert res["iterations"] == it and it < max_iters
    assert abs(res["norm"] - norm) <= 1e-12 * norm
    assert res["psi_mid"] == ora.as_grid(g, psi)[0, n // 2 + 1, n // 2 + 1]


@pytest.mark.gpu
def test_poisson3d_example_through_c_abi(host_bins):
    """The 3D mixed-BC config driven from C++ through the C ABI gives the same norm and checksum as
    the oracle on the same synthetic inputs (64^3, 20 sweeps after a 10-sweep warm-up)."""
    from oracle import oracle as ora
    r = subprocess.run([os.path.join(host_bins, "poisson3d"), "64", "20", "10"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    n = [64, 64, 64]
    g = ora.geom(n, hi=[64.0] * 3)
    spec = ([("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 0.25)], [("dirichlet", 1.0), ("neumann", -0.5), ("neumann", 0.5)])
    bc = ora.bcs(3, [spec])
    psi, rhs = ora.zeros(g), ora.zeros(g)
    ora.valid_view(g, psi)[0] = ora.splitmix_field(64 ** 3, 0x5EED0001).reshape(64, 64, 64)
    ora.valid_view(g, rhs)[0] = ora.splitmix_field(64 ** 3, 0x5EED0002).reshape(64, 64, 64)
    ora.fill_bc(g, psi, bc)
    l1 = ora.jacobi(g, psi, rhs, bc, 30)
    assert res["sweeps"] == 20 and res["glups"] > 0
    assert abs(res["l1_update"] - l1) <= 1e-12 * l1
    want = ora.lib().ora_sum_abs_valid_ld(C.byref(g), psi)
    assert abs(res["abs_sum"] - want) <= 1e-12 * want
