"""rnmf_b200.config.read_lua (the Python mirror of src/config.rs read_lua, config.rs:90-158): the values of the
shipped cases, Lua's arithmetic semantics on the supported subset, the reference's panic messages, and
agreement, value for value, with the C++ twin of the reader (host/include/rnmf/config.hpp)."""
import json
import os
import subprocess

import pytest

import __graft_entry__ as entry
from rnmf_b200 import config as cfgmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SUBSET = ("dim = 2 -- c\nlength = RealVec2(7 % 4 + 1.5, -(-2) ^ 2 * -1)\nn_cells = UIntVec2(7 // 2, (1 + 2) * 3)\n"
          "--[[ block\n comment ]] local t = { a = { b = 5 }, [2] = 9, 11 };\n"
          "model = Model({ H_far = RealVec2(t.a.b, t[2]), bubble_centre = RealVec2(t[1], math.sqrt(16)),\n"
          " mu = RealVec2(1, 2), n_iter = 1, n_sub_iter = 2, relax = 0.5, bubble_radius = math.pi, tol = 1e-3 })\n")


def _as_dict(c):
    m = c.model
    return {"dim": c.geom.dim, "length": list(c.geom.length), "n_cells": list(c.geom.n_cells), "residual_iters": c.residual_iters,
            "model": {"H_far": list(m.h_far), "bubble_centre": list(m.bubble_centre), "bubble_radius": m.bubble_radius,
                      "mu": list(m.mu), "n_iter": m.n_iter, "n_sub_iter": m.n_sub_iter, "relax": m.relax, "tol": m.tol}}


def test_shipped_case_values():
    c = cfgmod.read_lua(os.path.join(ROOT, "cases", "ferro_droplet.lua"))
    # the values of examples/magnetostatics/magnetostatics.lua:1-16, arithmetic evaluated as Lua does (IEEE double)
    assert _as_dict(c) == {"dim": 2, "length": [301.0, 301.0], "n_cells": [301, 301], "residual_iters": 1,
                           "model": {"H_far": [0.0, 1.0], "bubble_centre": [301.0 / 2, 301.0 / 2], "bubble_radius": 301.0 / 20,
                                     "mu": [3.2, 1.0], "n_iter": 10, "n_sub_iter": 550, "relax": 0.2, "tol": 0.1}}


def test_lua_semantics_on_the_subset(tmp_path):
    p = tmp_path / "t.lua"
    p.write_text(SUBSET)
    c = cfgmod.read_lua(str(p))
    assert c.geom.length == (4.5, 4.0) and c.geom.n_cells == (3, 9)          # % floors, -x^2 = -(x^2), // floors
    assert c.model.h_far == (5.0, 9.0) and c.model.bubble_centre == (11.0, 4.0)
    assert c.model.bubble_radius == 3.141592653589793 and c.model.tol == 1e-3


def test_panics_like_the_reference(tmp_path):
    with pytest.raises(cfgmod.ConfigPanic, match="Could not read lua file"):
        cfgmod.read_lua(str(tmp_path / "missing.lua"))
    p = tmp_path / "nomodel.lua"          # the reference's test.lua: no model block
    p.write_text("dim = 2\nlength = RealVec2(10.0, 10.0)\nn_cells = UIntVec2(10, 10)\nx = IntVec2(7, 8)\ny = RealVec2(3.0, 4.0)\n")
    with pytest.raises(cfgmod.ConfigPanic, match="failed reading model parameters"):
        cfgmod.read_lua(str(p))
    p2 = tmp_path / "nokey.lua"
    p2.write_text("dim = 2\nlength = RealVec2(1.0, 1.0)\nn_cells = UIntVec2(4, 4)\nmodel = Model({ H_far = RealVec2(0.0, 1.0) })\n")
    with pytest.raises(cfgmod.ConfigPanic, match="failed reading 'bubble_centre'"):
        cfgmod.read_lua(str(p2))
    p3 = tmp_path / "syntax.lua"
    p3.write_text("dim = = 2\n")
    with pytest.raises(cfgmod.ConfigPanic, match="failed executing lua file"):
        cfgmod.read_lua(str(p3))
    p4 = tmp_path / "neg.lua"
    p4.write_text("dim = 2\nlength = RealVec2(1.0, 1.0)\nn_cells = UIntVec2(-4, 4)\n")
    with pytest.raises(cfgmod.ConfigPanic, match="non-negative"):
        cfgmod.read_lua(str(p4))


def test_agrees_with_the_cpp_reader(tmp_path):
    entry.build_library()
    subprocess.run(["make", "-C", os.path.join(ROOT, "host"), "-s"], check=True)
    p = tmp_path / "t.lua"
    p.write_text(SUBSET)
    for path in (os.path.join(ROOT, "cases", "ferro_droplet.lua"), os.path.join(ROOT, "cases", "small_droplet.lua"), str(p)):
        r = subprocess.run([os.path.join(ROOT, "host", "bin", "config_dump"), path], capture_output=True, text=True, check=True)
        assert json.loads(r.stdout) == _as_dict(cfgmod.read_lua(path)), path
