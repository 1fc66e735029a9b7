"""Pins the CPU oracle against the reference's own unit-test vectors (tests/golden/, transcribed
from src/mesh/cartesian2d/dataframe.rs:862-1201 and mesh.rs:153-212).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as ora


def _ic(name):
    return {"x+1": lambda x, y, c: x + 1.0, "x+y": lambda x, y, c: x + y}[name]


def _type_code(s):
    kind, loc = s[:-1].split("(")
    return (ora.GHOST if kind == "Ghost" else 0) + ora.LOC[loc]


def _frame(case, ncomp=1):
    g = ora.geom(case["n"], hi=case["hi"], ng=case["ng"], ncomp=ncomp)
    data = ora.zeros(g)
    if case.get("ic", "none") != "none":
        ora.fill_ic(g, data, _ic(case["ic"]))
    return g, data


def _bc_of(case):
    pres = case.get("prescribed_values")
    fix = lambda f: (f[0], pres if f[0] == "prescribed" else f[1])
    return ora.bcs(2, [([fix(f) for f in case["lo"]], [fix(f) for f in case["hi_bc"]])])


def test_bc_fill_golden_vectors(golden):
    assert len(golden["bc_fill"]) == 5
    for case in golden["bc_fill"]:
        g, data = _frame(case)
        assert ora.fill_bc(g, data, _bc_of(case)) == 0
        np.testing.assert_array_equal(data, np.array(case["data"]), err_msg=case["name"])


def test_reflect_is_unimplemented():
    # dataframe.rs:375-377, :410-412: unimplemented!()
    g = ora.geom([3, 3], hi=[6.0, 6.0])
    faces = ora.bcs(2, [([("reflect", None), ("dirichlet", 0.0)], [("dirichlet", 0.0)] * 2)])
    assert ora.fill_bc(g, ora.zeros(g), faces) == -1


def test_node_type(golden):
    case = golden["node_type"]
    g = ora.geom(case["n"], hi=case["hi"], ng=case["ng"])
    ct = np.zeros(g.n_nodes, dtype=np.uint8)
    ora.lib().ora_classify_2d(C.byref(g), ct)
    assert ct.tolist() == [_type_code(s) for s in case["cell_type"]]


def test_valid_iteration_order(golden):
    case = golden["data_frame_iterator"]
    g, data = _frame(case)
    out = np.zeros(9)
    m = ora.lib().ora_valid_cells(C.byref(g), data, out)
    assert m == 9
    assert out.tolist() == case["valid_order"]


def _iter_sides(g, data, sides):
    codes = np.array([_type_code(s) for s in sides], dtype=np.uint8)
    out = np.zeros(g.n_nodes)
    outp = np.zeros(g.n_nodes, dtype=np.int64)
    m = ora.lib().ora_ghost_iter_2d(C.byref(g), data, codes, len(codes), -1, out, outp)
    return out[:m].tolist()


def test_ghost_iter_two_ghost(golden):
    case = golden["ghost_iter_two_ghost"]
    g, data = _frame(case)
    ora.fill_bc(g, data, _bc_of(case))
    for it in case["iters"]:
        assert _iter_sides(g, data, it["sides"]) == it["values"], it["sides"]


def test_ghost_collect_two_edge(golden):
    # collect_typed_cells (dataframe.rs:273-278): the boundary-adjacent valid strips = halo PACK order
    case = golden["ghost_collect_two_edge"]
    g, data = _frame(case)
    ora.fill_bc(g, data, _bc_of(case))
    for it in case["collect"]:
        assert _iter_sides(g, data, it["sides"]) == it["values"], it["sides"]


def test_mesh_node_pos(golden):
    case = golden["mesh_node_pos"]
    g = ora.geom(case["n"], hi=case["hi"])
    pos = np.zeros(2 * 25)
    ora.lib().ora_cell_centres_2d(C.byref(g), pos)
    assert pos.tolist() == case["node_pos"]


def test_mesh_indexing_and_iterator(golden):
    L = ora.lib()
    case = golden["mesh_indexing"]
    n0, n1 = case["n"]
    n_nodes = n0 * n1 * 2
    for q in case["get_ij"]:
        i, j, c = C.c_int64(), C.c_int64(), C.c_int64()
        ok = L.ora_get_ij_from_p(q["p"], n0, n1, n_nodes, 0, C.byref(i), C.byref(j), C.byref(c))
        if q["ijn"] is None:
            assert ok == 0
        else:
            assert ok == 1 and [i.value, j.value, c.value] == q["ijn"]
    g = ora.geom(case["n"], hi=case["hi"])
    pos = np.zeros(n_nodes)
    L.ora_cell_centres_2d(C.byref(g), pos)
    for q in case["index"]:
        i, j, n = q["ijn"]
        assert pos[i + n0 * j + n0 * n1 * n] == q["value"]     # mod.rs:31-37 with ng=0
    it = golden["mesh_iterator"]
    got = [[pos[p], pos[p + n0 * n1]] for p in range(n0 * n1)]
    assert got == it["positions"]
