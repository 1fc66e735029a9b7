"""Regenerates tests/golden/magnetostatics_trace.json by running the CPU oracle on
examples/magnetostatics/magnetostatics.lua's parameters (301x301, 10 x 550 GS sweeps).
The values equal, digit for digit, those of the independent survey-time C++ restatement
recorded in BASELINE.md S2.  (The Rust reference itself cannot be built in this image.)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as ora  # noqa: E402

g = ora.geom([301, 301], hi=[301.0, 301.0])
m = ora.model()
psi, tr, it = ora.zeros(g), np.zeros(10), C.c_int64()
assert ora.lib().ora_magnetostatics_run(C.byref(g), psi, C.byref(m), 0, tr, C.byref(it)) == 0
out = {"source": "oracle run of examples/magnetostatics/magnetostatics.lua:1-16 via main.rs:29-74, poisson.rs:13-100",
       "iters": it.value, "sum_diff": [float(v) for v in tr],
       "psi_150_150": float(ora.as_grid(g, psi)[0, 151, 151]),
       "psi_valid_abs_sum": float(ora.lib().ora_sum_abs_valid_ld(C.byref(g), psi))}
with open(os.path.join(os.path.dirname(__file__), "magnetostatics_trace.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
