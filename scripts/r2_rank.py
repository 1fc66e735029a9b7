#!/usr/bin/env python3
"""Run here (no GPU) after `gpurun -- 'bash scripts/gpu_r2_experiments.sh'`: ranks the variant benches in gpurun_out/r2exp_*.json per
workload, shows clocks / power, compares the e2e lines (value, checksums) and lists the parity-test outcome."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2exp_*.json"))):
    name = os.path.basename(f)[6:-5]
    try:
        d = json.load(open(f))
    except Exception as e:                       # a failed run leaves an empty / partial file
        err = open(f[:-5] + ".err").read()[-300:] if os.path.exists(f[:-5] + ".err") else ""
        print(f"{name:24s} FAILED ({e}) {err.strip().splitlines()[-1] if err.strip() else ''}")
        continue
    rows.setdefault(d["config"]["workload"], []).append((name, d))
for wl, lst in rows.items():
    print(f"== {wl}")
    base = next((d["value"] for n, d in lst if re.search(r"_v0$|_v30_|_plain$", n)), None)
    for name, d in sorted(lst, key=lambda x: -x[1]["value"]):
        c = d.get("clocks", {})
        e2e = d.get("e2e") or {}
        rel = f"{d['value'] / base:5.2f}x" if base else "     "
        line = f"  {name:22s} {d['value']:8.1f} GLUP/s {rel}  {d['ms_per_step']:8.2f} ms/step  sm {c.get('sm_mhz')} MHz {c.get('power_w_max')} W {c.get('reasons')}"
        if e2e.get("value"):
            line += f"  | e2e {e2e['value']:.1f} GLUP/s {e2e.get('ms_per_step', 0):.1f} ms {e2e.get('checks')}"
        print(line)
log = os.path.join(ROOT, "gpurun_out", "r2exp_pytest.log")
if os.path.exists(log):
    print("== parity (RNMF_EXPERIMENTAL=1)")
    print(open(log).read().rstrip())
