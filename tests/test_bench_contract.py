"""bench.py's CPU arm (`--impl reference`) runs here without a GPU and must print ONE JSON line with
the contract's keys; the GPU arm must refuse to run without a device (no fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GLUP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("stencil GLUP/s") and d["dtype"] == "f64" and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["config"]["workload"] == "3d7_768"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_without_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_roofline_accounting_single_and_double_sweep():
    """`roofline.achieved` = SURVEY 8d's 24 B per update x the updates one launch of the dominant kernel performs /
    that launch's share of the timed region: the double-sweep kernel counts two sweeps per launch, so frac may
    exceed 1; dram_frac puts the measured DRAM bytes of the committed ncu capture beside it."""
    b = _bench_module()
    cells = 768 ** 3
    # the committed default line (profiles/r1_bench_default_n1_tb2.json): 5 steps x 550 sweeps, 1370 double launches
    r = b.build_roofline({"ms": 663.694677734375 * 5, "doubles": 1370, "local_cells": cells}, 5, 550, "3d7_768", 6549.8, "MEASURED_PEAKS.json hbm_gbs")
    assert r["sweeps_per_launch"] == 2 and r["algorithmic_bytes_per_launch"] == 2 * 24 * cells
    assert abs(r["avg_launch_ms"] - 2 * 663.694677734375 / 550) < 1e-9
    assert abs(r["frac"] - 1.3755) < 1e-3 and r["frac"] == r["achieved"] / r["peak"]
    assert r["traffic"] and abs(r["dram_frac"] - r["traffic"] / (r["avg_launch_ms"] * 1e-3) / 1e9 / 6549.8) < 1e-12
    assert r["dram_frac"] < 1.0 < r["frac"]
    assert abs(r["frac_of_k_pass_ceiling"] - r["frac"] / 2) < 1e-12 and r["traffic_source"]
    # with the dominant kernel timed directly (CUDA events around back-to-back launches) that duration replaces the step share
    rk = b.build_roofline({"ms": 663.694677734375 * 5, "doubles": 1370, "local_cells": cells, "kernel_ms": 2.30, "kernel_launches_timed": 60},
                          5, 550, "3d7_768", 6549.8, "MEASURED_PEAKS.json hbm_gbs")
    assert rk["avg_launch_ms"] == 2.30 and abs(rk["avg_launch_ms_from_step"] - r["avg_launch_ms"]) < 1e-12 and "60 back-to-back" in rk["launch_timing"]
    assert abs(rk["achieved"] - 2 * 24 * cells / 2.30e-3 / 1e9) < 1e-6
    # single-sweep kernel only (variant 6): one sweep per launch, no note
    r1 = b.build_roofline({"ms": 965.5 * 5, "doubles": 0, "local_cells": cells}, 5, 550, "3d7_768", 6549.8, "MEASURED_PEAKS.json hbm_gbs")
    assert r1["sweeps_per_launch"] == 1 and r1["algorithmic_bytes_per_launch"] == 24 * cells and "note" not in r1
    assert abs(r1["frac"] - 0.9455) < 1e-3
    # the same line with the launch statistics the library now reports: unchanged accounting for the double-sweep kernel
    r2 = b.build_roofline({"ms": 663.694677734375 * 5, "doubles": 1370, "multi_launches": 1370, "multi_sweeps": 2740, "local_cells": cells},
                          5, 550, "3d7_768", 6549.8, "MEASURED_PEAKS.json hbm_gbs")
    assert r2["sweeps_per_launch"] == 2 and r2["frac"] == r["frac"] and r2["traffic"] == r["traffic"]
    # the four-sweep 2D kernel: 480 sweeps per step in launches of four; its DRAM bytes come from the round-2 capture
    c2 = 8192 ** 2
    r4 = b.build_roofline({"ms": 100.0, "doubles": 0, "multi_launches": 360, "multi_sweeps": 1440, "local_cells": c2}, 3, 480, "2d5_8192", 6549.8, "x")
    assert r4["sweeps_per_launch"] == 4 and r4["algorithmic_bytes_per_launch"] == 4 * 24 * c2 and "multi sweep (4 sweeps" in r4["kernel"]
    assert abs(r4["avg_launch_ms"] - 4 * 100.0 / 1440) < 1e-12 and r4["traffic"] and "r2_ncu_2d5_8192_tbk" in r4["traffic_source"]
    # the three-sweep 3D kernel (the 3D default): 549 of the 550 sweeps of a step in 183 launches of three; DRAM bytes from its capture
    r3 = b.build_roofline({"ms": 508.0 * 5, "doubles": 0, "multi_launches": 915, "multi_sweeps": 2745, "local_cells": cells, "kernel_ms": 2.78,
                           "kernel_launches_timed": 60}, 5, 550, "3d7_768", 6549.8, "MEASURED_PEAKS.json hbm_gbs")
    assert r3["sweeps_per_launch"] == 3 and r3["algorithmic_bytes_per_launch"] == 3 * 24 * cells and "multi sweep (3 sweeps" in r3["kernel"]
    assert r3["traffic"] and "r2_ncu_3d7_768_tb3" in r3["traffic_source"] and 0.5 < r3["dram_frac"] < 1.0 < r3["frac"]
    assert abs(r3["frac_of_k_pass_ceiling"] - r3["frac"] / 3) < 1e-12
    # both describe the same GLUP/s <-> GB/s relation: achieved / 24 B = updates per second
    assert abs(r["achieved"] / 24 - cells * 550 / (663.694677734375e-3) / 1e9) < 1e-6 * r["achieved"]


def test_gpu_suite_collects_on_the_cpu_box():
    """The GPU tests cannot run here, but a syntax or import error in them would only show at round end."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "--collect-only", "-q", "-m", "gpu"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "error" not in r.stdout.lower().splitlines()[-1]
    n = int(r.stdout.strip().splitlines()[-1].split("/")[0].split()[0])
    assert n >= 100
