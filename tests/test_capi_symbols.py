"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rnmf_b200.h declares,
the host-only entry points work, and compute entry points fail loudly without a device (no
fallback).  No GPU needed."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    entry.build_library()
    from rnmf_b200 import _capi
    return _capi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rnmf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rnmf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(capi):
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/rnmf_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == declared, "ctypes signature table and header disagree"


def test_abi_version_and_build_info(capi):
    L = capi.lib()
    assert L.rnmf_abi_version() == 1
    info = L.rnmf_build_info().decode()
    assert "sm_100a" in info and "fmad=false" in info


def test_slab_partition_matches_reference_rule(capi):
    # Domain::new_rectangular (cartesian2d/mod.rs:64-73): floor(n/P) each, remainder to the last
    from rnmf_b200 import slab_partition
    assert [slab_partition(1024, 8, r) for r in range(8)] == [(128 * r, 128) for r in range(8)]
    assert [slab_partition(10, 4, r) for r in range(4)] == [(0, 2), (2, 2), (4, 2), (6, 4)]
    assert slab_partition(5, 1, 0) == (0, 5)
    with pytest.raises(capi.RnmfError):
        slab_partition(5, 2, 2)


def test_poisson_coefs_host_function(capi):
    from rnmf_b200.poisson import poisson_coefs
    from oracle import oracle as ora
    for n, hi in (([301, 301], [301.0, 301.0]), ([10, 20], [3.0, 7.0]), ([8, 9, 10], [1.0, 2.0, 3.0])):
        g = ora.geom(n, hi=hi)
        assert poisson_coefs(g.dim, [g.dx[d] for d in range(g.dim)]) == ora.coefs(g).tolist()


def test_compute_fails_loudly_without_device(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from rnmf_b200 import Context
    with pytest.raises(capi.RnmfError) as e:
        Context(0)
    assert e.value.code == capi.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_code_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under rnmf_b200/ or host/ may reference it
    bad = []
    for base in ("rnmf_b200", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_obj" in dirpath:
                continue
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c")):
                    txt = open(os.path.join(dirpath, fn), errors="replace").read()
                    if re.search(r"\boracle\b|liboracle|ora_[a-z]", txt):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_the_kernel_emulator_is_never_part_of_the_product(capi):
    # tests/emu/ (a thread-per-lane SIMT emulator for the kernel sources) is test infrastructure: the product build never
    # defines RNMF_EMULATE, the emulator header lives outside the package, and the library exports / contains none of it
    import subprocess
    mk = open(os.path.join(ROOT, "rnmf_b200", "csrc", "Makefile")).read() + open(os.path.join(ROOT, "host", "Makefile")).read()
    assert "RNMF_EMULATE" not in mk and "tests/emu" not in mk
    assert not os.path.exists(os.path.join(ROOT, "rnmf_b200", "csrc", "emu"))
    so = os.path.join(ROOT, "rnmf_b200", "librnmf_b200.so")
    syms = subprocess.run(["nm", "-C", so], capture_output=True, text=True).stdout
    assert "emu::" not in syms and "emu_launch" not in syms
    for fn in os.listdir(os.path.join(ROOT, "rnmf_b200")):
        if fn.endswith(".py"):
            assert "emu_driver" not in open(os.path.join(ROOT, "rnmf_b200", fn)).read()
