"""The Rust binding crate (bindings/rust) cannot be compiled in this image (no cargo/rustc), so this test
keeps its raw layer honest against the C header: every function of include/rnmf_b200.h is declared in
src/ffi.rs with the same name, argument count and pointer-ness per argument; the POD structs have the same
fields in the same order; the enum constants agree."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def _c_functions():
    src = _strip_c_comments(open(os.path.join(ROOT, "include", "rnmf_b200.h")).read())
    out = {}
    for m in re.finditer(r"^\s*(?:const\s+)?[\w]+\s*\*?\s+(rnmf_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.M | re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        arglist = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        out[name] = [a.count("*") for a in arglist]
    return out


def _rust_functions():
    src = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    block = src[src.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (rnmf_\w+)\s*\(([^)]*)\)", block):
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = [a.count("*") for a in args]
    return out


def test_every_header_function_is_bound_with_the_same_shape():
    c, r = _c_functions(), _rust_functions()
    assert len(c) >= 50, sorted(c)
    assert sorted(c) == sorted(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name in c:
        assert c[name] == r[name], (name, c[name], r[name])        # same arity, same pointer depth per argument


def _c_struct_fields(name):
    src = _strip_c_comments(open(os.path.join(ROOT, "include", "rnmf_b200.h")).read())
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        for part in decl.split(",")[0:1] + [p for p in decl.split(",")[1:]]:
            fields.append(re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")))
    return fields


def _rust_struct_fields(name):
    src = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, flags=re.S).group(1)
    return re.findall(r"pub (\w+)\s*:", body)


def test_pod_structs_have_the_same_fields_in_the_same_order():
    for name in ("rnmf_bc_face", "rnmf_mu_params", "rnmf_frame_info"):
        assert _c_struct_fields(name) == _rust_struct_fields(name), name


def test_enum_constants_agree():
    hdr = _strip_c_comments(open(os.path.join(ROOT, "include", "rnmf_b200.h")).read())
    rs = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    c_vals = {k: int(v) for k, v in re.findall(r"(RNMF_(?:OK|ERR_\w+|BC_\w+))\s*=\s*(-?\d+)", hdr)}
    r_vals = {k: int(v) for k, v in re.findall(r"pub const (RNMF_(?:OK|ERR_\w+|BC_\w+)): i32 = (-?\d+);", rs)}
    assert c_vals and c_vals == r_vals
    assert int(re.search(r"#define RNMF_ABI_VERSION (\d+)", hdr).group(1)) == int(re.search(r"RNMF_ABI_VERSION: i32 = (\d+)", rs).group(1))
