"""The drop-in boundary: libflgp_b200.so loads here (no GPU) and exports every symbol include/flgp.h declares;
without a device it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "flgp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flgp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(flgp):
    lib = flgp._lib.load()
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    # and the ctypes mirror declares a signature for each of them
    assert set(flgp._lib.SIGNATURES) == set(syms)


def _classify_c(arg: str) -> str:
    """Coarse class of one C parameter of include/flgp.h."""
    a = re.sub(r"\bconst\b", "", arg).strip()
    if "*" in a or "flgp_objective_fn" in a:
        if re.search(r"\bchar\s*\*", a):
            return "str"
        return "ptr"
    base = a.rsplit(None, 1)[0].strip() if len(a.split()) > 1 else a
    return {"int": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "size_t": "u64"}[base]   # size_t = 64-bit here


def _classify_ctypes(t) -> str:
    if t is C.c_char_p:
        return "str"
    if t in (C.c_int, C.c_int32, C.c_uint):
        return "i32"
    if t is C.c_int64:
        return "i64"
    if t is C.c_uint64:
        return "u64"
    if t is C.c_double:
        return "f64"
    return "ptr"   # POINTER(...), c_void_p, CFUNCTYPE


def test_ctypes_signatures_match_the_header(flgp):
    """Every prototype of include/flgp.h against the ctypes mirror (flgp_b200/_lib.py): same number of parameters and
    the same class (pointer / string / int / int64 / uint64 or size_t / double) in every position — a mismatch here
    would corrupt the call frame silently."""
    flgp._lib.load()
    text = open(os.path.join(ROOT, "include", "flgp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b[A-Za-z_][A-Za-z0-9_ \*]*?\b(flgp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    seen = {}
    for name, args in protos:
        args = " ".join(args.split())
        seen[name] = [] if args in ("", "void") else [_classify_c(a) for a in args.split(",")]
    assert set(seen) == set(flgp._lib.SIGNATURES)
    for name, (_, argtypes) in flgp._lib.SIGNATURES.items():
        got = [_classify_ctypes(t) for t in argtypes]
        assert got == seen[name], (name, got, seen[name])


def _call_arities(text, name_re=r"flgp_[a-z0-9_]+"):
    """(name, number of top-level arguments) for every call `name(...)` in a C++ source text."""
    out = []
    for m in re.finditer(r"\b(" + name_re + r")\s*\(", text):
        depth, i, commas, empty = 1, m.end(), 0, True
        while i < len(text) and depth:
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                commas += 1
            if depth and not ch.isspace():
                empty = False
            i += 1
        out.append((m.group(1), 0 if empty else commas + 1))
    return out


def test_r_shim_calls_match_the_header_arity():
    """The Rcpp shim (flgp_b200/r_shim/flgp_shim.cpp) cannot be compiled here (no R toolchain): at least every call it
    makes into the C ABI passes as many arguments as include/flgp.h declares."""
    text = open(os.path.join(ROOT, "include", "flgp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    want = {}
    for name, args in re.findall(r"\b(flgp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = " ".join(args.split())
        want[name] = 0 if args in ("", "void") else len(args.split(","))
    shim = open(os.path.join(ROOT, "flgp_b200", "r_shim", "flgp_shim.cpp")).read()
    shim = re.sub(r"//[^\n]*", "", shim)
    shim = re.sub(r"/\*.*?\*/", "", shim, flags=re.S)
    calls = [(n, k) for n, k in _call_arities(shim) if n in want]
    assert len(calls) >= 30
    bad = [(n, k, want[n]) for n, k in calls if k != want[n]]
    assert not bad, bad


def test_r_shim_typechecks_against_the_reference_headers():
    """flgp_b200/r_shim/flgp_shim.cpp parsed and type-checked by g++ together with the reference's OWN headers
    (Spectrum.h, Utils.h, lae.h, Predict.h, train.h, MultiClassification.h: every [[Rcpp::export]] prototype the shim
    re-defines, EigenPair, ReturnValue, MultiClassifier) and include/flgp.h.  R, Rcpp and Eigen are not in the image:
    <RcppEigen.h> is the type-level stand-in of tests/r_mock/ (interfaces only, nothing is linked or run).  Catches
    undeclared names, clashes with the reference's declarations and wrong argument types at the C ABI."""
    import shutil
    import subprocess

    ref = "/root/reference/src"
    if not os.path.exists(os.path.join(ref, "Spectrum.h")) or not shutil.which("g++"):
        pytest.skip("the reference headers are not on this machine")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wmissing-declarations", "-I", os.path.join(ROOT, "tests", "r_mock"),
           "-I", ref, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "flgp_b200", "r_shim", "flgp_shim.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, LC_ALL="C"))
    assert p.returncode == 0, p.stderr[-3000:]
    # every function the shim defines at namespace scope is the definition of a prototype in the reference's headers
    # (same name AND same parameter types — otherwise it would be a new overload and R would still call the old body);
    # the only definitions without one are the three helpers the shim adds for the multi-class drivers
    undeclared = set(re.findall(r"no previous declaration for '[^']*?\b([A-Za-z_][A-Za-z0-9_]*)\(", p.stderr))
    assert undeclared == {"train_logit_mult_on_handle", "se_logit_mult_grid", "nystrom_logit_mult_grid"}, p.stderr[-3000:]


def test_header_is_plain_c99(tmp_path):
    """include/flgp.h is the C ABI: it must compile as C (cgo / .Call / ctypes bindings parse it as such)."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no C compiler")
    src = tmp_path / "abi.c"
    src.write_text('#include "flgp.h"\nint main(void) { return flgp_version(); }\n')
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                        os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every symbol of include/flgp.h to the reference seam it replaces (or says it has none)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in _header_symbols() if ("`" + s + "`") not in doc]
    assert not missing, missing


def test_no_silent_fallback_without_gpu(flgp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(flgp.FlgpError, match="no CPU fallback"):
        flgp.Context(0)


def test_default_init_is_deterministic_sorted_distinct(flgp):
    a = flgp.default_init(100000, 500, seed=7)
    b = flgp.default_init(100000, 500, seed=7)
    c = flgp.default_init(100000, 500, seed=8)
    assert (a == b).all() and (a != c).any()
    assert (a[1:] > a[:-1]).all() and a.min() >= 0 and a.max() < 100000
    with pytest.raises(flgp.FlgpError):
        flgp.default_init(10, 11)


def test_product_never_imports_the_oracle():
    """The product package and the C sources must not reference oracle/ (③: oracle is test infrastructure)."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "flgp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                t = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+oracle|#include\s*[<\"][^>\"]*oracle|libflgp_oracle|dlopen\([^)]*oracle",
                             t, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm) must put exactly one JSON line on stdout, whatever libraries print."""
    import json
    import subprocess
    import sys

    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "6000", "--iter-max", "3"], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0
