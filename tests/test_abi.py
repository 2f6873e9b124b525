"""The C-ABI shared library loads and exports every symbol include/amie_b200.h declares (no GPU)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "amie_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(amie_b200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_no_silent_cpu_path(pkg):
    assert b"sm_100a" in pkg.lib().amie_b200_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is only observable without one")
    # without a device the context cannot be created and the mirror raises: no fallback
    asm = pkg.Synth("S2-tri", 4).assembly()
    with pytest.raises(pkg.AmieB200Error):
        pkg.ConjugateGradient(asm).solve()


def test_stats_struct_layout_matches_header(pkg):
    txt = open(os.path.join(ROOT, "include", "amie_b200.h")).read()
    body = re.search(r"typedef struct amie_b200_stats\s*\{(.*?)\}\s*amie_b200_stats", txt, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, names = decl.split(None, 1)
        for n in names.split(","):
            fields.append((n.strip(), typ))
    assert [f for f, _ in fields] == [f for f, _ in pkg.Stats._fields_]
    assert ctypes.sizeof(pkg.Stats) == 8 * len(fields)


def test_nccl_load_order_keeps_torch_importable():
    """The library dlopens NCCL lazily; in a Python host it must pick the copy torch was built against,
    or a later `import torch` in the same process dies on an undefined NCCL symbol."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import __graft_entry__ as g\n"
        "pkg = g.load_package()\n"
        "assert len(pkg.nccl_unique_id()) == 128\n"
        "import torch\n"
        "import torch.distributed\n"
        "print('ok')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_ctypes_bindings_match_header_prototypes(pkg):
    """Every prototype of include/amie_b200.h is bound in the Python mirror with the same number of parameters and
    the same scalar/pointer kinds (a mismatch would corrupt the call silently)."""
    txt = open(os.path.join(ROOT, "include", "amie_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    L = pkg.lib()
    protos = re.findall(r"\b(amie_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt)
    assert len(protos) >= 45
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "int", ctypes.c_uint64: "u64",
             ctypes.c_int64: "i64", ctypes.c_double: "f64"}
    for name, params in protos:
        params = [p.strip() for p in params.split(",") if p.strip() and p.strip() != "void"]
        want = []
        for p in params:
            if "*" in p:
                want.append("ptr")
            elif p.startswith("uint64_t"):
                want.append("u64")
            elif p.startswith("int64_t"):
                want.append("i64")
            elif p.startswith("double"):
                want.append("f64")
            elif p.startswith("int"):
                want.append("int")
            else:
                raise AssertionError(f"{name}: unrecognised parameter {p!r}")
        fn = getattr(L, name)
        if not want:
            continue
        assert fn.argtypes is not None, f"{name} has no argtypes in the mirror"
        got = [kinds[t] for t in fn.argtypes]
        assert got == want, (name, got, want)
