"""CPU: the C-ABI library loads and exports every symbol include/tortoise_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "tortoise_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ts_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import tortoisesat.jl_b200 as tb
    lib = ctypes.CDLL(tb.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert lib.ts_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tortoisesat.jl_b200 as tb
    with pytest.raises(tb.TortoiseError):
        tb.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tortoisesat.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("liboracle", "orc_", "import oracle", "from oracle", "oracle/"):
                    assert needle not in src, (f, needle)
