"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports exactly the entry
points include/point2cyl.h declares.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

from point2cyl_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "point2cyl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2c_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_declares_something():
    syms = declared_symbols()
    assert "p2c_fps" in syms and "p2c_linear" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/point2cyl.h but not exported"


def test_binding_table_matches_header(lib_path):
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.p2c_version() >= 1
    assert lib.p2c_arch() == b"sm_100a"
    assert lib.p2c_segfit_stats_stride(8) == 8 * 8 + 19 * 8 + 2


def test_no_cpu_fallback():
    import torch
    from point2cyl_b200 import ops
    with pytest.raises(_lib.P2CError):
        ops.fps(torch.zeros(1, 16, 3), 4, torch.zeros(1, dtype=torch.long))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under point2cyl_b200/ or tools/ may import it (only tests/,
    __graft_entry__.smoke() and bench.py's CPU legs do)."""
    bad = []
    for top in ("point2cyl_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py"):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
