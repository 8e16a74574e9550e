"""CPU: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(c3d_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_entry_points():
    names = declared_symbols()
    assert "c3d_project_batch" in names and "c3d_knn_batch" in names
    assert len(names) >= 6


def test_library_exports_every_declared_symbol():
    from coarse3d_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    # every declared symbol has a ctypes signature in the binding
    unbound = [n for n in declared_symbols() if n not in _lib.SIGNATURES]
    assert not unbound, unbound
    assert lib.c3d_version() >= 100


def test_no_compute_without_gpu_is_loud():
    import pytest
    import torch
    from coarse3d_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("has GPU")
    pts = torch.zeros((4, 4))
    with pytest.raises(RuntimeError):
        ops.project_batch(pts, torch.tensor([0, 4], dtype=torch.int32), ops.Fov.from_degrees(), 4, 8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "coarse3d_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
