"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol include/structured_gpu.h declares."""
import ctypes
import os
import re

from structured_b200 import api, build

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "structured_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgpu_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(api.SYMBOLS) == syms


def test_create_fails_loudly_without_gpu_or_with_bad_args():
    """no CPU fallback: without a device sgpu_create returns an error and a message"""
    import numpy as np
    from structured_b200.cases import zoo_case
    L = api.load_library()
    case = zoo_case("A")
    d, keep = api.make_desc(case)
    d.order = 3
    h = ctypes.c_void_p()
    assert L.sgpu_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"Reconstruction not found" in L.sgpu_last_error(None)
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if not has:
        d, keep = api.make_desc(case)
        assert L.sgpu_create(ctypes.byref(d), ctypes.byref(h)) == -2        # SGPU_ERR_CUDA
        assert L.sgpu_last_error(None)
        try:
            api.GpuEulerEquation(case)
            raise AssertionError("expected SgpuError")
        except api.SgpuError:
            pass
