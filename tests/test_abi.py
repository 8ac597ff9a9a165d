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


def test_ctypes_struct_layouts_match_the_c_header(tmp_path):
    """the ctypes mirrors of sgpu_bc / sgpu_desc / sgpu_linsolve have the C compiler's sizes and field offsets, and the
    enum values the Python side hard-codes are the header's"""
    import subprocess
    structs = {"sgpu_bc": api.SgpuBc, "sgpu_desc": api.SgpuDesc, "sgpu_linsolve": api.SgpuLinsolve}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "structured_gpu.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    for e in ("SGPU_MAT_LHS", "SGPU_MAT_J", "SGPU_MAT_JT", "SGPU_MAT_LHS_T", "SGPU_PC_BLOCK_JACOBI", "SGPU_PC_LINE_J",
              "SGPU_FLUX_ROE", "SGPU_FLUX_AUSM", "SGPU_STATE_Q", "SGPU_STATE_Q_TMP"):
        lines.append('printf("enum %s %%d\\n", (int)%s);' % (e, e))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    enums = {}
    for ln in out:
        t = ln.split()
        if not t:
            continue
        if t[0] == "enum":
            enums[t[1]] = int(t[2])
        elif t[1] == "sizeof":
            assert ctypes.sizeof(structs[t[0]]) == int(t[2]), ln
        else:
            assert getattr(structs[t[0]], t[1]).offset == int(t[2]), ln
    assert [enums[k] for k in ("SGPU_MAT_LHS", "SGPU_MAT_J", "SGPU_MAT_JT", "SGPU_MAT_LHS_T")] == [api.MATRICES[k] for k in ("lhs", "J", "JT", "lhsT")]
    assert [enums["SGPU_PC_BLOCK_JACOBI"], enums["SGPU_PC_LINE_J"]] == [api.PRECONDS["block_jacobi"], api.PRECONDS["line_j"]]
    assert [enums["SGPU_FLUX_ROE"], enums["SGPU_FLUX_AUSM"]] == [api.FLUXES["roe"], api.FLUXES["ausm"]]
