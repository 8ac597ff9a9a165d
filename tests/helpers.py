"""Shared test helpers: golden fixtures, comparison rules, GPU availability."""
import os

import numpy as np

from structured_b200.cases import case_from_toml

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12  # north_star: residual / Jacobian parity to 1e-12 relative, fp64


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = case_from_toml(str(z["inp"]), z["xv"], z["yv"])
    return case, z


def golden_state(case, z):
    return z["q"] if "q" in z.files else case.perturbed_q()


def field_rel_err(a, b):
    """per-equation max|a-b| / max|b| (the scale of a residual field is its own max: where fluxes cancel,
    point-wise relative error is meaningless)"""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape
    if a.ndim == 3:
        return np.array([np.abs(a[..., k] - b[..., k]).max() / max(np.abs(b[..., k]).max(), 1e-300) for k in range(a.shape[-1])])
    return np.array([np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)])


def coo_to_dict_arrays(n, ri, ci, va):
    """sum duplicates, return sorted unique keys + values"""
    key = ri.astype(np.int64) * n + ci.astype(np.int64)
    order = np.argsort(key, kind="stable")
    key, va = key[order], va[order]
    uk, start = np.unique(key, return_index=True)
    return uk, np.add.reduceat(va, start) if len(va) else va


def jac_rel_err(n, a, b):
    """SURVEY.md Appendix B comparison rule: maps (row,col)->value with implicit zeros,
    |a-b| <= tol * max(|a|, |b|, s_row) with s_row = max |entry| of the row (of b)."""
    ka, va = coo_to_dict_arrays(n, *a)
    kb, vb = coo_to_dict_arrays(n, *b)
    keys = np.union1d(ka, kb)
    fa = np.zeros(len(keys)); fb = np.zeros(len(keys))
    fa[np.searchsorted(keys, ka)] = va
    fb[np.searchsorted(keys, kb)] = vb
    rows = keys // n
    srow = np.zeros(n)
    np.maximum.at(srow, rows, np.abs(fb))
    scale = np.maximum(np.maximum(np.abs(fa), np.abs(fb)), srow[rows])
    scale[scale == 0] = 1.0
    return float((np.abs(fa - fb) / scale).max()) if len(keys) else 0.0


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
