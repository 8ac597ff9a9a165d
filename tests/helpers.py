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


def crop_case(case, ia, ib, ja, jb, wall_distance=None):
    """Window [ia, ib) x [ja, jb) of a large case as a small case for the CPU oracle: physical boundary tables are kept
    (shifted, clipped) where the window touches that boundary, cut edges get `freestream` ghosts.  Residual and Jacobian
    rows of cells at least 3 cells away from every CUT edge equal those of the large case (stencil radius 2 + the corner
    ghosts a cut edge's BC cannot reproduce); the limiter constants stay those of the large grid (`global_counts`).
    Periodic / wake tables cannot be cropped: the window must stay 3 cells away from such boundaries."""
    import copy
    from structured_b200.cases import Boundary
    c = copy.copy(case)
    c.ni, c.nj = ib - ia + 1, jb - ja + 1
    c.xv = np.ascontiguousarray(case.xv[ia:ib + 1, ja:jb + 1]); c.yv = np.ascontiguousarray(case.yv[ia:ib + 1, ja:jb + 1])
    wd = case.wall_distance if wall_distance is None else wall_distance
    c.wall_distance = None if wd is None else np.ascontiguousarray(wd[ia:ib, ja:jb])
    c.beta = None if case.beta is None else np.ascontiguousarray(case.beta[ia:ib, ja:jb])
    c.global_counts = (case.nic, case.njc)
    c.window = None
    touch = {"bottom": ja == 0, "top": jb == case.njc, "left": ia == 0, "right": ib == case.nic}
    bcs = []
    for b in case.boundaries:
        assert b.type not in ("periodic", "wake") or not touch[b.face], "periodic / wake boundaries cannot be cropped"
        if not touch[b.face]:
            continue
        horiz = b.face in ("bottom", "top")
        n_old, off, n_new = (case.nic, ia, ib - ia) if horiz else (case.njc, ja, jb - ja)
        end = b.end if b.end >= 0 else n_old + 2 + b.end
        s, e = max(b.start - off, 0), min(end - off, n_new + 1)
        if e >= s:
            bcs.append(Boundary(b.type, b.face, s, e, b.u, b.v, b.T))
    for f in ("bottom", "top", "left", "right"):
        if not touch[f]:
            bcs.append(Boundary("freestream", f, 0, -1))
    c.boundaries = bcs
    return c


def rows_of_cells(njc, nv, cells):
    """flat Jacobian row numbers (i*njc + j)*nv + k of a list of cells"""
    cells = np.asarray(cells, dtype=np.int64).reshape(-1, 2)
    return ((cells[:, 0] * njc + cells[:, 1])[:, None] * nv + np.arange(nv)[None, :]).reshape(-1)


def crop_interior(case, box):
    """cells of a crop box that are >= 3 cells away from every cut edge: (i0, i1, j0, j1), half open"""
    ia, ib, ja, jb = box
    return (ia if ia == 0 else ia + 3, ib if ib == case.nic else ib - 3, ja if ja == 0 else ja + 3, jb if jb == case.njc else jb - 3)


def oracle_on_crop(case, q, box, wall_distance=None, lhs=True, want_jacobian=True):
    """CPU oracle on a window of a LARGE case: returns (residual of the window's trustworthy cells [i0:i1, j0:j1],
    (i0, i1, j0, j1), Jacobian rows of those cells as COO with GLOBAL numbering or None)."""
    from oracle.bindings import PortOracle
    ia, ib, ja, jb = box
    cc = crop_case(case, ia, ib, ja, jb, wall_distance=wall_distance)
    port = PortOracle(cc)
    qc = np.ascontiguousarray(q[ia:ib, ja:jb])
    i0, i1, j0, j1 = crop_interior(case, box)
    res = port.residual(qc, lhs)[i0 - ia:i1 - ia, j0 - ja:j1 - ja]
    coo = None
    if want_jacobian:
        ri, ci, va = port.jacobian(qc, lhs)
        nv, w = case.nv, jb - ja

        def to_global(idx):
            idx = idx.astype(np.int64)
            cell, k = idx // nv, idx % nv
            return (((cell // w) + ia) * case.njc + (cell % w) + ja) * nv + k
        gr, gc = to_global(ri), to_global(ci)
        cells = [(i, j) for i in range(i0, i1) for j in range(j0, j1)]
        keep = np.isin(gr, rows_of_cells(case.njc, nv, cells))
        coo = (gr[keep], gc[keep], va[keep])
    port.close()
    return res, (i0, i1, j0, j1), coo


def jac_worst(n, a, b, nv, njc, top=4):
    """diagnostics for a failed Jacobian comparison: the `top` worst entries under the Appendix-B rule as
    (row cell i, j, k, col cell i, j, k, ours, reference, scaled error)"""
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
    rel = np.abs(fa - fb) / scale
    out = []
    for k in np.argsort(rel)[::-1][:top]:
        r, c = divmod(int(keys[k]), n)
        out.append(((r // nv) // njc, (r // nv) % njc, r % nv, (c // nv) // njc, (c // nv) % njc, c % nv, float(fa[k]), float(fb[k]), float(rel[k])))
    return out


def jac_rel_err_split(n, a, b, nv):
    """Appendix-B error split by entry class: (all entries except d(mean-flow row)/d(q4), those SA-coupling entries).
    The coupling entries d rhs_k/d(rho nu~), k < 4, are (d flux/d mu_face) x (d mu_t/d q4): the first factor is the stress
    over mu -- a velocity-gradient combination that cancels to ~1e-3 of its terms on a smooth state -- and the second is
    ~1/mu_inf, so these entries are the largest of their rows AND carry the cancellation noise.  The oracle's own
    evaluation-order noise on them (the same source built with and without FMA contraction) is 3e-13 at Re 5e6 on the
    1024^2 flat plate; everything else sits at 1e-14."""
    ka, va = coo_to_dict_arrays(n, *a)
    kb, vb = coo_to_dict_arrays(n, *b)
    keys = np.union1d(ka, kb)
    fa = np.zeros(len(keys)); fb = np.zeros(len(keys))
    fa[np.searchsorted(keys, ka)] = va
    fb[np.searchsorted(keys, kb)] = vb
    rows, cols = keys // n, keys % n
    srow = np.zeros(n)
    np.maximum.at(srow, rows, np.abs(fb))
    scale = np.maximum(np.maximum(np.abs(fa), np.abs(fb)), srow[rows])
    scale[scale == 0] = 1.0
    rel = np.abs(fa - fb) / scale
    coupling = (cols % nv == 4) & (rows % nv < 4) if nv > 4 else np.zeros(len(keys), dtype=bool)
    e1 = float(rel[~coupling].max()) if (~coupling).any() else 0.0
    e2 = float(rel[coupling].max()) if coupling.any() else 0.0
    return e1, e2


TOL_SA_COUPLING = 1e-11    # bound for the d(mean flow)/d(rho nu~) entries on BASELINE-size grids (see jac_rel_err_split)
