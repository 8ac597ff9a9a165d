"""The reference's solution file formats (IOManager, src/utils/io.cpp), written from / read into the host AoS state
`q[nic][njc][nv]` that `sgpu_get_state` / `sgpu_set_state` exchange with the device planes.

  <label>.out   restart: the raw doubles of q, i outer, j, then k (write_restart / read_restart, io.cpp:136-180)
  <label>.npz   numpy archive with xc, yc, q, rho, u, v, p, T (write_npz, io.cpp:104-132); the primitives are
                FluidModel::primvars of q with shift 0 (io.cpp:41, src/model/fluid.cpp:50-67)
"""
from __future__ import annotations

import os

import numpy as np

GAMMA = 1.4  # src/common.h:40


def cell_centres(xv: np.ndarray, yv: np.ndarray):
    """Mesh::xc, yc (src/utils/mesh.cpp:199-200), same summation order"""
    xc = 0.25 * (((xv[:-1, :-1] + xv[1:, :-1]) + xv[:-1, 1:]) + xv[1:, 1:])
    yc = 0.25 * (((yv[:-1, :-1] + yv[1:, :-1]) + yv[:-1, 1:]) + yv[1:, 1:])
    return xc, yc


def primitives(case, q: np.ndarray):
    """FluidModel::primvars (src/model/fluid.cpp:50-67): rho, u, v, p, T = p/rho/R with R = p_inf/rho_inf/T_inf (fluid.cpp:12)"""
    rho = q[..., 0].copy()
    u = q[..., 1] / rho
    v = q[..., 2] / rho
    p = (q[..., 3] - 0.5 * rho * (u * u + v * v)) * (GAMMA - 1.0)
    R = case.p_inf / case.rho_inf / case.T_inf
    T = p / rho / R
    return rho, u, v, p, T


def write_restart(path: str, q: np.ndarray) -> None:
    np.ascontiguousarray(q, dtype=np.float64).tofile(path)


def read_restart(path: str, nic: int, njc: int, nv: int) -> np.ndarray:
    size, expected = os.path.getsize(path), nic * njc * nv * 8
    if size != expected:                                   # the reference asserts (io.cpp:168)
        raise ValueError("restart file %s holds %d bytes, expected %d for %d x %d x %d" % (path, size, expected, nic, njc, nv))
    return np.fromfile(path, dtype=np.float64).reshape(nic, njc, nv)


def write_npz(path: str, case, q: np.ndarray) -> None:
    """same keys and shapes as cnpy writes them; q keeps all nv components (the reference, with ntrans = 0, writes 4)"""
    xc, yc = cell_centres(case.xv, case.yv)
    rho, u, v, p, T = primitives(case, q)
    with open(path, "wb") as f:
        np.savez(f, xc=xc, yc=yc, q=np.ascontiguousarray(q), rho=rho, u=u, v=v, p=p, T=T)
