"""TEST INFRASTRUCTURE ONLY -- ctypes bindings of the two CPU checkers.

  RefOracle   oracle/_ref/libstructured_ref.so : the unmodified reference compiled as a library
              (oracle/ref/ref_unity.cpp).  Exists only where oracle/_ref was built (it travels to the
              GPU box as a prebuilt file; it is never rebuilt there).
  PortOracle  oracle/liboracle_port.so : our restatement (oracle/port/structured_port.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libstructured_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "structured_explicit")
PORT_LIB = os.path.join(HERE, "liboracle_port.so")

_P = ctypes.POINTER(ctypes.c_double)
_U = ctypes.POINTER(ctypes.c_uint)


def _dp(a):
    return a.ctypes.data_as(_P)


class SgpuBc(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("face", ctypes.c_int), ("start", ctypes.c_int), ("end", ctypes.c_int),
                ("u", ctypes.c_double), ("v", ctypes.c_double), ("T", ctypes.c_double)]


class SgpuDesc(ctypes.Structure):
    _fields_ = [("ni", ctypes.c_int), ("nj", ctypes.c_int), ("ntrans", ctypes.c_int), ("order", ctypes.c_int),
                ("lhs_order", ctypes.c_int), ("flux", ctypes.c_int),
                ("rho_inf", ctypes.c_double), ("u_inf", ctypes.c_double), ("v_inf", ctypes.c_double),
                ("p_inf", ctypes.c_double), ("T_inf", ctypes.c_double), ("mu_inf", ctypes.c_double),
                ("pr_inf", ctypes.c_double), ("dpdx", ctypes.c_double), ("dpdy", ctypes.c_double),
                ("n_bc", ctypes.c_int), ("bc", ctypes.POINTER(SgpuBc)), ("device", ctypes.c_int),
                ("j_begin", ctypes.c_int), ("j_end", ctypes.c_int)]


def make_desc(case, device=0, j_begin=0, j_end=0):
    """Case -> (sgpu_desc, keepalive)."""
    from structured_b200.cases import BC_TYPES, FACES, FLUXES
    n = len(case.boundaries)
    arr = (SgpuBc * max(n, 1))()
    for k, b in enumerate(case.boundaries):
        arr[k] = SgpuBc(BC_TYPES[b.type], FACES[b.face], b.start, b.end, b.u, b.v, b.T)
    d = SgpuDesc(case.ni, case.nj, case.ntrans, case.order, case.lhs_order if case.lhs_order is not None else case.order,
                 FLUXES[case.flux], case.rho_inf, case.u_inf, case.v_inf, case.p_inf, case.T_inf, case.mu_inf,
                 case.pr_inf, case.dpdx, case.dpdy, n, arr, device, j_begin, j_end)
    return d, arr


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def build_port() -> str:
    if not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(os.path.join(HERE, "port", "structured_port.hpp")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_LIB


def _coo_to_arrays(lib_free, nnz, r, c, v):
    n = nnz.value
    ri = np.ctypeslib.as_array(r, (max(n, 1),))[:n].copy()
    ci = np.ctypeslib.as_array(c, (max(n, 1),))[:n].copy()
    va = np.ctypeslib.as_array(v, (max(n, 1),))[:n].copy()
    lib_free(r); lib_free(c); lib_free(v)
    return ri, ci, va


class RefOracle:
    """The reference's own classes, driven from a `.inp` file (written from a Case if needed)."""

    def __init__(self, case=None, config_path=None):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libstructured_ref.so not built (needs /root/reference; `make -C oracle ref`)")
        L = self.L = ctypes.CDLL(REF_LIB)
        L.ref_create.restype = ctypes.c_void_p
        L.ref_create.argtypes = [ctypes.c_char_p]
        L.ref_time_residual.restype = ctypes.c_double
        L.ref_time_jacobian.restype = ctypes.c_double
        for name in ("ref_destroy", "ref_dims", "ref_get_grid", "ref_get_metrics", "ref_get_q", "ref_residual",
                     "ref_get_primitives", "ref_calc_dt", "ref_time_residual", "ref_jacobian", "ref_time_jacobian", "ref_free", "ref_surface"):
            getattr(L, name).argtypes = None
        self._tmp = None
        if config_path is None:
            from structured_b200.cases import write_case
            assert case.ntrans == 0, "the reference has no transport equation (ntrans = 0, src/solver/solution.cpp:9)"
            self._tmp = tempfile.TemporaryDirectory(prefix="sref_")
            config_path = write_case(case, self._tmp.name)
        self.dir = os.path.dirname(os.path.abspath(config_path))
        self.h = ctypes.c_void_p(L.ref_create(os.path.abspath(config_path).encode()))
        ni, nj, nv = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        L.ref_dims(self.h, ctypes.byref(ni), ctypes.byref(nj), ctypes.byref(nv))
        self.ni, self.nj, self.nv = ni.value, nj.value, nv.value
        self.nic, self.njc = self.ni - 1, self.nj - 1

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def grid(self):
        x, y = np.empty((self.ni, self.nj)), np.empty((self.ni, self.nj))
        self.L.ref_get_grid(self.h, _dp(x), _dp(y))
        return x, y

    def metrics(self):
        nchi, neta, vol = np.empty((self.ni, self.njc, 2)), np.empty((self.nic, self.nj, 2)), np.empty((self.nic, self.njc))
        self.L.ref_get_metrics(self.h, _dp(nchi), _dp(neta), _dp(vol))
        return nchi, neta, vol

    def initial_q(self):
        q = np.empty((self.nic, self.njc, self.nv))
        self.L.ref_get_q(self.h, _dp(q))
        return q

    def residual(self, q, lhs=False):
        q = np.ascontiguousarray(q, dtype=np.float64)
        rhs = np.empty_like(q)
        self.L.ref_residual(self.h, _dp(q), _dp(rhs), int(lhs))
        return rhs

    def primitives(self):
        out = [np.empty((self.nic + 2, self.njc + 2)) for _ in range(5)]
        self.L.ref_get_primitives(self.h, *[_dp(a) for a in out])
        return out

    def calc_dt(self, q, cfl):
        q = np.ascontiguousarray(q, dtype=np.float64)
        dt = np.empty_like(q)
        self.L.ref_calc_dt(self.h, _dp(q), ctypes.c_double(cfl), _dp(dt))
        return dt

    def surface(self, q_res, q_fin=None):
        """IOManager::write_surface run by the reference (src/utils/io.cpp:182-255): returns (rows of the text file
        [n][3] = xw cp cf at 6 significant digits, wall[6][nic] = the full-precision arrays the routine reads)."""
        q_res = np.ascontiguousarray(q_res, dtype=np.float64)
        q_fin = q_res if q_fin is None else np.ascontiguousarray(q_fin, dtype=np.float64)
        wall = np.empty((6, self.nic))
        label = ctypes.create_string_buffer(256)
        self.L.ref_surface(self.h, _dp(q_res), _dp(q_fin), _dp(wall), label, 256)
        path = os.path.join(self.dir, label.value.decode() + ".surface")
        text = open(path).read()
        os.remove(path)
        rows = np.array([[float(t) for t in line.split()] for line in text.splitlines() if line.strip()]).reshape(-1, 3)
        return rows, wall

    def time_residual(self, q, reps, lhs=False):
        q = np.ascontiguousarray(q, dtype=np.float64)
        return float(self.L.ref_time_residual(self.h, _dp(q), int(reps), int(lhs)))

    def jacobian(self, q, lhs=True):
        q = np.ascontiguousarray(q, dtype=np.float64)
        nnz, nc = ctypes.c_int(), ctypes.c_int()
        r, c, v = _U(), _U(), _P()
        self.L.ref_jacobian(self.h, _dp(q), int(lhs), ctypes.byref(nnz), ctypes.byref(r), ctypes.byref(c), ctypes.byref(v), ctypes.byref(nc))
        self.ncolors = nc.value
        return _coo_to_arrays(self.L.ref_free, nnz, r, c, v)

    def time_jacobian(self, q, lhs=True):
        q = np.ascontiguousarray(q, dtype=np.float64)
        nnz = ctypes.c_int()
        return float(self.L.ref_time_jacobian(self.h, _dp(q), int(lhs), ctypes.byref(nnz))), nnz.value


class PortOracle:
    """Our CPU restatement, driven from a Case through the same sgpu_desc the GPU library takes."""

    def __init__(self, case):
        L = self.L = ctypes.CDLL(build_port())
        L.port_create.restype = ctypes.c_void_p
        L.port_create.argtypes = [ctypes.POINTER(SgpuDesc)]
        L.port_time_residual.restype = ctypes.c_double
        L.port_time_jacobian.restype = ctypes.c_double
        self.case = case
        d, self._keep = make_desc(case)
        self.h = ctypes.c_void_p(L.port_create(ctypes.byref(d)))
        self.ni, self.nj, self.nv = case.ni, case.nj, case.nv
        self.nic, self.njc = case.nic, case.njc
        xv = np.ascontiguousarray(case.xv, dtype=np.float64)
        yv = np.ascontiguousarray(case.yv, dtype=np.float64)
        L.port_set_grid(self.h, _dp(xv), _dp(yv))
        if getattr(case, "global_counts", None):     # cropped window of a larger grid (tests/helpers.py::crop_case)
            L.port_set_global_counts(self.h, int(case.global_counts[0]), int(case.global_counts[1]))
        if case.ntrans:
            if case.wall_distance is not None:
                L.port_set_field(self.h, b"wall_distance", _dp(np.ascontiguousarray(case.wall_distance)))
            else:                                    # nearest wall edge, like sgpu_wall_distance_from_bcs
                L.port_wall_distance(self.h, None)
            if case.beta is not None:
                L.port_set_field(self.h, b"beta", _dp(np.ascontiguousarray(case.beta)))

    def close(self):
        if self.h:
            self.L.port_destroy(self.h)
            self.h = None

    def metrics(self):
        nchi, neta, vol = np.empty((self.ni, self.njc, 2)), np.empty((self.nic, self.nj, 2)), np.empty((self.nic, self.njc))
        self.L.port_get_metrics(self.h, _dp(nchi), _dp(neta), _dp(vol))
        return nchi, neta, vol

    def residual(self, q, lhs=False):
        q = np.ascontiguousarray(q, dtype=np.float64)
        rhs = np.empty_like(q)
        self.L.port_residual(self.h, _dp(q), _dp(rhs), int(lhs))
        return rhs

    def primitives(self):
        out = [np.empty((self.nic + 2, self.njc + 2)) for _ in range(6)]
        self.L.port_get_primitives(self.h, *[_dp(a) for a in out])
        return out

    def calc_dt(self, q, cfl):
        q = np.ascontiguousarray(q, dtype=np.float64)
        dt = np.empty_like(q)
        self.L.port_calc_dt(self.h, _dp(q), ctypes.c_double(cfl), _dp(dt))
        return dt

    def surface(self, q_res, q_fin=None, i_first=None, count=None, aoa=None):
        """IOManager::write_surface restated (port_surface): dict(xw, cp, cf, coeffs[6], wall[6][nic]).  Defaults are the
        reference's range j1 - 1 .. j1 - 1 + nb with j1 = geometry.tail, nb = ni - 2 j1 + 1 (src/utils/mesh.cpp:349-350)."""
        q_res = np.ascontiguousarray(q_res, dtype=np.float64)
        q_fin = q_res if q_fin is None else np.ascontiguousarray(q_fin, dtype=np.float64)
        if i_first is None:
            i_first, count = self.case.tail - 1, self.case.ni - 2*self.case.tail + 1
        aoa = self.case.aoa if aoa is None else aoa
        xw, cp, cf = np.empty(count), np.empty(count), np.empty(count)
        coeffs, wall = np.empty(6), np.empty((6, self.nic))
        self.L.port_surface(self.h, _dp(q_res), _dp(q_fin), int(i_first), int(count), ctypes.c_double(aoa), _dp(xw), _dp(cp), _dp(cf), _dp(coeffs), _dp(wall))
        return dict(xw=xw, cp=cp, cf=cf, coeffs=coeffs, wall=wall)

    def time_residual(self, q, reps, lhs=False):
        q = np.ascontiguousarray(q, dtype=np.float64)
        rhs = np.empty_like(q)
        return float(self.L.port_time_residual(self.h, _dp(q), _dp(rhs), int(reps), int(lhs)))

    def jacobian(self, q, lhs=True):
        q = np.ascontiguousarray(q, dtype=np.float64)
        nnz, nc = ctypes.c_int(), ctypes.c_int()
        r, c, v = _U(), _U(), _P()
        self.L.port_jacobian(self.h, _dp(q), int(lhs), ctypes.byref(nnz), ctypes.byref(r), ctypes.byref(c), ctypes.byref(v), ctypes.byref(nc))
        self.ncolors = nc.value
        return _coo_to_arrays(self.L.port_free, nnz, r, c, v)

    def time_jacobian(self, q, lhs=True):
        q = np.ascontiguousarray(q, dtype=np.float64)
        nnz = ctypes.c_int()
        return float(self.L.port_time_jacobian(self.h, _dp(q), int(lhs), ctypes.byref(nnz))), nnz.value

    def wall_distance(self):
        """nearest-wall-edge distance [nic][njc] from the case's wall / isothermalwall tables (also stored in the oracle)"""
        out = np.empty((self.nic, self.njc))
        self.L.port_wall_distance(self.h, _dp(out))
        return out

    def set_field(self, name, f):
        self.L.port_set_field(self.h, name.encode(), _dp(np.ascontiguousarray(f, dtype=np.float64)))

    def jacobian_rows(self, q, cells, lhs=True):
        """Rows of the Jacobian for sampled row cells on LARGE grids (static 5x5 colouring, 25 nv / lanes dual sweeps):
        cells [n][2] = (i, j), none within 2 cells of a periodic / wake boundary.  Returns COO (rind, cind, values) of the
        non-zero entries with the reference's flat numbering (i*njc + j)*nv + k."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 2)
        m, nv = len(cells), self.nv
        out = np.zeros((m, nv, 25, nv))
        self.L.port_jacobian_rows(self.h, _dp(q), int(lhs), cells.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), m, _dp(out))
        di, dj = np.meshgrid(np.arange(-2, 3), np.arange(-2, 3), indexing="ij")
        ci = cells[:, 0][:, None] + di.reshape(-1)[None, :]; cj = cells[:, 1][:, None] + dj.reshape(-1)[None, :]      # [m][25]
        rows = ((cells[:, 0].astype(np.int64)*self.njc + cells[:, 1])[:, None, None, None]*nv + np.arange(nv)[None, :, None, None])
        cols = ((ci.astype(np.int64)*self.njc + cj)[:, None, :, None]*nv + np.arange(nv)[None, None, None, :])
        rows, cols = np.broadcast_to(rows, out.shape), np.broadcast_to(cols, out.shape)
        keep = out != 0.0
        return rows[keep].astype(np.uint32), cols[keep].astype(np.uint32), out[keep]


def rk4_step_cpu(oracle, q, q_tmp, cfl):
    """The explicit rk4_jameson branch of Solver::step (src/solver/solver.cpp:66,107-116) on a CPU oracle."""
    dt = oracle.calc_dt(q, cfl)
    rhs = None
    for order in range(4):
        rhs = oracle.residual(q_tmp)
        q_tmp = q + rhs * dt / (4.0 - order)
    return q_tmp.copy(), q_tmp, rhs
