"""Host-side mirror of the reference's operator interface over the C ABI (include/structured_gpu.h).

`GpuEulerEquation` has the methods of `EulerEquation` (src/model/eulerequation.h:53-55: calc_residual,
calc_dt, initialize) plus the pieces of `Solver::step` (src/solver/solver.cpp:53-196) that the GPU path
takes over: the stage updates, the residual norms and the sparse Jacobian that replaces ADOL-C's
`sparse_jac`.  All compute goes through libstructured_gpu.so; if the library or a CUDA device is missing
every call raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from .cases import BC_TYPES, FACES, FLUXES, Case

_P = ctypes.POINTER(ctypes.c_double)
_U = ctypes.POINTER(ctypes.c_uint)

SYMBOLS = [
    "sgpu_create", "sgpu_destroy", "sgpu_last_error", "sgpu_set_stream", "sgpu_synchronize", "sgpu_dims",
    "sgpu_set_grid", "sgpu_set_grid_window", "sgpu_set_grid_file", "sgpu_set_field", "sgpu_set_field_window", "sgpu_get_field", "sgpu_compute_wall_distance",
    "sgpu_wall_distance_from_bcs", "sgpu_get_metrics", "sgpu_set_state", "sgpu_set_state_window", "sgpu_get_state", "sgpu_copy_state",
    "sgpu_get_rhs", "sgpu_get_rhs_window", "sgpu_get_dt", "sgpu_calc_dt", "sgpu_residual", "sgpu_residual_host", "sgpu_residual_host_window", "sgpu_rk_stage",
    "sgpu_forward_euler", "sgpu_explicit_step", "sgpu_jacobian_coo", "sgpu_jacobian_coo_rows", "sgpu_jacobian_device", "sgpu_jacobian_apply", "sgpu_dres_dbeta",
    "sgpu_wall_data", "sgpu_track_wall", "sgpu_surface", "sgpu_surface_gradient",
    "sgpu_linear_solve", "sgpu_implicit_step", "sgpu_adjoint_solve", "sgpu_adjoint_solve_ramp",
    "sgpu_vec_size", "sgpu_vec_from_rhs", "sgpu_vec_add_to_state", "sgpu_vec_halo_pack", "sgpu_vec_halo_unpack", "sgpu_vec_halo_pack_ghost", "sgpu_vec_halo_add",
    "sgpu_vec_from_host", "sgpu_vec_to_host", "sgpu_op_apply",
    "sgpu_precond_setup", "sgpu_precond_apply", "sgpu_vec_dots", "sgpu_vec_gs_update", "sgpu_vec_scale_rsqrt",
    "sgpu_halo_count", "sgpu_halo_pack", "sgpu_halo_unpack", "sgpu_halo_pack_ghost", "sgpu_halo_recv_buffer", "sgpu_halo_enable_peer", "sgpu_halo_set_peer", "sgpu_halo_ipc_handle", "sgpu_halo_open_peer",
    "sgpu_halo_push", "sgpu_halo_pull", "sgpu_launch_count", "sgpu_kernel_times", "sgpu_enable_kernel_timing",
]


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


class SgpuLinsolve(ctypes.Structure):
    """sgpu_linsolve of include/structured_gpu.h: GMRES controls in, iteration report out."""
    _fields_ = [("precond", ctypes.c_int), ("restart", ctypes.c_int), ("max_iter", ctypes.c_int), ("reorthogonalize", ctypes.c_int),
                ("rtol", ctypes.c_double),
                ("iterations", ctypes.c_int), ("converged", ctypes.c_int), ("rel_residual", ctypes.c_double),
                ("setup_ms", ctypes.c_float), ("solve_ms", ctypes.c_float), ("matvec_ms", ctypes.c_float), ("precond_ms", ctypes.c_float)]


MATRICES = {"lhs": 0, "J": 1, "JT": 2, "lhsT": 3}
PRECONDS = {"block_jacobi": 0, "line_j": 1}


class SgpuError(RuntimeError):
    pass


_LIB = None


def lib_path() -> str:
    # SGPU_LIB: load another build of the same library (A/B timing of kernel variants)
    return os.environ.get("SGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libstructured_gpu.so")


def load_library():
    """Load libstructured_gpu.so (building it if the sources are newer and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.environ.get("SGPU_LIB"):
        from . import build as _build
        if _build.stale(path):          # missing, or older than any source under csrc/ / include/: never load a stale library
            _build.build()
    L = ctypes.CDLL(path)
    L.sgpu_last_error.restype = ctypes.c_char_p
    L.sgpu_last_error.argtypes = [ctypes.c_void_p]
    L.sgpu_launch_count.restype = ctypes.c_longlong
    L.sgpu_launch_count.argtypes = [ctypes.c_void_p]
    L.sgpu_create.argtypes = [ctypes.POINTER(SgpuDesc), ctypes.POINTER(ctypes.c_void_p)]
    _LIB = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_P)


def make_desc(case: Case, device: int = 0, j_begin: int = 0, j_end: int = 0):
    n = len(case.boundaries)
    arr = (SgpuBc * max(n, 1))()
    for k, b in enumerate(case.boundaries):
        if b.type not in BC_TYPES:
            raise SgpuError("Wrong type of BC.")  # src/model/bc.cpp:523
        arr[k] = SgpuBc(BC_TYPES[b.type], FACES[b.face], b.start, b.end, b.u, b.v, b.T)
    if case.flux not in FLUXES:
        raise SgpuError("Flux not found.")  # src/model/eulerequation.cpp:128
    d = SgpuDesc(case.ni, case.nj, case.ntrans, case.order, case.lhs_order if case.lhs_order is not None else case.order,
                 FLUXES[case.flux], case.rho_inf, case.u_inf, case.v_inf, case.p_inf, case.T_inf, case.mu_inf,
                 case.pr_inf, case.dpdx, case.dpdy, n, arr, device, j_begin, j_end)
    return d, arr


class GpuEulerEquation:
    """EulerEquation + the explicit/implicit pieces of Solver::step on one B200 (or one j-slab of a grid)."""

    def __init__(self, case: Case, device: int = 0, j_begin: int = 0, j_end: int = 0, stream: Optional[int] = None,
                 window: Optional[tuple] = None):
        self.L = load_library()
        self.case = case
        d, self._keep = make_desc(case, device, j_begin, j_end)
        h = ctypes.c_void_p()
        rc = self.L.sgpu_create(ctypes.byref(d), ctypes.byref(h))
        if rc != 0:
            raise SgpuError("sgpu_create: %s" % self.L.sgpu_last_error(None).decode())
        self.h = h
        self.nic, self.njc, self.nv = case.nic, case.njc, case.nv
        self.j_begin, self.j_end = (j_begin, j_end) if (j_begin or j_end) else (0, case.njc)
        if stream is not None:
            self._ck(self.L.sgpu_set_stream(self.h, ctypes.c_void_p(stream)))
        if window is None:
            self._ck(self.L.sgpu_set_grid(self.h, _dp(np.ascontiguousarray(case.xv, dtype=np.float64)),
                                          _dp(np.ascontiguousarray(case.yv, dtype=np.float64))))
            if case.ntrans:
                if case.wall_distance is not None:
                    self.set_field("wall_distance", case.wall_distance)
                else:                                    # [turbulence] wall_distance = "compute": nearest wall edge, on the device
                    self.compute_wall_distance()
                if case.beta is not None:
                    self.set_field("beta", case.beta)
        else:
            # slab runs: case.xv / yv hold only vertex rows [jv_first, ...), the fields only cell rows [jc_first, ...)
            jv_first, jc_first = window
            xv = np.ascontiguousarray(case.xv, dtype=np.float64); yv = np.ascontiguousarray(case.yv, dtype=np.float64)
            self._ck(self.L.sgpu_set_grid_window(self.h, _dp(xv), _dp(yv), jv_first, xv.shape[1]))
            if case.ntrans:
                for name, f in (("wall_distance", case.wall_distance), ("beta", case.beta)):
                    if f is not None:
                        f = np.ascontiguousarray(f, dtype=np.float64)
                        self._ck(self.L.sgpu_set_field_window(self.h, name.encode(), _dp(f), jc_first, f.shape[1]))

    # ---- plumbing
    def _ck(self, rc: int):
        if rc != 0:
            raise SgpuError("libstructured_gpu error %d: %s" % (rc, self.L.sgpu_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.sgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._ck(self.L.sgpu_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.L.sgpu_launch_count(self.h))

    def _state_array(self):
        return np.zeros((self.nic, self.njc, self.nv))

    # ---- static inputs
    def set_field(self, name: str, field: np.ndarray):
        self._ck(self.L.sgpu_set_field(self.h, name.encode(), _dp(np.ascontiguousarray(field, dtype=np.float64))))

    def set_grid_file(self, path: str):
        """vertices from a binary file (cases.write_grid_bin); on a slab only its window of rows is read"""
        self._ck(self.L.sgpu_set_grid_file(self.h, os.fsencode(path)))

    def get_field(self, name: str) -> np.ndarray:
        out = np.zeros((self.nic, self.njc))
        self._ck(self.L.sgpu_get_field(self.h, name.encode(), _dp(out)))
        return out

    def compute_wall_distance(self, segments: Optional[np.ndarray] = None):
        """SA wall distance on the device: distance of every cell centre to the nearest wall edge.  segments [n][4] =
        x0 y0 x1 y1, default: the edges covered by the case's `wall` / `isothermalwall` tables (needs the full grid)."""
        if segments is None:
            if self.case.window is not None:
                raise SgpuError("wall edges from the [[boundary]] tables need the full vertex arrays; pass the segments")
            xv = np.ascontiguousarray(self.case.xv, dtype=np.float64); yv = np.ascontiguousarray(self.case.yv, dtype=np.float64)
            self._ck(self.L.sgpu_wall_distance_from_bcs(self.h, _dp(xv), _dp(yv)))
        else:
            seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
            self._ck(self.L.sgpu_compute_wall_distance(self.h, _dp(seg), int(seg.shape[0])))

    def metrics(self):
        nchi = np.zeros((self.case.ni, self.njc, 2)); neta = np.zeros((self.nic, self.case.nj, 2)); vol = np.zeros((self.nic, self.njc))
        self._ck(self.L.sgpu_get_metrics(self.h, _dp(nchi), _dp(neta), _dp(vol)))
        return nchi, neta, vol

    # ---- state
    def set_state(self, q: np.ndarray, which: int = 0):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == (self.nic, self.njc, self.nv)
        self._ck(self.L.sgpu_set_state(self.h, which, _dp(q)))
        self.synchronize()

    def set_state_window(self, q: np.ndarray, j_first: int, which: int = 0):
        """q holds only global rows [j_first, j_first + q.shape[1]) (must cover the owned rows)."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape[0] == self.nic and q.shape[2] == self.nv
        self._ck(self.L.sgpu_set_state_window(self.h, which, _dp(q), j_first, q.shape[1]))
        self.synchronize()

    def get_state(self, which: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = self._state_array() if out is None else out
        self._ck(self.L.sgpu_get_state(self.h, which, _dp(out)))
        return out

    def get_rhs(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = self._state_array() if out is None else out
        self._ck(self.L.sgpu_get_rhs(self.h, _dp(out)))
        return out

    def get_rhs_window(self, out: np.ndarray) -> np.ndarray:
        """owned rows only: out is [nic][j_end - j_begin][nv]"""
        assert out.shape == (self.nic, self.j_end - self.j_begin, self.nv) and out.flags.c_contiguous
        self._ck(self.L.sgpu_get_rhs_window(self.h, _dp(out)))
        return out

    def get_dt(self) -> np.ndarray:
        out = self._state_array()
        self._ck(self.L.sgpu_get_dt(self.h, _dp(out)))
        return out

    def initialize(self):
        """EulerEquation::initialize (src/model/eulerequation.cpp:262-291): q = q_tmp = freestream."""
        q = self.case.freestream_q()
        self.set_state(q, 0)
        self.set_state(q, 1)
        return q

    # ---- hot path
    def calc_dt(self, cfl: float):
        self._ck(self.L.sgpu_calc_dt(self.h, ctypes.c_double(cfl)))

    def residual_device(self, which: int = 0, lhs: bool = False, norms: bool = False):
        """Residual of the device-resident state; returns sum(rhs^2) per equation if norms."""
        if norms:
            l2 = np.zeros(self.nv)
            self._ck(self.L.sgpu_residual(self.h, which, int(lhs), _dp(l2)))
            return l2
        self._ck(self.L.sgpu_residual(self.h, which, int(lhs), None))
        return None

    def calc_residual(self, q: np.ndarray, lhs: bool = False, out: Optional[np.ndarray] = None) -> np.ndarray:
        """equation->calc_residual(q, rhs, lhs) with HOST arrays (src/solver/solver.cpp:80,93,104,110)."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = self._state_array() if out is None else out
        self._ck(self.L.sgpu_residual_host(self.h, _dp(q), _dp(out), int(lhs)))
        return out

    def calc_residual_window(self, q: np.ndarray, j_first: int, out: np.ndarray, lhs: bool = False) -> np.ndarray:
        """slab form of calc_residual: q holds rows [j_first, j_first + q.shape[1]), out the owned rows only"""
        assert q.flags.c_contiguous and out.flags.c_contiguous and out.shape == (self.nic, self.j_end - self.j_begin, self.nv)
        self._ck(self.L.sgpu_residual_host_window(self.h, _dp(q), j_first, q.shape[1], _dp(out), int(lhs)))
        return out

    def rk_stage(self, order: int):
        self._ck(self.L.sgpu_rk_stage(self.h, order))

    def forward_euler(self):
        self._ck(self.L.sgpu_forward_euler(self.h))

    def copy_state(self, dst: int, src: int):
        self._ck(self.L.sgpu_copy_state(self.h, dst, src))

    def explicit_step(self, cfl: float, scheme: Optional[str] = None) -> np.ndarray:
        """One explicit Solver::step on the device; returns the L2 norms of the last rhs (solver.cpp:125-134)."""
        scheme = self.case.scheme if scheme is None else scheme
        if scheme not in ("forward_euler", "rk4_jameson"):
            raise SgpuError("scheme not defined.")  # src/solver/solver.cpp:119
        l2 = np.zeros(self.nv)
        self._ck(self.L.sgpu_explicit_step(self.h, 0 if scheme == "forward_euler" else 1, ctypes.c_double(cfl), _dp(l2)))
        return np.sqrt(l2)

    # ---- Jacobian
    def jacobian_coo(self, apply_lhs_transform: bool = False, rows: Optional[tuple] = None):
        """Replacement of sparse_jac (src/solver/solver.cpp:156): returns (rind, cind, values) numpy copies.
        rows = (j_first, j_count): only those cell rows, from the Jacobian jacobian_device() left on the device."""
        nnz = ctypes.c_int(); r, c, v = _U(), _U(), _P()
        if rows is None:
            self._ck(self.L.sgpu_jacobian_coo(self.h, ctypes.byref(nnz), ctypes.byref(r), ctypes.byref(c), ctypes.byref(v), int(apply_lhs_transform)))
        else:
            self._ck(self.L.sgpu_jacobian_coo_rows(self.h, int(rows[0]), int(rows[1]), ctypes.byref(nnz), ctypes.byref(r), ctypes.byref(c),
                                                   ctypes.byref(v), int(apply_lhs_transform)))
        n = nnz.value
        libc = ctypes.CDLL(None)
        libc.free.argtypes = [ctypes.c_void_p]
        ri = np.ctypeslib.as_array(r, (max(n, 1),))[:n].copy(); ci = np.ctypeslib.as_array(c, (max(n, 1),))[:n].copy()
        va = np.ctypeslib.as_array(v, (max(n, 1),))[:n].copy()
        for p in (r, c, v):
            libc.free(ctypes.cast(p, ctypes.c_void_p))
        return ri, ci, va

    def jacobian_device(self):
        slots = ctypes.c_int(); ms = ctypes.c_float()
        self._ck(self.L.sgpu_jacobian_device(self.h, ctypes.byref(slots), ctypes.byref(ms)))
        return slots.value, ms.value

    def jacobian_apply(self, x: np.ndarray, transpose: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = self._state_array()
        self._ck(self.L.sgpu_jacobian_apply(self.h, int(transpose), _dp(x), _dp(y)))
        return y

    # ---- device linear solve / implicit step (replaces src/linearsolver/*)
    @staticmethod
    def _linsolve(precond, restart, max_iter, rtol, reorthogonalize) -> SgpuLinsolve:
        if precond not in PRECONDS:
            raise SgpuError("unknown preconditioner %r" % (precond,))
        return SgpuLinsolve(PRECONDS[precond], restart, max_iter, int(reorthogonalize), rtol, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0)

    @staticmethod
    def _linsolve_info(io: SgpuLinsolve) -> dict:
        return {"iterations": io.iterations, "converged": bool(io.converged), "rel_residual": io.rel_residual,
                "setup_ms": io.setup_ms, "solve_ms": io.solve_ms, "matvec_ms": io.matvec_ms, "precond_ms": io.precond_ms}

    def linear_solve(self, matrix: str = "lhs", b: Optional[np.ndarray] = None, precond: str = "block_jacobi", restart: int = 30,
                     max_iter: int = 500, rtol: float = 1e-10, reorthogonalize: bool = False, want_x: bool = True):
        """Solve A x = b on the device with the Jacobian of the last jacobian_device()/jacobian_coo()/implicit_step().
        matrix "lhs": A = -J + 1/dt (what linearsolver->set_lhs receives, solver.cpp:162-173), "J", "JT" (adjoint) or
        "lhsT" (transposed LHS: one pseudo-time step of the adjoint).
        b = None uses the device rhs of the last residual (linearsolver->set_rhs(solution->rhs), solver.cpp:174)."""
        if matrix not in MATRICES:
            raise SgpuError("unknown matrix %r" % (matrix,))
        io = self._linsolve(precond, restart, max_iter, rtol, reorthogonalize)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        x = self._state_array() if want_x else None
        self._ck(self.L.sgpu_linear_solve(self.h, MATRICES[matrix], None if bb is None else _dp(bb), None if x is None else _dp(x),
                                          ctypes.byref(io)))
        return x, self._linsolve_info(io)

    def implicit_step(self, cfl: float, under_relaxation: float = 1.0, precond: str = "line_j", restart: int = 30, max_iter: int = 500,
                      rtol: float = 1e-10, reorthogonalize: bool = False):
        """The ENABLE_ADOLC branch of Solver::step, device-resident (solver.cpp:66-101,154-175); returns (l2norm, info)."""
        io = self._linsolve(precond, restart, max_iter, rtol, reorthogonalize)
        l2 = np.zeros(self.nv)
        self._ck(self.L.sgpu_implicit_step(self.h, ctypes.c_double(cfl), ctypes.c_double(under_relaxation), ctypes.byref(io), _dp(l2)))
        return np.sqrt(l2), self._linsolve_info(io)

    # ---- device-vector building blocks of the slab-partitioned solve (driven by structured_b200/slab.py)
    def vec_size(self) -> int:
        n = ctypes.c_longlong()
        self._ck(self.L.sgpu_vec_size(self.h, ctypes.byref(n)))
        return n.value

    def vec_from_rhs(self, vec_ptr: int):
        self._ck(self.L.sgpu_vec_from_rhs(self.h, ctypes.c_void_p(vec_ptr)))

    def vec_add_to_state(self, vec_ptr: int, omega: float = 1.0, which: int = 0):
        self._ck(self.L.sgpu_vec_add_to_state(self.h, which, ctypes.c_void_p(vec_ptr), ctypes.c_double(omega)))

    def vec_halo_pack(self, vec_ptr: int, side: int, buf_ptr: int):
        self._ck(self.L.sgpu_vec_halo_pack(self.h, ctypes.c_void_p(vec_ptr), side, ctypes.c_void_p(buf_ptr)))

    def vec_halo_unpack(self, vec_ptr: int, side: int, buf_ptr: int):
        self._ck(self.L.sgpu_vec_halo_unpack(self.h, ctypes.c_void_p(vec_ptr), side, ctypes.c_void_p(buf_ptr)))

    def vec_halo_pack_ghost(self, vec_ptr: int, side: int, buf_ptr: int):
        self._ck(self.L.sgpu_vec_halo_pack_ghost(self.h, ctypes.c_void_p(vec_ptr), side, ctypes.c_void_p(buf_ptr)))

    def vec_halo_add(self, vec_ptr: int, side: int, buf_ptr: int):
        self._ck(self.L.sgpu_vec_halo_add(self.h, ctypes.c_void_p(vec_ptr), side, ctypes.c_void_p(buf_ptr)))

    def vec_from_host(self, host: np.ndarray, vec_ptr: int):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.shape == (self.nic, self.njc, self.nv)
        self._ck(self.L.sgpu_vec_from_host(self.h, _dp(host), ctypes.c_void_p(vec_ptr)))

    def vec_to_host(self, vec_ptr: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = self._state_array() if out is None else out
        self._ck(self.L.sgpu_vec_to_host(self.h, ctypes.c_void_p(vec_ptr), _dp(out)))
        return out

    def op_apply(self, matrix: str, x_ptr: int, y_ptr: int):
        self._ck(self.L.sgpu_op_apply(self.h, MATRICES[matrix], ctypes.c_void_p(x_ptr), ctypes.c_void_p(y_ptr)))

    def precond_setup(self, matrix: str, precond: str):
        self._ck(self.L.sgpu_precond_setup(self.h, MATRICES[matrix], PRECONDS[precond]))

    def precond_apply(self, matrix: str, precond: str, r_ptr: int, z_ptr: int):
        self._ck(self.L.sgpu_precond_apply(self.h, MATRICES[matrix], PRECONDS[precond], ctypes.c_void_p(r_ptr), ctypes.c_void_p(z_ptr)))

    def vec_dots(self, w_ptr: int, V_ptr: int, cnt: int, out_ptr: int):
        """out[0..cnt) = w . V_j on the device (the Gram-Schmidt projection kernel of sgpu_linear_solve)"""
        self._ck(self.L.sgpu_vec_dots(self.h, ctypes.c_void_p(w_ptr), ctypes.c_void_p(V_ptr), int(cnt), ctypes.c_void_p(out_ptr)))

    def vec_gs_update(self, w_ptr: int, V_ptr: int, cnt: int, h_ptr: int, normsq_ptr: int):
        """w -= sum_j h[j] V_j and *normsq = |w|^2 (this slab's rows), one pass"""
        self._ck(self.L.sgpu_vec_gs_update(self.h, ctypes.c_void_p(w_ptr), ctypes.c_void_p(V_ptr), int(cnt), ctypes.c_void_p(h_ptr),
                                           ctypes.c_void_p(normsq_ptr)))

    def vec_scale_rsqrt(self, dst_ptr: int, src_ptr: int, normsq_ptr: int):
        """dst = src / sqrt(*normsq)"""
        self._ck(self.L.sgpu_vec_scale_rsqrt(self.h, ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), ctypes.c_void_p(normsq_ptr)))

    def adjoint_solve(self, g: np.ndarray, cfl: float = 100.0, max_steps: int = 50, tol: float = 1e-8, precond: str = "line_j",
                      restart: int = 40, max_iter: int = 400, rtol: float = 1e-3, reorthogonalize: bool = False,
                      cfl_growth: float = 1.0, cfl_max: Optional[float] = None):
        """Steady adjoint J^T psi = -g (g = d objective / d q) by pseudo-time continuation with transposed-LHS GMRES solves;
        needs jacobian_device() at the converged state.  cfl_growth > 1 ramps the pseudo-time step like the forward solver's
        CFL ramp (CFL_k = min(cfl * growth^k, cfl_max)).  Returns (psi, info)."""
        io = self._linsolve(precond, restart, max_iter, rtol, reorthogonalize)
        gg = np.ascontiguousarray(g, dtype=np.float64)
        psi = self._state_array()
        steps = ctypes.c_int(); rel = ctypes.c_double()
        self._ck(self.L.sgpu_adjoint_solve_ramp(self.h, _dp(gg), _dp(psi), ctypes.c_double(cfl), ctypes.c_double(cfl_growth),
                                                ctypes.c_double(cfl if cfl_max is None else cfl_max), max_steps, ctypes.c_double(tol),
                                                ctypes.byref(io), ctypes.byref(steps), ctypes.byref(rel)))
        info = self._linsolve_info(io)
        info.update({"steps": steps.value, "rel_residual": rel.value})
        return psi, info

    def track_wall(self, on: bool = True):
        """keep the wall rows of every later residual evaluation (which_res = -1 then selects the last one)"""
        self._ck(self.L.sgpu_track_wall(self.h, int(on)))

    def wall_data(self, which_res: int = 0, which_q: int = 0):
        """(grad_u [nic][2], grad_v [nic][2], p_row0 [nic], p_row1 [nic]): what IOManager::write_surface reads
        (src/utils/io.cpp:182-255) -- eta-face gradients on the j = 0 faces of state which_res, pressure of the two lowest cell
        rows of state which_q."""
        gu, gv = np.empty((self.nic, 2)), np.empty((self.nic, 2))
        p0, p1 = np.empty(self.nic), np.empty(self.nic)
        self._ck(self.L.sgpu_wall_data(self.h, int(which_res), int(which_q), _dp(gu), _dp(gv), _dp(p0), _dp(p1)))
        return gu, gv, p0, p1

    def surface(self, which_res: int = 0, which_q: int = 0, i_first: Optional[int] = None, count: Optional[int] = None,
                aoa: Optional[float] = None) -> dict:
        """IOManager::write_surface (src/utils/io.cpp:182-255): xw, cp, cf per wall column and
        coeffs = [cl_pressure, cd_pressure, cl_viscous, cd_viscous, cl, cd].  Default range = the reference's
        j1 - 1 .. j1 - 1 + nb with j1 = geometry.tail, nb = ni - 2 j1 + 1 (src/utils/mesh.cpp:349-350)."""
        if i_first is None:
            i_first, count = self.case.tail - 1, self.case.ni - 2*self.case.tail + 1
        aoa = self.case.aoa if aoa is None else aoa
        xw, cp, cf, coeffs = np.empty(count), np.empty(count), np.empty(count), np.empty(6)
        self._ck(self.L.sgpu_surface(self.h, int(which_res), int(which_q), int(i_first), int(count), ctypes.c_double(aoa),
                                     _dp(xw), _dp(cp), _dp(cf), _dp(coeffs)))
        return dict(xw=xw, cp=cp, cf=cf, coeffs=coeffs)

    def surface_gradient(self, weights, which: int = 0, i_first: Optional[int] = None, count: Optional[int] = None,
                         aoa: Optional[float] = None) -> np.ndarray:
        """d/dq [nic][njc][nv] of sum_k weights[k]*coeffs[k], k = cl_pressure, cd_pressure, cl_viscous, cd_viscous of
        `surface(which, which)`: the adjoint right-hand side of a force objective."""
        if i_first is None:
            i_first, count = self.case.tail - 1, self.case.ni - 2*self.case.tail + 1
        aoa = self.case.aoa if aoa is None else aoa
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (4,)
        out = self._state_array()
        self._ck(self.L.sgpu_surface_gradient(self.h, int(which), int(i_first), int(count), ctypes.c_double(aoa), _dp(w), _dp(out)))
        return out

    def write_surface(self, path: str, **kw) -> dict:
        """the `<label>.surface` text file of the reference: `xw cp cf` per line, ostream default precision (io.cpp:225)"""
        s = self.surface(**kw)
        with open(path, "w") as f:
            for a, b, c in zip(s["xw"], s["cp"], s["cf"]):
                f.write("%g %g %g\n" % (a, b, c))
        return s

    # ---- the reference's solution files, fed from the device state (IOManager, src/utils/io.cpp:104-180)
    def write_restart(self, path: str, which: int = 0) -> None:
        from . import io as _io
        _io.write_restart(path, self.get_state(which))

    def read_restart(self, path: str, which: int = 0) -> None:
        from . import io as _io
        self.set_state(_io.read_restart(path, self.nic, self.njc, self.nv), which)

    def write_npz(self, path: str, which: int = 0) -> None:
        from . import io as _io
        _io.write_npz(path, self.case, self.get_state(which))

    def dres_dbeta(self) -> np.ndarray:
        """d rhs4 / d beta per cell at the device state (SA extension; field-inversion gradient building block)"""
        out = np.zeros((self.nic, self.njc))
        self._ck(self.L.sgpu_dres_dbeta(self.h, _dp(out)))
        return out

    # ---- halos (multi-GPU j-slabs)
    def halo_count(self) -> int:
        return int(self.L.sgpu_halo_count(self.h))

    def halo_pack(self, which: int, side: int, dev_ptr: int):
        self._ck(self.L.sgpu_halo_pack(self.h, which, side, ctypes.c_void_p(dev_ptr)))

    def halo_pack_ghost(self, which: int, side: int, dev_ptr: int):
        self._ck(self.L.sgpu_halo_pack_ghost(self.h, which, side, ctypes.c_void_p(dev_ptr)))

    def halo_unpack(self, which: int, side: int, dev_ptr: int):
        self._ck(self.L.sgpu_halo_unpack(self.h, which, side, ctypes.c_void_p(dev_ptr)))

    # peer-memory exchange (NVLink P2P)
    def halo_recv_buffer(self, side: int) -> int:
        p = ctypes.c_void_p()
        self._ck(self.L.sgpu_halo_recv_buffer(self.h, side, ctypes.byref(p)))
        return p.value

    def halo_enable_peer(self, peer_device: int):
        self._ck(self.L.sgpu_halo_enable_peer(self.h, peer_device))

    def halo_set_peer(self, side: int, dev_ptr: int):
        self._ck(self.L.sgpu_halo_set_peer(self.h, side, ctypes.c_void_p(dev_ptr)))

    def halo_ipc_handle(self, side: int) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._ck(self.L.sgpu_halo_ipc_handle(self.h, side, buf))
        return buf.raw

    def halo_open_peer(self, side: int, handle: bytes):
        self._ck(self.L.sgpu_halo_open_peer(self.h, side, ctypes.c_char_p(handle)))

    def halo_push(self, which: int = 0):
        self._ck(self.L.sgpu_halo_push(self.h, which))

    def halo_pull(self, which: int = 0):
        self._ck(self.L.sgpu_halo_pull(self.h, which))

    # ---- instrumentation
    def enable_kernel_timing(self, on: bool = True):
        self._ck(self.L.sgpu_enable_kernel_timing(self.h, int(on)))

    def kernel_times(self, n: int = 4096) -> np.ndarray:
        buf = (ctypes.c_float * n)()
        k = self.L.sgpu_kernel_times(self.h, buf, n)
        return np.array(buf[:k], dtype=np.float64)
