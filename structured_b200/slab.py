"""j-slab partition of a grid across ranks and the ghost-row exchange between neighbouring slabs.

Backend agnostic plumbing (torch.distributed: NCCL on the GPUs, gloo in the CPU tests).  The data path has
exactly one exchange step per residual evaluation -- two boundary cell rows of q to each neighbour, packed as
[nv][2][nic] doubles (the layout of sgpu_halo_pack / sgpu_halo_unpack) -- and one nv-double all-reduce per step for
the residual norms.  No other collective exists.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

LOW, HIGH = 0, 1  # side ids of the C ABI: 0 = low-j neighbour, 1 = high-j neighbour


def partition_rows(njc: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row ranges [j_begin, j_end) per rank; every slab keeps at least 2 rows (ghost depth)."""
    if world < 1 or njc < 2 * world:
        raise ValueError("need at least 2 cell rows per slab (njc=%d, world=%d)" % (njc, world))
    base, rem = divmod(njc, world)
    out, j = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((j, j + n))
        j += n
    return out


def neighbours(rank: int, world: int) -> dict:
    """side -> neighbour rank (absent at the physical bottom / top)"""
    nb = {}
    if rank > 0:
        nb[LOW] = rank - 1
    if rank < world - 1:
        nb[HIGH] = rank + 1
    return nb


def pack_rows_numpy(q_slab: np.ndarray, side: int) -> np.ndarray:
    """CPU statement of sgpu_halo_pack: q_slab is [nic][njl][nv] (owned rows); returns [nv][2][nic]."""
    rows = q_slab[:, :2, :] if side == LOW else q_slab[:, -2:, :]
    return np.ascontiguousarray(np.transpose(rows, (2, 1, 0)))


def unpack_rows_numpy(buf: np.ndarray, nic: int, nv: int) -> np.ndarray:
    """inverse layout change: [nv][2][nic] -> [nic][2][nv]"""
    return np.ascontiguousarray(np.transpose(buf.reshape(nv, 2, nic), (2, 1, 0)))


class HaloExchanger:
    """Owns the send/recv buffers of one rank and runs the neighbour exchange.

    pack(side, send_tensor) must fill the tensor with this slab's two boundary rows on that side;
    unpack(side, recv_tensor) must write the tensor into the slab's ghost rows on that side.
    """

    def __init__(self, rank: int, world: int, halo_count: int, device, dist_module=None):
        import torch
        self.rank, self.world = rank, world
        self.nb = neighbours(rank, world)
        self.dist = dist_module
        self.send = {s: torch.empty(halo_count, dtype=torch.float64, device=device) for s in self.nb}
        self.recv = {s: torch.empty(halo_count, dtype=torch.float64, device=device) for s in self.nb}
        # gloo moves host memory only: device buffers are staged through pinned host tensors (two ranks sharing ONE GPU in
        # the single-GPU form of the multi-process tests; NCCL refuses two ranks on one device)
        self.stage = None
        if dist_module is not None and torch.device(device).type == "cuda" and dist_module.is_initialized() and dist_module.get_backend() == "gloo":
            self.stage = {s: (torch.empty(halo_count, dtype=torch.float64).pin_memory(), torch.empty(halo_count, dtype=torch.float64).pin_memory())
                          for s in self.nb}

    def exchange(self, pack: Callable, unpack: Callable) -> None:
        if not self.nb:
            return
        dist = self.dist
        ops = []
        for side, nbr in self.nb.items():
            pack(side, self.send[side])
            if self.stage is None:
                ops.append(dist.P2POp(dist.isend, self.send[side], nbr))
                ops.append(dist.P2POp(dist.irecv, self.recv[side], nbr))
            else:
                hs, hr = self.stage[side]
                hs.copy_(self.send[side])                  # synchronous D2H on the current stream
                ops.append(dist.P2POp(dist.isend, hs, nbr))
                ops.append(dist.P2POp(dist.irecv, hr, nbr))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for side in self.nb:
            if self.stage is not None:
                self.recv[side].copy_(self.stage[side][1])
            unpack(side, self.recv[side])

    def allreduce_sum(self, values: np.ndarray, device) -> np.ndarray:
        """sum of the per-slab sums of rhs^2 (src/solver/solver.cpp:125-134 split over slabs)"""
        import torch
        if self.world == 1:
            return values
        t = torch.as_tensor(values, dtype=torch.float64, device=device).clone()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


# ---------------------------------------------------------------------------------------------------
# Slab-partitioned Krylov solve (SURVEY.md 8(f) N1 across GPUs).  Host logic only: the operator and the
# preconditioner are the CUDA kernels behind sgpu_op_apply / sgpu_precond_apply; what is distributed is
#   * the operand's two ghost rows per interior slab edge (HaloExchanger, before every product), and
#   * every inner product (one all-reduce of k+1 doubles per Gram-Schmidt pass).
# The same restarted, right-preconditioned GMRES as sgpu_linear_solve (csrc/linsolve_api.inl), written on torch
# tensors so that the CPU tests run it over gloo with a numpy-defined operator.
# ---------------------------------------------------------------------------------------------------
def distributed_gmres(apply_op: Callable, apply_pc: Callable, b, restart: int = 30, max_iter: int = 500, rtol: float = 1e-10,
                      allreduce: Callable = None, kernels=None):
    """Solve A x = b for the row block this rank owns.

    apply_op(x, out) / apply_pc(r, out) act on this rank's block (apply_op may exchange halos inside);
    allreduce(t) sums a small tensor over the ranks in place (None: single rank).
    kernels: an object with vec_dots / vec_gs_update / vec_scale_rsqrt taking device pointers (GpuEulerEquation): the
    Gram-Schmidt step then runs on the kernels of sgpu_linear_solve -- projections in one pass, update + |w|^2 in one pass,
    scaling from the device value -- with the all-reduces in between; None: plain torch operations (the CPU tests).
    Returns (x, info) with info = {iterations, converged, rel_residual}; rel_residual is the TRUE residual.
    """
    import math

    import torch
    n, m = b.numel(), int(restart)
    V = torch.zeros((m + 1, n), dtype=b.dtype, device=b.device)
    x = torch.zeros_like(b); w = torch.zeros_like(b); z = torch.zeros_like(b); u = torch.zeros_like(b)
    hbuf = torch.zeros(m + 2, dtype=b.dtype, device=b.device)

    def gsum(t):
        if allreduce is not None:
            allreduce(t)
        return t

    def norm(a):
        return math.sqrt(float(gsum(torch.dot(a, a).reshape(1))[0]))

    bnorm = norm(b)
    info = {"iterations": 0, "converged": True, "rel_residual": 0.0}
    if bnorm == 0.0:
        return x, info
    iters, beta, first = 0, bnorm, True
    while True:
        if first:
            w.copy_(b); first = False
        else:
            apply_op(x, w)
            torch.sub(b, w, out=w)
            beta = norm(w)
        info["rel_residual"] = beta / bnorm
        if beta <= rtol * bnorm or iters >= max_iter:
            break
        V[0].copy_(w).div_(beta)
        H = np.zeros((m + 1, m)); cs = np.zeros(m); sn = np.zeros(m); g = np.zeros(m + 1); g[0] = beta
        k, done = 0, False
        while k < m and iters < max_iter and not done:
            apply_pc(V[k], z)
            apply_op(z, w)
            if kernels is not None:
                kernels.vec_dots(w.data_ptr(), V.data_ptr(), k + 1, hbuf.data_ptr())
                gsum(hbuf[:k + 1])                                 # classical Gram-Schmidt, one all-reduce
                kernels.vec_gs_update(w.data_ptr(), V.data_ptr(), k + 1, hbuf.data_ptr(), hbuf[k + 1:].data_ptr())
                gsum(hbuf[k + 1:k + 2])
                kernels.vec_scale_rsqrt(V[k + 1].data_ptr(), w.data_ptr(), hbuf[k + 1:].data_ptr())
                hh = hbuf[:k + 2].cpu().numpy().copy()             # the one host synchronisation of the iteration
                hk1 = hh[k + 1] = math.sqrt(hh[k + 1])
            else:
                h = gsum(torch.mv(V[:k + 1], w))                   # classical Gram-Schmidt, one all-reduce
                w.addmv_(V[:k + 1].t(), h, alpha=-1.0)
                hk1 = norm(w)
                hh = np.concatenate([h.cpu().numpy(), [hk1]])
                if hk1 > 0.0:
                    V[k + 1].copy_(w).div_(hk1)
            for j in range(k):                                     # Givens rotations
                t = cs[j] * hh[j] + sn[j] * hh[j + 1]
                hh[j + 1] = -sn[j] * hh[j] + cs[j] * hh[j + 1]
                hh[j] = t
            den = math.hypot(hh[k], hh[k + 1])
            cs[k], sn[k] = (1.0, 0.0) if den == 0.0 else (hh[k] / den, hh[k + 1] / den)
            hh[k] = cs[k] * hh[k] + sn[k] * hh[k + 1]
            g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k]
            H[:k + 1, k] = hh[:k + 1]
            if abs(g[k + 1]) <= rtol * bnorm or hk1 == 0.0:
                done = True
            k += 1; iters += 1
        if k == 0:
            break
        y = np.zeros(k)
        for j in range(k - 1, -1, -1):
            s = g[j] - H[j, j + 1:k] @ y[j + 1:k]
            y[j] = s / H[j, j] if H[j, j] != 0.0 else 0.0
        torch.mv(V[:k].t(), torch.as_tensor(y, dtype=b.dtype, device=b.device), out=u)
        apply_pc(u, z)
        x.add_(z)
    info["iterations"] = iters
    info["converged"] = info["rel_residual"] <= rtol
    return x, info


class SlabLinearSolver:
    """linearsolver->set_lhs / set_rhs / solve_and_update (src/solver/solver.cpp:172-175) for ONE j-slab of a grid:
    every rank holds a GpuEulerEquation built with (j_begin, j_end) and runs the same calls."""

    def __init__(self, eq, rank: int, world: int, dist_module=None, device=None):
        import torch
        self.eq, self.rank, self.world, self.dist = eq, rank, world, dist_module
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.halo = HaloExchanger(rank, world, eq.halo_count(), self.device, dist_module)
        self.n = eq.vec_size()

    def _allreduce(self, t):
        if self.world > 1:
            if self.halo.stage is not None:                # gloo with device tensors: reduce a host copy
                h = t.cpu()
                self.dist.all_reduce(h, op=self.dist.ReduceOp.SUM)
                t.copy_(h)
            else:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)

    def apply_op(self, matrix: str, x, out) -> None:
        """out = A x on this slab's rows.  Forward matrices ("lhs", "J") need the operand's two ghost rows per interior edge
        BEFORE the product; transposed ones ("JT", "lhsT") need none, but each slab's rows also contribute to the
        neighbour's boundary cells: those contributions are left in the ghost rows of `out`, sent over and ADDED there --
        the transpose of the operand exchange (one more exchange of two block rows)."""
        eq = self.eq
        if matrix in ("JT", "lhsT"):
            eq.op_apply(matrix, x.data_ptr(), out.data_ptr())
            self.halo.exchange(lambda side, t: eq.vec_halo_pack_ghost(out.data_ptr(), side, t.data_ptr()),
                               lambda side, t: eq.vec_halo_add(out.data_ptr(), side, t.data_ptr()))
        else:
            self.halo.exchange(lambda side, t: eq.vec_halo_pack(x.data_ptr(), side, t.data_ptr()),
                               lambda side, t: eq.vec_halo_unpack(x.data_ptr(), side, t.data_ptr()))
            eq.op_apply(matrix, x.data_ptr(), out.data_ptr())

    def solve(self, matrix: str = "lhs", precond: str = "line_j", restart: int = 30, max_iter: int = 500, rtol: float = 1e-10,
              b=None, setup: bool = True):
        """A x = b; b = None: the device rhs of the last residual (what linearsolver->set_rhs receives).  Uses the Jacobian
        of the last jacobian_device(); returns (x as a device tensor laid out as state planes, info)."""
        import torch
        eq = self.eq
        if b is None:
            b = torch.zeros(self.n, dtype=torch.float64, device=self.device)
            eq.vec_from_rhs(b.data_ptr())
        if setup:
            eq.precond_setup(matrix, precond)

        def apply_pc(r, out):
            eq.precond_apply(matrix, precond, r.data_ptr(), out.data_ptr())

        return distributed_gmres(lambda x, out: self.apply_op(matrix, x, out), apply_pc, b, restart, max_iter, rtol,
                                 self._allreduce if self.world > 1 else None, kernels=eq if b.is_cuda else None)

    def adjoint_solve(self, g_host: np.ndarray, cfl: float = 100.0, max_steps: int = 50, tol: float = 1e-8, precond: str = "line_j",
                      restart: int = 40, max_iter: int = 400, rtol: float = 1e-3):
        """sgpu_adjoint_solve on a slab partition: J^T psi = -g by pseudo-time continuation, (delta/dt - J^T) dpsi = g + J^T psi
        per step with transposed-LHS GMRES solves, psi += dpsi (A22; no reference code -- "parity unpinned").
        g_host: GLOBAL [nic][njc][nv] objective gradient (every rank passes the same array, only its rows are read).
        Returns (psi as a device tensor of this slab, info)."""
        import math

        import torch
        eq = self.eq
        g = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        eq.vec_from_host(g_host, g.data_ptr())
        psi = torch.zeros_like(g); r = torch.zeros_like(g)
        eq.calc_dt(cfl)

        def gnorm(a):
            t = torch.dot(a, a).reshape(1)
            self._allreduce(t)
            return math.sqrt(float(t[0]))
        g0 = gnorm(g)
        info = {"steps": 0, "iterations": 0, "rel_residual": 0.0, "converged": True}
        if g0 == 0.0:
            return psi, info
        steps = 0
        while True:
            self.apply_op("JT", psi, r)
            r.add_(g)
            rel = gnorm(r) / g0
            info["rel_residual"] = rel
            if rel <= tol or steps >= max_steps:
                break
            dpsi, gi = self.solve("lhsT", precond, restart, max_iter, rtol, b=r, setup=(steps == 0))
            info["iterations"] += gi["iterations"]
            psi.add_(dpsi)
            steps += 1
        info["steps"] = steps
        info["converged"] = info["rel_residual"] <= tol
        return psi, info

    def implicit_step(self, cfl: float, under_relaxation: float = 1.0, exchange_state: Callable = None, **kw):
        """The implicit branch of Solver::step on a slab partition: dt, residual, Jacobian, distributed GMRES, update.
        exchange_state() must refresh the ghost rows of q (the caller's residual halo exchange) before the residual."""
        eq = self.eq
        if exchange_state is not None:
            exchange_state()
        eq.calc_dt(cfl)
        l2 = eq.residual_device(0, norms=True)
        l2 = self.halo.allreduce_sum(np.asarray(l2), self.device)
        eq.jacobian_device()
        x, info = self.solve("lhs", **kw)
        eq.vec_add_to_state(x.data_ptr(), under_relaxation)
        return np.sqrt(l2), info
