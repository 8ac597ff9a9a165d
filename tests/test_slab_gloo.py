"""CPU suite: the N > 1 host logic -- slab partition + ghost-row exchange + norm reduction -- with two gloo ranks."""
import os
import socket

import numpy as np
import pytest

from structured_b200.slab import HIGH, LOW, HaloExchanger, neighbours, pack_rows_numpy, partition_rows, unpack_rows_numpy


def test_partition_rows():
    assert partition_rows(8192, 8) == [(k * 1024, (k + 1) * 1024) for k in range(8)]
    p = partition_rows(150, 4)
    assert p[0][0] == 0 and p[-1][1] == 150 and all(a[1] == b[0] for a, b in zip(p, p[1:]))
    assert max(b - a for a, b in p) - min(b - a for a, b in p) <= 1
    with pytest.raises(ValueError):
        partition_rows(5, 3)
    assert neighbours(0, 4) == {HIGH: 1} and neighbours(3, 4) == {LOW: 2} and neighbours(1, 4) == {LOW: 0, HIGH: 2}


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nic, njc, nv, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        q = rng.standard_normal((nic, njc, nv))                  # the global state (every rank can regenerate it)
        j0, j1 = partition_rows(njc, world)[rank]
        own = np.ascontiguousarray(q[:, j0:j1, :])
        ghosts = {}
        ex = HaloExchanger(rank, world, 2 * nv * nic, "cpu", dist)

        def pack(side, t):
            t.copy_(torch.from_numpy(pack_rows_numpy(own, side).reshape(-1)))

        def unpack(side, t):
            ghosts[side] = unpack_rows_numpy(t.numpy(), nic, nv)

        ex.exchange(pack, unpack)
        ok = True
        if LOW in ex.nb:
            ok &= np.array_equal(ghosts[LOW], q[:, j0 - 2:j0, :])     # the low neighbour's top two rows
        if HIGH in ex.nb:
            ok &= np.array_equal(ghosts[HIGH], q[:, j1:j1 + 2, :])    # the high neighbour's bottom two rows
        part = (own ** 2).sum(axis=(0, 1))
        tot = ex.allreduce_sum(part, "cpu")
        ok &= np.allclose(tot, (q ** 2).sum(axis=(0, 1)), rtol=1e-13)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_norms_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 37, 23, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _poisson_like(n):
    """a non-symmetric, diagonally dominant pentadiagonal matrix (radius-2 stencil like the j-direction of the Jacobian)"""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    d = 4.0 + rng.random(n)
    return sp.diags([0.3 * rng.random(n - 2), -1.0 - 0.2 * rng.random(n - 1), d, -0.7 + 0.1 * rng.random(n - 1), 0.2 * rng.random(n - 2)],
                    [-2, -1, 0, 1, 2]).tocsr()


def test_distributed_gmres_single_rank_matches_direct_solve():
    import scipy.sparse.linalg as spla
    import torch
    from structured_b200.slab import distributed_gmres
    n = 300
    A = _poisson_like(n)
    b = np.random.default_rng(4).standard_normal(n)
    dinv = 1.0 / A.diagonal()

    def op(x, out):
        out.copy_(torch.from_numpy(A @ x.numpy()))

    def pc(r, out):
        out.copy_(r * torch.from_numpy(dinv))

    x, info = distributed_gmres(op, pc, torch.from_numpy(b), restart=25, max_iter=400, rtol=1e-12)
    assert info["converged"] and np.abs(x.numpy() - spla.spsolve(A.tocsc(), b)).max() <= 1e-9


def _gmres_worker(rank, world, port, n, out):
    import scipy.sparse.linalg as spla
    import torch
    import torch.distributed as dist
    from structured_b200.slab import distributed_gmres
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = _poisson_like(n)
        b = np.random.default_rng(4).standard_normal(n)
        r0, r1 = partition_rows(n, world)[rank]
        Aloc = A[r0:r1, :]
        dinv = torch.from_numpy(1.0 / A.diagonal()[r0:r1])
        nb = neighbours(rank, world)

        def op(x, out):
            # the operand's two boundary entries per side go to the neighbours (the vector form of the ghost-row exchange)
            ops, recv = [], {}
            for side, peer in nb.items():
                send = (x[:2] if side == LOW else x[-2:]).clone()
                recv[side] = torch.empty(2, dtype=torch.float64)
                ops += [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv[side], peer)]
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            full = np.zeros(n)
            full[r0:r1] = x.numpy()
            if LOW in nb:
                full[r0 - 2:r0] = recv[LOW].numpy()
            if HIGH in nb:
                full[r1:r1 + 2] = recv[HIGH].numpy()
            out.copy_(torch.from_numpy(Aloc @ full))

        def pc(r, o):
            o.copy_(r * dinv)

        def allreduce(t):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

        x, info = distributed_gmres(op, pc, torch.from_numpy(b[r0:r1].copy()), restart=25, max_iter=400, rtol=1e-12, allreduce=allreduce)
        want = spla.spsolve(A.tocsc(), b)[r0:r1]
        out.put((rank, bool(info["converged"] and np.abs(x.numpy() - want).max() <= 1e-9), info["iterations"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_gmres_gloo(world):
    """row-block partitioned GMRES: halo exchange inside the operator, all-reduced inner products; every rank must see
    the same iteration count and its block of the direct solution"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gmres_worker, args=(r, world, port, 301, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[:2] for r in res) == [(r, True) for r in range(world)]
    assert len({r[2] for r in res}) == 1


class _NumpySlabEq:
    """Stand-in for GpuEulerEquation's vector entry points on one row block of a NON-SYMMETRIC banded matrix (half
    bandwidth 2 = the two ghost rows of the slab exchange): lets the CPU suite run SlabLinearSolver's transposed products
    and adjoint continuation over gloo.  A vector = [2 ghost | owned rows | 2 ghost] entries."""

    def __init__(self, A, r0, r1, dt):
        self.A, self.r0, self.r1, self.n = A.tocsr(), r0, r1, A.shape[0]
        self.dt = dt
        self.nloc = r1 - r0

    def vec_size(self):
        return self.nloc + 4

    def halo_count(self):
        return 2

    def _view(self, ptr, n):
        import ctypes
        return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))

    def calc_dt(self, cfl):
        pass

    def vec_from_host(self, host, ptr):
        v = self._view(ptr, self.vec_size()); v[:] = 0.0; v[2:2 + self.nloc] = host[self.r0:self.r1]

    def _rows(self, side, ghost):
        if side == 0:
            return slice(0, 2) if ghost else slice(2, 4)
        return slice(self.nloc + 2, self.nloc + 4) if ghost else slice(self.nloc, self.nloc + 2)

    def vec_halo_pack(self, vptr, side, bptr):
        self._view(bptr, 2)[:] = self._view(vptr, self.vec_size())[self._rows(side, False)]

    def vec_halo_unpack(self, vptr, side, bptr):
        self._view(vptr, self.vec_size())[self._rows(side, True)] = self._view(bptr, 2)

    def vec_halo_pack_ghost(self, vptr, side, bptr):
        v = self._view(vptr, self.vec_size())
        self._view(bptr, 2)[:] = v[self._rows(side, True)]; v[self._rows(side, True)] = 0.0

    def vec_halo_add(self, vptr, side, bptr):
        self._view(vptr, self.vec_size())[self._rows(side, False)] += self._view(bptr, 2)

    def _matrix(self, name):
        import scipy.sparse as sp
        A = self.A if name in ("J", "JT") else (sp.diags(1.0 / self.dt) - self.A).tocsr()
        return A, name in ("JT", "lhsT")

    def op_apply(self, name, xptr, yptr):
        A, tr = self._matrix(name)
        x, y = self._view(xptr, self.vec_size()), self._view(yptr, self.vec_size())
        lo, hi = max(self.r0 - 2, 0), min(self.r1 + 2, self.n)
        full = np.zeros(self.n)
        if not tr:                                             # needs the operand's ghost entries
            full[lo:hi] = x[2 - (self.r0 - lo):2 + self.nloc + (hi - self.r1)]
            y[:] = 0.0; y[2:2 + self.nloc] = (A[self.r0:self.r1, :] @ full)
        else:                                                  # own rows' contributions, incl. to the neighbours' entries
            full[self.r0:self.r1] = x[2:2 + self.nloc]
            c = A[self.r0:self.r1, :].T @ x[2:2 + self.nloc]
            y[:] = 0.0; y[2 - (self.r0 - lo):2 + self.nloc + (hi - self.r1)] = c[lo:hi]

    def precond_setup(self, name, precond):
        A, _ = self._matrix(name)
        self.dinv = 1.0 / A.diagonal()[self.r0:self.r1]

    def precond_apply(self, name, precond, rptr, zptr):
        r, z = self._view(rptr, self.vec_size()), self._view(zptr, self.vec_size())
        z[:] = 0.0; z[2:2 + self.nloc] = r[2:2 + self.nloc] * self.dinv


def _adjoint_worker(rank, world, port, n, out):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    import torch.distributed as dist
    from structured_b200.slab import SlabLinearSolver
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(9)
        J = -(_poisson_like(n) + sp.diags([0.3 * rng.standard_normal(n - 1), 0.2 * rng.standard_normal(n - 2)], [1, -2]))   # stable, non-symmetric
        g = rng.standard_normal(n)
        dt = np.full(n, 50.0)
        r0, r1 = partition_rows(n, world)[rank]
        eq = _NumpySlabEq(J, r0, r1, dt)
        solver = SlabLinearSolver(eq, rank, world, dist, torch.device("cpu"))
        # transposed product across the blocks
        x = rng.standard_normal(n)
        xv = torch.zeros(eq.vec_size(), dtype=torch.float64); yv = torch.zeros_like(xv)
        eq.vec_from_host(x, xv.data_ptr())
        solver.apply_op("JT", xv, yv)
        ok_t = np.abs(yv.numpy()[2:2 + eq.nloc] - (J.T @ x)[r0:r1]).max() <= 1e-12 and np.all(yv.numpy()[:2] == 0) and np.all(yv.numpy()[-2:] == 0)
        psi, info = solver.adjoint_solve(g, cfl=1.0, max_steps=60, tol=1e-10, precond="block_jacobi", restart=30, max_iter=300, rtol=1e-8)
        want = spla.spsolve(J.T.tocsc(), -g)[r0:r1]
        ok = bool(ok_t and info["converged"] and np.abs(psi.numpy()[2:2 + eq.nloc] - want).max() <= 1e-7 * np.abs(want).max())
        out.put((rank, ok, info["steps"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_transposed_products_and_adjoint_continuation_gloo(world):
    """SlabLinearSolver.apply_op("JT") (reverse halo: pack ghost contributions, add at the neighbour) and adjoint_solve
    (J^T psi = -g by pseudo-time continuation with transposed-LHS GMRES) on row blocks of a non-symmetric banded matrix"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_adjoint_worker, args=(r, world, port, 240, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[:2] for r in res) == [(r, True) for r in range(world)], res
    assert len({r[2] for r in res}) == 1
