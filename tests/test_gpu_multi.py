"""Multi-GPU: two j-slabs on two B200s exchanging their ghost rows through NVLink peer memory (stores straight into the
neighbour's buffer + device-side sequence flags) reproduce the single-GPU residual bit for bit (needs >= 2 devices:
`gpurun --gpus 2`).  The two-PROCESS tests (implicit step, adjoint solve over torch.distributed) also run on a box with ONE
GPU: both ranks then share cuda:0 and talk over gloo with host-staged buffers (NCCL refuses two ranks on one device) --
same slab contexts, same kernels, same host logic, only the transport differs."""
import numpy as np
import pytest

from structured_b200.cases import turbulent_channel_case
from structured_b200.slab import HIGH, LOW

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("ntrans", [0, 1])
def test_two_gpus_peer_memory_halo_bit_for_bit(ntrans):
    from structured_b200.api import GpuEulerEquation
    nic, njc, split = 300, 96, 40
    case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=1e5)
    q = case.perturbed_q()
    one = GpuEulerEquation(case, device=0)
    want = one.calc_residual(q)
    lo = GpuEulerEquation(case, device=0, j_begin=0, j_end=split)
    hi = GpuEulerEquation(case, device=1, j_begin=split, j_end=njc)
    lo.halo_enable_peer(1); hi.halo_enable_peer(0)
    lo.halo_set_peer(HIGH, hi.halo_recv_buffer(LOW))
    hi.halo_set_peer(LOW, lo.halo_recv_buffer(HIGH))
    got = np.zeros_like(want)
    for rep in range(3):                                     # several exchanges: both receive slots and the flags cycle
        qq = q * (1.0 + 0.001 * rep)
        want = one.calc_residual(qq)
        # each slab uploads ONLY its own rows: the ghost rows can only come from the peer exchange
        lo.set_state_window(np.ascontiguousarray(qq[:, :split, :]), 0)
        hi.set_state_window(np.ascontiguousarray(qq[:, split:, :]), split)
        lo.halo_push(0); hi.halo_push(0)
        lo.halo_pull(0); hi.halo_pull(0)
        for s in (lo, hi):
            s.residual_device(0)
            s.get_rhs(out=got)
        assert np.array_equal(got, want), rep
    for s in (one, lo, hi):
        s.close()


def _init_rank(rank, world):
    """one rank per GPU over NCCL when the box has enough devices, else all ranks on cuda:0 over gloo"""
    import torch
    import torch.distributed as dist
    ndev = torch.cuda.device_count()
    idev = rank if ndev >= world else 0
    torch.cuda.set_device(idev)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", idev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return idev


def _implicit_worker(rank, world, port, nic, njc, ntrans, cfl, out):
    import os

    import torch
    import torch.distributed as dist
    from structured_b200.api import GpuEulerEquation
    from structured_b200.slab import HaloExchanger, SlabLinearSolver, partition_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    idev = _init_rank(rank, world)
    try:
        case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=2e4)
        q = case.perturbed_q(0.02)
        j0, j1 = partition_rows(njc, world)[rank]
        eq = GpuEulerEquation(case, device=idev, j_begin=j0, j_end=j1)
        eq.set_state_window(np.ascontiguousarray(q[:, j0:j1, :]), j0)        # own rows only: ghosts must come from the exchange
        dev = torch.device("cuda", idev)
        qhalo = HaloExchanger(rank, world, eq.halo_count(), dev, dist)
        solver = SlabLinearSolver(eq, rank, world, dist, dev)

        def exchange_state():
            qhalo.exchange(lambda s, t: eq.halo_pack(0, s, t.data_ptr()), lambda s, t: eq.halo_unpack(0, s, t.data_ptr()))

        l2, info = solver.implicit_step(cfl, 1.0, exchange_state=exchange_state, precond="line_j", restart=50, max_iter=1500, rtol=1e-12)
        rows = eq.get_state()[:, j0:j1, :]
        out.put((rank, j0, j1, rows, l2, info["converged"], info["iterations"]))
        eq.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 1, reason="needs a GPU")
@pytest.mark.parametrize("ntrans", [0, 1])
def test_two_gpu_implicit_step_over_nccl_equals_one_gpu(ntrans):
    """one implicit Solver::step on two j-slabs, two processes, NCCL (one GPU: both ranks on cuda:0 over gloo): ghost rows of q and of every Krylov operand
    exchanged, inner products all-reduced -- same new state as sgpu_implicit_step on one GPU"""
    import socket

    import torch.multiprocessing as mp
    from structured_b200.api import GpuEulerEquation
    nic, njc, cfl = 128, 80, 8.0
    case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=2e4)
    q = case.perturbed_q(0.02)
    one = GpuEulerEquation(case, device=0)
    one.set_state(q)
    l2_one, info = one.implicit_step(cfl, 1.0, precond="line_j", restart=50, max_iter=1500, rtol=1e-12)
    assert info["converged"]
    want = one.get_state()
    one.close()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_implicit_worker, args=(r, 2, port, nic, njc, ntrans, cfl, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    got = np.zeros_like(want)
    for rank, j0, j1, rows, l2, conv, iters in res:
        assert conv
        got[:, j0:j1, :] = rows
        assert np.abs(l2 - l2_one).max() <= 1e-10 * np.abs(l2_one).max()
    dq = want - q
    assert np.abs(got - want).max() <= 1e-8 * np.abs(dq).max(), np.abs(got - want).max() / np.abs(dq).max()


def _adjoint_worker(rank, world, port, nic, njc, ntrans, g, out):
    import os

    import torch
    import torch.distributed as dist
    from structured_b200.api import GpuEulerEquation
    from structured_b200.slab import SlabLinearSolver, partition_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    idev = _init_rank(rank, world)
    try:
        case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=2e4, periodic=False)
        q = case.perturbed_q(0.02)
        j0, j1 = partition_rows(njc, world)[rank]
        eq = GpuEulerEquation(case, device=idev, j_begin=j0, j_end=j1)
        eq.set_state(q)
        eq.jacobian_device()
        solver = SlabLinearSolver(eq, rank, world, dist, torch.device("cuda", idev))
        psi, info = solver.adjoint_solve(g, cfl=1e4, max_steps=12, tol=1e-9, precond="line_j", restart=60, max_iter=600, rtol=1e-6)
        rows = eq.vec_to_host(psi.data_ptr())[:, j0:j1, :]
        out.put((rank, j0, j1, rows, info))
        eq.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 1, reason="needs a GPU")
def test_two_gpu_adjoint_solve_equals_one_gpu():
    """J^T psi = -g on two j-slabs over NCCL (transposed products: each slab's share of the neighbour's rows is exchanged
    and added; inner products all-reduced) against sgpu_adjoint_solve on one GPU.  Laminar: the well-posed case."""
    import socket

    import torch.multiprocessing as mp
    from structured_b200.api import GpuEulerEquation
    nic, njc, ntrans = 96, 64, 0
    case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=2e4, periodic=False)
    q = case.perturbed_q(0.02)
    rng = np.random.default_rng(5)
    g = rng.standard_normal(q.shape)
    one = GpuEulerEquation(case, device=0)
    one.set_state(q); one.jacobian_device()
    want, winfo = one.adjoint_solve(g, cfl=1e4, max_steps=12, tol=1e-9, precond="line_j", restart=60, max_iter=600, rtol=1e-6)
    assert winfo["rel_residual"] <= 1e-9, winfo
    one.close()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_adjoint_worker, args=(r, 2, port, nic, njc, ntrans, g, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    got = np.zeros_like(want)
    for rank, j0, j1, rows, info in res:
        assert info["rel_residual"] <= 1e-9, info
        got[:, j0:j1, :] = rows
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()


def _ipc_halo_worker(rank, world, port, nic, njc, ntrans, out):
    import os

    import torch
    import torch.distributed as dist
    from structured_b200.api import GpuEulerEquation
    from structured_b200.slab import partition_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    idev = rank if ndev >= world else 0
    torch.cuda.set_device(idev)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # handles only; the rows travel through peer memory
    try:
        case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=1e5)
        j0, j1 = partition_rows(njc, world)[rank]
        eq = GpuEulerEquation(case, device=idev, j_begin=j0, j_end=j1)
        mine = (eq.halo_ipc_handle(LOW), eq.halo_ipc_handle(HIGH))
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        if rank > 0:
            eq.halo_open_peer(LOW, allh[rank - 1][HIGH])
        if rank < world - 1:
            eq.halo_open_peer(HIGH, allh[rank + 1][LOW])
        got = []
        for rep in range(3):                                         # several exchanges: both receive slots and the flags cycle
            q = case.perturbed_q() * (1.0 + 0.001 * rep)
            eq.set_state_window(np.ascontiguousarray(q[:, j0:j1, :]), j0)   # own rows only
            dist.barrier()
            eq.halo_push(0)
            eq.halo_pull(0)
            eq.residual_device(0)
            rows = np.zeros((nic, j1 - j0, case.nv))
            eq.get_rhs_window(rows)
            got.append(rows)
            eq.synchronize()
            dist.barrier()
        out.put((rank, j0, j1, got))
        dist.barrier()
        eq.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 1, reason="needs a GPU")
@pytest.mark.parametrize("ntrans", [0, 1])
def test_three_processes_cuda_ipc_halo_bit_for_bit(ntrans):
    """the transport bench.py uses at N > 1: one process per slab, receive buffers opened through CUDA IPC handles, boundary rows
    stored straight into the neighbour's memory, device-side sequence flags (sgpu_halo_ipc_handle / open_peer / push / pull).
    Three slabs (one has two neighbours) must reproduce the one-context residual bit for bit; with fewer than three GPUs the
    ranks share cuda:0 -- the same IPC path, the stores then stay on one device."""
    import socket

    import torch.multiprocessing as mp
    from structured_b200.api import GpuEulerEquation
    nic, njc, world = 200, 90, 3
    case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=1e5)
    one = GpuEulerEquation(case, device=0)
    want = [one.calc_residual(case.perturbed_q() * (1.0 + 0.001 * rep)) for rep in range(3)]
    one.close()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_ipc_halo_worker, args=(r, world, port, nic, njc, ntrans, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, j0, j1, got in res:
        for rep in range(3):
            assert np.array_equal(got[rep], want[rep][:, j0:j1, :]), (rank, rep)


def _explicit_slab_worker(rank, world, port, nic, njc, ntrans, nsteps, cfl, out):
    import os

    import torch
    import torch.distributed as dist
    from structured_b200.api import GpuEulerEquation
    from structured_b200.slab import partition_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    idev = rank if ndev >= world else 0
    torch.cuda.set_device(idev)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # handles and norms only; the rows travel through peer memory
    try:
        case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=1e5)
        j0, j1 = partition_rows(njc, world)[rank]
        eq = GpuEulerEquation(case, device=idev, j_begin=j0, j_end=j1)
        mine = (eq.halo_ipc_handle(LOW), eq.halo_ipc_handle(HIGH))
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        if rank > 0:
            eq.halo_open_peer(LOW, allh[rank - 1][HIGH])
        if rank < world - 1:
            eq.halo_open_peer(HIGH, allh[rank + 1][LOW])
        q = case.perturbed_q(0.01)
        own = np.ascontiguousarray(q[:, j0:j1, :])
        eq.set_state_window(own, j0, 0); eq.set_state_window(own, j0, 1)   # own rows only: every ghost row comes from the exchange
        dist.barrier()
        l2 = []
        for _ in range(nsteps):
            part = eq.explicit_step(cfl, "rk4_jameson") ** 2               # this slab's sums of rhs^2
            t = torch.as_tensor(part, dtype=torch.float64)
            dist.all_reduce(t)
            l2.append(np.sqrt(t.numpy()))
        rows = eq.get_state(0)[:, j0:j1, :]
        out.put((rank, j0, j1, rows, np.array(l2)))
        dist.barrier()
        eq.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 1, reason="needs a GPU")
@pytest.mark.parametrize("ntrans", [0, 1])
def test_explicit_steps_on_three_slabs_equal_one_context_bitwise(ntrans):
    """sgpu_explicit_step on j-slabs: one call per rank and step runs calc_dt and the four fused Runge-Kutta stages, each preceded
    by the ghost-row exchange over CUDA-IPC peer buffers (device-side flags, no host synchronisation inside the step).  Three
    processes reproduce the one-context states BIT FOR BIT after several steps; the norms agree to summation order."""
    import socket

    import torch.multiprocessing as mp
    from structured_b200.api import GpuEulerEquation
    nic, njc, world, nsteps, cfl = 150, 60, 3, 3, 0.5
    case = turbulent_channel_case(nic, njc, ntrans=ntrans, reynolds=1e5)
    q = case.perturbed_q(0.01)
    one = GpuEulerEquation(case, device=0)
    one.set_state(q, 0); one.set_state(q, 1)
    l2_one = np.array([one.explicit_step(cfl, "rk4_jameson") for _ in range(nsteps)])
    want = one.get_state(0)
    one.close()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_explicit_slab_worker, args=(r, world, port, nic, njc, ntrans, nsteps, cfl, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, j0, j1, rows, l2 in res:
        assert np.array_equal(rows, want[:, j0:j1, :]), rank
        assert np.abs(l2 - l2_one).max() <= 1e-12 * np.abs(l2_one).max()
