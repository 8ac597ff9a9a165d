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
