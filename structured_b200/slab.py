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

    def exchange(self, pack: Callable, unpack: Callable) -> None:
        if not self.nb:
            return
        dist = self.dist
        ops = []
        for side, nbr in self.nb.items():
            pack(side, self.send[side])
            ops.append(dist.P2POp(dist.isend, self.send[side], nbr))
            ops.append(dist.P2POp(dist.irecv, self.recv[side], nbr))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for side in self.nb:
            unpack(side, self.recv[side])

    def allreduce_sum(self, values: np.ndarray, device) -> np.ndarray:
        """sum of the per-slab sums of rhs^2 (src/solver/solver.cpp:125-134 split over slabs)"""
        import torch
        if self.world == 1:
            return values
        t = torch.as_tensor(values, dtype=torch.float64, device=device).clone()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()
