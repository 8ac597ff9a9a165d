"""Multi-GPU (needs >= 2 devices; run with `gpurun --gpus 2`): two j-slabs on two B200s exchanging their ghost rows
through NVLink peer memory (stores straight into the neighbour's buffer + device-side sequence flags) reproduce the
single-GPU residual bit for bit."""
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
