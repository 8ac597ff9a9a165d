"""GPU parity tests proper: the CUDA path (through the C ABI) against the reference's golden outputs and the CPU oracle."""
import os

import numpy as np
import pytest

from helpers import TOL, field_rel_err, golden, golden_state
from structured_b200.cases import ZOO, case_from_toml, turbulent_channel_case, zoo_case

pytestmark = pytest.mark.gpu

CASES = ["channel", "naca0012"] + ["zoo_" + z for z in ZOO]


def gpu_eq(case, **kw):
    from structured_b200.api import GpuEulerEquation
    return GpuEulerEquation(case, **kw)


@pytest.mark.parametrize("name", CASES)
def test_residual_matches_reference_golden(name):
    case, z = golden(name)
    eq = gpu_eq(case)
    q = golden_state(case, z)
    for lhs, key in ((False, "rhs"), (True, "rhs_lhs")):
        rhs = eq.calc_residual(q, lhs=lhs)
        err = field_rel_err(rhs, z[key])
        assert err.max() <= TOL, (name, key, err)
    eq.close()


@pytest.mark.parametrize("name", CASES)
def test_metrics_and_dt_match_reference_golden(name):
    case, z = golden(name)
    eq = gpu_eq(case)
    from oracle.bindings import PortOracle
    port = PortOracle(case)                           # port metrics are bit-identical to the reference's (test_oracle)
    for a, b in zip(eq.metrics(), port.metrics()):
        assert np.array_equal(a, b)
    q = golden_state(case, z)
    eq.set_state(q)
    eq.calc_dt(float(z["cfl_dt"]))
    dt = eq.get_dt()
    assert field_rel_err(dt[..., 0], z["dt"]).max() <= TOL
    for k in range(1, case.nv):
        assert np.array_equal(dt[..., k], dt[..., 0])
    eq.close(); port.close()


@pytest.mark.parametrize("name", ["channel", "naca0012"])
def test_explicit_steps_match_stock_reference_binary(name):
    case, z = golden(name)
    ex = case_from_toml(str(z["explicit_inp"]), z["xv"], z["yv"])
    eq = gpu_eq(ex)
    eq.initialize()
    l2 = None
    for _ in range(int(z["explicit_steps"])):
        l2 = eq.explicit_step(ex.cfl)
    q = eq.get_state()
    # rounding differences are amplified over steps; 10-20 rk4 steps stay far below 1e-10
    assert field_rel_err(q, z["explicit_q"]).max() <= 1e-10
    last = str(z["explicit_history"]).strip().splitlines()[-1].split()
    for k in range(4):
        assert abs(float(last[-4 + k]) - l2[k]) <= 0.006 * l2[k]
    eq.close()


@pytest.mark.parametrize("nic,njc,order,flux", [(40, 24, 2, "roe"), (131, 37, 2, "roe"), (300, 70, 1, "ausm"), (257, 129, 2, "ausm")])
def test_sa_residual_matches_oracle(nic, njc, order, flux):
    """SA extension ("parity unpinned": the oracle is our own CPU statement of the spec); sizes straddle strip/chunk seams"""
    from oracle.bindings import PortOracle
    case = turbulent_channel_case(nic, njc, ntrans=1, order=order, flux=flux, reynolds=2e4)
    port = PortOracle(case); eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    for lhs in (False, True):
        err = field_rel_err(eq.calc_residual(q, lhs=lhs), port.residual(q, lhs))
        assert err.max() <= TOL, err
    eq.close(); port.close()


@pytest.mark.parametrize("ntrans", [0, 1])
def test_laminar_large_matches_oracle_and_norms(ntrans):
    from oracle.bindings import PortOracle
    case = turbulent_channel_case(520, 300, ntrans=ntrans, reynolds=1e5)
    port = PortOracle(case); eq = gpu_eq(case)
    q = case.perturbed_q()
    ref = port.residual(q)
    eq.set_state(q)
    l2 = eq.residual_device(0, norms=True)
    rhs = eq.get_rhs()
    assert field_rel_err(rhs, ref).max() <= TOL
    want = (ref ** 2).sum(axis=(0, 1))
    assert np.abs(l2 - want).max() <= 1e-12 * want.max()
    eq.close(); port.close()


def test_freestream_preservation():
    """uniform state + freestream BCs on a curvilinear grid: the inviscid rhs is 0 to round-off (SURVEY.md section 4
    item 3).  The reference's viscous dual cells are not closed on a curvilinear grid, so with mu > 0 the
    (non-zero) result must simply equal the oracle's."""
    from oracle.bindings import PortOracle
    from structured_b200.cases import Boundary
    case = zoo_case("C", 96, 40)
    case.boundaries = [Boundary("freestream", f, 0, -1) for f in ("bottom", "top", "left", "right")]
    eq = gpu_eq(case)
    rhs = eq.calc_residual(case.freestream_q())
    nchi, neta, vol = eq.metrics()
    scale = (case.p_inf + case.rho_inf * case.u_inf ** 2) * np.abs(nchi).max() / vol.min()
    assert np.abs(rhs).max() <= 1e-13 * scale
    eq.close()
    case.mu_inf = 1e-3
    eq = gpu_eq(case); port = PortOracle(case)
    want = port.residual(case.freestream_q())
    assert np.abs(want).max() > 1e-3
    # pure cancellation noise of O(u n / V) terms: compare at the scale of the cancelling terms, not of their tiny sum
    assert field_rel_err(eq.calc_residual(case.freestream_q())[..., 1:], want[..., 1:]).max() <= 1e-9
    eq.close(); port.close()


@pytest.mark.parametrize("ntrans,split", [(0, 17), (1, 40), (1, 2)])
def test_two_slabs_equal_one_slab_bit_for_bit(ntrans, split):
    """j-slab partition on ONE device: two contexts + halo exchange reproduce the single-context residual exactly"""
    import torch
    case = turbulent_channel_case(150, 64, ntrans=ntrans, reynolds=1e5)
    q = case.perturbed_q()
    one = gpu_eq(case)
    want = one.calc_residual(q)
    lo = gpu_eq(case, j_begin=0, j_end=split)
    hi = gpu_eq(case, j_begin=split, j_end=case.njc)
    # set_state already reads the neighbour rows from the global host array ...
    got = np.zeros_like(want)
    for s in (lo, hi):
        s.set_state(q)
        s.residual_device(0)
        s.get_rhs(out=got)
    assert np.array_equal(got, want)
    # ... and the device-side halo exchange must deliver the same ghost rows
    n = lo.halo_count()
    buf = torch.empty(n, dtype=torch.float64, device="cuda")
    qz = q.copy(); qz[:, split - 2:split + 2, :] *= 1.01           # host arrays the slabs have NOT seen
    lo.set_state(qz); hi.set_state(qz)
    want2 = one.calc_residual(qz)
    # corrupt ghosts, then repair them through pack/unpack
    junk = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    lo.halo_unpack(0, 1, junk.data_ptr()); hi.halo_unpack(0, 0, junk.data_ptr())
    lo.halo_pack(0, 1, buf.data_ptr()); hi.halo_unpack(0, 0, buf.data_ptr())
    hi.halo_pack(0, 0, buf.data_ptr()); lo.halo_unpack(0, 1, buf.data_ptr())
    got2 = np.zeros_like(want2)
    for s in (lo, hi):
        s.residual_device(0)
        s.get_rhs(out=got2)
    assert np.array_equal(got2, want2)
    for s in (one, lo, hi):
        s.close()


def test_windowed_slab_inputs_equal_global_inputs():
    """a rank that only generates / uploads its own rows of grid, fields and state gets the same residual rows"""
    nic, njc, split = 96, 48, 20
    full = turbulent_channel_case(nic, njc, ntrans=1, reynolds=1e5)
    one = gpu_eq(full)
    want = one.calc_residual(full.perturbed_q())
    for (j0, j1) in ((0, split), (split, njc)):
        ja, jb = max(j0 - 2, 0), min(j1 + 2, njc)
        part = turbulent_channel_case(nic, njc, ntrans=1, reynolds=1e5, cell_rows=(ja, jb))
        eq = gpu_eq(part, j_begin=j0, j_end=j1, window=part.window)
        eq.set_state_window(part.perturbed_q(j_first=ja, j_count=jb - ja), ja)
        eq.residual_device(0)
        got = np.zeros((nic, j1 - j0, 5))
        eq.get_rhs_window(got)
        assert np.array_equal(got, want[:, j0:j1, :])
        eq.close()
    one.close()


@pytest.mark.parametrize("nic,njc", [(2, 2), (3, 5), (60, 9), (61, 7), (121, 3), (124, 9), (125, 7), (249, 3)])
def test_ragged_and_minimum_sizes(nic, njc):
    """edge sizes: the smallest grid the ABI accepts, strips that end exactly at / one past a strip seam, few rows"""
    from oracle.bindings import PortOracle
    from structured_b200.cases import Boundary
    case = zoo_case("A", nic, njc)
    case.boundaries = [Boundary("wall", "bottom", 1, -2), Boundary("freestream", "top", 0, -1),
                       Boundary("freestream", "left", 0, -1), Boundary("outflow", "right", 0, -1)]
    port = PortOracle(case); eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    for lhs in (False, True):
        assert field_rel_err(eq.calc_residual(q, lhs=lhs), port.residual(q, lhs)).max() <= TOL
    eq.close(); port.close()


def test_error_behaviour_mirrors_reference_messages():
    from structured_b200.api import GpuEulerEquation, SgpuError
    from structured_b200.cases import Boundary
    case = zoo_case("A")
    case.boundaries.append(Boundary("slipwall", "left", 1, -2))          # "Boundary condition not implemented!" (bc.cpp:116-126)
    with pytest.raises(SgpuError, match="not implemented"):
        GpuEulerEquation(case)
    case = zoo_case("A"); case.flux = "hllc"
    with pytest.raises(SgpuError, match="Flux not found"):
        GpuEulerEquation(case)
    case = zoo_case("A")
    eq = GpuEulerEquation(case)
    with pytest.raises(SgpuError, match="scheme not defined"):
        eq.explicit_step(0.1, "rk3")
    with pytest.raises(SgpuError, match="calc_dt"):
        eq.set_state(case.perturbed_q()); eq.jacobian_coo(apply_lhs_transform=True)
    eq.close()


def test_full_size_4096_matches_oracle_and_properties():
    """BASELINE.json's 1-GPU size (4096 x 4096, SA), the bench workload itself: the residual of EVERY cell against the
    CPU oracle (one oracle evaluation, ~10 s), then size-independent properties -- device norms == norms of the downloaded
    field; two j-slabs reproduce the one-slab field bit for bit (checksum of checksums + sampled rows)."""
    from oracle.bindings import PortOracle
    n = 4096
    case = turbulent_channel_case(n, n, ntrans=1)
    q = case.perturbed_q()
    eq = gpu_eq(case)
    eq.set_state(q)
    l2 = eq.residual_device(0, norms=True)
    rhs = eq.get_rhs()
    assert np.isfinite(rhs).all()
    port = PortOracle(case)
    want_rhs = port.residual(q)
    port.close()
    err = field_rel_err(rhs, want_rhs)
    assert err.max() <= TOL, err
    del want_rhs, port
    # the Jacobian of the same 16.8 M-cell state (43.6 GB on the device): rows of ~2 600 cells -- bottom wall, top wall, interior --
    # against the oracle evaluated on crops of the grid (tests/helpers.py::crop_case)
    from helpers import TOL_SA_COUPLING, jac_rel_err_split, jac_worst, oracle_on_crop, rows_of_cells
    from test_gpu_jacobian import check_pattern
    slots, ms = eq.jacobian_device()
    assert slots == 13 and ms > 0
    ncells = 0
    for box in ((2000, 2034, 0, 28), (777, 811, n - 28, n), (3000, 3034, 2040, 2074)):
        _, (i0, i1, j0, j1), ref = oracle_on_crop(case, q, box)
        ri, ci, va = eq.jacobian_coo(rows=(j0, j1 - j0))
        cells = [(i, j) for i in range(i0, i1) for j in range(j0, j1)]
        keep = np.isin(ri, rows_of_cells(n, 5, cells))
        ours = (ri[keep], ci[keep], va[keep])
        e1, e2 = jac_rel_err_split(q.size, ours, ref, 5)
        assert e1 <= TOL and e2 <= TOL_SA_COUPLING, (box, e1, e2, jac_worst(q.size, ours, ref, 5, n))
        check_pattern(q.size, ours, ref)
        ncells += len(cells)
    assert ncells >= 2000
    want = np.einsum("ijk,ijk->k", rhs, rhs)
    assert np.abs(l2 - want).max() <= 1e-12 * want.max()
    col_sum = rhs.sum(axis=0)                                  # [njc][nv] checksum per row
    sample = rhs[::511, :, :].copy()
    eq.close()
    del rhs
    split = 1500
    got_sum = np.zeros_like(col_sum)
    for (j0, j1) in ((0, split), (split, n)):
        ja, jb = max(j0 - 2, 0), min(j1 + 2, n)
        s = gpu_eq(case, j_begin=j0, j_end=j1)
        s.set_state_window(np.ascontiguousarray(q[:, ja:jb, :]), ja)
        s.residual_device(0)
        part = np.zeros((n, j1 - j0, 5))
        s.get_rhs_window(part)
        got_sum[j0:j1] = part.sum(axis=0)
        assert np.array_equal(part[::511], sample[:, j0:j1, :])
        s.close()
    assert np.array_equal(got_sum, col_sum)


@pytest.mark.parametrize("periodic", ["left", "right", None])
@pytest.mark.parametrize("ntrans", [0, 1])
def test_host_column_pipeline_equals_device_path_bitwise(ntrans, periodic):
    """sgpu_residual_host pipelines column chunks (H2D | BCs + kernel | D2H); every chunking must give the bits of the
    device-resident path, including the corner ghosts a periodic side copies from the wrap-around column -- whichever
    face the periodic table names (the reference fills both ghost columns for either, bc.cpp:329-365).  The pipelined
    path runs on a context whose ghost cells hold ANOTHER state's values, so a boundary condition skipped in some
    chunk cannot hide behind ghosts left over from the device-path evaluation."""
    case = turbulent_channel_case(430, 52, ntrans=ntrans, reynolds=2e4, periodic=periodic is not None)
    if periodic == "right":
        for b in case.boundaries:
            if b.type == "periodic":
                b.face = "right"
    q = case.perturbed_q(0.02)
    ref = gpu_eq(case)
    ref.set_state(q)
    ref.residual_device()
    dev = ref.get_rhs()
    ref.close()
    eq = gpu_eq(case)
    for chunks in ("2", "3", "7", "70"):
        eq.set_state(case.perturbed_q(0.05) * 1.1)             # poison: ghosts of a different state
        eq.residual_device()
        os.environ["SGPU_PIPE_CHUNKS"] = chunks
        try:
            assert np.array_equal(eq.calc_residual(q), dev), chunks
        finally:
            del os.environ["SGPU_PIPE_CHUNKS"]
    eq.close()


def test_solution_files_roundtrip_through_device(tmp_path):
    """`.out` / `.npz` (IOManager::write_restart, read_restart, write_npz, src/utils/io.cpp:104-180) fed from the device planes"""
    from structured_b200 import io as sio
    case = turbulent_channel_case(37, 21, ntrans=1)
    q = case.perturbed_q(0.02)
    eq = gpu_eq(case)
    eq.set_state(q)
    eq.write_restart(str(tmp_path / "a.out"))
    assert (tmp_path / "a.out").read_bytes() == np.ascontiguousarray(q).tobytes()
    eq.set_state(case.freestream_q())
    eq.read_restart(str(tmp_path / "a.out"))
    assert np.array_equal(eq.get_state(), q)
    eq.write_npz(str(tmp_path / "a.npz"))
    z = np.load(tmp_path / "a.npz")
    rho, u, v, p, T = sio.primitives(case, q)
    assert np.array_equal(z["q"], q) and np.array_equal(z["p"], p) and np.array_equal(z["T"], T)
    assert z["xc"].shape == (case.nic, case.njc)
    eq.close()


def test_grid_from_binary_file_equals_host_arrays(tmp_path):
    """sgpu_set_grid_file (N3): vertex rows straight from the binary file into the device planes -- whole grid and a slab
    that reads only its window -- must give the metrics and the residual of sgpu_set_grid bit for bit"""
    from structured_b200.api import SgpuError
    from structured_b200.cases import write_grid_bin
    case = turbulent_channel_case(150, 64, ntrans=1, reynolds=1e5)
    path = str(tmp_path / "grid.bin")
    write_grid_bin(path, case.xv, case.yv)
    q = case.perturbed_q()
    ref = gpu_eq(case)
    want_m, want_r = ref.metrics(), ref.calc_residual(q)
    ref.close()
    for (j0, j1) in ((0, 0), (0, 30), (30, 64)):
        eq = gpu_eq(case, j_begin=j0, j_end=j1)
        eq.set_grid_file(path)
        a, b = (0, case.njc) if j1 == 0 else (j0, j1)
        got = eq.metrics()
        assert np.array_equal(got[0][:, a:b], want_m[0][:, a:b]) and np.array_equal(got[1][:, a:b + 1], want_m[1][:, a:b + 1])
        assert np.array_equal(got[2][:, a:b], want_m[2][:, a:b])
        eq.set_state(q)
        eq.residual_device(0)
        out = np.zeros_like(want_r); eq.get_rhs(out=out)
        assert np.array_equal(out[:, a:b], want_r[:, a:b])
        eq.close()
    eq = gpu_eq(case)
    (tmp_path / "bad.bin").write_bytes(b"garbage!" + bytes(64))
    with pytest.raises(SgpuError, match="file format not found"):
        eq.set_grid_file(str(tmp_path / "bad.bin"))
    eq.close()


def test_halo_exchange_refuses_a_missing_peer_instead_of_hanging():
    """advisor finding: a slab with a neighbour but no registered peer buffer used to leave the neighbour's wait kernel
    spinning forever; push / pull now fail loudly (and the device-side wait is bounded)"""
    from structured_b200.api import SgpuError
    case = turbulent_channel_case(64, 32, ntrans=0, reynolds=1e5)
    lo = gpu_eq(case, j_begin=0, j_end=16)
    lo.set_state(case.perturbed_q())
    with pytest.raises(SgpuError, match="no peer receive buffer"):
        lo.halo_push(0)
    with pytest.raises(SgpuError, match="before the first halo push"):
        lo.halo_pull(0)
    lo.close()


@pytest.mark.parametrize("ntrans", [0, 1])
@pytest.mark.parametrize("which", ["channel", "zooA", "zooD", "zooE"])
def test_fused_runge_kutta_stages_equal_the_two_kernel_form_bitwise(which, ntrans, monkeypatch):
    """sgpu_explicit_step with the stage update q_tmp' = q + rhs*dt/(4 - order) (update_rk4, src/solver/solver.cpp:4-13) written
    from the residual kernel's epilogue (ping-pong q_tmp buffers, division by the stage constant as a correctly rounded fma
    sequence) against the residual + axpy_dt_div_kernel form: states, rhs and norms bit for bit over several steps"""
    from structured_b200.cases import zoo_case
    case = turbulent_channel_case(190, 70, ntrans=ntrans, reynolds=1e5) if which == "channel" else zoo_case(which[-1], 70, 33, ntrans=ntrans)
    q0 = case.perturbed_q(0.01 if which == "channel" else 0.002)
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("SGPU_RK_FUSED", fused)
        eq = gpu_eq(case)
        eq.set_state(q0, 0); eq.set_state(q0, 1)
        l2 = [eq.explicit_step(0.5 if which == "channel" else 0.02, "rk4_jameson") for _ in range(3)]
        out[fused] = (eq.get_state(0), eq.get_state(1), eq.get_rhs(), np.array(l2))
        eq.close()
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b)
    assert np.isfinite(out["1"][0]).all() and not np.array_equal(out["1"][0], q0)
