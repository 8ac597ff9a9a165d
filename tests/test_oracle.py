"""CPU suite: the oracle against the reference's own outputs (golden fixtures) and, where oracle/_ref
exists (build container), against the reference compiled as a library."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, TOL, field_rel_err, golden, golden_state, jac_rel_err
from oracle.bindings import PortOracle, RefOracle, have_ref, rk4_step_cpu
from structured_b200.cases import ZOO, zoo_case

CASES = ["channel", "naca0012"] + ["zoo_" + z for z in ZOO]


@pytest.mark.parametrize("name", CASES)
def test_port_matches_golden_residual_and_dt(name):
    case, z = golden(name)
    port = PortOracle(case)
    q = golden_state(case, z)
    # the port is a restatement of the same operation sequence: it reproduces the reference bit for bit
    assert np.array_equal(port.residual(q, False), z["rhs"])
    assert np.array_equal(port.residual(q, True), z["rhs_lhs"])
    assert np.array_equal(port.calc_dt(q, float(z["cfl_dt"]))[..., 0], z["dt"])
    port.close()


@pytest.mark.parametrize("name", ["channel"] + ["zoo_" + z for z in ZOO])
def test_port_matches_golden_jacobian(name):
    case, z = golden(name)
    port = PortOracle(case)
    q = golden_state(case, z)
    ri, ci, va = port.jacobian(q, True)
    assert len(va) == len(z["jac_values"])          # same STRUCTURAL pattern as the reference-derived one
    assert np.array_equal(ri, z["jac_rind"]) and np.array_equal(ci, z["jac_cind"])
    assert np.array_equal(va, z["jac_values"])
    port.close()


def test_port_matches_golden_jacobian_naca_sample():
    case, z = golden("naca0012")
    port = PortOracle(case)
    ri, ci, va = port.jacobian(case.perturbed_q(), True)
    assert len(va) == int(z["jac_nnz"])
    assert abs(np.sqrt((va * va).sum()) - float(z["jac_fro"])) <= 1e-13 * float(z["jac_fro"])
    keep = np.isin(ri, z["jac_rows"])
    assert np.array_equal(ri[keep], z["jac_rind"]) and np.array_equal(ci[keep], z["jac_cind"])
    assert np.array_equal(va[keep], z["jac_values"])
    port.close()


@pytest.mark.parametrize("name", ["channel", "naca0012"])
def test_port_explicit_run_matches_stock_binary(name):
    """rk4_jameson steps with the port == the stock reference binary's .npz after the same number of steps"""
    case, z = golden(name)
    from structured_b200.cases import case_from_toml
    ex = case_from_toml(str(z["explicit_inp"]), z["xv"], z["yv"])
    port = PortOracle(ex)
    q = ex.freestream_q(); q_tmp = q.copy()
    assert np.array_equal(q, z["q0"])
    for _ in range(int(z["explicit_steps"])):
        q, q_tmp, rhs = rk4_step_cpu(port, q, q_tmp, ex.cfl)
    assert np.array_equal(q, z["explicit_q"])
    # last history line = L2 norms of the last rhs (src/solver/solver.cpp:125-141)
    last = str(z["explicit_history"]).strip().splitlines()[-1].split()
    l2 = [np.sqrt((rhs[..., k] ** 2).sum()) for k in range(4)]
    for k in range(4):
        assert "%.2e" % l2[k] == last[-4 + k].replace("info]", "").strip() or abs(float(last[-4 + k]) - l2[k]) <= 0.006 * l2[k]
    port.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("z", ZOO)
def test_port_matches_reference_library_on_zoo(z):
    case = zoo_case(z)                               # 24x16, larger than the golden zoo fixtures
    ref = RefOracle(case); port = PortOracle(case)
    q = case.perturbed_q(0.02)
    for lhs in (False, True):
        assert np.array_equal(port.residual(q, lhs), ref.residual(q, lhs))
    a = ref.jacobian(q, True); b = port.jacobian(q, True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ref.close(); port.close()


def test_sa_port_properties():
    """SA extension has no reference: pin what can be pinned -- with nu~ -> 0 the mean-flow rows reduce to the
    laminar residual, and the Jacobian matches central differences."""
    from structured_b200.cases import turbulent_channel_case
    case = turbulent_channel_case(12, 10, ntrans=1)
    lam = turbulent_channel_case(12, 10, ntrans=0)
    ps, pl = PortOracle(case), PortOracle(lam)
    q = case.perturbed_q(0.02)
    q0 = q.copy(); q0[..., 4] = 0.0
    r_sa = ps.residual(q0); r_lam = pl.residual(np.ascontiguousarray(q0[..., :4]))
    assert field_rel_err(r_sa[..., :4], r_lam).max() < 1e-13
    ri, ci, va = ps.jacobian(q, True)
    n = q.size
    rng = np.random.default_rng(7)
    v = rng.standard_normal(q.shape) * np.abs(q).mean(axis=(0, 1))
    Jv = np.zeros(n); np.add.at(Jv, ri, va * v.reshape(-1)[ci])
    h = 1e-6
    fd = (ps.residual(q + h * v) - ps.residual(q - h * v)).reshape(-1) / (2 * h)
    assert np.abs(Jv - fd).max() / np.abs(fd).max() < 1e-6
    ps.close(); pl.close()


def test_port_surface_matches_reference_golden():
    """IOManager::write_surface (src/utils/io.cpp:182-255): the arrays it reads are bit-identical to the reference's,
    the text rows agree to the 6 significant digits the reference prints."""
    case, _ = golden("naca0012")
    z = np.load(os.path.join(GOLDEN, "naca0012_surface.npz"))
    port = PortOracle(case)
    s = port.surface(case.perturbed_q(float(z["amp_res"])), case.perturbed_q(float(z["amp_fin"])))
    assert np.array_equal(s["wall"], z["wall"])
    rows = np.stack([s["xw"], s["cp"], s["cf"]], axis=1)
    assert rows.shape == z["rows"].shape == (case.ni - 2*case.tail + 1, 3)
    assert (np.abs(rows - z["rows"]) <= 5.1e-6*np.abs(rows) + 1e-300).all()
    # coefficient sums: independent numpy restatement of io.cpp:226-249
    gux, guy, gvx, gvy, p0, p1 = s["wall"][:, case.tail - 1:case.tail - 1 + len(rows)]
    qinf = 0.5*case.rho_inf*(case.u_inf**2 + case.v_inf**2)
    i0 = case.tail - 1
    dx = np.diff(case.xv[i0:i0 + len(rows) + 1, 0]); dy = np.diff(case.yv[i0:i0 + len(rows) + 1, 0])
    tau = case.mu_inf*(guy - gvx)/qinf; sf = 2.0/3.0*(gux + gvy)
    sxx = case.mu_inf*(2*gux - sf)/qinf; syy = case.mu_inf*(2*gvy - sf)/qinf
    fn_p, fc_p = -(s["cp"]*dx).sum(), (s["cp"]*dy).sum()
    fn_v, fc_v = (-tau*dy + syy*dx).sum(), (tau*dx - sxx*dy).sum()
    ca, sa = np.cos(case.aoa), np.sin(case.aoa)
    want = np.array([-fc_p*sa + fn_p*ca, fc_p*ca + fn_p*sa, -fc_v*sa + fn_v*ca, fc_v*ca + fn_v*sa])
    scale = np.abs(np.array([fn_p, fc_p, fn_v, fc_v])).max()
    assert np.abs(s["coeffs"][:4] - want).max() <= 1e-12*scale
    assert np.allclose(s["coeffs"][4:], [want[0] + want[2], want[1] + want[3]], rtol=1e-13)
    port.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_surface_matches_reference_library_channel():
    case, _ = golden("channel")
    ref = RefOracle(case); port = PortOracle(case)
    q = case.perturbed_q(0.02)
    rows, wall = ref.surface(q)
    s = port.surface(q)
    assert np.array_equal(s["wall"], wall)
    assert len(rows) == len(s["xw"])
    ref.close(); port.close()


def test_crop_reproduces_the_full_grid_oracle_bit_for_bit():
    """tests/helpers.py::crop_case is what lets the GPU tests compare Jacobian rows at BASELINE's sizes: a window of a
    large case (physical boundaries kept, cut edges frozen, the large grid's limiter constants) reproduces the
    full-grid oracle on every cell >= 3 cells away from a cut edge -- residual and Jacobian rows, exactly."""
    from helpers import oracle_on_crop, rows_of_cells
    from structured_b200.cases import flat_plate_case
    case = flat_plate_case(90, 70, reynolds=1e5)
    port = PortOracle(case)
    case.wall_distance = port.wall_distance()
    q = case.perturbed_q(0.02)
    res = port.residual(q, True)
    full = port.jacobian(q, True)
    ile = 18
    for box in ((0, 22, 0, 20), (ile - 10, ile + 12, 0, 18), (68, 90, 0, 20), (40, 62, 25, 47), (0, 20, 50, 70), (70, 90, 48, 70)):
        r, (i0, i1, j0, j1), coo = oracle_on_crop(case, q, box)
        assert np.array_equal(r, res[i0:i1, j0:j1])
        rows = rows_of_cells(case.njc, 5, [(i, j) for i in range(i0, i1) for j in range(j0, j1)])
        keep = np.isin(full[0], rows)
        assert jac_rel_err(q.size, coo, (full[0][keep], full[1][keep], full[2][keep])) == 0.0
        assert keep.sum() == len(coo[2])
    port.close()


def test_static_colouring_rows_equal_the_pattern_coloured_jacobian():
    """port_jacobian_rows (colour = (i mod 5, j mod 5, k), no pattern pass) against the sparse_jac stand-in"""
    from helpers import rows_of_cells
    from structured_b200.cases import zoo_case
    case = zoo_case("A", 33, 21, ntrans=1)
    port = PortOracle(case)
    q = case.perturbed_q(0.02)
    full = port.jacobian(q, True)
    cells = [(i, j) for i in range(0, 33, 4) for j in range(0, 21, 3)]
    rows = port.jacobian_rows(q, cells)
    keep = np.isin(full[0], rows_of_cells(case.njc, 5, cells)) & (full[2] != 0.0)
    assert jac_rel_err(q.size, rows, (full[0][keep], full[1][keep], full[2][keep])) == 0.0
    port.close()


def test_port_wall_distance_against_numpy():
    from structured_b200.cases import flat_plate_case, wall_segments, zoo_case, ZOO_SA
    for case in [flat_plate_case(60, 40)] + [zoo_case(z, ntrans=1) for z in ZOO_SA]:
        port = PortOracle(case)
        got = port.wall_distance()
        seg = wall_segments(case)
        xc = 0.25 * (case.xv[:-1, :-1] + case.xv[1:, :-1] + case.xv[:-1, 1:] + case.xv[1:, 1:])
        yc = 0.25 * (case.yv[:-1, :-1] + case.yv[1:, :-1] + case.yv[:-1, 1:] + case.yv[1:, 1:])
        a, ab = seg[:, :2], seg[:, 2:] - seg[:, :2]
        d = np.stack([xc, yc], axis=-1)[:, :, None, :] - a[None, None, :, :]
        t = np.clip((d * ab).sum(-1) / (ab * ab).sum(-1), 0.0, 1.0)
        want = np.sqrt((((d - t[..., None] * ab) ** 2).sum(-1)).min(axis=-1))
        assert np.abs(got - want).max() <= 1e-13 * want.max()
        assert got.min() > 0
        port.close()
