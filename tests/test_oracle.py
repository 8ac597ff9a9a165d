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
