"""GPU parity tests of the Jacobian path: device block-stencil Jacobian -> COO export (the sparse_jac
replacement) against the reference-derived golden Jacobians and the CPU oracle."""
import numpy as np
import pytest

from helpers import TOL, coo_to_dict_arrays, golden, golden_state, jac_rel_err
from structured_b200.cases import ZOO, turbulent_channel_case, zoo_case

pytestmark = pytest.mark.gpu


def gpu_eq(case, **kw):
    from structured_b200.api import GpuEulerEquation
    return GpuEulerEquation(case, **kw)


def check_pattern(n, ours, ref):
    """every structural entry of the reference pattern must be present in ours (ours may carry a few explicit
    zeros more); returns (nnz_ours, nnz_ref)"""
    ko, _ = coo_to_dict_arrays(n, *ours)
    kr, _ = coo_to_dict_arrays(n, *ref)
    missing = np.setdiff1d(kr, ko)
    assert len(missing) == 0, "entries of the reference pattern missing: %d, first (row, col) = %s" % (len(missing), divmod(int(missing[0]), n))
    assert len(ko) == len(ours[2]), "COO export must not contain duplicate (row, col)"
    assert np.all(np.diff(ours[0].astype(np.int64) * n + ours[1].astype(np.int64)) > 0), "COO export must be sorted by (row, col)"
    return len(ko), len(kr)


@pytest.mark.parametrize("name", ["channel"] + ["zoo_" + z for z in ZOO])
def test_jacobian_matches_reference_golden(name):
    case, z = golden(name)
    eq = gpu_eq(case)
    q = golden_state(case, z)
    eq.set_state(q)
    ours = eq.jacobian_coo()
    ref = (z["jac_rind"], z["jac_cind"], z["jac_values"])
    n = q.size
    err = jac_rel_err(n, ours, ref)
    assert err <= TOL, (name, err)
    no, nr = check_pattern(n, ours, ref)
    assert no <= 1.10 * nr, (no, nr)
    eq.close()


def test_jacobian_matches_reference_golden_naca_sample():
    case, z = golden("naca0012")
    eq = gpu_eq(case)
    q = case.perturbed_q()
    eq.set_state(q)
    ri, ci, va = eq.jacobian_coo()
    n = q.size
    keep = np.isin(ri, z["jac_rows"])
    ours = (ri[keep], ci[keep], va[keep])
    ref = (z["jac_rind"], z["jac_cind"], z["jac_values"])
    assert jac_rel_err(n, ours, ref) <= TOL
    check_pattern(n, ours, ref)
    assert abs(np.sqrt((va * va).sum()) - float(z["jac_fro"])) <= 1e-12 * float(z["jac_fro"])
    assert abs(va[ri == ci].sum() - float(z["jac_trace"])) <= 1e-12 * abs(float(z["jac_trace"]))
    eq.close()


@pytest.mark.parametrize("nic,njc,order,lhs_order,flux,periodic", [(33, 21, 2, 2, "roe", True), (40, 18, 2, 1, "roe", False), (26, 30, 1, 1, "ausm", True), (21, 17, 2, 2, "ausm", False)])
def test_sa_jacobian_matches_oracle(nic, njc, order, lhs_order, flux, periodic):
    from oracle.bindings import PortOracle
    case = turbulent_channel_case(nic, njc, ntrans=1, order=order, lhs_order=lhs_order, flux=flux, reynolds=2e4, periodic=periodic)
    port = PortOracle(case); eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    eq.set_state(q)
    ours = eq.jacobian_coo(); ref = port.jacobian(q, True)
    err = jac_rel_err(q.size, ours, ref)
    assert err <= TOL, err
    check_pattern(q.size, ours, ref)
    eq.close(); port.close()


def test_lhs_transform_and_ownership():
    """values = -values, diagonal += 1/dt (src/solver/solver.cpp:162-171)"""
    case, z = golden("channel")
    eq = gpu_eq(case)
    q = golden_state(case, z)
    eq.set_state(q)
    eq.calc_dt(3.0)
    ri, ci, va = eq.jacobian_coo()
    ri2, ci2, va2 = eq.jacobian_coo(apply_lhs_transform=True)
    assert np.array_equal(ri, ri2) and np.array_equal(ci, ci2)
    dt = eq.get_dt().reshape(-1)
    want = -va
    d = ri == ci
    want[d] += 1.0 / dt[ri[d]]
    assert np.abs(va2 - want).max() <= 1e-13 * np.abs(want).max()
    eq.close()


@pytest.mark.parametrize("ntrans", [0, 1])
def test_device_jacobian_products_and_adjoint_identity(ntrans):
    """y = J x agrees with the COO matrix and with a directional derivative of the residual; (J^T psi).v = psi.(J v)"""
    case = turbulent_channel_case(48, 40, ntrans=ntrans, reynolds=2e4)
    eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    eq.set_state(q)
    ri, ci, va = eq.jacobian_coo()
    slots, ms = eq.jacobian_device()
    assert slots == 13 and ms > 0
    rng = np.random.default_rng(3)
    v = rng.standard_normal(q.shape) * np.abs(q).mean(axis=(0, 1))
    psi = rng.standard_normal(q.shape)
    Jv = eq.jacobian_apply(v)
    want = np.zeros(q.size); np.add.at(want, ri, va * v.reshape(-1)[ci])
    assert np.abs(Jv.reshape(-1) - want).max() <= 1e-12 * np.abs(want).max()
    h = 1e-6
    fd = (eq.calc_residual(q + h * v) - eq.calc_residual(q - h * v)) / (2 * h)
    assert np.abs(Jv - fd).max() <= 2e-6 * np.abs(fd).max()
    JTpsi = eq.jacobian_apply(psi, transpose=True)
    lhs, rhs = float((JTpsi * v).sum()), float((psi * Jv).sum())
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs))
    eq.close()


@pytest.mark.parametrize("ntrans,split", [(0, 9), (1, 14)])
def test_slab_jacobians_concatenate_to_the_global_jacobian(ntrans, split):
    """each j-slab builds its own rows with GLOBAL row / column indices; no exchange beyond the q ghost rows"""
    case = turbulent_channel_case(37, 30, ntrans=ntrans, reynolds=2e4, periodic=(ntrans == 0))
    q = case.perturbed_q(0.02)
    one = gpu_eq(case); one.set_state(q)
    want = one.jacobian_coo()
    parts = []
    for (j0, j1) in ((0, split), (split, case.njc)):
        eq = gpu_eq(case, j_begin=j0, j_end=j1)
        eq.set_state(q)
        parts.append(eq.jacobian_coo())
        eq.close()
    got = tuple(np.concatenate([p[k] for p in parts]) for k in range(3))
    order = np.lexsort((got[1], got[0]))
    got = tuple(g[order] for g in got)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(got[2], want[2])
    one.close()


def test_dres_dbeta_matches_oracle_difference():
    """the SA source is linear in beta: dR4/dbeta = R4(beta+1) - R4(beta), exact up to round-off (oracle statement)"""
    from oracle.bindings import PortOracle
    case = turbulent_channel_case(40, 28, ntrans=1, reynolds=2e4)
    q = case.perturbed_q(0.02)
    eq = gpu_eq(case); eq.set_state(q)
    got = eq.dres_dbeta()
    r0 = PortOracle(case).residual(q)
    case2 = turbulent_channel_case(40, 28, ntrans=1, reynolds=2e4); case2.beta = case.beta + 1.0
    r1 = PortOracle(case2).residual(q)
    want = r1[..., 4] - r0[..., 4]
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert np.abs(r1[..., :4] - r0[..., :4]).max() == 0.0
    eq.close()


def test_field_inversion_size_2048x1024_properties():
    """BASELINE.json config 4 (2048 x 1024, SA + beta field): device Jacobian, J v against a central difference of the
    residual, the adjoint identity psi.(J v) = (J^T psi).v and dR/dbeta against the (linear) beta dependence."""
    case = turbulent_channel_case(2048, 1024, ntrans=1)
    eq = gpu_eq(case)
    q = case.perturbed_q()
    eq.set_state(q)
    slots, ms = eq.jacobian_device()
    assert slots == 13 and ms > 0
    # sampled rows against the oracle evaluated on crops of the grid (bottom wall, top wall, interior; the i-periodic
    # seam is covered at small sizes): > 2000 row cells, 1e-12 with the Appendix-B rule
    from helpers import TOL_SA_COUPLING, jac_rel_err_split, jac_worst, oracle_on_crop, rows_of_cells
    ncells = 0
    for box in ((100, 126, 0, 24), (1000, 1026, 0, 24), (1900, 1926, 1000, 1024), (40, 66, 1000, 1024), (700, 726, 500, 526), (1500, 1526, 40, 66)):
        _, (i0, i1, j0, j1), ref = oracle_on_crop(case, q, box)
        ri, ci, va = eq.jacobian_coo(rows=(j0, j1 - j0))
        cells = [(i, j) for i in range(i0, i1) for j in range(j0, j1)]
        keep = np.isin(ri, rows_of_cells(case.njc, 5, cells))
        ours = (ri[keep], ci[keep], va[keep])
        err, err_cpl = jac_rel_err_split(q.size, ours, ref, 5)
        assert err <= TOL and err_cpl <= TOL_SA_COUPLING, (box, err, err_cpl, jac_worst(q.size, ours, ref, 5, case.njc))
        check_pattern(q.size, ours, ref)
        ncells += len(cells)
    assert ncells >= 2000
    rng = np.random.default_rng(5)
    v = rng.standard_normal(q.shape) * np.abs(q).mean(axis=(0, 1))
    psi = rng.standard_normal(q.shape)
    Jv = eq.jacobian_apply(v)
    JTpsi = eq.jacobian_apply(psi, transpose=True)
    a, b = float((JTpsi * v).sum()), float((psi * Jv).sum())
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
    h = 1e-7
    fd = (eq.calc_residual(q + h * v) - eq.calc_residual(q - h * v)) / (2 * h)
    # AD differentiates the taken branch of the non-smooth pieces (|.| in the Roe flux, the upwind switch of the SA
    # flux where the wall-normal mass flux changes sign, min/max in the SA source); a central difference straddles
    # them in a few wall cells, so the comparison is in the L2 norm plus a bound on the share of outliers
    for k in range(5):
        d = Jv[..., k] - fd[..., k]
        assert np.sqrt((d * d).sum()) <= 1e-3 * np.sqrt((fd[..., k] ** 2).sum()), k
        assert (np.abs(d) > 1e-5 * np.abs(fd[..., k]).max()).mean() <= 1e-3, k
    eq.set_state(q)
    dRdb = eq.dres_dbeta()
    r0 = eq.calc_residual(q)
    eq.set_field("beta", case.beta + 1.0)
    r1 = eq.calc_residual(q)
    assert np.abs(dRdb - (r1[..., 4] - r0[..., 4])).max() <= 1e-10 * np.abs(dRdb).max()
    eq.close()


def test_row_window_export_equals_rows_of_the_full_export():
    """sgpu_jacobian_coo_rows: the COO rows of a window of cell rows from the resident Jacobian = the same rows of
    sgpu_jacobian_coo (global numbering, sorted, LHS transform included)"""
    case = turbulent_channel_case(45, 28, ntrans=1, reynolds=2e4)
    eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    eq.set_state(q); eq.calc_dt(4.0)
    for lhs in (False, True):
        ri, ci, va = eq.jacobian_coo(apply_lhs_transform=lhs)
        for (j0, n) in ((0, 3), (11, 9), (25, 3), (0, 28)):
            r2, c2, v2 = eq.jacobian_coo(apply_lhs_transform=lhs, rows=(j0, n))
            cell_j = (ri // 5) % case.njc
            keep = (cell_j >= j0) & (cell_j < j0 + n)
            assert np.array_equal(r2, ri[keep]) and np.array_equal(c2, ci[keep]) and np.array_equal(v2, va[keep])
    from structured_b200.api import SgpuError
    with pytest.raises(SgpuError, match="do not intersect"):
        eq.jacobian_coo(rows=(40, 5))
    eq.close()
