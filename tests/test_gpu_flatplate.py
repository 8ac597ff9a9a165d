"""BASELINE.json config 2 -- the SA flat plate (slipwall -> wall junction, freestream / outflow), the SA x BC matrix (zoo
cases with ntrans = 1) and the device wall distance -- against the CPU oracle.  The SA extension has no reference code
("parity unpinned"): the oracle is our own CPU statement of the spec."""
import numpy as np
import pytest

from helpers import TOL, TOL_SA_COUPLING, field_rel_err, jac_rel_err, jac_rel_err_split, jac_worst, oracle_on_crop, rows_of_cells
from structured_b200.cases import ZOO_SA, flat_plate_case, zoo_case
from test_gpu_jacobian import check_pattern

pytestmark = pytest.mark.gpu


def gpu_eq(case, **kw):
    from structured_b200.api import GpuEulerEquation
    return GpuEulerEquation(case, **kw)


@pytest.mark.parametrize("make", [lambda: flat_plate_case(200, 120), lambda: zoo_case("A", 70, 33, ntrans=1), lambda: zoo_case("B", 31, 50, ntrans=1),
                                  lambda: zoo_case("D", 45, 20, ntrans=1), lambda: zoo_case("E", 64, 24, ntrans=1)])
def test_device_wall_distance_matches_oracle(make):
    """sgpu_wall_distance_from_bcs: nearest wall edge over `wall` / `isothermalwall` tables on any face"""
    from oracle.bindings import PortOracle
    case = make()
    eq = gpu_eq(case)
    got = eq.get_field("wall_distance")
    want = PortOracle(case).wall_distance()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert (np.abs(got - want) <= 1e-10 * want).all()          # cell-wise too (nearest cells sit half a cell off the wall)
    eq.close()


def test_wall_distance_without_walls_and_explicit_segments():
    from structured_b200.cases import Boundary
    case = zoo_case("A", 30, 20, ntrans=1)
    case.boundaries = [Boundary("freestream", f, 0, -1) for f in ("bottom", "top", "left", "right")]
    eq = gpu_eq(case)
    assert (eq.get_field("wall_distance") == 1e30).all()       # no wall: destruction term vanishes
    eq.compute_wall_distance(np.array([[0.0, -1.0, 1.0, -1.0]]))
    xc = 0.25 * (case.xv[:-1, :-1] + case.xv[1:, :-1] + case.xv[:-1, 1:] + case.xv[1:, 1:])
    yc = 0.25 * (case.yv[:-1, :-1] + case.yv[1:, :-1] + case.yv[:-1, 1:] + case.yv[1:, 1:])
    want = np.hypot(xc - np.clip(xc, 0.0, 1.0), yc + 1.0)
    assert np.abs(eq.get_field("wall_distance") - want).max() <= 1e-14
    eq.close()


@pytest.mark.parametrize("name", ZOO_SA)
def test_sa_bc_matrix_residual_and_jacobian_match_oracle(name):
    """nu~ ghost rules of every BC type / face: slipwall (A), side walls + periodic bottom/top (B), isothermal walls +
    periodic left/right (D), wake + wall segment + freestream on three faces (E)"""
    from oracle.bindings import PortOracle
    case = zoo_case(name, 37, 22, ntrans=1)
    port = PortOracle(case); eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    for lhs in (False, True):
        err = field_rel_err(eq.calc_residual(q, lhs=lhs), port.residual(q, lhs))
        assert err.max() <= TOL, (name, lhs, err)
    eq.set_state(q)
    ours, ref = eq.jacobian_coo(), port.jacobian(q, True)
    err = jac_rel_err(q.size, ours, ref)
    assert err <= TOL, (name, err)
    check_pattern(q.size, ours, ref)
    eq.close(); port.close()


@pytest.mark.parametrize("nic,njc,order,lhs_order,flux", [(64, 48, 2, 2, "roe"), (125, 37, 2, 1, "roe"), (40, 60, 1, 1, "ausm")])
def test_flat_plate_matches_oracle(nic, njc, order, lhs_order, flux):
    from oracle.bindings import PortOracle
    case = flat_plate_case(nic, njc, order=order, lhs_order=lhs_order, flux=flux, reynolds=1e5)
    port = PortOracle(case); eq = gpu_eq(case)
    q = case.perturbed_q(0.02)
    for lhs in (False, True):
        err = field_rel_err(eq.calc_residual(q, lhs=lhs), port.residual(q, lhs))
        assert err.max() <= TOL, (lhs, err)
    eq.set_state(q)
    ours, ref = eq.jacobian_coo(), port.jacobian(q, True)
    assert jac_rel_err(q.size, ours, ref) <= TOL
    check_pattern(q.size, ours, ref)
    eq.close(); port.close()


def test_flat_plate_1024_residual_all_cells_and_sampled_jacobian_rows():
    """BASELINE.json config 2 at its named size (~1 M cells, SA): the residual of EVERY cell against the oracle, and the
    Jacobian rows of > 2000 cells -- windows on every boundary, the four corners, the slipwall -> wall junction and the
    interior -- against the oracle evaluated on crops of the grid (tests/helpers.py::crop_case reproduces the full-grid
    oracle bit for bit on the cells it reports; checked on the CPU in tests/test_oracle.py)."""
    from oracle.bindings import PortOracle
    n = 1024
    case = flat_plate_case(n, n)
    eq = gpu_eq(case)
    wd = eq.get_field("wall_distance")                        # device wall distance -> the oracle gets the same field
    case_o = flat_plate_case(n, n); case_o.wall_distance = wd
    port = PortOracle(case_o)
    q = case.perturbed_q()
    for lhs in (False, True):
        err = field_rel_err(eq.calc_residual(q, lhs=lhs), port.residual(q, lhs))
        assert err.max() <= TOL, (lhs, err)
    port.close()
    eq.set_state(q)
    slots, ms = eq.jacobian_device()
    assert slots == 13
    ile = int(round(0.2 * n))                                 # first wall cell column
    boxes = [(0, 24, 0, 24), (ile - 12, ile + 12, 0, 24), (500, 524, 0, 24), (n - 24, n, 0, 24),           # bottom: corner, junction, plate, outflow corner
             (0, 24, n - 24, n), (500, 524, n - 24, n), (n - 24, n, n - 24, n),                         # top
             (0, 24, 500, 524), (n - 24, n, 500, 524), (300, 324, 300, 324), (700, 724, 60, 84)]           # sides, interior
    ncells = 0
    for box in boxes:
        _, (i0, i1, j0, j1), ref = oracle_on_crop(case_o, q, box)
        ri, ci, va = eq.jacobian_coo(rows=(j0, j1 - j0))
        cells = [(i, j) for i in range(i0, i1) for j in range(j0, j1)]
        keep = np.isin(ri, rows_of_cells(n, 5, cells))
        ours = (ri[keep], ci[keep], va[keep])
        err, err_cpl = jac_rel_err_split(q.size, ours, ref, 5)
        assert err <= TOL and err_cpl <= TOL_SA_COUPLING, (box, err, err_cpl, jac_worst(q.size, ours, ref, 5, case.njc))
        check_pattern(q.size, ours, ref)
        ncells += len(cells)
    assert ncells >= 2000
    eq.close()
