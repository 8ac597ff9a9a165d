"""CPU suite: host-side logic -- the reference's file formats and the synthetic workload generators."""
import os

import numpy as np

from helpers import golden
from structured_b200.cases import (Boundary, Case, bump_channel_grid, case_from_toml, load_case, read_grid,
                                   turbulent_channel_case, write_case, write_grid_p3d, zoo_case)


def test_write_then_load_case_roundtrip(tmp_path):
    c = zoo_case("A")
    path = write_case(c, str(tmp_path), "rt")
    d = load_case(path)
    assert (d.ni, d.nj, d.order, d.lhs_order, d.flux) == (c.ni, c.nj, c.order, c.lhs_order, c.flux)
    assert np.array_equal(d.xv, c.xv) and np.array_equal(d.yv, c.yv)       # %.17e round-trips doubles
    assert [(b.type, b.face, b.start, b.end) for b in d.boundaries] == [(b.type, b.face, b.start, b.end) for b in c.boundaries]
    assert d.mu_inf == c.mu_inf and d.u_inf == c.u_inf


def test_grid_formats(tmp_path):
    x, y = bump_channel_grid(5, 4)
    write_grid_p3d(str(tmp_path / "g.p3d"), x, y)
    x2, y2 = read_grid(str(tmp_path / "g.p3d"), 6, 5, "p3d")
    assert np.array_equal(x, x2) and np.array_equal(y, y2)
    with open(tmp_path / "g.simple", "w") as f:                            # "simple": x y per line, j outer
        for j in range(5):
            for i in range(6):
                f.write("%.17e %.17e\n" % (x[i, j], y[i, j]))
    x3, y3 = read_grid(str(tmp_path / "g.simple"), 6, 5, "simple")
    assert np.array_equal(x, x3) and np.array_equal(y, y3)


def test_golden_configs_parse_like_the_reference():
    case, z = golden("naca0012")
    assert (case.ni, case.nj, case.order, case.lhs_order, case.flux) == (281, 151, 2, 1, "roe")
    assert [b.type for b in case.boundaries] == ["freestream", "freestream", "freestream", "wake", "wall"]
    assert np.array_equal(case.freestream_q(), z["q0"])
    case, z = golden("channel")
    assert case.dpdx == -0.02592 and case.boundaries[2].type == "periodic" and case.boundaries[0].end == -2


def test_synthetic_workload():
    c = turbulent_channel_case(32, 16)
    assert c.nv == 5 and c.viscous and c.wall_distance.shape == (32, 16) and (c.wall_distance > 0).all()
    q = c.perturbed_q()
    assert q.shape == (32, 16, 5) and (q[..., 0] > 0).all() and (q[..., 4] > 0).all()
