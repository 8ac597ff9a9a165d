"""GPU: the C++ drop-in (integration/structured_gpu_explicit = the reference's unmodified main.cpp / Config / Mesh /
IOManager + our Solver::step over the C ABI) must reproduce the stock reference binary's outputs."""
import os
import subprocess
import tomllib

import numpy as np
import pytest

from helpers import field_rel_err, golden
from structured_b200.cases import write_grid_p3d, write_grid_simple

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BIN = os.path.join(ROOT, "integration", "structured_gpu_explicit")


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/structured_gpu_explicit not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", ["channel", "naca0012"])
def test_dropin_binary_reproduces_stock_binary(name, tmp_path):
    case, z = golden(name)
    inp = str(z["explicit_inp"])
    t = tomllib.loads(inp)
    grid = t["geometry"]["filename"]
    (write_grid_p3d if t["geometry"]["format"] == "p3d" else write_grid_simple)(str(tmp_path / os.path.basename(grid)), z["xv"], z["yv"])
    (tmp_path / "run.inp").write_text(inp)
    res = subprocess.run([BIN, "-c", "run.inp"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    label = t["io"]["label"]
    out = np.load(tmp_path / (label + ".npz"))
    assert field_rel_err(out["q"], z["explicit_q"]).max() <= 1e-10
    # the reference's own output pipeline ran on our state: primitive fields written by IOManager agree too
    ours_hist = (tmp_path / (label + ".history")).read_text().strip().splitlines()[-1].split()
    ref_hist = str(z["explicit_history"]).strip().splitlines()[-1].split()
    assert ours_hist[0] == ref_hist[0]                       # same step counter
    for a, b in zip(ours_hist[-4:], ref_hist[-4:]):          # L2 norms printed with 3 significant digits
        assert abs(float(a) - float(b)) <= 0.011 * abs(float(b))
