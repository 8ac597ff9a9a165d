"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference and the oracle/_ref build):
    make -C oracle ref refbin && python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own, so these are outputs of the
unmodified reference sources: the unity-TU library (oracle/ref/ref_unity.cpp) for residual / dt /
Jacobian at a fixed perturbed state, and the stock `structured_explicit` binary for short explicit runs.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from structured_b200.cases import ZOO, load_case, write_case, zoo_case  # noqa: E402
from oracle.bindings import REF_BIN, RefOracle  # noqa: E402

REF = "/root/reference"


def jac_sample_rows(case, n=400, seed=1234):
    """rows kept for the big NACA Jacobian: all equations of cells along the boundaries' first two layers at a
    few stations (wake cut, wall, far field corners) plus a seeded random set"""
    rng = np.random.default_rng(seed)
    cells = set()
    for i in (0, 1, 5, 39, 40, 41, 42, 139, 140, 238, 239, 240, 241, 278, 279):
        for j in (0, 1, 2, case.njc - 2, case.njc - 1):
            if i < case.nic:
                cells.add((i, j))
    while len(cells) < n:
        cells.add((int(rng.integers(case.nic)), int(rng.integers(case.njc))))
    rows = sorted((i * case.njc + j) * case.nv + k for (i, j) in cells for k in range(case.nv))
    return np.array(rows, dtype=np.int64)


def run_stock_binary(inp_text, grid_src, grid_name, label, workdir):
    """stock reference binary on a config; returns (q from <label>.npz, history text)"""
    shutil.copy(grid_src, os.path.join(workdir, grid_name))
    with open(os.path.join(workdir, "run.inp"), "w") as f:
        f.write(inp_text)
    subprocess.run([REF_BIN, "-c", "run.inp"], cwd=workdir, check=True, stdout=subprocess.DEVNULL)
    z = np.load(os.path.join(workdir, label + ".npz"))
    hist = open(os.path.join(workdir, label + ".history")).read()
    return z["q"].copy(), hist


def fixture_from_config(name, cfg, cfl_dt, jac="full", explicit=None):
    text = open(cfg).read()
    case = load_case(cfg)
    ref = RefOracle(config_path=cfg)
    q = case.perturbed_q()
    out = dict(inp=np.array(text), xv=case.xv, yv=case.yv, q0=ref.initial_q(),
               rhs=ref.residual(q, False), rhs_lhs=ref.residual(q, True), dt=ref.calc_dt(q, cfl_dt)[..., 0], cfl_dt=cfl_dt)
    ri, ci, va = ref.jacobian(q, True)
    out["jac_nnz"] = len(va); out["jac_fro"] = np.sqrt((va * va).sum()); out["jac_trace"] = va[ri == ci].sum()
    if jac == "full":
        out.update(jac_rind=ri, jac_cind=ci, jac_values=va)
    else:
        rows = jac_sample_rows(case)
        keep = np.isin(ri, rows)
        out.update(jac_rows=rows, jac_rind=ri[keep], jac_cind=ci[keep], jac_values=va[keep])
    if explicit is not None:
        with tempfile.TemporaryDirectory() as wd:
            qn, hist = run_stock_binary(explicit["inp"], explicit["grid_src"], explicit["grid_name"], explicit["label"], wd)
        out.update(explicit_inp=np.array(explicit["inp"]), explicit_q=qn, explicit_history=np.array(hist), explicit_steps=explicit["steps"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "nnz", len(va), "steps", None if explicit is None else explicit["steps"])
    ref.close()


def surface_fixture():
    """IOManager::write_surface of the reference on the NACA case (src/utils/io.cpp:182-255): gradients from the state the
    last residual saw (amp 0.01), pressure from a different final state (amp 0.012) -- the two states the routine mixes."""
    na = os.path.join(REF, "examples/laminar/naca0012")
    with tempfile.TemporaryDirectory() as wd:
        for f in os.listdir(na):
            shutil.copy(os.path.join(na, f), wd)
        cfg = os.path.join(wd, "config.inp")
        case = load_case(cfg)
        ref = RefOracle(config_path=cfg)
        rows, wall = ref.surface(case.perturbed_q(0.01), case.perturbed_q(0.012))
        ref.close()
    np.savez_compressed(os.path.join(HERE, "naca0012_surface.npz"), rows=rows, wall=wall, amp_res=0.01, amp_fin=0.012)
    print("surface", rows.shape)


def explicit_surface_fixture():
    """`<label>.surface` of the stock binary after the short explicit runs stored in channel.npz / naca0012.npz (same inputs):
    the text IOManager::write_surface produced (src/utils/io.cpp:182-255) -- final pressure, gradients of the last RK stage."""
    out = {}
    for name in ("channel", "naca0012"):
        z = np.load(os.path.join(HERE, name + ".npz"))
        inp = str(z["explicit_inp"])
        import tomllib
        t = tomllib.loads(inp)
        src = os.path.join(REF, "examples/laminar", "channel" if name == "channel" else "naca0012", os.path.basename(t["geometry"]["filename"]))
        with tempfile.TemporaryDirectory() as wd:
            shutil.copy(src, wd)
            with open(os.path.join(wd, "run.inp"), "w") as f:
                f.write(inp)
            subprocess.run([REF_BIN, "-c", "run.inp"], cwd=wd, check=True, stdout=subprocess.DEVNULL)
            q = np.load(os.path.join(wd, t["io"]["label"] + ".npz"))["q"]
            assert np.array_equal(q, z["explicit_q"])
            out[name] = np.array(open(os.path.join(wd, t["io"]["label"] + ".surface")).read())
    np.savez_compressed(os.path.join(HERE, "explicit_surface.npz"), **out)
    print("explicit surface", {k: len(str(v).splitlines()) for k, v in out.items()})


def main():
    if not os.path.isdir(REF):
        sys.exit("needs the reference tree at /root/reference")
    ch = os.path.join(REF, "examples/laminar/channel")
    # Solver::solve runs steps counter = 0 .. iteration_max+1 (the stop test `counter > iteration_max`
    # comes after the update, src/solver/solver.cpp:103-143)
    n_it = 18
    ex = open(os.path.join(ch, "explicit.inp")).read().replace("iteration_max = 1000000", "iteration_max = %d" % n_it)
    fixture_from_config("channel", os.path.join(ch, "implicit.inp"), 0.5, "full",
                        dict(inp=ex, grid_src=os.path.join(ch, "channel_3x101.unf2"), grid_name="channel_3x101.unf2", label="explicit", steps=n_it + 2))
    na = os.path.join(REF, "examples/laminar/naca0012")
    n_it = 8
    ex = open(os.path.join(na, "config.inp")).read().replace("cfl = 2.0", "cfl = 0.5").replace("cfl_ramp = true", "cfl_ramp = false") \
        .replace("iteration_max = 20000", "iteration_max = %d" % n_it).replace("stdout_frequency = 1", "stdout_frequency = 1000") \
        .replace("fileout_frequency = 1", "fileout_frequency = 1000")
    fixture_from_config("naca0012", os.path.join(na, "config.inp"), 0.5, "sample",
                        dict(inp=ex, grid_src=os.path.join(na, "grid.unf2"), grid_name="grid.unf2", label="implicit", steps=n_it + 2))
    surface_fixture()
    explicit_surface_fixture()
    for z in ZOO:
        case = zoo_case(z, 14, 10)
        with tempfile.TemporaryDirectory() as wd:
            cfg = write_case(case, wd, "zoo")
            ref = RefOracle(config_path=cfg)
            q = case.perturbed_q(0.02)
            ri, ci, va = ref.jacobian(q, True)
            np.savez_compressed(os.path.join(HERE, "zoo_%s.npz" % z), inp=np.array(open(cfg).read()), xv=case.xv, yv=case.yv, q=q,
                                rhs=ref.residual(q, False), rhs_lhs=ref.residual(q, True), dt=ref.calc_dt(q, 0.7)[..., 0], cfl_dt=0.7,
                                jac_rind=ri, jac_cind=ci, jac_values=va)
            print("zoo", z, "nnz", len(va))
            ref.close()


if __name__ == "__main__":
    main()
