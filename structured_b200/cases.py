"""Host-side case description: the reference's config/grid formats, unchanged.

A `Case` holds exactly what `Config` + `Mesh` + `BoundaryContainer` of the reference read from a
`.inp` TOML file and its grid file (src/utils/config.cpp:32-87, src/model/bc.cpp:465-526,
src/utils/mesh.cpp:134-169).  `load_case()` parses a stock reference config; `write_case()` emits one
(plus a p3d grid) that the unmodified reference can run, which is how the parity tests drive the
reference on synthetic grids.  The synthetic grids / states are the ones SURVEY.md section 8(d) names.
"""
from __future__ import annotations

import dataclasses
import math
import os
import tomllib
from typing import List, Optional

import numpy as np

GAMMA = 1.4  # src/common.h:40

BC_TYPES = {"freestream": 0, "slipwall": 1, "wall": 2, "isothermalwall": 3, "wake": 4, "outflow": 5, "periodic": 6}
BC_NAMES = {v: k for k, v in BC_TYPES.items()}
FACES = {"bottom": 0, "right": 1, "top": 2, "left": 3}  # src/model/bc.h:8-11
FACE_NAMES = {v: k for k, v in FACES.items()}
FLUXES = {"roe": 0, "ausm": 1}


@dataclasses.dataclass
class Boundary:
    type: str
    face: str
    start: int = 0
    end: int = 0
    u: float = 0.0
    v: float = 0.0
    T: float = 0.0


@dataclasses.dataclass
class Case:
    ni: int
    nj: int
    xv: Optional[np.ndarray] = None  # [ni][nj]
    yv: Optional[np.ndarray] = None
    tail: int = 1
    # [freestream]  (defaults of src/utils/config.cpp:33-40)
    rho_inf: float = 1.0
    u_inf: float = 0.0
    v_inf: float = 0.0
    p_inf: float = 1.0 / 1.4
    T_inf: float = 1.0 / 1.4
    mu_inf: float = 0.0
    pr_inf: float = 0.7
    aoa: float = 0.0
    # [solver]
    order: int = 1
    lhs_order: Optional[int] = None
    scheme: str = "forward_euler"
    flux: str = "ausm"  # the reference's default when solver.flux is missing (config.cpp:54)
    cfl: float = 1.0
    iteration_max: int = 1
    # [source]
    dpdx: float = 0.0
    dpdy: float = 0.0
    boundaries: List[Boundary] = dataclasses.field(default_factory=list)
    # extension (no reference counterpart): Spalart-Allmaras transport equation
    ntrans: int = 0
    wall_distance: Optional[np.ndarray] = None  # [nic][njc]
    beta: Optional[np.ndarray] = None
    label: str = "case"
    window: Optional[tuple] = None  # (first vertex row, first cell row) when xv/yv/fields hold only a row window

    @property
    def nic(self) -> int:
        return self.ni - 1

    @property
    def njc(self) -> int:
        return self.nj - 1

    @property
    def nv(self) -> int:
        return 4 + self.ntrans

    @property
    def viscous(self) -> bool:
        return self.mu_inf > 1e-15  # src/utils/config.cpp:41

    def freestream_q(self) -> np.ndarray:
        """EulerEquation::initialize (src/model/eulerequation.cpp:262-276); SA slot = 3 nu_inf * rho."""
        q = np.empty((self.nic, self.njc, self.nv)) if self.nic * self.njc < (1 << 26) else None
        if q is None:  # huge grids: callers only need the (uniform) cell value, see perturbed_q
            q = np.empty((1, 1, self.nv))
        q[..., 0] = self.rho_inf
        q[..., 1] = self.rho_inf * self.u_inf
        q[..., 2] = self.rho_inf * self.v_inf
        q[..., 3] = self.p_inf / (GAMMA - 1.0) + 0.5 * self.rho_inf * (self.u_inf**2 + self.v_inf**2)
        if self.ntrans:
            q[..., 4] = 3.0 * self.mu_inf
        return q

    def perturbed_q(self, amp: float = 0.01, j_first: int = 0, j_count: Optional[int] = None) -> np.ndarray:
        """freestream x (1 + amp sin(1 + 0.7 i + 0.3 j + k)): SURVEY.md section 8(d) synthetic state.
        With j_first/j_count only the window of global rows [j_first, j_first + j_count) is generated."""
        j_count = self.njc - j_first if j_count is None else j_count
        fs = self.freestream_q()[0, 0]
        i = np.arange(self.nic, dtype=np.float64)[:, None, None]
        j = (j_first + np.arange(j_count, dtype=np.float64))[None, :, None]
        k = np.arange(self.nv, dtype=np.float64)[None, None, :]
        q = fs[None, None, :] * (1.0 + amp * np.sin(1.0 + 0.7 * i + 0.3 * j + k))
        # momentum components that are exactly zero at freestream get an additive perturbation so that
        # every Jacobian column is exercised
        for kk in (1, 2):
            if fs[kk] == 0.0:
                q[..., kk] = amp * self.rho_inf * 0.1 * np.sin(2.0 + 0.5 * i[..., 0] + 0.9 * j[..., 0])
        return np.ascontiguousarray(q)


# ------------------------------------------------------------------------------------------------
# reference file formats
# ------------------------------------------------------------------------------------------------
GRID_BIN_MAGIC = b"SGRIDF64"   # binary vertex file: magic, int32 ni, int32 nj, x[nj][ni] float64, y[nj][ni] float64 (j outer, i inner)


def write_grid_bin(filename: str, xv: np.ndarray, yv: np.ndarray) -> None:
    """Binary variant of the p3d file for large grids (parsing 134 M ASCII vertices takes minutes):
    the same j-outer / i-inner order as Mesh::plot3d_loader (src/utils/mesh.cpp:146-169), raw little-endian float64."""
    ni, nj = xv.shape
    with open(filename, "wb") as f:
        f.write(GRID_BIN_MAGIC)
        np.array([ni, nj], dtype="<i4").tofile(f)
        np.ascontiguousarray(xv.T, dtype="<f8").tofile(f)
        np.ascontiguousarray(yv.T, dtype="<f8").tofile(f)


def read_grid_bin(filename: str, ni: int, nj: int, rows: Optional[tuple] = None):
    """-> xv, yv [ni][rows]; rows = (first vertex row, count) reads only that window (what one rank of a slab run needs)"""
    with open(filename, "rb") as f:
        if f.read(8) != GRID_BIN_MAGIC:
            raise ValueError("file format not found!")  # src/utils/mesh.cpp:364
        hdr = np.fromfile(f, dtype="<i4", count=2)
    assert int(hdr[0]) == ni and int(hdr[1]) == nj, "binary grid header mismatch"
    j0, n = (0, nj) if rows is None else rows
    mm = np.memmap(filename, dtype="<f8", mode="r", offset=16, shape=(2, nj, ni))
    return np.ascontiguousarray(mm[0, j0:j0 + n, :].T), np.ascontiguousarray(mm[1, j0:j0 + n, :].T)


def read_grid(filename: str, ni: int, nj: int, fmt: str):
    """Mesh::simple_loader / plot3d_loader (src/utils/mesh.cpp:134-169): j outer, i inner; "bin": the binary variant."""
    if fmt == "bin":
        return read_grid_bin(filename, ni, nj)
    tok = np.array(open(filename).read().split(), dtype=np.float64)
    if fmt == "simple":
        xy = tok[: 2 * ni * nj].reshape(nj, ni, 2)
        return np.ascontiguousarray(xy[..., 0].T), np.ascontiguousarray(xy[..., 1].T)
    if fmt == "p3d":
        assert int(tok[0]) == 1 and int(tok[1]) == ni and int(tok[2]) == nj, "p3d header mismatch"
        x = tok[3 : 3 + ni * nj].reshape(nj, ni)
        y = tok[3 + ni * nj : 3 + 2 * ni * nj].reshape(nj, ni)
        return np.ascontiguousarray(x.T), np.ascontiguousarray(y.T)
    raise ValueError("file format not found!")  # src/utils/mesh.cpp:364


def write_grid_p3d(filename: str, xv: np.ndarray, yv: np.ndarray) -> None:
    ni, nj = xv.shape
    with open(filename, "w") as f:
        f.write("1\n%d %d\n" % (ni, nj))
        for a in (xv, yv):
            np.savetxt(f, a.T.reshape(-1), fmt="%.17e")


def write_grid_simple(filename: str, xv: np.ndarray, yv: np.ndarray) -> None:
    """the reference's "simple" format: `x y` per line, j outer / i inner (src/utils/mesh.cpp:134-143)"""
    ni, nj = xv.shape
    with open(filename, "w") as f:
        np.savetxt(f, np.stack([xv.T.reshape(-1), yv.T.reshape(-1)], axis=1), fmt="%.17e")


def case_from_toml(text: str, xv: Optional[np.ndarray] = None, yv: Optional[np.ndarray] = None, base_dir: str = ".") -> Case:
    """Parse reference `.inp` TOML text like Config / BoundaryContainer do; the grid is read from
    geometry.filename unless vertex arrays are passed."""
    t = tomllib.loads(text)
    g, fs, so = t.get("geometry", {}), t.get("freestream", {}), t.get("solver", {})
    src, io = t.get("source", {}), t.get("io", {})
    ni, nj = int(g.get("ni", 0)), int(g.get("nj", 0))
    c = Case(ni=ni, nj=nj, tail=int(g.get("tail", 0)))
    for key in ("rho_inf", "u_inf", "v_inf", "p_inf", "T_inf", "mu_inf", "pr_inf"):
        if key in fs:
            setattr(c, key, float(fs[key]))
    c.aoa = float(fs.get("aoa", 0.0)) * math.pi / 180.0
    c.order = int(so.get("order", 1))
    c.lhs_order = int(so.get("lhs_order", c.order))
    c.scheme = so.get("scheme", "forward_euler")
    c.flux = so.get("flux", "ausm")
    c.cfl = float(so.get("cfl", 1.0))
    c.iteration_max = int(so.get("iteration_max", 1))
    c.dpdx, c.dpdy = float(src.get("dpdx", 0.0)), float(src.get("dpdy", 0.0))
    c.label = io.get("label", "flow")
    tb = t.get("turbulence", {})  # new optional table; stock files do not have it
    c.ntrans = int(tb.get("ntrans", 0))
    if c.ntrans:
        nc = (ni - 1, nj - 1)
        wd = tb.get("wall_distance", "compute")
        if wd != "compute":
            c.wall_distance = np.fromfile(wd if os.path.isabs(wd) else os.path.join(base_dir, wd), dtype=np.float64).reshape(nc)
        if tb.get("beta_file"):
            bf = tb["beta_file"]
            c.beta = np.fromfile(bf if os.path.isabs(bf) else os.path.join(base_dir, bf), dtype=np.float64).reshape(nc)
    for b in t.get("boundary", []):
        c.boundaries.append(Boundary(type=b.get("type", ""), face=b.get("face", ""), start=int(b.get("start", 0)),
                                     end=int(b.get("end", 0)), u=float(b.get("u", 0.0)), v=float(b.get("v", 0.0)),
                                     T=float(b.get("T", 0.0))))
    if xv is not None:
        c.xv, c.yv = np.ascontiguousarray(xv, dtype=np.float64), np.ascontiguousarray(yv, dtype=np.float64)
    else:
        fn = g.get("filename", "grid.unf2")
        if not os.path.isabs(fn):
            fn = os.path.join(base_dir, fn)
        c.xv, c.yv = read_grid(fn, ni, nj, g.get("format", "grid.unf2"))
    return c


def load_case(config_path: str) -> Case:
    """Parse a stock reference `.inp` (TOML) + its grid, like Config / Mesh / BoundaryContainer do."""
    with open(config_path, "r") as f:
        text = f.read()
    return case_from_toml(text, base_dir=os.path.dirname(os.path.abspath(config_path)))


def write_case(case: Case, directory: str, name: str = "case") -> str:
    """Write `<name>.inp` + `<name>.p3d` that the unmodified reference accepts; returns the .inp path."""
    os.makedirs(directory, exist_ok=True)
    grid = os.path.join(directory, name + ".p3d")
    write_grid_p3d(grid, case.xv, case.yv)
    lines = ["[geometry]", 'filename = "./%s.p3d"' % name, "ni = %d" % case.ni, "nj = %d" % case.nj,
             "tail = %d" % case.tail, 'format = "p3d"', "", "[freestream]"]
    for key in ("rho_inf", "u_inf", "v_inf", "p_inf", "T_inf", "mu_inf", "pr_inf"):
        lines.append("%s = %s" % (key, repr(float(getattr(case, key)))))
    lines += ["aoa = %s" % repr(float(case.aoa * 180.0 / math.pi)), "", "[source]", "dpdx = %s" % repr(float(case.dpdx)),
              "dpdy = %s" % repr(float(case.dpdy)), "", "[solver]", "order = %d" % case.order,
              "lhs_order = %d" % (case.lhs_order if case.lhs_order is not None else case.order),
              "cfl = %s" % repr(float(case.cfl)), 'scheme = "%s"' % case.scheme, 'flux = "%s"' % case.flux,
              "iteration_max = %d" % case.iteration_max, "", "[io]", "stdout_frequency = 1000000",
              "fileout_frequency = 1000000", "restart = false", 'label = "%s"' % name, ""]
    if case.ntrans:              # optional table of the GPU drop-in (integration/solver_gpu.cpp); the stock reference ignores it
        lines += ["[turbulence]", "ntrans = %d" % case.ntrans]
        if case.wall_distance is None:
            lines.append('wall_distance = "compute"')
        else:
            np.ascontiguousarray(case.wall_distance, dtype=np.float64).tofile(os.path.join(directory, name + ".wdist.bin"))
            lines.append('wall_distance = "./%s.wdist.bin"' % name)
        if case.beta is not None:
            np.ascontiguousarray(case.beta, dtype=np.float64).tofile(os.path.join(directory, name + ".beta.bin"))
            lines.append('beta_file = "./%s.beta.bin"' % name)
        lines.append("")
    for b in case.boundaries:
        lines += ["[[boundary]]", 'name = ""', 'type = "%s"' % b.type, 'face = "%s"' % b.face, "start = %d" % b.start,
                  "end = %d" % b.end, "u = %s" % repr(float(b.u)), "v = %s" % repr(float(b.v)), "T = %s" % repr(float(b.T)), ""]
    path = os.path.join(directory, name + ".inp")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return path


# ------------------------------------------------------------------------------------------------
# synthetic grids
# ------------------------------------------------------------------------------------------------
def bump_channel_grid(nic: int, njc: int, L: float = 1.0, H: float = 0.25, s: float = 2.5, bump: float = 0.05,
                      skew: float = 0.0, j_first: int = 0, j_count: Optional[int] = None):
    """x_i = L i/nic ; y_j = H [1 + tanh(s(2j/njc - 1))/tanh(s)]/2 with a Gaussian bump on the lower wall.
    `skew` shears the interior grid lines so that chi normals are not axis aligned either.
    j_first / j_count select a window of vertex rows (slab runs generate only their own rows)."""
    j_count = njc + 1 - j_first if j_count is None else j_count
    i = np.arange(nic + 1, dtype=np.float64)[:, None]
    j = (j_first + np.arange(j_count, dtype=np.float64))[None, :]
    x = L * i / nic + 0.0 * j
    y = H * (1.0 + np.tanh(s * (2.0 * j / njc - 1.0)) / math.tanh(s)) / 2.0 + 0.0 * i
    y = y + bump * H * (1.0 - y / H) * np.exp(-(((x - L / 2.0) / (0.1 * L)) ** 2))
    if skew:
        x = x + skew * L / nic * np.sin(2.0 * np.pi * y / H) * np.sin(np.pi * x / L) ** 2
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def _cell_centres(xv, yv):
    xc = 0.25 * (xv[:-1, :-1] + xv[1:, :-1] + xv[:-1, 1:] + xv[1:, 1:])
    yc = 0.25 * (yv[:-1, :-1] + yv[1:, :-1] + yv[:-1, 1:] + yv[1:, 1:])
    return xc, yc


def channel_wall_distance(case: "Case", bottom=None, top=None) -> np.ndarray:
    """Distance of each cell centre to the nearer of the bottom (j=0) / top (j=nj-1) grid lines.
    bottom / top = (x, y) vertex rows of the walls when case.xv only holds a window of rows."""
    xc, yc = _cell_centres(case.xv, case.yv)
    bx, by = (case.xv[:, 0], case.yv[:, 0]) if bottom is None else bottom
    tx, ty = (case.xv[:, -1], case.yv[:, -1]) if top is None else top
    xb = 0.5 * (bx[:-1] + bx[1:])[:, None]; yb = 0.5 * (by[:-1] + by[1:])[:, None]
    xt = 0.5 * (tx[:-1] + tx[1:])[:, None]; yt = 0.5 * (ty[:-1] + ty[1:])[:, None]
    d = np.minimum(np.hypot(xc - xb, yc - yb), np.hypot(xc - xt, yc - yt))
    return np.ascontiguousarray(d)


def synthetic_beta(case: "Case", L: float = 1.0, H: float = 0.25) -> np.ndarray:
    xc, yc = _cell_centres(case.xv, case.yv)
    return np.ascontiguousarray(1.0 + 0.1 * np.sin(2.0 * np.pi * xc / L) * np.cos(np.pi * yc / H))


def turbulent_channel_case(nic: int, njc: int, ntrans: int = 1, order: int = 2, lhs_order: Optional[int] = None,
                           flux: str = "roe", mach: float = 0.2, reynolds: float = 5e6, periodic: bool = True,
                           cell_rows: Optional[tuple] = None) -> Case:
    """The C2/C3/C5 synthetic workload: bump channel, isothermal bottom wall, adiabatic top wall,
    periodic (or freestream/outflow) in i, M = 0.2, Re_L = 5e6, MUSCL + Roe + viscous (+ SA).
    cell_rows = (ja, jb): generate only the window of cell rows [ja, jb) (vertex rows [ja, jb]) of the GLOBAL
    nic x njc grid -- what one rank of a slab run needs; the case then carries `window = (ja, ja)`."""
    if cell_rows is None:
        xv, yv = bump_channel_grid(nic, njc)
    else:
        ja, jb = cell_rows
        xv, yv = bump_channel_grid(nic, njc, j_first=ja, j_count=jb - ja + 1)
    c = Case(ni=nic + 1, nj=njc + 1, xv=xv, yv=yv)
    c.rho_inf, c.u_inf, c.v_inf, c.p_inf, c.T_inf = 1.0, mach, 0.0, 1.0 / 1.4, 1.0 / 1.4
    c.mu_inf = c.rho_inf * c.u_inf * 1.0 / reynolds
    c.order, c.lhs_order, c.flux, c.scheme = order, (order if lhs_order is None else lhs_order), flux, "rk4_jameson"
    c.ntrans = ntrans
    c.boundaries = [
        Boundary("isothermalwall", "bottom", 1, -2, T=c.T_inf),
        Boundary("wall", "top", 1, -2),
    ]
    if periodic:
        c.boundaries.append(Boundary("periodic", "left", 0, -1))
    else:
        c.boundaries += [Boundary("freestream", "left", 0, -1), Boundary("outflow", "right", 0, -1)]
    if ntrans:
        if cell_rows is None:
            c.wall_distance = channel_wall_distance(c)
        else:
            bx, by = bump_channel_grid(nic, njc, j_first=0, j_count=1)
            tx, ty = bump_channel_grid(nic, njc, j_first=njc, j_count=1)
            c.wall_distance = channel_wall_distance(c, (bx[:, 0], by[:, 0]), (tx[:, 0], ty[:, 0]))
        c.beta = synthetic_beta(c)
    c.window = None if cell_rows is None else (cell_rows[0], cell_rows[0])
    c.label = "channel_%dx%d" % (nic, njc)
    return c


def wall_segments(case: "Case") -> np.ndarray:
    """[n][4] = x0 y0 x1 y1 of every boundary edge covered by a `wall` / `isothermalwall` table (ranges as in
    BoundaryContainer::get_index, src/model/bc.cpp:436-457: padded index p <-> cell p - 1).  Full grids only."""
    assert case.window is None, "wall edges need the full vertex arrays"
    nic, njc = case.nic, case.njc
    out = []
    for b in case.boundaries:
        if b.type not in ("wall", "isothermalwall"):
            continue
        horiz = b.face in ("bottom", "top")
        n = nic if horiz else njc
        end = b.end if b.end >= 0 else n + 2 + b.end
        for p in range(max(b.start, 1), min(end, n) + 1):
            if horiz:
                j = 0 if b.face == "bottom" else case.nj - 1
                out.append((case.xv[p - 1, j], case.yv[p - 1, j], case.xv[p, j], case.yv[p, j]))
            else:
                i = 0 if b.face == "left" else case.ni - 1
                out.append((case.xv[i, p - 1], case.yv[i, p - 1], case.xv[i, p], case.yv[i, p]))
    return np.array(out, dtype=np.float64).reshape(-1, 4)


def flat_plate_grid(nic: int, njc: int, L: float = 1.0, H: float = 0.25, s: float = 3.0, skew: float = 0.15):
    """Flat-plate grid: x uniform, y clustered at the (flat) lower boundary, y_j = H [1 - tanh(s (1 - j/njc))/tanh(s)];
    `skew` shears the interior grid lines (zero at j = 0 and j = njc) so that the chi normals are not axis aligned."""
    i = np.arange(nic + 1, dtype=np.float64)[:, None]
    j = np.arange(njc + 1, dtype=np.float64)[None, :]
    y = H * (1.0 - np.tanh(s * (1.0 - j / njc)) / math.tanh(s)) + 0.0 * i
    x = L * i / nic + skew * L / nic * np.sin(np.pi * y / H) * np.sin(2.0 * np.pi * i / nic)
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def flat_plate_case(nic: int, njc: int, ntrans: int = 1, order: int = 2, lhs_order: Optional[int] = None, flux: str = "roe",
                    mach: float = 0.2, reynolds: float = 5e6, leading_edge: float = 0.2) -> Case:
    """BASELINE.json config 2: turbulent flat plate -- bottom boundary `slipwall` up to
    x = leading_edge * L, adiabatic `wall` from there on, `freestream` on top and at the inflow, `outflow` on the right;
    M = 0.2, Re_L = 5e6, MUSCL + Roe + viscous + SA.  The wall distance is left to the device
    (`wall_distance = None` -> nearest edge of the wall segment, sgpu_wall_distance_from_bcs)."""
    xv, yv = flat_plate_grid(nic, njc)
    c = Case(ni=nic + 1, nj=njc + 1, xv=xv, yv=yv)
    c.rho_inf, c.u_inf, c.v_inf, c.p_inf, c.T_inf = 1.0, mach, 0.0, 1.0 / 1.4, 1.0 / 1.4
    c.mu_inf = c.rho_inf * c.u_inf * 1.0 / reynolds
    c.order, c.lhs_order, c.flux, c.scheme = order, (order if lhs_order is None else lhs_order), flux, "rk4_jameson"
    c.ntrans = ntrans
    ile = max(1, min(nic - 1, int(round(leading_edge * nic))))        # cells 0 .. ile-1 slip, ile .. nic-1 wall
    c.boundaries = [
        Boundary("slipwall", "bottom", 1, ile),
        Boundary("wall", "bottom", ile + 1, -2),
        Boundary("freestream", "top", 0, -1),
        Boundary("freestream", "left", 0, -1),
        Boundary("outflow", "right", 0, -1),
    ]
    if ntrans:
        c.beta = synthetic_beta(c)
    c.label = "flatplate_%dx%d" % (nic, njc)
    return c


def zoo_case(name: str, nic: int = 24, njc: int = 16, ntrans: int = 0) -> Case:
    """Small synthetic cases that together exercise every BC type / face the reference implements,
    both fluxes, both reconstruction orders and the inviscid switch (parity-test cases, not bench lines)."""
    xv, yv = bump_channel_grid(nic, njc, L=1.0, H=0.5, s=1.2, bump=0.1, skew=0.3)
    c = Case(ni=nic + 1, nj=njc + 1, xv=xv, yv=yv, label="zoo_" + name)
    c.rho_inf, c.u_inf, c.v_inf, c.p_inf, c.T_inf = 1.0, 0.3, 0.02, 1.0 / 1.4, 1.0 / 1.4
    c.mu_inf, c.scheme = 2e-3, "rk4_jameson"
    if name == "A":      # slipwall bottom, adiabatic wall top, freestream in, outflow out; MUSCL + Roe, viscous
        c.order, c.lhs_order, c.flux = 2, 2, "roe"
        c.boundaries = [Boundary("slipwall", "bottom", 1, -2), Boundary("wall", "top", 1, -2, u=0.05),
                        Boundary("freestream", "left", 0, -1), Boundary("outflow", "right", 0, -1)]
    elif name == "B":    # walls on left/right, periodic bottom/top; first order + AUSM, viscous, source term
        c.order, c.lhs_order, c.flux = 1, 1, "ausm"
        c.dpdx, c.dpdy = -0.01, 0.005
        c.boundaries = [Boundary("wall", "left", 1, -2), Boundary("wall", "right", 1, -2, v=0.01),
                        Boundary("periodic", "bottom", 0, -1)]
    elif name == "C":    # inviscid, MUSCL rhs with first-order lhs (the NACA setting), Roe
        c.mu_inf = 0.0
        c.order, c.lhs_order, c.flux = 2, 1, "roe"
        c.boundaries = [Boundary("slipwall", "bottom", 1, -2), Boundary("freestream", "top", 0, -1),
                        Boundary("freestream", "left", 0, -1), Boundary("freestream", "right", 0, -1)]
    elif name == "D":    # isothermal walls + periodic in i (the channel setting) with AUSM second order
        c.order, c.lhs_order, c.flux = 2, 2, "ausm"
        c.dpdx = -0.02
        c.boundaries = [Boundary("isothermalwall", "top", 1, -2, T=0.75), Boundary("isothermalwall", "bottom", 1, -2, T=0.70, u=0.02),
                        Boundary("periodic", "left", 0, -1)]
    elif name == "E":    # wake cut on the bottom (C-grid style) + wall, freestream elsewhere; first-order lhs
        c.order, c.lhs_order, c.flux = 2, 1, "roe"
        nw = nic // 4
        c.boundaries = [Boundary("freestream", "top", 0, -1), Boundary("freestream", "left", 0, -1),
                        Boundary("freestream", "right", 0, -1), Boundary("wake", "bottom", 1, nw),
                        Boundary("wall", "bottom", nw + 1, nic - nw)]
    else:
        raise ValueError(name)
    if ntrans:                   # SA x BC matrix: nu~ ghost rules of every BC type (wall distance: nearest wall edge, if any)
        assert c.viscous, "the SA extension needs a viscous case"
        c.ntrans = ntrans
        c.label += "_sa"
        c.beta = synthetic_beta(c, L=1.0, H=0.5)
    return c


ZOO = ("A", "B", "C", "D", "E")
ZOO_SA = ("A", "B", "D", "E")    # C is inviscid
