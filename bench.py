#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: Mcell residual evals/s (+ Jacobian build ms, % HBM roofline).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   the CPU implementation of the same path on the host cores

A "step" is one residual evaluation (boundary conditions + the fused residual kernel + norm partials;
with N > 1 preceded by the ghost-row exchange) over the whole synthetic grid, state resident in HBM.
Workload (BASELINE.json configs[2] / configs[4]): bump-channel grid, SA turbulent (nv = 5), MUSCL + Roe +
viscous, fp64.  N = 1: 4096 x 4096 cells.  N > 1: 16384 x (1024 N) cells in j-slabs of 16384 x 1024 per GPU
(N = 8 is the named 16384 x 8192 grid) -- the same 16.8 M cells per GPU, i.e. weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_CELL = {4: 104.0, 5: 136.0}          # compulsory bytes per cell-eval
JAC_B_PER_CELL = {4: 1736.0, 5: 2696.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nic", type=int, default=0, help="override grid (cells in i)")
    ap.add_argument("--njc", type=int, default=0, help="override grid (cells in j, per GPU)")
    ap.add_argument("--ntrans", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-jacobian", action="store_true")
    ap.add_argument("--no-linsolve", action="store_true", help="skip the device GMRES sample on the built Jacobian (N=1 only)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="ghost-row transport for N > 1")
    return ap.parse_args()


def workload_dims(args, n_gpus):
    if args.nic and args.njc:
        return args.nic, args.njc * n_gpus, args.njc
    if n_gpus == 1:
        return 4096, 4096, 4096
    return 16384, 1024 * n_gpus, 1024


def workload_name(nic, njc_total, nv, n_gpus, njc_per):
    """identical in both arms (the driver pairs the lines by config); sampling details go to cpu_baseline.sample"""
    return "SA turbulent bump channel %dx%d cells, MUSCL+Roe+viscous, nv=%d, fp64%s" % (
        nic, njc_total, nv, "" if n_gpus == 1 else ", j-slabs of %dx%d per GPU" % (nic, njc_per))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_cpus(local, world):
    """Pin this rank to its GPU's CPU set (`nvidia-smi topo -m`, column "CPU Affinity"); when several ranks share one set
    the set is split evenly between them.  Returns a short description for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        sets = {}
        for ln in out.splitlines():
            f = ln.split()
            if not f or not f[0].startswith("GPU") or not f[0][3:].isdigit():
                continue
            rng = [x for x in f[1:] if x[0].isdigit() and ("-" in x or "," in x or x.isdigit())]
            if not rng:
                continue
            cpus = set()
            for part in rng[0].split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            sets[int(f[0][3:])] = sorted(cpus)
        mine = sets.get(local)
        if not mine:
            return "unbound (no CPU affinity reported)"
        sharers = sorted(g for g, c in sets.items() if c == mine and g < world)
        if len(sharers) > 1 and local in sharers:
            k, n = sharers.index(local), len(sharers)
            per = max(1, len(mine) // n)
            mine = mine[k * per:(k + 1) * per] or mine
        allowed = sorted(set(mine) & os.sched_getaffinity(0)) or sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, allowed)
        return "cpus %d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))
    except Exception as ex:  # noqa: BLE001
        return "unbound (%s)" % str(ex)[:60]


def copy_ceiling(torch, dist, world, h2d_bytes, d2h_bytes, reps=3):
    """What the PCIe / host-memory path allows with NO kernel in between: every rank copies its e2e step's bytes
    host->device and device->host concurrently on two streams; returns ms per step (max over ranks)."""
    hin = torch.empty(h2d_bytes // 8, dtype=torch.float64).pin_memory()
    hout = torch.empty(d2h_bytes // 8, dtype=torch.float64).pin_memory()
    din = torch.empty(h2d_bytes // 8, dtype=torch.float64, device="cuda")
    dout = torch.empty(d2h_bytes // 8, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def once():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        once()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def build_case(nic, njc_total, ntrans, cell_rows=None):
    from structured_b200.cases import turbulent_channel_case
    return turbulent_channel_case(nic, njc_total, ntrans=ntrans, order=2, flux="roe", mach=0.2, reynolds=5e6, periodic=True,
                                  cell_rows=cell_rows)


# -------------------------------------------------------------------------------------------------
# CPU baseline (the checker, timed): port of the same workload on a bounded sample + the reference itself
# -------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, nic, njc, ntrans, reps = args
    sys.path.insert(0, ROOT)
    from oracle.bindings import PortOracle, RefOracle
    case = build_case(nic, njc, ntrans)
    q = case.perturbed_q()
    orc = RefOracle(case) if kind == "reference" else PortOracle(case)
    orc.time_residual(q, 1)
    t = orc.time_residual(q, reps)
    return t


def _cpu_jacobian_worker(args):
    nic, njc, ntrans = args
    sys.path.insert(0, ROOT)
    from oracle.bindings import PortOracle
    case = build_case(nic, njc, ntrans)
    orc = PortOracle(case)
    t, nnz = orc.time_jacobian(case.perturbed_q(), True)
    return t, nnz, orc.ncolors if hasattr(orc, "ncolors") else None


def cpu_jacobian(nic, njc, ntrans):
    """the oracle's stand-in for trace_on .. sparse_jac (src/solver/solver.cpp:72-90,154-157) on a bounded sample, 1 core"""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(1) as pool:
        return pool.map(_cpu_jacobian_worker, [(nic, njc, ntrans)])[0]


def cpu_throughput(kind, nic, njc, ntrans, reps, procs):
    """procs independent evaluations run concurrently (the reference is single threaded: src/SConstruct:34-36
    OpenMP option is broken and scales negatively); returns Mcell-evals/s over all procs."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.time()
    with ctx.Pool(procs) as pool:
        ts = pool.map(_cpu_worker, [(kind, nic, njc, ntrans, reps)] * procs)
    wall = max(ts)
    return procs * nic * njc * reps / wall / 1e6, wall, time.time() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.bindings import have_ref
    nic, njc_total, _ = workload_dims(args, args.gpus)
    cores = os.cpu_count() or 1
    sn = 512                                        # sample: 512 x 512 cells of the same workload per core
    reps = max(1, args.steps // 4)
    for _ in range(max(1, min(args.warmup, 1))):
        pass
    val, wall, total = cpu_throughput("port", sn, sn, args.ntrans, reps, cores)
    extra = {}
    if have_ref() and not args.no_cpu_baseline:
        v1, _, _ = cpu_throughput("reference", sn, sn, 0, reps, 1)
        extra = {"reference_laminar_1core_mcells": round(v1, 4)}
    sample = ("%d concurrent single-thread evaluations (one per host core) x %d reps of a %dx%d-cell sample of the workload "
              "(same grid generator, state, BCs, nv=%d); CPU cost per cell is size independent") % (cores, reps, sn, sn, 4 + args.ntrans)
    out = {"impl": "reference", "metric": "residual_mcell_evals_per_s", "value": round(val, 4), "unit": "Mcell-evals/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * wall / reps, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(nic, njc_total, 4 + args.ntrans, args.gpus, workload_dims(args, args.gpus)[2])},
           "cpu_baseline": {"value": round(val, 4), "unit": "Mcell-evals/s", "cores": cores, "kind": "port", "sample": sample, **extra},
           "e2e": {"value": round(val, 4), "unit": "Mcell-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from structured_b200.api import GpuEulerEquation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_cpus(local, world)            # before any pinned allocation (first-touch locality of the staging pages)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    nic, njc_total, njc_per = workload_dims(args, n_gpus)
    nv = 4 + args.ntrans
    from structured_b200.slab import partition_rows
    j0, j1 = partition_rows(njc_total, world)[rank]
    # each rank generates only the grid rows its slab needs (own rows + 2 ghost rows per interior side)
    win = (max(j0 - 2, 0), min(j1 + 2, njc_total)) if world > 1 else None
    t_setup = time.time()
    case = build_case(nic, njc_total, args.ntrans, cell_rows=win)
    t_case = time.time() - t_setup
    eq = GpuEulerEquation(case, device=local, j_begin=j0 if world > 1 else 0, j_end=j1 if world > 1 else 0, window=case.window)
    # each rank generates only its own rows (+2 ghost rows each side) of the synthetic state
    jw0, jw1 = (max(j0 - 2, 0), min(j1 + 2, njc_total)) if world > 1 else (0, njc_total)
    q = case.perturbed_q(j_first=jw0, j_count=jw1 - jw0)
    eq.set_state_window(q, jw0, 0)
    eq.synchronize()
    # setup per rank = grid window + context (vertices, metrics, wall distance, beta) + state upload
    setup = {"total_s": round(time.time() - t_setup, 2), "grid_window_generation_s": round(t_case, 2),
             "what": "per rank: vertex rows of the slab generated on the host (a binary vertex file is read the same way: sgpu_set_grid_file), context + metrics on the device, state upload"}
    cells_local = nic * njc_per
    cells_total = nic * njc_total

    from structured_b200.slab import HIGH, LOW, HaloExchanger
    halo = HaloExchanger(rank, world, eq.halo_count(), "cuda", dist) if world > 1 else None
    halo_mode = "none"
    if world > 1:
        halo_mode = args.halo
        if halo_mode == "p2p":
            # NVLink peer memory: every rank opens its neighbours' receive buffers through CUDA IPC handles
            try:
                mine = (eq.halo_ipc_handle(LOW), eq.halo_ipc_handle(HIGH))
                allh = [None] * world
                dist.all_gather_object(allh, mine)
                if rank > 0:
                    eq.halo_open_peer(LOW, allh[rank - 1][HIGH])
                if rank < world - 1:
                    eq.halo_open_peer(HIGH, allh[rank + 1][LOW])
                ok = torch.tensor([1.0], device="cuda")
            except Exception:  # noqa: BLE001
                ok = torch.tensor([0.0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1.0:
                halo_mode = "nccl"

    def exchange():
        if halo_mode == "p2p":
            eq.halo_push(0)          # boundary rows stored straight into the neighbours' memory + sequence flag
            eq.halo_pull(0)          # device-side wait on the neighbours' flags, then unpack into the ghost rows
        else:
            halo.exchange(lambda side, t: eq.halo_pack(0, side, t.data_ptr()), lambda side, t: eq.halo_unpack(0, side, t.data_ptr()))

    # ---- N > 1 self-check (outside the timed region): after corrupting the ghost rows, ONE exchange must deliver the rows
    #      the host generated for them bit for bit, and the slab's rhs must equal the rhs of the window-uploaded state
    halo_check = None
    if world > 1:
        ok = True
        want = {}
        for side, rows in ((LOW, (j0 - 2, j0)), (HIGH, (j1, j1 + 2))):
            if (side == LOW and rank > 0) or (side == HIGH and rank < world - 1):
                want[side] = np.ascontiguousarray(np.transpose(q[:, rows[0] - jw0:rows[1] - jw0, :], (2, 1, 0))).reshape(-1)   # [nv][2][nic]
        eq.residual_device(0)
        rhs_ref = np.zeros((nic, j1 - j0, nv)); eq.get_rhs_window(rhs_ref)      # ghosts uploaded from the host window
        junk = torch.full((eq.halo_count(),), 7.0, dtype=torch.float64, device="cuda")
        for side in want:
            eq.halo_unpack(0, side, junk.data_ptr())
        torch.cuda.synchronize(); dist.barrier()
        exchange()
        got = torch.empty(eq.halo_count(), dtype=torch.float64, device="cuda")
        for side, w in want.items():
            eq.halo_pack_ghost(0, side, got.data_ptr())
            torch.cuda.synchronize()
            ok = ok and np.array_equal(got.cpu().numpy(), w)
        eq.residual_device(0)
        rhs_x = np.zeros_like(rhs_ref); eq.get_rhs_window(rhs_x)
        ok = ok and np.array_equal(rhs_x, rhs_ref)
        del rhs_x, rhs_ref
        t = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        halo_check = "bit-exact" if t.item() >= 1.0 else "FAILED"

    l2 = np.zeros(nv)

    def step():
        if world > 1:
            exchange()
        eq.residual_device(0, lhs=False, norms=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    eq.enable_kernel_timing(True)
    launches0 = eq.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eq.launch_count - launches0
    ktimes = eq.kernel_times()
    eq.enable_kernel_timing(False)
    clocks = None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = cells_total / (ms_step * 1e-3) / 1e6

    # residual norms once (outside the timed loop: the reference computes them once per step, not per evaluation)
    l2 = eq.residual_device(0, norms=True)

    # ---- e2e: the call-site form through HOST buffers (pinned), H2D + kernel + D2H inside the timed region
    e2e = None
    # ---- one explicit Solver::step (rk4_jameson, src/solver/solver.cpp:109-114) on the device: dt + 4 x (BCs + residual + stage update).
    # Measured here, before the Jacobian / linear-solve legs allocate their 60 GB: an explicit run never holds those (with them resident
    # the fused step measures 0.5 ms slower on the same box)
    explicit = None
    if (world == 1 or halo_mode == "p2p") and not args.no_jacobian:
        try:
            eq.copy_state(1, 0)
            res = {}
            for name, env in (("rk4_step_ms", None), ("rk4_step_two_kernel_ms", "0")):
                if env is None:
                    os.environ.pop("SGPU_RK_FUSED", None)
                else:
                    os.environ["SGPU_RK_FUSED"] = env
                eq.explicit_step(1e-3, "rk4_jameson")
                eq.synchronize()
                t0 = time.time()
                for _ in range(5):
                    eq.explicit_step(1e-3, "rk4_jameson")
                eq.synchronize()
                ms = (time.time() - t0) / 5 * 1e3
                if world > 1:                               # every rank runs the same steps (ghost rows over peer memory inside the call)
                    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    ms = float(tt[0].item())
                res[name] = round(ms, 3)
            os.environ.pop("SGPU_RK_FUSED", None)
            eq.set_state_window(q, jw0, 0)                     # the later legs see the state the headline loop ran on
            eq.synchronize()
            res["what"] = ("calc_dt + 4 x (%sboundary conditions + residual with the stage update fused into its epilogue) + q <- q_tmp; two_kernel = residual + separate update pass"
                           % ("ghost-row exchange over peer memory + " if world > 1 else ""))
            explicit = res
        except Exception as ex:  # noqa: BLE001
            explicit = {"unavailable": str(ex)[:160]}

    if not args.no_e2e:
        qh = torch.from_numpy(q).pin_memory()
        qn = qh.numpy()
        rh = torch.empty((nic, njc_per, nv), dtype=torch.float64).pin_memory()
        rn = rh.numpy()
        k_e2e = max(2, min(args.steps, 5))
        dev_rhs = eq.get_rhs() if world == 1 else None      # device-resident path's result, before the host path overwrites it

        def e2e_step():
            # host q (pinned) -> device, residual, owned rhs rows -> host: the call-site form of calc_residual
            if world > 1:
                eq.calc_residual_window(qn, jw0, rn)
            else:
                eq.calc_residual(qn, out=rn)

        e2e_step()
        barrier()
        e0.record()
        for _ in range(k_e2e):
            e2e_step()
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        rows_in = njc_per + (0 if world == 1 else (2 if rank in (0, world - 1) else 4))
        # the host-buffer path and the device-resident path must produce the same rhs (they run the same kernel on
        # chunks): compare the last e2e result with the device rhs of the timed loop
        chk = None
        if dev_rhs is not None:
            chk = float(np.abs(rn - dev_rhs).max() / max(np.abs(dev_rhs).max(), 1e-300))
            dev_rhs = None
        h2d_b, d2h_b = int(nic * rows_in * nv * 8), int(nic * njc_per * nv * 8)
        ms_copy = copy_ceiling(torch, dist, world, h2d_b, d2h_b)
        e2e = {"value": round(cells_total / (ms_e / k_e2e * 1e-3) / 1e6, 3), "unit": "Mcell-evals/s", "rel_diff_vs_device_path": chk,
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "steps": k_e2e, "note": "per-GPU bytes; sgpu_residual_host(q_host, rhs_host) from pinned host arrays",
               "copy_ceiling": {"value": round(cells_total / (ms_copy * 1e-3) / 1e6, 1), "unit": "Mcell-evals/s",
                                "GBps_per_gpu_each_way": round(0.5 * (h2d_b + d2h_b) / (ms_copy * 1e-3) / 1e9, 1),
                                "what": "the same bytes moved by bare concurrent cudaMemcpyAsync H2D + D2H on all ranks at once, no kernel: the PCIe / host-memory ceiling of this box"},
               "cpu_affinity": affinity}

    if rank == 0:
        clocks = sampler.stop()                          # sampled across the timed residual loop and the e2e loop
    # ---- Jacobian build (second half of the metric); skipped when the device form is unavailable
    jac = None
    if not args.no_jacobian:
        try:
            slots, ms = eq.jacobian_device()
            slots, ms = eq.jacobian_device()
            jac = {"build_ms": round(ms, 4), "slots": slots, "Mcell_per_s": round(cells_local / (ms * 1e-3) / 1e6, 2),
                   "roofline_frac": round(cells_local * JAC_B_PER_CELL[nv] / (ms * 1e-3) / 1e9 / measured_peak()[0], 4)}
        except Exception as ex:  # noqa: BLE001
            jac = {"unavailable": str(ex)[:120]}

    # ---- device linear solve on that Jacobian: a bounded GMRES sample, N = 1 only
    lin = None
    if jac and "build_ms" in jac and world == 1 and not args.no_linsolve:
        try:
            eq.calc_dt(20.0)
            _, info = eq.linear_solve("lhs", precond="line_j", restart=20, max_iter=40, rtol=1e-8, want_x=False)
            # bytes the product must move: the block entries that are not structurally zero (45 of 325 at nv = 5, 13 slots:
            # mass row x corner cells, q4 column of the mass row and of the radius-2 arms -- the kernel does not read them) + x, y, dt
            zero_planes = (20 + 5 + 20) if (nv == 5 and jac["slots"] == 13) else 0
            mv_bytes = cells_local * (jac["slots"] * nv * nv - zero_planes + 2 * nv + 1) * 8
            lin = {"system": "(-J + 1/dt) dq = rhs at CFL 20 (src/solver/solver.cpp:162-175)", "method": "GMRES(20), right j-line block-tridiagonal preconditioner",
                   "iterations": info["iterations"], "rel_residual": info["rel_residual"],
                   "ms_per_iteration": round(info["solve_ms"] / max(info["iterations"], 1), 3), "setup_ms": round(info["setup_ms"], 3),
                   "matvec_ms": round(info["matvec_ms"], 4), "precond_ms": round(info["precond_ms"], 4),
                   "matvec_GBps": round(mv_bytes / (info["matvec_ms"] * 1e-3) / 1e9, 1) if info["matvec_ms"] > 0 else None,
                   "matvec_roofline_frac": round(mv_bytes / (info["matvec_ms"] * 1e-3) / 1e9 / measured_peak()[0], 4) if info["matvec_ms"] > 0 else None}
        except Exception as ex:  # noqa: BLE001
            lin = {"unavailable": str(ex)[:120]}

    # ---- N > 1: the same sample through the slab-partitioned solver (operand ghost rows exchanged per product, inner
    #      products all-reduced; structured_b200/slab.py)
    if jac and "build_ms" in jac and world > 1 and not args.no_linsolve:
        try:
            from structured_b200.slab import SlabLinearSolver
            eq.calc_dt(20.0)
            eq.residual_device(0)
            sl = SlabLinearSolver(eq, rank, world, dist)
            x = torch.zeros(sl.n, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
            eq.vec_from_rhs(x.data_ptr())

            def matvec():
                sl.halo.exchange(lambda side, t: eq.vec_halo_pack(x.data_ptr(), side, t.data_ptr()),
                                 lambda side, t: eq.vec_halo_unpack(x.data_ptr(), side, t.data_ptr()))
                eq.op_apply("J", x.data_ptr(), y.data_ptr())
            matvec(); barrier()
            e0.record()
            for _ in range(5):
                matvec()
            e1.record(); barrier()
            mv_ms = e0.elapsed_time(e1) / 5
            sl.solve("lhs", precond="line_j", restart=20, max_iter=2, rtol=1e-8)     # warm-up: workspace allocations, first launches
            barrier(); e0.record()
            _, info = sl.solve("lhs", precond="line_j", restart=20, max_iter=20, rtol=1e-8)
            e1.record(); barrier()
            sv_ms = e0.elapsed_time(e1)
            t = torch.tensor([mv_ms, sv_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mv_ms, sv_ms = float(t[0].item()), float(t[1].item())
            zero_planes = (20 + 5 + 20) if (nv == 5 and jac["slots"] == 13) else 0
            mv_bytes = cells_local * (jac["slots"] * nv * nv - zero_planes + 2 * nv) * 8
            lin = {"system": "(-J + 1/dt) dq = rhs at CFL 20 (src/solver/solver.cpp:162-175), j-slabs", "method": "GMRES(20) over torch.distributed, slab-local j-line preconditioner",
                   "iterations": info["iterations"], "rel_residual": info["rel_residual"], "ms_per_iteration_incl_setup": round(sv_ms / max(info["iterations"], 1), 3),
                   "jacobian_apply_ms": round(mv_ms, 4), "jacobian_apply_GBps_per_gpu": round(mv_bytes / (mv_ms * 1e-3) / 1e9, 1),
                   "jacobian_apply_roofline_frac": round(mv_bytes / (mv_ms * 1e-3) / 1e9 / measured_peak()[0], 4)}
            del x, y, sl
        except Exception as ex:  # noqa: BLE001
            lin = {"unavailable": str(ex)[:160]}

    # ---- one implicit Solver::step on the device (src/solver/solver.cpp:66-101,154-175): dt, residual, Jacobian, line factors, GMRES, update
    implicit = None
    if world == 1 and not args.no_jacobian and not args.no_linsolve:
        try:
            eq.set_state_window(q, jw0, 0)
            eq.synchronize()
            t0 = time.time()
            l2i, info = eq.implicit_step(20.0, 1.0, precond="line_j", restart=20, max_iter=40, rtol=1e-2)
            eq.synchronize()
            implicit = {"ms": round((time.time() - t0) * 1e3, 1), "gmres_iterations": info["iterations"], "gmres_rel_residual": info["rel_residual"],
                        "setup_ms": round(info["setup_ms"], 2), "solve_ms": round(info["solve_ms"], 1),
                        "what": "sgpu_implicit_step at CFL 20: calc_dt + residual + Jacobian build + twisted line factors + GMRES(20) to 1e-2 + update, device resident"}
        except Exception as ex:  # noqa: BLE001
            implicit = {"unavailable": str(ex)[:160]}

    # ---- the laminar rows the reference itself has (nv = 4, 104 B/cell): same grid, device resident, beside the COMPILED
    #      REFERENCE (oracle/_ref, kind "reference") on a bounded sample -- the one pairing against the reference's own code
    laminar = None
    if world == 1 and args.ntrans == 1 and not args.no_cpu_baseline:
        try:
            lcase = build_case(nic, njc_total, 0)
            leq = GpuEulerEquation(lcase, device=local)
            leq.set_state(lcase.perturbed_q())
            for _ in range(3):
                leq.residual_device(0)
            leq.enable_kernel_timing(True)
            for _ in range(10):
                leq.residual_device(0)
            lk = float(np.mean(leq.kernel_times()))
            leq.close()
            laminar = {"workload": "laminar bump channel %dx%d cells, MUSCL+Roe+viscous, nv=4, fp64" % (nic, njc_total),
                       "kernel_ms": round(lk, 4), "value": round(cells_local / (lk * 1e-3) / 1e6, 1), "unit": "Mcell-evals/s",
                       "roofline_frac": round(cells_local * B_PER_CELL[4] / (lk * 1e-3) / 1e9 / measured_peak()[0], 4), "bytes_per_cell": B_PER_CELL[4]}
        except Exception as ex:  # noqa: BLE001
            laminar = {"unavailable": str(ex)[:160]}

    # ---- BASELINE.json config 2: SA flat plate, ~1 M cells (slipwall -> wall junction, freestream / outflow, wall distance
    #      computed on the device): residual kernel and Jacobian build on one GPU
    plate = None
    if world == 1 and args.ntrans == 1 and not args.no_jacobian:
        try:
            from structured_b200.cases import flat_plate_case
            pcase = flat_plate_case(1024, 1024)
            t0w = time.time()
            peq = GpuEulerEquation(pcase, device=local)            # includes sgpu_wall_distance_from_bcs (820 wall edges x 1 M cells)
            peq.synchronize()
            setup_s = time.time() - t0w
            peq.set_state(pcase.perturbed_q())
            for _ in range(3):
                peq.residual_device(0)
            peq.enable_kernel_timing(True)
            for _ in range(20):
                peq.residual_device(0)
            pk = float(np.mean(peq.kernel_times()))
            peq.jacobian_device()
            _, pj = peq.jacobian_device()
            peq.close()
            pc = 1024 * 1024
            plate = {"workload": "SA flat plate 1024x1024 cells (BASELINE.json config 2), MUSCL+Roe+viscous, nv=5, fp64",
                     "residual_kernel_ms": round(pk, 4), "residual_Mcell_per_s": round(pc / (pk * 1e-3) / 1e6, 1),
                     "residual_roofline_frac": round(pc * B_PER_CELL[5] / (pk * 1e-3) / 1e9 / measured_peak()[0], 4),
                     "jacobian_build_ms": round(pj, 4), "jacobian_roofline_frac": round(pc * JAC_B_PER_CELL[5] / (pj * 1e-3) / 1e9 / measured_peak()[0], 4),
                     "context_setup_s_incl_wall_distance": round(setup_s, 3)}
        except Exception as ex:  # noqa: BLE001
            plate = {"unavailable": str(ex)[:160]}

    peak, peak_src = measured_peak()
    kt = float(np.mean(ktimes)) if len(ktimes) else None
    achieved = cells_local * B_PER_CELL[nv] / (kt * 1e-3) / 1e9 if kt else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "residual_traffic.json")
    if os.path.exists(tp) and nv == 5 and nic * njc_per == 4096 * 4096:
        try:                                             # ncu --set full capture of this kernel on the 16.8 M-cell SA workload
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(achieved, 2) if achieved else None, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic,
                "traffic_source": "ncu --set full capture, committed (profiles/residual_traffic.json); not measured by this run" if traffic else None,
                "kernel": "residual_kernel<nv=%d,order=2,roe,viscous>" % nv, "kernel_ms": round(kt, 4) if kt else None,
                "bytes_per_cell": B_PER_CELL[nv], "peak_source": peak_src}
    # the kernel is fp64-ISSUE bound, not HBM bound: the second roofline.  instr_per_cell = fp64 instructions of
    # the SASS row loop per thread x (threads per strip / cells per strip), from the committed static count
    roofline_fp64 = None
    fp = os.path.join(ROOT, "profiles", "residual_fp64.json")
    if kt and os.path.exists(fp) and nv == 5:
        try:
            ipc = float(json.load(open(fp))["fp64_instr_per_cell"])
            sms, clk = torch.cuda.get_device_properties(local).multi_processor_count, (clocks or {}).get("sm_mhz") or 1965.0
            floor_ms = ipc * cells_local / (sms * 64.0 * clk * 1e6) * 1e3          # 64 fp64 lanes per SM and clock
            roofline_fp64 = {"instr_per_cell": ipc, "floor_ms": round(floor_ms, 4), "frac": round(floor_ms / kt, 4),
                             "hbm_frac_at_floor": round(cells_local * B_PER_CELL[nv] / (floor_ms * 1e-3) / 1e9 / peak, 4),
                             "source": "static SASS count, profiles/residual_fp64.json; 64 fp64 lanes/SM/clk at the sampled SM clock"}
        except Exception:
            roofline_fp64 = None

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        from oracle.bindings import have_ref
        sn, reps = 512, 4
        v, wall, _ = cpu_throughput("port", sn, sn, args.ntrans, reps, 1)
        cpu = {"value": round(v, 4), "unit": "Mcell-evals/s", "cores": 1, "kind": "port",
               "sample": "%d reps of a %dx%d-cell sample of the same workload (nv=%d), 1 thread" % (reps, sn, sn, nv)}
        if have_ref():
            v2, _, _ = cpu_throughput("reference", sn, sn, 0, reps, 1)
            cpu["reference_laminar_1core"] = round(v2, 4)
            if laminar and "value" in laminar:
                laminar["cpu_baseline"] = {"value": round(v2, 4), "unit": "Mcell-evals/s", "cores": 1, "kind": "reference",
                                           "sample": "%d reps of EulerEquation::calc_residual (oracle/_ref, the reference compiled unmodified) on a %dx%d-cell sample of the same laminar workload" % (reps, sn, sn)}
        # Jacobian CPU baseline: the oracle's dual-number coloured Jacobian, 1 core, bounded sample
        try:
            jn = 160
            tj, nnzj, ncol = cpu_jacobian(jn, jn, args.ntrans)
            cpu["jacobian"] = {"ms_per_Mcell": round(tj * 1e3 / (jn * jn / 1e6), 1), "Mcell_per_s": round(jn * jn / tj / 1e6, 5), "cores": 1, "kind": "port",
                               "what": "oracle stand-in for trace_on .. sparse_jac (src/solver/solver.cpp:72-90,154-157): index-domain pattern + greedy colouring (%s colours) + dual-number sweeps; NOT ADOL-C (un-vendored, absent from this image)" % ncol,
                               "sample": "one build on a %dx%d-cell sample of the same workload (nv=%d), nnz %d, %.2f s" % (jn, jn, nv, nnzj, tj),
                               "extrapolated_ms_at_workload": round(tj * 1e3 * cells_local / (jn * jn), 0)}
        except Exception as ex:  # noqa: BLE001
            cpu["jacobian"] = {"unavailable": str(ex)[:160]}

    if rank == 0:
        out = {"metric": "residual_mcell_evals_per_s", "value": round(value, 2), "unit": "Mcell-evals/s", "n_gpus": n_gpus,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(nic, njc_total, nv, n_gpus, njc_per),
                          "l2_flush": "inputs (q %.0f MB + rhs %.0f MB per GPU) exceed the 126 MB L2" % (cells_local * nv * 8 / 1e6, cells_local * nv * 8 / 1e6),
                          "step": "ghost-row exchange (N>1) + boundary conditions + fused residual kernel", "halo": halo_mode},
               "roofline": roofline, "roofline_fp64": roofline_fp64, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
               "jacobian": jac, "linear_solve": lin, "laminar": laminar, "flat_plate": plate, "halo_check": halo_check, "setup": setup, "explicit_step": explicit, "implicit_step": implicit,
               "l2norm": [float(x) for x in np.sqrt(l2)]}
        print(json.dumps(out), flush=True)
    eq.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
