#!/usr/bin/env python
"""bench.py -- IFunction+IJacobian assembled DOF/s on the headline configuration C4
(3-D heat, P2 tetrahedra, UnitCube n=184: 37,377,024 cells, 50,243,409 DOFs, 1,437,462,465
non-zeros).

    python bench.py --gpus N --steps K --warmup W            # the B200 engine
    python bench.py --impl reference --gpus N ...            # restated CPU assembly (oracle port)

One "step" = one fdts_ifunction + one fdts_ijacobian over the whole (rank-local) mesh, i.e. what
PETSc TS asks of _TSContext.form_function / form_jacobian per Newton iteration
(firedrake_ts/solving_utils.py:236-317).  `value` is global DOFs / time per step with inputs
resident in HBM; `e2e` is the same pair through the host-buffer C ABI (fdts_*_host: pinned host
x, xdot in, F and the CSR values out, copies inside the timed region).  Inputs are far larger
than L2 (126 MB), so no flush between iterations is needed.  N > 1: strong scaling, the same
global problem box-partitioned over the ranks (2x1x1 / 2x2x1 / 2x2x2), one halo exchange of
u and udot per step over NCCL send/recv inside the timed region (hidden behind the interior
cells of the residual; ms_ifunction includes it), max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IFunction+IJacobian assembled DOF/s"
UNIT = "DOF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("-n", type=int, default=184, help="cubes per edge of the unit cube (184 = headline C4)")
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--cpu-n", type=int, default=None,
                    help="mesh size of the CPU runs (default: --impl reference runs the full configuration -n; "
                         "the cpu_baseline sample inside the B200 arm uses n=96)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C2/C3/C5 record (N=1 only)")
    ap.add_argument("--halo", default="native", choices=["native", "python"],
                    help="N > 1: ghost exchange inside the library (fdts_ifunction_exchange, default) or driven from "
                         "Python through torch.distributed P2P ops (round-1 path, kept for comparison)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-nodict", action="store_true")
    return ap.parse_args()


def workload_name(n, degree):
    nodes = (degree * n + 1) ** 3
    tag = "C4: " if (n == 184 and degree == 2) else ""
    return f"{tag}3-D heat P{degree} tets UnitCube n={n}, {6 * n ** 3} cells, {nodes} DOFs"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (multiprocessing, slab partition, slowest worker)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    rank, size, n, degree, steps, warmup = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import torch

    torch.set_num_threads(1)
    from firedrake_ts_b200 import mesh as M
    from firedrake_ts_b200.workloads import make_problem
    from oracle import Oracle

    part = M.Partition((size, 1, 1), rank)
    form, bcs, u, udot = make_problem("heat", 3, degree, n, partition=part)
    orc = Oracle.from_form(form, bcs)
    orc.pattern()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.ifunction(0.0, u.data, udot.data)
        orc.ijacobian(0.0, u.data, udot.data, 10.0)
        times.append(time.perf_counter() - t0)
    return sum(times[warmup:])


def cpu_pair_seconds(n, degree, steps, warmup, cores):
    """wall time of `steps` IFunction+IJacobian pairs on `cores` worker processes (slowest worker)."""
    import multiprocessing as mp

    cores = max(1, min(cores, n))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(r, cores, n, degree, steps, warmup) for r in range(cores)])
    return max(res), cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import build as build_oracle

    build_oracle()
    cores = os.cpu_count() or 1
    # the SAME configuration as the B200 arm (-n, default 184 = C4); bounded by the number of timed pairs:
    # one pair of the oracle port takes ~35 s at C4 on 16 cores, so 3 timed pairs + 1 warm-up + set-up
    # stay within a few minutes
    n = a.cpu_n or a.n
    steps, warmup = max(1, min(a.steps, 3 if n > 100 else 4)), min(a.warmup, 1)
    t, cores = cpu_pair_seconds(n, a.degree, steps, warmup, cores)
    dofs = (a.degree * n + 1) ** 3
    value = dofs * steps / t
    sample = (f"oracle port (gcc -O3), {cores} processes x 1 thread, slab partition of the same form on "
              f"UnitCube n={n} ({6 * n ** 3} cells, {dofs} DOFs), {steps} timed pairs, slowest worker")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, a.degree), "sample": f"n={n}, {steps} timed pairs",
                   "note": "restated CPU baseline (not Firedrake: the reference's Firedrake/PETSc stack is not "
                           "installable in this image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.05:
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        note = None
        if not sm and self.samples:  # timed region shorter than the sampling period: use the nearest samples
            near = sorted(self.samples, key=lambda s_: min(abs(s_[0] - t0), abs(s_[0] - t1)))[:3]
            for ts, line in near:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                except (ValueError, IndexError):
                    continue
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            note = "no sample fell inside the timed region; nearest samples used"
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------
def algorithmic_bytes(nc, nd, dim, nv, ndofs, nnz, linear=True):
    """compulsory traffic only."""
    maps = 4 * nc * nd + 4 * nc * (dim + 1) + 8 * dim * nv
    ifun = maps + 2 * 8 * ndofs + 8 * ndofs
    ijac = maps + (0 if linear else 8 * ndofs) + 8 * nnz
    return ifun, ijac


C5_FLOP_PER_CELL = (11040, 110400)  # dense-quadrature convention nq = 46, nd = 20


def time_other_configs(dev, hbm_peak, reps=5):
    """C1, C2, C3, C5 of SURVEY.md section 8(d) on one GPU (the headline line stays C4): ms per call, DOF/s and
    the roofline fraction of each (HBM bytes for C1-C3, the 8(d) flop convention against the fp64 peak measured
    on this box for C5)."""
    import torch

    from firedrake_ts_b200.engine import Engine, measure_fp64_peak
    from firedrake_ts_b200.workloads import CONFIGS, make_problem

    peak_dfma = measure_fp64_peak(dev.index or 0, False) / 1e3
    peak_dmma = measure_fp64_peak(dev.index or 0, True) / 1e3
    out = {"fp64_peak": {"dfma_tflops": peak_dfma, "dmma_tflops": peak_dmma, "how": "fdts_measure_fp64_peak: "
                         "dependent-free FMA / mma.sync.m8n8k4.f64 chains, all SMs, CUDA events"}}
    for name in ("C1", "C2", "C3", "C3q", "C5"):
        try:
            out[name] = _time_config(name, dev, hbm_peak, peak_dfma, peak_dmma, reps)
        except Exception as exc:  # a failing side configuration must not take the headline line down
            out[name] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
    return out


def _time_config(name, dev, hbm_peak, peak_dfma, peak_dmma, reps):
    import torch

    from firedrake_ts_b200.engine import Engine
    from firedrake_ts_b200.workloads import CONFIGS, make_problem

    if True:
        family, dim, degree, n, ftile, foffset, mtile = CONFIGS[name]
        quad = name.endswith("q")
        form, bcs, u, udot = make_problem(family, dim, degree, n, tile=mtile, device=dev, ftile=ftile or None,
                                          foffset=foffset, quadrilateral=quad)
        V = form.space
        eng = Engine(form, bcs, device=dev, scatter=2 if (family == "heat" and ftile) else None)
        x, xd = u.data, udot.data
        F = torch.empty(eng.num_rows, dtype=torch.float64, device=dev)
        J = eng.new_values()
        for _ in range(2):
            eng.ifunction(0.0, x, xd, out=F)
            eng.ijacobian(0.0, x, xd, 10.0, out=J)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(reps):
            eng.ifunction(0.0, x, xd, out=F)
        ev[1].record()
        for _ in range(reps):
            eng.ijacobian(0.0, x, xd, 10.0, out=J)
        ev[2].record()
        torch.cuda.synchronize()
        t_if, t_ij = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        nc, nv = V.mesh.num_cells, V.mesh.num_coord_nodes
        maps = 4 * nc * V.nd + 4 * nc * V.mesh.nverts + 8 * dim * nv
        b_if = maps + 3 * 8 * V.num_dofs
        b_ij = maps + (0 if family == "heat" else 8 * V.num_dofs) + 8 * eng.nnz
        rec = {"workload": f"{family} {'Q' if quad else 'P'}{degree} dim {dim} n={n}: {nc} cells, {V.num_dofs} DOFs, {eng.nnz} non-zeros",
               "ms_ifunction": t_if, "ms_ijacobian": t_ij, "value": V.num_dofs / ((t_if + t_ij) * 1e-3), "unit": UNIT,
               "algorithmic_bytes": {"ifunction": b_if, "ijacobian": b_ij}}
        if name == "C5":
            fl = [c * nc for c in C5_FLOP_PER_CELL]
            ach = sum(fl) / ((t_if + t_ij) * 1e-3) / 1e12
            rec["roofline"] = {"bound": "fp64", "achieved": ach, "peak": max(peak_dfma, peak_dmma), "unit": "TFLOP/s",
                               "frac": ach / max(peak_dfma, peak_dmma), "algorithmic_flop": {"ifunction": fl[0], "ijacobian": fl[1]},
                               "ifunction_frac": fl[0] / (t_if * 1e-3) / 1e12 / max(peak_dfma, peak_dmma),
                               "ijacobian_frac": fl[1] / (t_ij * 1e-3) / 1e12 / max(peak_dfma, peak_dmma)}
        else:
            ach = (b_if + b_ij) / ((t_if + t_ij) * 1e-3) / 1e9
            rec["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                               "ifunction_frac": b_if / (t_if * 1e-3) / 1e9 / hbm_peak,
                               "ijacobian_frac": b_ij / (t_ij * 1e-3) / 1e9 / hbm_peak}
        rec["checksum"] = {"sum_F": float(F.sum()), "sum_J": float(J.sum())}
        return rec


def run_b200(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu:
        # timed before CUDA work starts (spawned workers; GPU idle)
        cn = a.cpu_n or min(a.n, 96)
        t, cores = cpu_pair_seconds(cn, a.degree, 2, 1, os.cpu_count() or 1)
        d = (a.degree * cn + 1) ** 3
        cpu_base = {"value": d * 2 / t, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"oracle port (gcc -O3), {cores} processes x 1 thread, slab partition, same form on "
                              f"UnitCube n={cn} ({6 * cn ** 3} cells, {d} DOFs), 2 timed pairs, slowest worker; "
                              f"the full configuration is timed by --impl reference"}

    from firedrake_ts_b200 import mesh as M
    from firedrake_ts_b200.engine import Engine
    from firedrake_ts_b200.halo import HaloExchange
    from firedrake_ts_b200.workloads import make_problem

    part = M.Partition(M.default_grid(world, 3), rank) if world > 1 else None
    ftile = 6 if a.degree == 2 else (6 if a.degree == 1 else 0)
    mtile = 3 if a.degree == 2 else 6
    form, bcs, u, udot = make_problem("heat", 3, a.degree, a.n, tile=mtile, partition=part, device=dev, ftile=ftile)
    V = form.space
    eng = Engine(form, bcs, device=dev, scatter=2 if ftile else 1)
    halo = HaloExchange(V, eng, group=group, device=dev) if world > 1 else None
    native = world > 1 and a.halo == "native"
    if native:
        eng.halo_setup(group)  # the library's own NCCL communicator: exchange fused with the residual

    def residual_with_exchange():
        if native:
            eng.ifunction_exchange(0.0, x, xd, out=F)
        else:
            eng.ifunction_overlapped(0.0, x, xd, halo, out=F)
    x, xd = u.data, udot.data
    F = torch.empty(eng.num_rows, dtype=torch.float64, device=dev)
    J = eng.new_values()
    shift = 10.0  # 1/dt with PETSc's default dt = 0.1

    def step():
        if halo is not None:  # ghost exchange hidden behind the interior cells of the residual
            residual_with_exchange()
        else:
            eng.ifunction(0.0, x, xd, out=F)
        eng.ijacobian(0.0, x, xd, shift, out=J)

    # clocks are sampled from before the warm-up on (nvidia-smi needs a few 100 ms to start streaming)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wait = time.time()
    while sampler is not None and not sampler.samples and time.time() - t_wait < 3.0:
        time.sleep(0.05)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    l0 = eng.launch_count
    torch.cuda.synchronize()
    t_start = time.time()
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record()
    for k in range(a.steps):
        ev[k][0].record()
        if halo is not None:
            residual_with_exchange()
        else:
            eng.ifunction(0.0, x, xd, out=F)
        ev[k][1].record()
        eng.ijacobian(0.0, x, xd, shift, out=J)
        ev[k][2].record()
    e_end.record()
    torch.cuda.synchronize()
    t_end = time.time()
    launches = eng.launch_count - l0
    total_ms = e_beg.elapsed_time(e_end)
    t_if = sum(e[0].elapsed_time(e[1]) for e in ev) / a.steps
    t_ij = sum(e[1].elapsed_time(e[2]) for e in ev) / a.steps
    if world > 1:
        t = torch.tensor([total_ms, t_if, t_ij], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, t_if, t_ij = t.tolist()
        dist.barrier()
    clocks = sampler.summary(t_start, t_end) if sampler else None
    # partition-independent checksums of the assembled objects (compare N = 1 with N > 1)
    chk = torch.stack([F.sum(), F.abs().sum(), J.sum(), J.abs().sum()])
    if world > 1:
        dist.all_reduce(chk)
    chk = [float(v) for v in chk.tolist()]
    ndofs = V.global_num_dofs
    ms_per_step = total_ms / a.steps
    value = ndofs / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (IJacobian), algorithmic bytes of this rank's share
    nnz_local = eng.nnz
    b_if, b_ij = algorithmic_bytes(V.mesh.num_cells, V.nd, 3, V.mesh.num_coord_nodes, V.num_dofs, nnz_local)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = b_ij / (t_ij * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        key = f"n{a.n}_p{a.degree}_g{world}"
        traffic = tj.get(key)
    roofline = {"bound": "hbm", "kernel": "IJacobian (heat_jacobian_gather2)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": b_ij, "ms_per_launch": t_ij,
                "pair": {"algorithmic_bytes": b_if + b_ij, "achieved": (b_if + b_ij) / (ms_per_step * 1e-3) / 1e9,
                         "frac": (b_if + b_ij) / (ms_per_step * 1e-3) / 1e9 / peak},
                "ifunction": {"algorithmic_bytes": b_if, "ms": t_if, "frac": b_if / (t_if * 1e-3) / 1e9 / peak}}

    # the same Jacobian with the pattern dictionary switched off: every row block reads its own index
    # lists from HBM, which is what an unstructured mesh (no repeated blocks) would do
    if world == 1 and ftile and not a.no_nodict:
        eng_nd = Engine(form, bcs, device=dev, scatter=2, dedup=0)
        for _ in range(2):
            eng_nd.ijacobian(0.0, x, xd, shift, out=J)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng_nd.ijacobian(0.0, x, xd, shift, out=J)
        e1.record()
        torch.cuda.synchronize()
        roofline["ijacobian_no_dictionary"] = {"ms": e0.elapsed_time(e1) / 5, "index_bytes_per_launch":
                                               2 * eng_nd.get_option("plan_contributions")}
        del eng_nd
        torch.cuda.empty_cache()

    # end-to-end through the host-buffer C ABI (single rank only: host Vec/Mat of one process)
    e2e = None
    if not a.no_e2e and world == 1:
        xh = x.cpu().pin_memory()
        xdh = xd.cpu().pin_memory()
        Fh = torch.empty(eng.num_rows, dtype=torch.float64).pin_memory()
        Jh = torch.empty(eng.nnz, dtype=torch.float64).pin_memory()
        del J
        torch.cuda.empty_cache()
        ksteps = max(1, min(a.steps, 3))
        eng.ifunction_host(0.0, xh, xdh, Fh)
        eng.ijacobian_host(0.0, xh, xdh, shift, Jh)
        torch.cuda.synchronize()
        moved0 = (eng.get_option("h2d_bytes"), eng.get_option("d2h_bytes"))  # counted by the library per copy
        t0 = time.perf_counter()
        for _ in range(ksteps):
            eng.ifunction_host(0.0, xh, xdh, Fh)
            eng.ijacobian_host(0.0, xh, xdh, shift, Jh)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        e2e = {"value": ndofs / dt, "unit": UNIT,
               "h2d_bytes_per_step": (eng.get_option("h2d_bytes") - moved0[0]) // ksteps,
               "d2h_bytes_per_step": (eng.get_option("d2h_bytes") - moved0[1]) // ksteps,
               "steps": ksteps, "ms_per_step": dt * 1e3,
               "checksum": float(Fh.sum()) + float(Jh[:1000].sum())}
    elif world > 1 and not a.no_e2e:
        # every rank owns host copies of its Vec/Mat (the reference's VECMPI/MATMPIAIJ in host memory): per step the
        # owned part of x, xdot goes host->device, ghosts are exchanged on the device, F and the rank's CSR values
        # come back device->host; ranks run concurrently, wall clock between barriers, max over ranks
        nown = eng.num_rows
        xh, xdh = x[:nown].cpu().pin_memory(), xd[:nown].cpu().pin_memory()
        Fh = torch.empty(nown, dtype=torch.float64).pin_memory()
        Jh = torch.empty(eng.nnz, dtype=torch.float64).pin_memory()

        def e2e_step():
            x[:nown].copy_(xh, non_blocking=True)
            xd[:nown].copy_(xdh, non_blocking=True)
            residual_with_exchange()
            Fh.copy_(F, non_blocking=True)
            eng.ijacobian(0.0, x, xd, shift, out=J)
            Jh.copy_(J, non_blocking=True)

        e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        ksteps = max(1, min(a.steps, 3))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / ksteps], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nb = torch.tensor([2 * 8 * nown, 8 * nown + 8 * eng.nnz], dtype=torch.float64, device=dev)
        dist.all_reduce(nb)
        dt = float(dt.item())
        e2e = {"value": ndofs / dt, "unit": UNIT, "h2d_bytes_per_step": int(nb[0].item()),
               "d2h_bytes_per_step": int(nb[1].item()), "steps": ksteps, "ms_per_step": dt * 1e3,
               "note": "per-rank pinned host Vec/Mat; bytes summed over ranks; halo exchange on the device"}
    elif world > 1:
        e2e = None

    plan_info = (eng.get_option("plan_blocks"), eng.get_option("plan_patterns")) if ftile else (0, 0)
    configs = None
    if world == 1 and not a.no_configs and a.n == 184 and a.degree == 2:
        del eng, F, x, xd, u, udot, form, bcs, V
        torch.cuda.empty_cache()
        configs = time_other_configs(dev, peak)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a.n, a.degree), "partition": "x".join(map(str, M.default_grid(world, 3))),
                       "l2": "inputs >> L2 (126 MB), no flush", "halo": (a.halo if world > 1 else None), "scatter": "closed-form residual + atomics, "
                       "owner-computes gather Jacobian (plan 2: sector SELL + pattern dictionary)", "shift": shift,
                       "row_blocks": plan_info[0], "index_patterns": plan_info[1]},
            "ms_ifunction": t_if, "ms_ijacobian": t_ij,
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
            "checksum": {"sum_F": chk[0], "sum_absF": chk[1], "sum_J": chk[2], "sum_absJ": chk[3]},
            "clocks": clocks, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    # stdout carries exactly one JSON line: native libraries that print to file descriptor 1 (NCCL's
    # version banner under torchrun) are sent to stderr, Python's stdout keeps the original descriptor
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
