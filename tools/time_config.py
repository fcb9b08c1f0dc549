"""ad-hoc timing of one configuration (development aid; bench.py is the contract)."""
import argparse
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from helpers import make_problem
from firedrake_ts_b200.engine import Engine, measure_fp64_peak

ap = argparse.ArgumentParser()
ap.add_argument("--family", default="heat")
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("--degree", type=int, default=2)
ap.add_argument("-n", type=int, default=184)
ap.add_argument("--tile", default="4")
ap.add_argument("--scatter", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--peak", action="store_true")
ap.add_argument("--ftile", default=None)
ap.add_argument("--plan", type=int, default=2)
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--sector", type=int, default=4)
ap.add_argument("--optimize", type=int, default=1)
ap.add_argument("--threads", type=int, default=512)
ap.add_argument("--rtiles", type=int, default=0)
ap.add_argument("--dedup", type=int, default=1)
ap.add_argument("--foffset", type=int, default=0)
ap.add_argument("--cache-metric", type=int, default=1)
ap.add_argument("--numbering", default=None)
ap.add_argument("--block-rows", type=int, default=192)
ap.add_argument("--dmma", type=int, default=-1)
a = ap.parse_args()

if a.peak:
    print("fp64 DFMA GFLOP/s", measure_fp64_peak(0, False), "DMMA GFLOP/s", measure_fp64_peak(0, True), flush=True)
t0 = time.time()
def _tile(v):
    if v is None:
        return None
    t = tuple(int(x) for x in str(v).split(","))
    return t[0] if len(t) == 1 else t


form, bcs, u, udot = make_problem(a.family, a.dim, a.degree, a.n, tile=_tile(a.tile), device="cuda:0",
                                  ftile=_tile(a.ftile), foffset=a.foffset, numbering=a.numbering, block_rows=a.block_rows)
torch.cuda.synchronize()
print("mesh+space", time.time() - t0, "s; cells", form.space.mesh.num_cells, "dofs", form.space.num_dofs, flush=True)
t0 = time.time()
eng = Engine(form, bcs, scatter=a.scatter, plan_version=a.plan, dedup=a.dedup,
             options={"sector": a.sector, "optimize": a.optimize, "threads": a.threads, "residual_tiles": a.rtiles,
                      "cache_metric": a.cache_metric})
if a.rtiles:
    eng.set_option("residual_tiles", a.rtiles)
torch.cuda.synchronize()
print("engine create", time.time() - t0, "s; nnz", eng.nnz, "maxrow", eng.max_row_nodes, "maxdeg", eng.max_degree, flush=True)
if a.scatter == 2:
    print("plan: version", eng.get_option("plan_version"), "blocks", eng.get_option("plan_blocks"), "cells_max",
          eng.get_option("plan_cells_max"), "contrib", eng.get_option("plan_contributions"), "patterns",
          eng.get_option("plan_patterns"), "shared contrib", eng.get_option("plan_shared_contributions"),
          "vmax", eng.get_option("plan_vertices_max"), "residual tiles", eng.get_option("residual_tiles"),
          "patterns", eng.get_option("residual_patterns"), flush=True)
x, xd = u.data, udot.data
if a.debug:
    eng.set_option("debug", a.debug)  # needs the development build (FDTS_B200_LIB=.../libfdts_b200_dev.so)
if a.dmma >= 0:
    eng.set_option("dmma", a.dmma)
F = torch.empty(eng.num_rows, dtype=torch.float64, device="cuda")
J = eng.new_values()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for r in range(a.reps):
    ev[0].record()
    eng.ifunction(0.0, x, xd, out=F)
    ev[1].record()
    eng.ijacobian(0.0, x, xd, 10.0, out=J)
    ev[2].record()
    torch.cuda.synchronize()
    tf, tj = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"rep {r}: ifunction {tf:.3f} ms  ijacobian {tj:.3f} ms  pair {tf+tj:.3f} ms  "
          f"{form.space.global_num_dofs/((tf+tj)*1e-3)/1e9:.3f} GDOF/s", flush=True)
print("F norm", float(F.norm()), "J sum", float(J.sum()))
