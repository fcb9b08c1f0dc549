"""A/B of the host-buffer Jacobian read-back on one box: monolithic D2H after the kernel against the
pipelined variant (option pipelined_readback).  Development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from firedrake_ts_b200.engine import Engine
from firedrake_ts_b200.workloads import make_problem

form, bcs, u, udot = make_problem("heat", 3, 2, 184, tile=3, ftile=6, device="cuda:0")
eng = Engine(form, bcs, device="cuda:0", scatter=2)
xh, xdh = u.data.cpu().pin_memory(), udot.data.cpu().pin_memory()
Fh = torch.empty(eng.num_rows, dtype=torch.float64).pin_memory()
Jh = torch.empty(eng.nnz, dtype=torch.float64).pin_memory()
for rep in range(2):
    for mode in (0, 1):
        eng.set_option("pipelined_readback", mode)
        eng.ifunction_host(0.0, xh, xdh, Fh)
        eng.ijacobian_host(0.0, xh, xdh, 10.0, Jh)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):  # alternating, as bench.py's e2e leg does
            eng.ifunction_host(0.0, xh, xdh, Fh)
            eng.ijacobian_host(0.0, xh, xdh, 10.0, Jh)
        torch.cuda.synchronize()
        print(f"pipelined {mode}: alternating pair {(time.perf_counter()-t0)/3*1e3:.1f} ms", flush=True)
for rep in range(2):
    for mode in (0, 1):
        eng.set_option("pipelined_readback", mode)
        eng.ijacobian_host(0.0, xh, xdh, 10.0, Jh)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.ifunction_host(0.0, xh, xdh, Fh)
        t1 = time.perf_counter()
        for _ in range(3):
            eng.ijacobian_host(0.0, xh, xdh, 10.0, Jh)
        t2 = time.perf_counter()
        print(f"pipelined {mode}: ifunction_host {(t1-t0)/3*1e3:.1f} ms  ijacobian_host {(t2-t1)/3*1e3:.1f} ms "
              f"({8*eng.nnz/((t2-t1)/3)/1e9:.1f} GB/s incl. kernel)", flush=True)
