"""examples/cahn-hilliard.py of the reference on the B200 engine (reference examples/cahn-hilliard.py:11-117): mixed
Q1 x Q1 Cahn-Hilliard on UnitSquareMesh(96, 96, quadrilateral=True), lambda = 1e-2, dt = 5e-6, two steps from the
seeded random concentration; Newton tolerances as in the reference; direct solve ("pc_type": "lu", the
reference's default branch).  The Schur-complement fieldsplit branch (lines 73-112) needs a user Schur
preconditioner and is out of scope; `pc_type fieldsplit` with additive / multiplicative splits is available."""
import os

import numpy as np
import torch

import firedrake_ts_b200 as firedrake_ts
from firedrake_ts_b200 import Function, FunctionSpace, UnitSquareMesh, forms

lmbda = 1.0e-02  # surface parameter
dt = 5.0e-06  # time step
n = int(os.environ.get("FDTS_EXAMPLE_N", 96))

mesh = UnitSquareMesh(n, n, quadrilateral=True)
ME = FunctionSpace(mesh, 1, ncomp=2, layout="blocked")  # V * V

u = Function(ME)  # current solution
u_t = Function(ME)  # time derivative
F = forms.cahn_hilliard(u, u_t, lmbda=lmbda)

rng = np.random.default_rng(11)
u.sub(0).copy_(torch.from_numpy(0.63 + 0.2 * (0.5 - rng.random(ME.num_nodes))))

params = {"mat_type": "aij", "pc_type": "lu", "ksp_type": "preonly", "snes_rtol": 1e-9, "snes_atol": 1e-10,
          "snes_stol": 1e-16, "ksp_rtol": 1e-6, "ksp_atol": 1e-15}


def monitor(ts, step, t, x):
    cdat = u.sub(0)
    print(f"step {step}: t = {t:.2e}  c in [{float(cdat.min()):.4f}, {float(cdat.max()):.4f}]  mean {float(cdat.mean()):.6f}")


problem = firedrake_ts.DAEProblem(F, u, u_t, (0.0, 2 * dt))
solver = firedrake_ts.DAESolver(problem, solver_parameters=params, monitor_callback=monitor)
solver.ts.setTimeStep(dt)
solver.solve()
assert solver.ts.getStepNumber() == 2 and torch.isfinite(u.data).all()
