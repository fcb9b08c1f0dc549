"""examples/heat-explicit.py of the reference on the B200 engine (reference examples/heat-explicit.py:5-36): the
IMEX split F(u_t, u, t) = G(u, t) with F = (u_t, v) and G = -((grad u, grad v) - (1, v)), Dirichlet values 1 and 0
at the two ends, ARKIMEX by default (ts_solver.py:252-253); G is projected with the mass matrix as in
solving_utils.py:447-458."""
import os

import torch

import firedrake_ts_b200 as firedrake_ts
from firedrake_ts_b200 import DirichletBC, Function, FunctionSpace, UnitIntervalMesh, forms

n = int(os.environ.get("FDTS_EXAMPLE_N", 10))
mesh = UnitIntervalMesh(n)
V = FunctionSpace(mesh, 1)

u = Function(V)
u_t = Function(V)
F, G = forms.heat_explicit(u, u_t, source=1.0)

x = V.node_coords[:, 0]
left = torch.nonzero(x == 0.0).reshape(-1).to(torch.int32)
right = torch.nonzero(x == 1.0).reshape(-1).to(torch.int32)
bcs = [DirichletBC(V, 1.0, left), DirichletBC(V, 0.0, right)]  # subdomains 1 and 2 of the interval

u.interpolate(lambda x: (x[:, 0] < 0.5).double())
print(f"u0 = {u.data.cpu().numpy()}")


def monitor(ts, step, t, x):
    if step % 200 == 0:
        print(f"step {step}: t = {t:.4f}")


problem = firedrake_ts.DAEProblem(F, u, u_t, (0.0, 0.4), bcs=bcs, G=G)
solver = firedrake_ts.DAESolver(problem, options_prefix="", monitor_callback=monitor)
solver.ts.setTimeStep(0.1 / n ** 2)  # the stand-in ARKIMEX is fixed-step; PETSc's adapts the step
solver.solve()
print(u.data.cpu().numpy())
assert torch.isfinite(u.data).all() and abs(float(u.data[left[0].item()]) - 1.0) < 1e-12
