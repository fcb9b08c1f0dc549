"""examples/heat.py of the reference on the B200 engine (reference examples/heat.py:5-47): P1 heat equation on
UnitIntervalMesh(10), bump initial state, homogeneous Dirichlet condition, default implicit TS to T = 1, run
twice like the reference does.  (The reference's NonlinearVariationalSolver warm-up, lines 22-30, is plain
Firedrake and has no counterpart here.)"""
import os

import torch

import firedrake_ts_b200 as firedrake_ts
from firedrake_ts_b200 import DirichletBC, Function, FunctionSpace, UnitIntervalMesh, forms

n = int(os.environ.get("FDTS_EXAMPLE_N", 10))
mesh = UnitIntervalMesh(n)
V = FunctionSpace(mesh, 1)

u = Function(V)
u_t = Function(V)
F = forms.heat(u, u_t, source=1.0)  # inner(u_t, v)*dx + inner(grad(u), grad(v))*dx - 1.0*v*dx

bc = DirichletBC(V, 0.0, "on_boundary")
u.interpolate(lambda x: (x[:, 0] < 0.5).double())
print(f"u0 = {u.data.cpu().numpy()}")


def monitor(ts, step, t, x):
    print(f"step {step}: t = {t:.3f}  max u = {float(u.data.max()):.6f}")


problem = firedrake_ts.DAEProblem(F, u, u_t, (0.0, 1.0), bcs=bc)
solver = firedrake_ts.DAESolver(problem, options_prefix="", monitor_callback=monitor)
solver.solve()
solver.solve()  # a second solve restarts from the current state, as in the reference (lines 40-44)

print(solver.ts.getTime())
x = V.node_coords[:, 0].to(u.data.device)
steady = 0.5 * x * (1.0 - x)  # -u'' = 1, u(0) = u(1) = 0
print(u.data.cpu().numpy())
assert float((u.data - steady).abs().max()) < 1e-3, "the heat equation did not reach its steady state"
