"""examples/burgers.py of the reference on the B200 engine (reference examples/burgers.py:5-48): vector CG2
Burgers equation, nu = 1e-4, u0 = (sin(pi x), 0), implicit default TS with dt = 1/n to T = 0.5.  (VTK output,
lines 27-32, is out of scope: the monitor prints the kinetic energy instead.)"""
import math
import os

import torch

import firedrake_ts_b200 as firedrake_ts
from firedrake_ts_b200 import Function, FunctionSpace, UnitSquareMesh, forms

n = int(os.environ.get("FDTS_EXAMPLE_N", 30))
mesh = UnitSquareMesh(n, n)
V = FunctionSpace(mesh, 2, ncomp=2, layout="interleaved")  # VectorFunctionSpace(mesh, "CG", 2)

u = Function(V, name="Velocity")
u_t = Function(V, name="VelocityTimeDerivative")
ic = torch.zeros(V.num_nodes, 2, dtype=torch.float64)
ic[:, 0] = torch.sin(math.pi * V.node_coords[:, 0])
u.data.copy_(ic.reshape(-1))

nu = 0.0001
F = forms.burgers(u, u_t, nu=nu)  # (inner(u_t, v) + inner(dot(u, nabla_grad(u)), v) + nu*inner(grad(u), grad(v)))*dx


def ts_monitor(ts, steps, time, X):
    print(f"step {steps}: t = {time:.4f}  |u|^2 = {float((u.data ** 2).sum()) / V.num_nodes:.6f}")


problem = firedrake_ts.DAEProblem(F, u, u_t, (0.0, 0.5))
solver = firedrake_ts.DAESolver(problem, monitor_callback=ts_monitor,
                                solver_parameters={"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-10})
solver.ts.setTimeStep(1.0 / n)
solver.solve()
assert torch.isfinite(u.data).all()
