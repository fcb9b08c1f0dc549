"""Timing of the IMEX callbacks (reference solving_utils.py:320-384) on one GPU through the public
API: form_rhs_function = G assembly + mass-matrix projection (CG + Jacobi on the device CSR),
form_rhs_jacobian = dG/du assembly, and one ARKIMEX step of examples/heat-explicit.py.  Prints one
JSON line.  Development / profiles aid; bench.py stays the contract for the headline pair."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import firedrake_ts_b200 as fts
from firedrake_ts_b200 import dmhooks, forms
from firedrake_ts_b200.ts import Vec

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=96)
ap.add_argument("--degree", type=int, default=2)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--scatter", type=int, default=None)
a = ap.parse_args()

mesh = fts.UnitCubeMesh(a.n, tile=3, device="cuda")
V = fts.FunctionSpace(mesh, a.degree, tile=6 if a.degree == 2 else None)
u, u_t = fts.Function(V), fts.Function(V)
F, G = forms.heat_explicit(u, u_t)
bc = fts.DirichletBC(V, 0.0, "on_boundary")
u.interpolate(lambda x: torch.prod(torch.sin(3.141592653589793 * x), dim=1))
problem = fts.DAEProblem(F, u, u_t, (0.0, 1.0), bcs=bc, G=G)
solver = fts.DAESolver(problem, solver_parameters={"ksp_type": "cg", "pc_type": "jacobi"},
                       rhs_projection_parameters={"ksp_rtol": a.rtol},
                       appctx={"scatter": a.scatter} if a.scatter is not None else None)
ctx, ts = solver._ctx, solver.ts
X = Vec(u.owned().clone().to(ctx._device))
Gv = Vec(torch.zeros_like(X.array))


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with dmhooks.add_hooks(ts.getDM(), solver, appctx=ctx):
    ms_G = timed(lambda: ctx._assemble_rhs_residual(ctx._G_vec), a.reps)
    ms_rhs = timed(lambda: ctx.form_rhs_function(ts, 0.0, X, Gv), a.reps)
    its = ctx._rhs_projection_solver.ksp.its

    def rhs_jac():
        ctx._rhs_jacobian_assembled = False
        ctx.form_rhs_jacobian(ts, 0.0, X, ctx._rhs_jac.petscmat, ctx._rhs_pjac.petscmat)

    ms_jac = timed(rhs_jac, a.reps)
    M = ctx._rhs_projection_solver.mass_matrix
    ms_spmv = timed(lambda: M.mult(X.array), a.reps)
    from firedrake_ts_b200 import linalg

    ms_spmv_warp = timed(lambda: linalg.csr_spmv(M.rowptr, M.colidx, M.values, X.array, kernel="warp"), a.reps)
    # residual of the projection: || M p - G || / || G ||
    res = float((M.mult(ctx._projected_G.array) - ctx._G_vec.array).norm() / ctx._G_vec.array.norm())

# one full ARKIMEX (ars122) step through DAESolver.solve
h = 0.05 / (a.n * a.degree) ** 2
problem.tspan = (0.0, 2 * h)
ts.setTime(0.0)
ts.setMaxTime(2 * h)
ts.setTimeStep(h)
solver.solve()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts.setTime(0.0)
e0.record()
solver.solve()
e1.record()
torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / max(1, ts.getStepNumber())
nnz = int(M.values.numel())
print(json.dumps({
    "scatter": a.scatter, "workload": f"heat-explicit IMEX split, 3-D P{a.degree} tets UnitCube n={a.n}", "dofs": int(V.num_dofs), "nnz": nnz,
    "ms_G_assembly": ms_G, "ms_form_rhs_function": ms_rhs, "projection_cg_iterations": int(its),
    "projection_rtol": a.rtol, "projection_residual": res, "ms_form_rhs_jacobian": ms_jac,
    "ms_mass_spmv": ms_spmv, "ms_mass_spmv_warp_per_row": ms_spmv_warp, "spmv_GBps": (12.0 * nnz + 16.0 * V.num_dofs) / ms_spmv * 1e-6,
    "ms_arkimex_ars122_step": ms_step, "steps": int(ts.getStepNumber()),
    "rhs_dof_per_s": V.num_dofs / ms_rhs * 1e3}))
