"""GPU: the drop-in DAESolver on the B200 engine reproduces oracle-driven trajectories (same time
stepper, same callbacks) to 1e-10 at final time (BASELINE.json north_star), keeps two solvers
apart, and satisfies size-independent properties at the headline size."""
import math

import numpy as np
import pytest
import torch

import firedrake_ts_b200 as fts
from firedrake_ts_b200 import forms
from firedrake_ts_b200.workloads import make_problem
from helpers import OracleAssembler

pytestmark = pytest.mark.gpu
TRAJ_TOL = 1e-10


def run_both(build, tspan, params, dt=None, scatter=None):
    finals = []
    for which in ("oracle", "b200"):
        form, bcs, u, udot = build()
        problem = fts.DAEProblem(form, u, udot, tspan, bcs=bcs or None)
        kw = dict(solver_parameters=dict(params))
        if which == "oracle":
            kw["assembler"] = OracleAssembler(form, bcs)
        elif scatter is not None:
            kw["appctx"] = {"scatter": scatter}
        solver = fts.DAESolver(problem, **kw)
        if dt:
            solver.ts.setTimeStep(dt)
        solver.solve()
        finals.append((u.data.detach().cpu().clone(), solver.ts.getStepNumber(), solver))
    return finals


DIRECT = {"pc_type": "lu", "snes_rtol": 1e-13, "snes_atol": 1e-15, "snes_stol": 0.0, "snes_max_it": 30}


@pytest.mark.parametrize("scatter", [0, 1, 2])
def test_heat_trajectory(scatter):
    def build():
        form, bcs, u, udot = make_problem("heat", 2, 2, 8, tile=2, ftile=6 if scatter == 2 else None)
        udot.data.zero_()
        return form, bcs, u, udot

    (uo, so, _), (ug, sg, s) = run_both(build, (0.0, 0.5), DIRECT, scatter=scatter)
    assert so == sg == 5
    assert (uo - ug).abs().max() < TRAJ_TOL
    assert s._ctx.engine.launch_count > 0 and s._ctx._x.data.is_cuda


def test_burgers_trajectory():
    """examples/burgers.py: vector CG2, nu = 1e-4, implicit default (BEULER), dt = 1/n."""
    n = 6

    def build():
        msh = fts.UnitSquareMesh(n, n)
        V = fts.FunctionSpace(msh, 2, ncomp=2, layout="interleaved")
        u, u_t = fts.Function(V, name="Velocity"), fts.Function(V, name="VelocityTimeDerivative")
        x = V.node_coords
        ic = torch.zeros(V.num_nodes, 2, dtype=torch.float64)
        ic[:, 0] = torch.sin(math.pi * x[:, 0])
        u.data.copy_(ic.reshape(-1))
        return forms.burgers(u, u_t, nu=1e-4), [], u, u_t

    (uo, so, _), (ug, sg, _) = run_both(build, (0.0, 0.5), DIRECT, dt=1.0 / n)
    assert so == sg == 3
    assert (uo - ug).abs().max() < TRAJ_TOL


def test_cahn_hilliard_trajectory():
    def build():
        msh = fts.UnitSquareMesh(8, 8)
        ME = fts.FunctionSpace(msh, 1, ncomp=2, layout="blocked")
        u, u_t = fts.Function(ME), fts.Function(ME)
        rng = np.random.default_rng(11)
        u.sub(0).copy_(torch.from_numpy(0.63 + 0.2 * (0.5 - rng.random(ME.num_nodes))))
        return forms.cahn_hilliard(u, u_t, lmbda=1.0e-2), [], u, u_t

    dt = 5.0e-6
    # Newton tolerances of examples/cahn-hilliard.py:50-61 (the residual's scale makes 1e-13 unreachable)
    params = {"pc_type": "lu", "snes_rtol": 1e-9, "snes_atol": 1e-10, "snes_stol": 1e-16, "snes_max_it": 30}
    (uo, so, _), (ug, sg, _) = run_both(build, (0.0, 2 * dt), params, dt=dt)
    assert so == sg == 2
    assert (uo - ug).abs().max() < TRAJ_TOL * max(1.0, float(uo.abs().max()))


def test_cahn_hilliard_literal_quadrilateral_trajectory():
    """examples/cahn-hilliard.py:14-16 as written: UnitSquareMesh(n, n, quadrilateral=True), Q1 x Q1."""
    def build():
        msh = fts.UnitSquareMesh(10, 10, quadrilateral=True)
        ME = fts.FunctionSpace(msh, 1, ncomp=2, layout="blocked")
        u, u_t = fts.Function(ME), fts.Function(ME)
        rng = np.random.default_rng(11)
        u.sub(0).copy_(torch.from_numpy(0.63 + 0.2 * (0.5 - rng.random(ME.num_nodes))))
        return forms.cahn_hilliard(u, u_t, lmbda=1.0e-2), [], u, u_t

    dt = 5.0e-6
    params = {"pc_type": "lu", "snes_rtol": 1e-9, "snes_atol": 1e-10, "snes_stol": 1e-16, "snes_max_it": 30}
    (uo, so, _), (ug, sg, s) = run_both(build, (0.0, 2 * dt), params, dt=dt)
    assert so == sg == 2
    assert (uo - ug).abs().max() < TRAJ_TOL * max(1.0, float(uo.abs().max()))
    assert s._ctx.engine.launch_count > 0


def test_bbm_trajectory():
    def build():
        msh = fts.PeriodicIntervalMesh(100, 100.0)
        V = fts.FunctionSpace(msh, 1)
        u, u_t = fts.Function(V), fts.Function(V)
        c = 0.5
        arg = 0.5 * (c * V.node_coords[:, 0] - c * 30.0)
        u.data.copy_(3 * c ** 2 / (1 - c ** 2) * (2 / (torch.exp(arg) + torch.exp(-arg))) ** 2)
        return forms.bbm(u, u_t), [], u, u_t

    params = dict(DIRECT, ts_type="theta", ts_theta_theta=0.5, ts_exact_final_time="matchstep")
    (uo, so, _), (ug, sg, _) = run_both(build, (0.0, 3.0), params, dt=1.0)
    assert so == sg == 3
    assert (uo - ug).abs().max() < TRAJ_TOL


def test_device_krylov_solve_matches_direct():
    """Jacobi-GMRES on the device-resident CSR (fdts_csr_spmv / fdts_csr_diagonal)."""
    def build():
        form, bcs, u, udot = make_problem("heat", 3, 1, 6, tile=2)
        udot.data.zero_()
        return form, bcs, u, udot

    form, bcs, u, udot = build()
    p = fts.DAEProblem(form, u, udot, (0.0, 0.2), bcs=bcs)
    fts.DAESolver(p, solver_parameters={"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-13,
                                        "snes_rtol": 1e-12}).solve()
    form2, bcs2, u2, udot2 = build()
    p2 = fts.DAEProblem(form2, u2, udot2, (0.0, 0.2), bcs=bcs2)
    fts.DAESolver(p2, assembler=OracleAssembler(form2, bcs2), solver_parameters=DIRECT).solve()
    assert (u.data.cpu() - u2.data).abs().max() < 1e-10


def test_two_solvers_on_gpu():
    sols = []
    for _ in range(2):
        form, bcs, u, udot = make_problem("heat", 1, 1, 10)
        u.interpolate(lambda x: (torch.abs(x[:, 0] - 0.5) < 0.1).double())
        udot.data.zero_()
        sols.append((fts.DAESolver(fts.DAEProblem(form, u, udot, (0.0, 1.0), bcs=bcs), solver_parameters=DIRECT), u))
    sols[0][0].solve()
    sols[1][0].solve()
    assert torch.equal(sols[0][1].data, sols[1][1].data)
    assert sols[0][0]._ctx.engine is not sols[1][0]._ctx.engine


@pytest.mark.parametrize("n,degree", [(184, 2)])
def test_headline_size_properties(n, degree):
    """C4 at full size (37.4 M cells, 50.2 M DOFs): counts of SURVEY.md section 8, and properties that do
    not need an oracle at this size: sum(J) = shift*|Omega| without BCs (stiffness annihilates
    constants, the mass matrix sums to the volume), sum(F) = int udot_h - |Omega| for u = const,
    run-to-run bitwise reproducibility of the gather path, unit Dirichlet rows with BCs."""
    from firedrake_ts_b200.engine import Engine

    form, bcs, u, udot = make_problem("heat", 3, degree, n, tile=3, ftile=6, device="cuda:0")
    V = form.space
    assert V.mesh.num_cells == 37_377_024 and V.num_dofs == 50_243_409
    eng = Engine(form, (), device="cuda:0", scatter=2)
    assert eng.nnz == 1_437_462_465 and eng.max_row_nodes == 65
    shift = 10.0
    J = eng.ijacobian(0.0, u.data, udot.data, shift)
    assert abs(float(J.sum()) - shift) < 1e-8
    J2 = eng.ijacobian(0.0, u.data, udot.data, shift)
    assert torch.equal(J, J2)
    eng.set_option("scatter", 1)
    J3 = eng.ijacobian(0.0, u.data, udot.data, shift)  # atomic path: same values up to summation order
    assert float((J - J3).abs().max()) < 1e-12 * float(J.abs().max())
    del J2, J3
    u.data.fill_(3.0)
    udot.data.fill_(2.0)
    F = eng.ifunction(0.0, u.data, udot.data)
    assert abs(float(F.sum()) - (2.0 - 1.0)) < 1e-9
    del eng, J, F
    torch.cuda.empty_cache()
    eng = Engine(form, bcs, device="cuda:0", scatter=2)
    J = eng.ijacobian(0.0, u.data, udot.data, shift)
    rp, ci = eng.csr()
    b = bcs[0].nodes.long().cuda()
    # Dirichlet rows: exactly one non-zero (1.0, on the diagonal)
    rows = torch.repeat_interleave(torch.arange(eng.num_rows, device="cuda"), rp[1:] - rp[:-1])
    isb = torch.zeros(eng.num_rows, dtype=torch.bool, device="cuda")
    isb[b] = True
    sel = isb[rows]
    assert float(J[sel].sum()) == float(b.numel())
    assert float(J[sel].abs().max()) == 1.0
    assert float(J[isb[ci.long()] & ~sel].abs().max()) == 0.0  # Dirichlet columns dropped
    F = eng.ifunction(0.0, u.data, udot.data)
    assert float(F[b].abs().max()) == 0.0


@pytest.mark.parametrize("arktype,dim,n,scatter", [("ars122", 1, 10, None), ("1", 2, 6, None), ("ars122", 2, 6, 2)])
def test_heat_explicit_imex_trajectory(arktype, dim, n, scatter):
    """examples/heat-explicit.py (F = (u_t, v), G = -((grad u, grad v) - (1, v)), ts_type arkimex by
    default) on the B200 engine vs the same stepper driven by the oracle: form_rhs_function with the
    mass-matrix projection (CG + Jacobi on the device CSR vs LU on the CPU), reference
    solving_utils.py:320-347, 429-458."""
    dt, nsteps = (5.0e-4, 20) if dim == 1 else (2.0e-4, 10)
    finals = []
    for which in ("oracle", "b200"):
        if dim == 1:
            mesh = fts.UnitIntervalMesh(n)
        else:
            mesh = fts.UnitSquareMesh(n, n)
        V = fts.FunctionSpace(mesh, 1 if dim == 1 else 2)
        u, u_t = fts.Function(V), fts.Function(V)
        F, G = forms.heat_explicit(u, u_t)
        bcs = [fts.DirichletBC(V, 0.0, "on_boundary")]
        u.interpolate(lambda x: torch.prod(torch.sin(math.pi * x), dim=1) + (x[:, 0] < 0.5).double())
        problem = fts.DAEProblem(F, u, u_t, (0.0, dt * nsteps), bcs=bcs, G=G)
        kw = dict(solver_parameters={"pc_type": "lu", "ts_arkimex_type": arktype})
        if which == "oracle":
            kw.update(assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs),
                      rhs_projection_parameters={"pc_type": "lu"})
        else:
            kw.update(rhs_projection_parameters={"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-14})
            if scatter is not None:  # owner-computes gather for F and for G (dG/du through the scaling wrapper)
                kw["appctx"] = {"scatter": scatter}
        solver = fts.DAESolver(problem, **kw)
        if which == "b200" and scatter is not None:
            ctx = solver._ctx
            assert ctx.rhs_engine.get_option("scatter") == scatter and ctx.rhs_engine.get_option("plan_blocks") > 0
            from firedrake_ts_b200 import dmhooks
            from firedrake_ts_b200.ts import Vec

            with dmhooks.add_hooks(solver.ts.getDM(), solver, appctx=ctx):
                ctx.form_rhs_jacobian(solver.ts, 0.0, Vec(u.owned().clone().to(ctx._device)), ctx._rhs_jac.petscmat,
                                      ctx._rhs_pjac.petscmat)
            Jg = ctx._rhs_jac.petscmat.values.cpu().numpy()
            Jo = OracleAssembler(G, bcs).ijacobian(0.0, u.data.cpu(), u_t.data.cpu(), 0.0).numpy()
            assert np.abs(Jg - Jo).max() < 1e-12 * np.abs(Jo).max()
        assert solver.ts.type == "arkimex"
        solver.ts.setTimeStep(dt)
        solver.solve()
        assert solver.ts.getStepNumber() == nsteps
        finals.append((u.data.detach().cpu().clone(), solver))
    (uo, _), (ug, s) = finals
    assert (uo - ug).abs().max() < TRAJ_TOL
    assert s._ctx.rhs_engine.launch_count > 0 and s._ctx._projected_G.array.is_cuda
    # G is not the residual of a steady problem: the state really moved
    assert (ug - s._problem.u_restrict.function_space().interpolate(
        lambda x: torch.prod(torch.sin(math.pi * x), dim=1) + (x[:, 0] < 0.5).double())).abs().max() > 1e-3


@pytest.mark.parametrize("family,dim,degree,n", [("heat", 3, 2, 7), ("heat", 1, 1, 300), ("burgers", 2, 2, 9),
                                                 ("cahn_hilliard", 2, 1, 11), ("bbm", 3, 3, 3)])
def test_csr_spmv_kernels(family, dim, degree, n):
    """fdts_csr_spmv (CSR-stream) and fdts_csr_spmv_warp (warp per row) against scipy on assembled
    Jacobians: short rows, long rows (P3, mixed), row counts that do not fill the last CTA."""
    import scipy.sparse as sp
    from firedrake_ts_b200 import linalg
    from firedrake_ts_b200.engine import Engine

    form, bcs, u, udot = make_problem(family, dim, degree, n)
    eng = Engine(form, bcs, device="cuda:0")
    rp, ci = eng.csr()
    J = eng.ijacobian(0.0, u.data.cuda(), udot.data.cuda(), 3.0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(form.space.num_dofs, generator=g, dtype=torch.float64)
    A = sp.csr_matrix((J.cpu().numpy(), ci.cpu().numpy(), rp.cpu().numpy()), shape=(rp.numel() - 1, x.numel()))
    ref = A @ x.numpy()
    scale = np.abs(A).dot(np.abs(x.numpy())).max()
    for kernel in ("stream", "warp"):
        y = linalg.csr_spmv(rp, ci, J, x.cuda(), kernel=kernel).cpu().numpy()
        assert np.abs(y - ref).max() < 1e-13 * scale, kernel


def test_split_subcontexts_on_gpu():
    """_TSContext.split on the B200 engine: rows of F / diagonal blocks of J of each field of the mixed
    Cahn-Hilliard space against the oracle (reference solving_utils.py:145-233)."""
    import scipy.sparse as sp

    from firedrake_ts_b200.ts import DM, TS, Vec
    from oracle import Oracle

    form, bcs, u, udot = make_problem("cahn_hilliard", 2, 1, 7, quadrilateral=True)
    orc = Oracle.from_form(form, bcs)
    x0, xd0 = u.data.clone(), udot.data.clone()
    Ffull = orc.ifunction(0.0, x0, xd0)
    rp, ci = orc.pattern()
    nn = form.space.num_nodes
    A = sp.csr_matrix((orc.ijacobian(0.0, x0, xd0, 3.0), ci, rp), shape=(2 * nn, 2 * nn)).toarray()
    solver = fts.DAESolver(fts.DAEProblem(form, u, udot, (0.0, 1e-5)), solver_parameters={"ts_dt": 5e-6})
    ctx = solver._ctx
    for m, sc in enumerate(ctx.split(((0,), (1,)))):
        ts, dm = TS(), DM()
        dm._appctx.append(sc)
        ts.setDM(dm)
        X, Xd = Vec(x0[m * nn:(m + 1) * nn].cuda()), Vec(xd0[m * nn:(m + 1) * nn].cuda())
        Fv = Vec(torch.zeros(nn, dtype=torch.float64, device="cuda"))
        sc.form_function(ts, 0.0, X, Xd, Fv)
        assert np.abs(Fv.array.cpu().numpy() - Ffull[m * nn:(m + 1) * nn]).max() < 1e-12 * np.abs(Ffull).max()
        sc.form_jacobian(ts, 0.0, X, Xd, 3.0, sc._jac.petscmat, sc._pjac.petscmat)
        B = sc._jac.petscmat.to_scipy().toarray()
        assert np.abs(B - A[m * nn:(m + 1) * nn, m * nn:(m + 1) * nn]).max() < 1e-12 * np.abs(A).max()
        assert sc._jac.petscmat.values.is_cuda


def test_fieldsplit_solve_on_gpu():
    def build():
        form, bcs, u, udot = make_problem("cahn_hilliard", 2, 1, 8, quadrilateral=True)
        u.data[form.space.num_nodes:] = 0.0
        return form, bcs, u, udot

    dt = 5.0e-6
    params = {"snes_rtol": 1e-10, "snes_atol": 1e-11, "ksp_type": "gmres", "ksp_rtol": 1e-12,
              "pc_type": "fieldsplit", "pc_fieldsplit_type": "multiplicative", "fieldsplit_0_pc_type": "lu",
              "fieldsplit_1_pc_type": "lu"}
    (uo, so, _), (ug, sg, s) = run_both(build, (0.0, 2 * dt), params, dt=dt)
    assert so == sg == 2
    assert (uo - ug).abs().max() < 1e-8 * max(1.0, float(uo.abs().max()))
