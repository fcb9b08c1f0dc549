"""The IMEX branch F(u_t, u, t) = G(u, t): DAEProblem(G=...), _TSContext.form_rhs_function /
form_rhs_jacobian with the mass-matrix projection (reference firedrake_ts/solving_utils.py:320-384,
429-458; examples/heat-explicit.py), driven on the CPU through the oracle assembler, and the
stand-in's additive Runge-Kutta integrator (PETSc's TSARKIMEX formulation)."""
import math

import numpy as np
import pytest
import torch

import firedrake_ts_b200 as fts
from firedrake_ts_b200 import forms
from firedrake_ts_b200.ts import ARKIMEX_TABLEAUS, TS, Mat, Vec
from helpers import OracleAssembler
from oracle import Oracle


def heat_explicit_1d(n=10, degree=1):
    """examples/heat-explicit.py:5-21 (node sets instead of boundary ids 1 and 2)."""
    mesh = fts.UnitIntervalMesh(n)
    V = fts.FunctionSpace(mesh, degree)
    u, u_t = fts.Function(V), fts.Function(V)
    F, G = forms.heat_explicit(u, u_t)
    x = V.node_coords[:, 0]
    left = torch.nonzero(x < 1e-12).flatten().to(torch.int32)
    right = torch.nonzero(x > 1 - 1e-12).flatten().to(torch.int32)
    bcs = [fts.DirichletBC(V, 1.0, left), fts.DirichletBC(V, 0.0, right)]
    u.interpolate(lambda x: (x[:, 0] < 0.5).double())
    return V, u, u_t, F, G, bcs


def test_diffusion_family_reduces_to_heat():
    """a = k = s = 1 is examples/heat.py:11; F and G of the split add up to it (linearity)."""
    mesh = fts.UnitSquareMesh(5, 4)
    V = fts.FunctionSpace(mesh, 2)
    u, u_t = fts.Function(V), fts.Function(V)
    rng = np.random.default_rng(3)
    x, xd = rng.standard_normal(V.num_dofs), rng.standard_normal(V.num_dofs)
    heat = Oracle.from_form(forms.heat(u, u_t), [])
    full = Oracle.from_form(forms.diffusion(u, u_t, 1.0, 1.0, 1.0), [])
    F, G = forms.heat_explicit(u, u_t)
    oF, oG = Oracle.from_form(F, []), Oracle.from_form(G, [])
    assert np.abs(full.ifunction(0.0, x, xd) - heat.ifunction(0.0, x, xd)).max() < 1e-14
    assert np.abs(full.ijacobian(0.0, x, xd, 2.5) - heat.ijacobian(0.0, x, xd, 2.5)).max() < 1e-14
    # F - G == heat residual, shift*dF/dudot - dG/du == heat Jacobian
    assert np.abs(oF.ifunction(0.0, x, xd) - oG.ifunction(0.0, x, xd) - heat.ifunction(0.0, x, xd)).max() < 1e-13
    assert np.abs(oF.ijacobian(0.0, x, xd, 2.5) - oG.ijacobian(0.0, x, xd, 0.0)
                  - heat.ijacobian(0.0, x, xd, 2.5)).max() < 1e-13
    # G does not depend on udot, F does not depend on u
    assert np.array_equal(oG.ifunction(0.0, x, xd), oG.ifunction(0.0, x, 0 * xd))
    assert np.array_equal(oF.ifunction(0.0, x, xd), oF.ifunction(0.0, 0 * x, xd))


def test_problem_with_G_defaults_to_arkimex_and_checks_args():
    V, u, u_t, F, G, bcs = heat_explicit_1d()
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.01), bcs=bcs, G=G)
    assert problem.dGdu is not None and len(problem.dGdu.arguments()) == 2
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs))
    assert s.ts.type == "arkimex"  # reference ts_solver.py:252-253
    assert s.ts.equation_type is None  # only set to IMPLICIT without G (reference ts_solver.py:249-250)
    with pytest.raises(TypeError):
        fts.DAEProblem(F, u, u_t, (0.0, 1.0), G="not a form")
    with pytest.raises(ValueError):
        fts.DAEProblem(F, u, u_t, (0.0, 1.0), G=forms.derivative(G, 0.0))
    with pytest.raises(ValueError):
        fts.DAESolver(problem, assembler=OracleAssembler(F, bcs))  # test assembler for F but none for G


def _mkb(V, u, u_t, bcs):
    """dense M, K, load b and the Dirichlet mask from the heat-family oracle (independent of the
    diffusion family): J(shift) = shift M + K, F(0, 0) = -b."""
    o = Oracle.from_form(forms.heat(u, u_t), [])
    z = np.zeros(V.num_dofs)
    rp, ci = o.pattern()
    n = V.num_dofs

    def dense(vals):
        A = np.zeros((n, n))
        for r in range(n):
            A[r, ci[rp[r]:rp[r + 1]]] = vals[rp[r]:rp[r + 1]]
        return A

    K = dense(o.ijacobian(0.0, z, z, 0.0))
    M = dense(o.ijacobian(0.0, z, z, 1.0)) - K
    b = -o.ifunction(0.0, z, z)
    isbc = np.zeros(n, bool)
    for bc in bcs:
        isbc[bc.nodes.numpy()] = True
    return M, K, b, isbc


@pytest.mark.parametrize("project", [True, False])
def test_rhs_callbacks_against_dense_algebra(project):
    """form_rhs_function returns M_bc^-1 G_bc (or G_bc), form_rhs_jacobian returns dG/du with unit
    Dirichlet rows (bcs_G / bcs_dGdu, reference solving_utils.py:101-107)."""
    V, u, u_t, F, G, bcs = heat_explicit_1d(8, 2)
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.01), bcs=bcs, G=G)
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs),
                      project_rhs=project, rhs_projection_parameters={"pc_type": "lu"})
    ctx = s._ctx
    M, K, b, isbc = _mkb(V, u, u_t, bcs)
    rng = np.random.default_rng(5)
    xv = rng.standard_normal(V.num_dofs)
    X, Gv = Vec(torch.from_numpy(xv.copy())), Vec(torch.zeros(V.num_dofs, dtype=torch.float64))
    from firedrake_ts_b200 import dmhooks

    with dmhooks.add_hooks(s.ts.getDM(), s, appctx=ctx):
        ctx.form_rhs_function(s.ts, 0.3, X, Gv)
        ctx.form_rhs_jacobian(s.ts, 0.3, X, ctx._rhs_jac.petscmat, ctx._rhs_pjac.petscmat)
    g = -K @ xv + b
    g[isbc] = 0.0  # zero_bc_nodes=True
    if project:
        Mb = M.copy()
        Mb[isbc, :] = 0.0
        Mb[:, isbc] = 0.0
        Mb[isbc, isbc] = 1.0
        g = np.linalg.solve(Mb, g)
    assert np.abs(Gv.array.numpy() - g).max() < 1e-11 * max(1.0, np.abs(g).max())
    Jd = ctx._rhs_jac.petscmat.to_scipy().toarray()
    Kb = -K.copy()
    Kb[isbc, :] = 0.0
    Kb[:, isbc] = 0.0
    Kb[isbc, isbc] = 1.0
    assert np.abs(Jd - Kb).max() < 1e-11 * np.abs(K).max()
    assert float(ctx._time) == 0.3


def test_heat_explicit_trajectory_matches_hand_rolled_midpoint():
    """examples/heat-explicit.py through DAESolver (ts_type arkimex, ars122) == explicit midpoint on
    M u' = -K u + b written out with dense algebra."""
    V, u, u_t, F, G, bcs = heat_explicit_1d(10, 1)
    u0 = u.data.clone().numpy()
    for bc in bcs:
        u0[bc.nodes.numpy()] = float(bc.value)
    dt, nsteps = 5.0e-4, 40
    seen = []
    problem = fts.DAEProblem(F, u, u_t, (0.0, dt * nsteps), bcs=bcs, G=G)
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs),
                      solver_parameters={"pc_type": "lu", "ts_arkimex_type": "ars122"},
                      rhs_projection_parameters={"pc_type": "lu"},
                      monitor_callback=lambda ts, step, t, x: seen.append((step, t)))
    s.ts.setTimeStep(dt)
    s.solve()
    assert s.ts.getStepNumber() == nsteps and seen[-1][0] == nsteps and abs(seen[-1][1] - dt * nsteps) < 1e-12
    assert s.ts.rhs_evals == 2 * nsteps

    M, K, b, isbc = _mkb(V, u, u_t, bcs)
    free = ~isbc

    def slope(v):
        g = (-K @ v + b)[free]
        out = np.zeros_like(v)
        out[free] = np.linalg.solve(M[np.ix_(free, free)], g)
        return out

    v = u0.copy()
    for _ in range(nsteps):
        mid = v + 0.5 * dt * slope(v)
        v = v + dt * slope(mid)
    assert np.abs(u.data.numpy() - v).max() < 1e-10
    assert np.abs(u.data.numpy()[isbc] - u0[isbc]).max() == 0.0


def test_implicit_F_explicit_source_split():
    """stiff part implicit, source explicit: F = (u_t, v) + (grad u, grad v), G = (1, v); the steady
    state is the one of examples/heat.py:11."""
    mesh = fts.UnitIntervalMesh(10)
    V = fts.FunctionSpace(mesh, 2)
    u, u_t = fts.Function(V), fts.Function(V)
    F = forms.diffusion(u, u_t, mass=1.0, diffusivity=1.0, source=0.0)
    G = forms.diffusion(u, u_t, mass=0.0, diffusivity=0.0, source=-1.0)
    bc = fts.DirichletBC(V, 0.0, "on_boundary")
    problem = fts.DAEProblem(F, u, u_t, (0.0, 3.0), bcs=bc, G=G)
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, [bc]), rhs_assembler=OracleAssembler(G, [bc]),
                      mass_assembler=OracleAssembler(forms.mass_form(F), [bc]),
                      solver_parameters={"pc_type": "lu", "snes_rtol": 1e-12, "snes_atol": 1e-14,
                                         "ts_arkimex_type": "1"},  # L-stable implicit part
                      rhs_projection_parameters={"pc_type": "lu"})
    s.solve()  # dt = 0.1: far beyond the explicit limit of the diffusion term, fine because it is implicit
    x = V.node_coords[:, 0]
    assert (u.data - x * (1 - x) / 2).abs().max() < 1e-6


@pytest.mark.parametrize("name,order", [("1", 1), ("ars122", 2)])
def test_arkimex_order_on_scalar_ode(name, order):
    """u' = -lam u + cos t with the decay implicit and the forcing explicit: observed order of the
    tableau (protects the coefficients)."""
    lam = 3.0
    assert name in ARKIMEX_TABLEAUS

    def run(nsteps):
        ts = TS()
        J = Mat(torch.tensor([0, 1]), torch.tensor([0], dtype=torch.int32), torch.zeros(1, dtype=torch.float64), 1)

        def ifunc(ts_, t, X, Xdot, Fv):
            Fv.array.copy_(Xdot.array + lam * X.array)

        def ijac(ts_, t, X, Xdot, shift, Jm, Pm):
            Jm.values[0] = shift + lam

        def rhs(ts_, t, X, Gv):
            Gv.array[0] = math.cos(t)

        ts.setIFunction(ifunc, Vec(torch.zeros(1, dtype=torch.float64)))
        ts.setIJacobian(ijac, J, J)
        ts.setRHSFunction(rhs, Vec(torch.zeros(1, dtype=torch.float64)))
        ts.setFromOptions({"ts_type": "arkimex", "ts_arkimex_type": name, "pc_type": "lu",
                           "snes_rtol": 1e-14, "snes_atol": 1e-15})
        ts.setMaxTime(1.0)
        ts.setTimeStep(1.0 / nsteps)
        uv = Vec(torch.ones(1, dtype=torch.float64))
        ts.solve(uv)
        assert ts.getStepNumber() == nsteps
        return float(uv.array[0])

    exact = (1 - lam / (1 + lam ** 2)) * math.exp(-lam) + (lam * math.cos(1.0) + math.sin(1.0)) / (1 + lam ** 2)
    e1, e2 = abs(run(40) - exact), abs(run(80) - exact)
    assert e2 < e1 and abs(math.log2(e1 / e2) - order) < 0.25, (e1, e2)


def test_implicit_ts_type_with_G_is_rejected():
    V, u, u_t, F, G, bcs = heat_explicit_1d()
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.01), bcs=bcs, G=G)
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs),
                      solver_parameters={"ts_type": "beuler", "pc_type": "lu"})
    with pytest.raises(NotImplementedError):
        s.solve()


def test_pcg_device_matches_pcg():
    """the synchronisation-free PCG (device scalars, periodic convergence check) follows the same
    recurrence as the reference PCG."""
    from firedrake_ts_b200 import linalg

    rng = np.random.default_rng(2)
    n = 60
    Q = rng.standard_normal((n, n))
    A = torch.from_numpy(Q @ Q.T + n * np.eye(n))
    b = torch.from_numpy(rng.standard_normal(n))
    dinv = 1.0 / torch.diagonal(A)
    mv, pc = (lambda v: A @ v), (lambda r: r * dinv)
    x1, it1 = linalg.pcg(mv, b, pc, lambda a, c: float(torch.dot(a, c)), 1e-12, 1e-50, 500)
    x2, it2 = linalg.pcg_device(mv, b, pc, 1e-12, 1e-50, 500, check_every=4)
    assert it1 <= it2 < it1 + 4
    assert (A @ x2 - b).norm() <= 1e-12 * b.norm()
    assert (x1 - x2).abs().max() < 1e-12
