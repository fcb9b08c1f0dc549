"""DAEProblem / DAESolver behaviour (the reference's public API, firedrake_ts/ts_solver.py) driven
on the CPU through the oracle assembler: the reference's own two test scenarios
(tests/test_two_daesolver.py, the examples' soft oracles) and the error / callback contract."""
import math

import numpy as np
import pytest
import torch

import firedrake_ts_b200 as fts
from firedrake_ts_b200 import forms
from helpers import OracleAssembler

LU = {"pc_type": "lu", "snes_rtol": 1e-12, "snes_atol": 1e-14}


def heat_1d(n=10, degree=1):
    mesh = fts.UnitIntervalMesh(n)
    V = fts.FunctionSpace(mesh, degree)
    u, u_t = fts.Function(V), fts.Function(V)
    F = forms.heat(u, u_t)
    bc = fts.DirichletBC(V, 0.0, "on_boundary")
    return V, u, u_t, F, bc


def test_two_ts():
    """reference tests/test_two_daesolver.py:6-40 (issue #4): two solvers coexist."""
    V, u, u_t, F, bc = heat_1d()
    u.interpolate(lambda x: (torch.abs(x[:, 0] - 0.5) < 0.1).double())
    problem = fts.DAEProblem(F, u, u_t, (0.0, 1.0), bcs=bc)
    solver = fts.DAESolver(problem, assembler=OracleAssembler(F, [bc]), solver_parameters=LU)
    V2, u2, u_t2, F2, bc2 = heat_1d()
    problem2 = fts.DAEProblem(F2, u2, u_t2, (0.0, 1.0), bcs=bc2)
    solver2 = fts.DAESolver(problem2, assembler=OracleAssembler(F2, [bc2]), solver_parameters=LU)  # noqa: F841
    solver.solve()
    solver2.solve()
    assert solver.ts.getStepNumber() == 10 and abs(solver.ts.getTime() - 1.0) < 1e-12
    assert solver._ctx is not solver2._ctx and problem.dm is not problem2.dm


def test_heat_relaxes_to_steady_state():
    """examples/heat.py:22-29: the t -> infinity limit of -u'' = 1, u(0)=u(1)=0 is x(1-x)/2."""
    V, u, u_t, F, bc = heat_1d(10, 2)
    problem = fts.DAEProblem(F, u, u_t, (0.0, 3.0), bcs=bc)
    fts.DAESolver(problem, assembler=OracleAssembler(F, [bc]), solver_parameters=LU).solve()
    x = V.node_coords[:, 0]
    assert (u.data - x * (1 - x) / 2).abs().max() < 1e-9


def test_backward_euler_matches_hand_rolled_step():
    """one BE step == (M/dt + K) u1 = M/dt u0 + b on the interior (PETSc theta scheme)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    V, u, u_t, F, bc = heat_1d(8, 2)
    u.interpolate(lambda x: torch.sin(math.pi * x[:, 0]))
    u0 = u.data.clone()
    asm = OracleAssembler(F, [bc])
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.1), bcs=bc)
    s = fts.DAESolver(problem, assembler=asm, solver_parameters=LU)
    s.solve()
    rp, ci = asm.pattern()
    shift = 10.0
    J = sp.csr_matrix((asm.o.ijacobian(0, u0, 0 * u0, shift), ci, rp))
    # F(u1, shift (u1-u0)) = 0 with F linear: J u1 = J u0 - F(u0, 0)
    rhs = J @ u0.numpy() - asm.o.ifunction(0.0, u0, 0 * u0)
    u1 = spla.spsolve(J.tocsc(), rhs)
    assert np.abs(u1 - u.data.numpy()).max() < 1e-12


def test_callbacks_and_monitor_fire():
    V, u, u_t, F, bc = heat_1d()
    calls = {"pre_f": 0, "post_f": 0, "pre_j": 0, "post_j": 0, "mon": 0}
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.3), bcs=bc)

    def post_j(X, Xdot, J):
        calls["post_j"] += 1
        assert J.handle == solver._ctx._jac.petscmat.handle

    solver = fts.DAESolver(
        problem, assembler=OracleAssembler(F, [bc]), solver_parameters=LU,
        pre_function_callback=lambda X, Xd: calls.__setitem__("pre_f", calls["pre_f"] + 1),
        post_function_callback=lambda X, Xd, F_: calls.__setitem__("post_f", calls["post_f"] + 1),
        pre_jacobian_callback=lambda X, Xd: calls.__setitem__("pre_j", calls["pre_j"] + 1),
        post_jacobian_callback=post_j,
        monitor_callback=lambda ts, step, t, x: calls.__setitem__("mon", calls["mon"] + 1))
    solver.solve()
    assert calls["mon"] == 4 and calls["pre_f"] == calls["post_f"] >= 6 and calls["pre_j"] == calls["post_j"] >= 3
    # time / shift Constants are updated by the callbacks (solving_utils.py:253,304)
    assert abs(float(problem.shift) - 10.0) < 1e-12 and abs(float(problem.time) - 0.3) < 1e-12


def test_argument_checking_matches_reference():
    V, u, u_t, F, bc = heat_1d()
    with pytest.raises(TypeError, match="not a Function"):
        fts.DAEProblem(F, u.data, u_t, (0, 1))
    with pytest.raises(TypeError, match="not a Function"):
        fts.DAEProblem(F, u, None, (0, 1))
    with pytest.raises(TypeError, match="not a BaseForm"):
        fts.DAEProblem("F", u, u_t, (0, 1))
    with pytest.raises(ValueError, match="not a linear form"):
        fts.DAEProblem(forms.derivative(F, fts.Constant(1.0)), u, u_t, (0, 1))
    problem = fts.DAEProblem(F, u, u_t, (0, 1), bcs=bc)
    with pytest.raises(TypeError, match="solver_parameters"):
        fts.DAESolver(problem, parameters={})
    assert len(problem.J.arguments()) == 2 and problem.J.shift is problem.shift


def test_time_constant_shared_with_user(capsys=None):
    """examples/heat.py:36: the user's Constant passed as `time` is the one the callbacks assign."""
    V, u, u_t, F, bc = heat_1d()
    tconst = fts.Constant(0.0)
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc, time=tconst)
    fts.DAESolver(problem, assembler=OracleAssembler(F, [bc]), solver_parameters=LU).solve()
    assert problem.time is tconst and abs(float(tconst) - 0.2) < 1e-12


def test_bbm_solitary_wave():
    """examples/bbm.py: P1 on PeriodicIntervalMesh, implicit midpoint; the exact solitary wave
    (bbm.py:27-32) is tracked and the invariant I1 = int u is conserved."""
    N, L = 200, 100.0
    msh = fts.PeriodicIntervalMesh(N, L)
    V = fts.FunctionSpace(msh, 1)
    u, u_t = fts.Function(V), fts.Function(V)
    c, center = 0.5, 30.0

    def uexact(x, t):
        arg = 0.5 * (c * x - c * t / (1 - c ** 2) - c * center)
        return 3 * c ** 2 / (1 - c ** 2) * (2 / (torch.exp(arg) + torch.exp(-arg))) ** 2

    x = V.node_coords[:, 0]
    u.data.copy_(uexact(x, 0.0))
    I1_0 = float(u.data.sum()) * (L / N)
    F = forms.bbm(u, u_t)
    problem = fts.DAEProblem(F, u, u_t, (0.0, 6.0))
    params = dict(LU, ts_type="theta", ts_theta_theta=0.5, ts_exact_final_time="matchstep")
    solver = fts.DAESolver(problem, assembler=OracleAssembler(F, []), solver_parameters=params)
    solver.ts.setTimeStep(0.5)
    solver.solve()
    err = float((u.data - uexact(x, solver.ts.getTime())).norm() / uexact(x, 6.0).norm())
    assert err < 0.02, err
    assert abs(float(u.data.sum()) * (L / N) - I1_0) < 1e-9


def test_cahn_hilliard_two_steps():
    """examples/cahn-hilliard.py with the seeded initial state (rng 11), on triangles (BASELINE configs)."""
    msh = fts.UnitSquareMesh(8, 8)
    ME = fts.FunctionSpace(msh, 1, ncomp=2, layout="blocked")
    u, u_t = fts.Function(ME), fts.Function(ME)
    rng = np.random.default_rng(11)
    u.sub(0).copy_(torch.from_numpy(0.63 + 0.2 * (0.5 - rng.random(ME.num_nodes))))
    mass0 = float(u.sub(0).sum())
    F = forms.cahn_hilliard(u, u_t, lmbda=1.0e-2)
    dt = 5.0e-6
    problem = fts.DAEProblem(F, u, u_t, (0.0, 2 * dt))
    params = {"pc_type": "lu", "snes_rtol": 1e-9, "snes_atol": 1e-10, "snes_stol": 1e-16}
    solver = fts.DAESolver(problem, assembler=OracleAssembler(F, []), solver_parameters=params)
    solver.ts.setTimeStep(dt)
    solver.solve()
    assert solver.ts.getStepNumber() == 2
    assert torch.isfinite(u.data).all() and float(u.sub(1).abs().max()) > 0
    assert abs(float(u.sub(0).sum()) - mass0) < 1.0  # nodal sum stays close (mass is conserved weakly)


def test_krylov_path_agrees_with_direct():
    V, u, u_t, F, bc = heat_1d(16, 2)
    u.interpolate(lambda x: torch.sin(math.pi * x[:, 0]))
    u0 = u.data.clone()
    p1 = fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc)
    fts.DAESolver(p1, assembler=OracleAssembler(F, [bc]), solver_parameters=LU).solve()
    ref = u.data.clone()
    u.data.copy_(u0)
    for ksp in ("gmres", "cg"):
        u.data.copy_(u0)
        p2 = fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc)
        fts.DAESolver(p2, assembler=OracleAssembler(F, [bc]),
                      solver_parameters={"ksp_type": ksp, "pc_type": "jacobi", "ksp_rtol": 1e-13,
                                         "snes_rtol": 1e-12}).solve()
        assert (u.data - ref).abs().max() < 1e-10


def test_user_jacobian_replaces_dF_du_only():
    """reference ts_solver.py:108: J = shift*derivative(F, udot) + (J or derivative(F, u)).  A user J
    built from a different stiffness (an approximate Jacobian) changes the matrix but, being only the
    Newton matrix of a linear problem, not the converged solution."""
    V, u, u_t, F, bc = heat_1d(12, 2)
    u.interpolate(lambda x: torch.sin(math.pi * x[:, 0]))
    u0 = u.data.clone()
    fts.DAESolver(fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc), assembler=OracleAssembler(F, [bc]),
                  solver_parameters=LU).solve()
    ref = u.data.clone()
    u.data.copy_(u0)
    Ku = forms.diffusion(u, u_t, mass=0.0, diffusivity=1.3)  # 30 % too stiff
    Ju = forms.derivative(Ku, fts.Constant(0.0))
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc, J=Ju)
    assert problem.J_du is Ju and len(problem.J.arguments()) == 2
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, [bc]), jac_du_assembler=OracleAssembler(Ku, [bc]),
                      mass_assembler=OracleAssembler(forms.mass_form(F), [bc]),
                      solver_parameters={"pc_type": "lu", "snes_rtol": 1e-14, "snes_atol": 1e-16, "snes_stol": 0.0,
                                         "snes_max_it": 200})
    s.solve()
    assert (u.data - ref).abs().max() < 1e-10
    assert s.ts.newton_its > s.ts.getStepNumber()  # the inexact matrix needs more than one Newton step
    # the assembled matrix is shift*M + 1.3 K with unit Dirichlet diagonal
    A = s._ctx._jac.petscmat.to_scipy().toarray()
    shift = float(problem.shift)
    Mm = OracleAssembler(forms.mass_form(F), [bc]).ijacobian(0.0, u.data, u_t.data, 1.0)
    Kk = OracleAssembler(Ku, [bc]).ijacobian(0.0, u.data, u_t.data, 0.0)
    import scipy.sparse as sp
    rp, ci = OracleAssembler(F, [bc]).pattern()
    want = sp.csr_matrix((shift * Mm.numpy() + Kk.numpy(), ci, rp)).toarray()
    b = bc.nodes.long()
    want[b, b] = 1.0
    assert np.abs(A - want).max() < 1e-12 * np.abs(want).max()
    with pytest.raises(TypeError, match="Provided Jacobian"):
        fts.DAEProblem(F, u, u_t, (0, 1), J="J")
    with pytest.raises(ValueError, match="not a bilinear form"):
        fts.DAEProblem(F, u, u_t, (0, 1), J=Ku)


def test_preconditioner_form_Jp_fills_P():
    """reference solving_utils.py:120-122, 310-312: Jp is assembled into its own Mat P after J, and the
    KSP builds its preconditioner from P (here: LU of a lumped-stiffness-like operator inside GMRES)."""
    V, u, u_t, F, bc = heat_1d(12, 2)
    u.interpolate(lambda x: torch.sin(math.pi * x[:, 0]))
    u0 = u.data.clone()
    fts.DAESolver(fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc), assembler=OracleAssembler(F, [bc]),
                  solver_parameters=LU).solve()
    ref = u.data.clone()
    u.data.copy_(u0)
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc)
    Pf = forms.diffusion(u, u_t, mass=1.0, diffusivity=0.8)
    problem = fts.DAEProblem(F, u, u_t, (0.0, 0.2), bcs=bc, Jp=forms.derivative(Pf, None))
    assert not problem.Jp_eq_J
    seen = []
    s = fts.DAESolver(problem, assembler=OracleAssembler(F, [bc]), pjac_assembler=OracleAssembler(Pf, [bc]),
                      post_jacobian_callback=lambda X, Xdot, J: seen.append(J),
                      solver_parameters={"ksp_type": "gmres", "pc_type": "lu", "ksp_rtol": 1e-14, "snes_rtol": 1e-13})
    ctx = s._ctx
    assert ctx._pjac is not ctx._jac and ctx._pjac.petscmat.handle != ctx._jac.petscmat.handle
    s.solve()
    assert (u.data - ref).abs().max() < 1e-10
    assert s.ts.linear_its > s.ts.newton_its  # P != J: GMRES needs more than one iteration
    shift = float(problem.shift)
    Pw = OracleAssembler(Pf, [bc]).ijacobian(0.0, u.data, u_t.data, shift).numpy()
    assert np.abs(ctx._pjac.petscmat.values.numpy() - Pw).max() < 1e-13 * np.abs(Pw).max()
    assert seen and all(J is ctx._jac.petscmat for J in seen)
    with pytest.raises(ValueError, match="preconditioner is not a bilinear form"):
        fts.DAEProblem(F, u, u_t, (0, 1), Jp=Pf)


def test_unknown_solver_options_and_nullspaces_are_rejected():
    V, u, u_t, F, bc = heat_1d(8, 1)
    for bad in ({"pc_type": "hypre"}, {"pc_type": "fieldsplit", "pc_fieldsplit_type": "schur"},
                {"pc_type": "fieldsplit", "fieldsplit_0_pc_type": "hypre"}, {"ksp_type": "bcgs", "pc_type": "jacobi"}):
        u.data.zero_()
        s = fts.DAESolver(fts.DAEProblem(F, u, u_t, (0.0, 0.1), bcs=bc), assembler=OracleAssembler(F, [bc]),
                          solver_parameters=bad)
        with pytest.raises(NotImplementedError):
            s.solve()
    ns = object()
    s = fts.DAESolver(fts.DAEProblem(F, u, u_t, (0.0, 0.1), bcs=bc), assembler=OracleAssembler(F, [bc]),
                      nullspace=ns, near_nullspace="near", solver_parameters={"ksp_type": "cg", "pc_type": "jacobi"})
    m = s._ctx._jac.petscmat
    assert m._nullspace is ns and m._near_nullspace == "near" and m._transpose_nullspace is None
    with pytest.raises(NotImplementedError, match="nullspace"):
        s.solve()


def test_negative_time_span_takes_the_right_number_of_steps():
    V, u, u_t, F, bc = heat_1d(8, 1)
    s = fts.DAESolver(fts.DAEProblem(F, u, u_t, (-1.0, 0.0), bcs=bc), assembler=OracleAssembler(F, [bc]),
                      solver_parameters=LU)
    s.solve()
    assert s.ts.getStepNumber() == 10 and abs(s.ts.getTime()) < 1e-12


def test_mat_diagonal_of_blocked_space_on_a_partition():
    """local column of owned dof (m, i) is m*num_nodes + i, not m*num_owned + i"""
    from firedrake_ts_b200 import mesh as M
    from firedrake_ts_b200.ts import Mat
    from firedrake_ts_b200.workloads import make_problem

    form, bcs, u, udot = make_problem("cahn_hilliard", 2, 1, 6, partition=M.Partition((2, 1), 0))
    V = form.space
    asm = OracleAssembler(form, bcs)
    rp, ci = asm.pattern()
    vals = asm.ijacobian(0.0, u.data, udot.data, 3.0)
    A = Mat(torch.from_numpy(rp), torch.from_numpy(ci), vals, V.num_dofs, space=V)
    d = A.diagonal().numpy()
    dense = A.to_scipy().toarray()
    nown, nn = V.num_owned_nodes, V.num_nodes
    want = np.array([dense[m * nown + i, m * nn + i] for m in range(2) for i in range(nown)])
    assert np.array_equal(d, want) and np.all(d[:nown] != 0)


def test_split_field_subcontexts():
    """reference solving_utils.py:145-233: sub-contexts on the fields of the mixed Cahn-Hilliard space assemble
    the field's rows of F and the diagonal block of J with the other field as a coefficient; cached per
    field tuple; the sub-Functions write through to the parent's."""
    from firedrake_ts_b200.ts import DM, Vec
    from firedrake_ts_b200.workloads import make_problem
    from oracle import Oracle

    form, bcs, u, udot = make_problem("cahn_hilliard", 2, 1, 5)
    problem = fts.DAEProblem(form, u, udot, (0.0, 1e-5))
    solver = fts.DAESolver(problem, assembler=OracleAssembler(form, bcs), solver_parameters={"ts_dt": 5e-6})
    ctx = solver._ctx
    splits = ctx.split(((0,), (1,)))
    assert ctx.split([[0], [1]]) is splits and len(splits) == 2
    with pytest.raises(ValueError):
        ctx.split(((2,),))
    orc = Oracle.from_form(form, bcs)
    V = form.space
    nn = V.num_nodes
    x0, xd0 = u.data.clone(), udot.data.clone()
    Ffull = orc.ifunction(0.0, x0, xd0)
    Jfull = orc.ijacobian(0.0, x0, xd0, 3.0)
    rp, ci = orc.pattern()
    import scipy.sparse as sp

    A = sp.csr_matrix((Jfull, ci, rp), shape=(2 * nn, 2 * nn)).toarray()
    for m, sc in enumerate(splits):
        assert sc.space.ncomp == 1 and sc.space.num_nodes == nn
        ts = solver.ts.__class__()
        dm = DM()
        dm._appctx.append(sc)
        ts.setDM(dm)
        X, Xd = Vec(x0[m * nn:(m + 1) * nn].clone()), Vec(xd0[m * nn:(m + 1) * nn].clone())
        Fv = Vec(torch.zeros(nn, dtype=torch.float64))
        sc.form_function(ts, 0.0, X, Xd, Fv)
        assert np.allclose(Fv.array.numpy(), Ffull[m * nn:(m + 1) * nn], rtol=1e-13, atol=1e-15)
        sc.form_jacobian(ts, 0.0, X, Xd, 3.0, sc._jac.petscmat, sc._pjac.petscmat)
        B = sc._jac.petscmat.to_scipy().toarray()
        assert np.allclose(B, A[m * nn:(m + 1) * nn, m * nn:(m + 1) * nn], rtol=1e-13, atol=1e-15)
        # a changed sub-state is seen by the parent (shared data) and changes the other field's residual
        X.array += 0.01
        sc.form_function(ts, 0.0, X, Xd, Fv)
        assert torch.allclose(u.data[m * nn:(m + 1) * nn], X.array)
        u.data.copy_(x0)
        udot.data.copy_(xd0)
    # the pair of fields is the whole problem
    both = ctx.split(((0, 1),))[0]
    assert both._jac.petscmat.values.numel() == len(Jfull)


@pytest.mark.parametrize("kind", ["additive", "multiplicative"])
def test_fieldsplit_preconditioner_solves_cahn_hilliard(kind):
    """GMRES preconditioned by PCFIELDSPLIT over (c, mu) with LU on the diagonal blocks reproduces the direct
    solve of the literal example's first steps."""
    from firedrake_ts_b200.workloads import make_problem

    def run(params):
        form, bcs, u, udot = make_problem("cahn_hilliard", 2, 1, 6, quadrilateral=True)
        u.data[form.space.num_nodes:] = 0.0
        problem = fts.DAEProblem(form, u, udot, (0.0, 1e-5))
        s = fts.DAESolver(problem, assembler=OracleAssembler(form, bcs), solver_parameters=dict(params, ts_dt=5e-6))
        s.solve()
        return u.data.clone(), s

    base = {"snes_rtol": 1e-10, "snes_atol": 1e-11}
    ud, _ = run(dict(base, pc_type="lu"))
    uf, s = run(dict(base, ksp_type="gmres", ksp_rtol=1e-12, pc_type="fieldsplit", pc_fieldsplit_type=kind,
                     fieldsplit_0_pc_type="lu", fieldsplit_1_pc_type="lu"))
    assert s.ts.getStepNumber() == 2
    assert (ud - uf).abs().max() < 1e-8 * max(1.0, float(ud.abs().max()))
    with pytest.raises(NotImplementedError):
        run(dict(base, pc_type="fieldsplit", pc_fieldsplit_type="schur"))
