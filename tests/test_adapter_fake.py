"""GPU: the Firedrake adapter's callbacks executed with PETSc-/Firedrake-shaped stand-ins (fake_petsc.py).
B200TSContext.form_function / form_jacobian write F and the Mat values in place through the device handles;
results against the oracle at 1e-12, reference side effects (state copied into ctx._x/_xdot, time/shift
assigned, callbacks, nullspace re-attached, Jp assembled by the inherited path, pattern check) verified."""
import numpy as np
import pytest
import torch

from firedrake_ts_b200 import adapter_firedrake as A
from firedrake_ts_b200 import forms
from fake_petsc import (FakeBC, FakeFiredrakeSpace, FakeMat, FakeProblem, FakeTS, FakeVec, StubRefTSContext,
                        fake_dmhooks)
from helpers import make_problem, rel_err

pytestmark = pytest.mark.gpu
CASES = [("heat", 2, 2, 6, False), ("heat", 3, 2, 3, False), ("burgers", 2, 2, 4, False),
         ("cahn_hilliard", 2, 1, 7, False), ("cahn_hilliard", 2, 1, 7, True), ("bbm", 3, 2, 3, False),
         ("diffusion", 2, 1, 8, False)]


def build_context(family, dim, degree, n, quad, Jp=None, **cb):
    from oracle import Oracle

    form, bcs, u, udot = make_problem(family, dim, degree, n, quadrilateral=quad)
    V = form.space
    orc = Oracle.from_form(form, bcs)
    rp, ci = orc.pattern()
    Vfd = FakeFiredrakeSpace(V)
    nown = V.num_owned_dofs
    fbcs = [FakeBC(bc.nodes.numpy()) for bc in bcs]
    problem = FakeProblem(Vfd, nown, bcs=fbcs, Jp=Jp)
    jac = FakeMat(rp, ci)
    cls = A.make_context_class(StubRefTSContext, hooks=fake_dmhooks, values_pointer=lambda m: m.values.data_ptr())
    ctx = cls(problem, appctx={"b200_form": (form.family, form.params)}, jac=jac,
              pjac=FakeMat(rp, ci) if Jp is not None else None, **cb)
    return ctx, form, bcs, u, udot, orc, jac


@pytest.mark.parametrize("family,dim,degree,n,quad", CASES)
def test_callbacks_match_oracle(family, dim, degree, n, quad):
    ctx, form, bcs, u, udot, orc, jac = build_context(family, dim, degree, n, quad)
    assert ctx._b200 is not None
    ts = FakeTS(ctx)
    X, Xd = FakeVec(array=u.data.cuda()), FakeVec(array=udot.data.cuda())
    F = FakeVec(u.data.numel())
    type(ctx).form_function(ts, 0.25, X, Xd, F)
    assert rel_err(F.array.cpu().numpy(), orc.ifunction(0.25, u.data, udot.data)) < 1e-12
    assert ctx._time.value == 0.25
    assert torch.equal(ctx._x.dat._vec.array, X.array) and torch.equal(ctx._xdot.dat._vec.array, Xd.array)
    assert X.handles_out == Xd.handles_out == F.handles_out == 0
    type(ctx).form_jacobian(ts, 0.5, X, Xd, 10.0, jac, jac)
    assert rel_err(jac.values.cpu().numpy(), orc.ijacobian(0.5, u.data, udot.data, 10.0)) < 1e-12
    assert ctx._shift.value == 10.0 and ctx._time.value == 0.5 and jac.assembled == 1
    assert ctx.nullspace_calls == [(False, False), (True, False), (False, True)]
    assert ctx._jacobian_assembled and ctx.pjac_assembled == 0 and ctx.reference_path == []


def test_callbacks_order_and_jp():
    seen = []
    ctx, form, bcs, u, udot, orc, jac = build_context(
        "heat", 2, 1, 5, False, Jp=object(),
        pre_function_callback=lambda X, Xd: seen.append("preF"),
        post_function_callback=lambda X, Xd, F: seen.append(("postF", float(F.array.abs().sum()) > 0)),
        pre_jacobian_callback=lambda X, Xd: seen.append("preJ"),
        post_jacobian_callback=lambda X, Xd, J: seen.append(("postJ", float(J.values.abs().sum()) > 0)))
    ts = FakeTS(ctx)
    X, Xd, F = FakeVec(array=u.data.cuda()), FakeVec(array=udot.data.cuda()), FakeVec(u.data.numel())
    type(ctx).form_function(ts, 0.0, X, Xd, F)
    type(ctx).form_jacobian(ts, 0.0, X, Xd, 1.0, jac, ctx._pjac.petscmat)
    assert seen == ["preF", ("postF", True), "preJ", ("postJ", True)]
    assert ctx.pjac_assembled == 1  # the preconditioner form went through the inherited assembler
    with pytest.raises(AssertionError):  # a foreign Mat is refused, as in the reference
        type(ctx).form_jacobian(ts, 0.0, X, Xd, 1.0, FakeMat(jac.rowptr, jac.colidx), jac)


def test_pattern_mismatch_is_refused_and_unknown_forms_fall_through():
    ctx, form, bcs, u, udot, orc, jac = build_context("heat", 2, 1, 5, False)
    ts = FakeTS(ctx)
    X, Xd = FakeVec(array=u.data.cuda()), FakeVec(array=udot.data.cuda())
    bad = FakeMat(jac.rowptr, jac.colidx[::-1].copy())
    ctx._jac.petscmat = bad
    with pytest.raises(RuntimeError, match="pattern differs"):
        type(ctx).form_jacobian(ts, 0.0, X, Xd, 1.0, bad, bad)
    # a form outside the catalogue keeps the reference's own callbacks
    ctx._b200 = None
    type(ctx).form_function(ts, 0.0, X, Xd, FakeVec(u.data.numel()))
    type(ctx).form_jacobian(ts, 0.0, X, Xd, 1.0, bad, bad)
    assert ctx.reference_path == ["F", "J"]


def test_use_b200_swaps_context_and_types():
    ctx, form, bcs, u, udot, orc, jac = build_context("heat", 2, 1, 4, False)
    solver = type("S", (), {})()
    old = StubRefTSContext(ctx._problem, appctx={"b200_form": (forms.HEAT, form.params)}, jac=jac)
    solver._ctx, solver._work = old, FakeVec(4)

    class Ctx(A.make_context_class(StubRefTSContext, hooks=fake_dmhooks, values_pointer=lambda m: m.values.data_ptr())):
        def __init__(self, problem, **kw):
            super().__init__(problem, jac=jac, **kw)

    A.use_b200(solver, context_class=Ctx)
    assert isinstance(solver._ctx, Ctx) and solver._ctx._b200 is not None
    assert jac.type == "aijcusparse" and solver._work.type == "cuda"
