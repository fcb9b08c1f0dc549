"""DAEProblem / DAESolver: the public API of firedrake-ts, unchanged
(reference firedrake_ts/ts_solver.py:39-393), on top of the B200 engine.

``DAEProblem(F, u, udot, tspan, time=None, bcs=None, J=None, Jp=None,
form_compiler_parameters=None, is_linear=False, G=None)`` and
``DAESolver(problem, **kwargs)`` take the same arguments with the same meaning
and raise the same errors as the reference.  ``F`` is a catalogue form
(firedrake_ts_b200.forms) instead of a UFL form because no UFL/TSFC exists
here; with a real Firedrake install the adapter in adapter_firedrake.py is used
instead and UFL forms are recognised by family.
"""
from __future__ import annotations

from contextlib import ExitStack
from itertools import chain

from . import dmhooks, forms
from .function import Constant, DirichletBC, Function
from .solving_utils import _TSContext, check_ts_convergence
from .ts import DM, TS, Vec


def check_pde_args(F, G, J, Jp):
    """reference ts_solver.py:13-29"""
    if not isinstance(F, forms.Form):
        raise TypeError("Provided residual is a '%s', not a BaseForm or Slate Tensor" % type(F).__name__)
    if len(F.arguments()) != 1:
        raise ValueError("Provided residual is not a linear form")
    if G is not None and not isinstance(G, forms.Form):
        raise TypeError(f"Provided G residual is a '{type(G).__name__}', not a BaseForm or Slate Tensor")
    if G is not None and len(G.arguments()) != 1:
        raise ValueError("Provided G residual is not a linear form")
    if not isinstance(J, forms.Form):
        raise TypeError("Provided Jacobian is a '%s', not a BaseForm or Slate Tensor" % type(J).__name__)
    if len(J.arguments()) != 2:
        raise ValueError("Provided Jacobian is not a bilinear form")
    if Jp is not None and not isinstance(Jp, forms.Form):
        raise TypeError("Provided preconditioner is a '%s', not a BaseForm or Slate Tensor" % type(Jp).__name__)
    if Jp is not None and len(Jp.arguments()) != 2:
        raise ValueError("Provided preconditioner is not a bilinear form")


def _extract_bcs(bcs):
    if bcs is None:
        return ()
    if isinstance(bcs, DirichletBC):
        return (bcs,)
    return tuple(bcs)


class DAEProblem(object):
    r"""Nonlinear variational problem in DAE form F(u̇, u, t) = G(u, t)."""

    def __init__(self, F, u, udot, tspan, time=None, bcs=None, J=None, Jp=None,
                 form_compiler_parameters=None, is_linear=False, G=None):
        r"""
        :param F: the nonlinear form
        :param u: the :class:`.Function` to solve for
        :param udot: the :class:`.Function` for time derivative
        :param tspan: the tuple for start time and end time
        :param time: the :class:`.Constant` for time-dependent weak forms
        :param bcs: the boundary conditions (optional)
        :param J: the Jacobian J = sigma*dF/du̇ + dF/du (optional)
        :param Jp: a form used for preconditioning the linear system (optional)
        :param dict form_compiler_parameters: accepted and stored (no form compiler here)
        :is_linear: internally used to check form style consistency
        :param G: G(t, u) term treated explicitly by IMEX methods: F(u̇, u, t) = G(u, t)
            (reference ts_solver.py:111; examples/heat-explicit.py)
        """
        self.bcs = _extract_bcs(bcs)
        self.is_linear = is_linear
        self.Jp_eq_J = Jp is None

        self.u_restrict = u
        self.udot = udot
        self.tspan = tspan
        self.F = F
        self.G = G
        self.Jp = Jp

        if not isinstance(self.u_restrict, Function):
            raise TypeError("Provided solution is a '%s', not a Function" % type(self.u_restrict).__name__)
        if not isinstance(self.udot, Function):
            raise TypeError("Provided time derivative is a '%s', not a Function" % type(self.udot).__name__)

        # current value of time that may be used in weak form
        self.time = time or Constant(0.0)
        # timeshift value provided by the solver
        self.shift = Constant(1.0)

        # Use the user-provided Jacobian. If none is provided, derive the Jacobian from the residual
        # (reference ts_solver.py:108: J = shift*derivative(F, udot) + derivative(F, u)).
        # A user-provided J replaces dF/du only; the shift * dF/dudot term always comes from F.  Catalogue
        # forms carry no symbolic sum, so the user's part is kept beside the total (J_du) and the context
        # assembles shift * M_F + J_du.
        self.J_du = J
        if J is not None and not isinstance(J, forms.Form):
            raise TypeError("Provided Jacobian is a '%s', not a BaseForm or Slate Tensor" % type(J).__name__)
        if J is not None and len(J.arguments()) != 2:
            raise ValueError("Provided Jacobian is not a bilinear form")
        self.J = forms.derivative(F, self.shift) if isinstance(F, forms.Form) else J
        # reference ts_solver.py:111: dGdu = derivative(G, u)  (no udot term: shift 0)
        self.dGdu = forms.derivative(G, Constant(0.0)) if isinstance(G, forms.Form) else None

        check_pde_args(self.F, self.G, self.J, self.Jp)

        self.form_compiler_parameters = form_compiler_parameters
        self._constant_jacobian = False
        self._constant_rhs_jacobian = False

    def dirichlet_bcs(self):
        for bc in self.bcs:
            yield from bc.dirichlet_bcs()

    @property
    def dm(self):
        V = self.u_restrict.function_space()
        if not hasattr(V, "_dm"):
            V._dm = DM()
        return V._dm


class OptionsManager(object):
    """the slice of firedrake.petsc.OptionsManager the reference uses (ts_solver.py:130,202-231)."""
    _count = 0

    def __init__(self, parameters, options_prefix):
        self.parameters = dict(parameters or {})
        if options_prefix is None:
            options_prefix = "firedrake_ts_%d_" % OptionsManager._count
            OptionsManager._count += 1
        self.options_prefix = options_prefix

    def set_default_parameter(self, key, val):
        self.parameters.setdefault(key, val)

    def set_from_options(self, ts):
        ts.setFromOptions(self.parameters)

    def inserted_options(self):
        return ExitStack()


class DAESolver(OptionsManager):
    r"""Solves a :class:`DAEProblem` (reference ts_solver.py:130-393)."""

    def __init__(self, problem, **kwargs):
        r"""
        :arg problem: A :class:`DAEProblem` to solve.
        :kwarg nullspace / transpose_nullspace / near_nullspace: stored and re-attached
               after every Jacobian assembly, as in the reference.
        :kwarg solver_parameters: dict of PETSc-style options (``ts_type``, ``ts_theta_theta``,
               ``snes_rtol``, ``ksp_type``, ``pc_type``, ``mat_type`` ...).
        :kwarg appctx, options_prefix: as in the reference.
        :kwarg pre_jacobian_callback / pre_function_callback / post_jacobian_callback /
               post_function_callback: user hooks called with the Vecs / Mat of the callback.
        :kwarg monitor_callback: ``monitor(ts, step, t, x)`` called after every step.
        :kwarg comm: torch.distributed process group (one rank per GPU) for partitioned meshes.
        :kwarg assembler: "b200" (default); tests may pass a CPU oracle object.
        """
        assert isinstance(problem, DAEProblem)

        parameters = kwargs.get("solver_parameters")
        if "parameters" in kwargs:
            raise TypeError("Use solver_parameters, not parameters")
        nullspace = kwargs.get("nullspace")
        nullspace_T = kwargs.get("transpose_nullspace")
        near_nullspace = kwargs.get("near_nullspace")
        options_prefix = kwargs.get("options_prefix")
        pre_j_callback = kwargs.get("pre_jacobian_callback")
        pre_f_callback = kwargs.get("pre_function_callback")
        post_j_callback = kwargs.get("post_jacobian_callback")
        post_f_callback = kwargs.get("post_function_callback")
        monitor_callback = kwargs.get("monitor_callback")
        comm = kwargs.get("comm")

        self.nullspace = nullspace
        self.near_nullspace = near_nullspace
        self.nullspace_T = nullspace_T

        super(DAESolver, self).__init__(parameters, options_prefix)

        mat_type = self.parameters.get("mat_type")
        pmat_type = self.parameters.get("pmat_type")
        if mat_type == "matfree" or pmat_type == "matfree":
            raise NotImplementedError("matrix-free operators are outside the engine's scope")

        appctx = kwargs.get("appctx")

        ctx = _TSContext(problem, mat_type=mat_type, pmat_type=pmat_type, appctx=appctx,
                         pre_jacobian_callback=pre_j_callback, pre_function_callback=pre_f_callback,
                         post_jacobian_callback=post_j_callback, post_function_callback=post_f_callback,
                         options_prefix=self.options_prefix, assembler=kwargs.get("assembler", "b200"), comm=comm,
                         project_rhs=kwargs.get("project_rhs", True),
                         rhs_projection_parameters=kwargs.get("rhs_projection_parameters"),
                         rhs_assembler=kwargs.get("rhs_assembler"), mass_assembler=kwargs.get("mass_assembler"),
                         jac_du_assembler=kwargs.get("jac_du_assembler"), pjac_assembler=kwargs.get("pjac_assembler"))

        if ctx.is_mixed:
            # Mixed problem, use jacobi pc if user has not supplied one.
            self.set_default_parameter("pc_type", "jacobi")
        self.set_default_parameter("ts_max_snes_failures", -1)

        self.ts = TS().create(comm=comm)
        self.snes = self.ts.getSNES()

        self._problem = problem
        self._ctx = ctx
        # work vector in the layout (and on the device) of the owned dofs: VECCUDA in PETSc terms
        self._work = Vec(problem.u_restrict.owned().clone().to(ctx._device), comm)

        self.ts.setDM(problem.dm)
        self.ts.setMonitor(monitor_callback)
        self.ts.setTime(problem.tspan[0])
        self.ts.setMaxTime(problem.tspan[1])
        if problem.G is None:
            self.ts.setEquationType(TS.EquationType.IMPLICIT)
        else:
            # If G is provided use the arkimex solver (reference ts_solver.py:252-253)
            self.set_default_parameter("ts_type", "arkimex")
        self.set_default_parameter("ts_exact_final_time", "stepover")

        self._set_problem_eval_funcs(ctx, problem, nullspace, nullspace_T, near_nullspace)

        dm = self.ts.getDM()
        with dmhooks.add_hooks(dm, self, appctx=self._ctx, save=False):
            self.set_from_options(self.ts)

        self._transfer_operators = ()
        self._setup = False

    def _set_problem_eval_funcs(self, ctx, problem, nullspace, nullspace_T, near_nullspace):
        r"""(re-)register the callbacks and nullspaces on the TS (reference ts_solver.py:274-312;
        re-done at every solve(), which tests/test_two_daesolver.py of the reference pins)."""
        ctx.set_ifunction(self.ts)
        ctx.set_ijacobian(self.ts)
        ctx.set_rhs_function(self.ts)
        ctx.set_rhs_jacobian(self.ts)
        ctx.set_nullspace(nullspace, None, transpose=False, near=False)
        ctx.set_nullspace(nullspace_T, None, transpose=True, near=False)
        ctx.set_nullspace(near_nullspace, None, transpose=False, near=True)
        ctx._nullspace = nullspace
        ctx._nullspace_T = nullspace_T
        ctx._near_nullspace = near_nullspace

    def set_transfer_manager(self, manager):
        self._ctx.transfer_manager = manager

    def solve(self, bounds=None):
        r"""Solve the time-dependent variational problem (reference ts_solver.py:323-367)."""
        if bounds is not None:
            raise NotImplementedError("variable bounds need PETSc's vinewton solvers")
        self._set_problem_eval_funcs(self._ctx, self._problem, self.nullspace, self.nullspace_T,
                                     self.near_nullspace)
        dm = self.ts.getDM()
        for dbc in self._problem.dirichlet_bcs():
            dbc.apply(self._problem.u_restrict)

        work = self._work
        u = self._problem.u_restrict
        work.array.copy_(u.owned().to(work.array.device))
        with ExitStack() as stack:
            for ctx in chain((self.inserted_options(), dmhooks.add_hooks(dm, self, appctx=self._ctx)),
                             self._transfer_operators):
                stack.enter_context(ctx)
            self.ts.solve(work)
        u.set_owned(work.array.to(u.data.device))
        if self._ctx._halo is not None:
            self._ctx._halo.forward(u.data)
        self._setup = True
        check_ts_convergence(self.ts)

    def adjoint_solve(self):
        raise NotImplementedError("TSAdjoint lives in PETSc; it re-uses the same IJacobian callback "
                                  "(SURVEY.md section 2.1 row 6) and is available through adapter_firedrake only")
