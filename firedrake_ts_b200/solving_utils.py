"""_TSContext: the callback/context layer (reference firedrake_ts/solving_utils.py).

Same names, argument meaning and error behaviour as the reference's
``_TSContext`` (solving_utils.py:42-317) for the IFunction / IJacobian path:
``set_ifunction``, ``set_ijacobian``, ``form_function(ts, t, X, Xdot, F)``,
``form_jacobian(ts, t, X, Xdot, shift, J, P)``, the four user callbacks, the
``J.handle`` assertion, the ``_jacobian_assembled`` flag and the
``_time``/``_shift`` Constants.  What changes is what sits behind
``_assemble_residual`` / ``_assemble_jac``: the B200 engine (engine.Engine)
writing straight into the device-resident Vec / AIJ value arrays, plus a halo
exchange in place of PyOP2's.  The IMEX RHS branch (solving_utils.py:320-384),
``split`` and ``compute_operators`` are out of scope.
"""
from __future__ import annotations

import torch

from . import dmhooks, forms
from .ts import Mat, Vec


class ConvergenceError(Exception):
    pass


TSReasons = {1: "CONVERGED_TIME", 2: "CONVERGED_ITS", 0: "CONVERGED_ITERATING",
             -1: "DIVERGED_NONLINEAR_SOLVE", -2: "DIVERGED_STEP_REJECTED"}


def check_ts_convergence(ts):
    """reference solving_utils.py:22-39"""
    r = ts.getConvergedReason()
    if r == -3:
        raise ConvergenceError("TS solve failed to converge. Reason: TSFORWARD_DIVERGED_LINEAR_SOLVE")
    if r == -4:
        raise ConvergenceError("TS solve failed to converge. Reason: TSADJOINT_DIVERGED_LINEAR_SOLVE")
    reason = TSReasons.get(r, str(r))
    if r < 0:
        raise ConvergenceError(
            f"TS solve failed to converge after {ts.getStepNumber()} iterations. Reason: {reason}")


class _Assembled:
    """holder with ``.petscmat`` like a Firedrake Matrix (solving_utils.py:122,284)."""

    def __init__(self, petscmat):
        self.petscmat = petscmat


class _TSContext:
    r"""Context holding information for TS callbacks (reference solving_utils.py:42-111).

    :arg problem: a :class:`DAEProblem`.
    :arg mat_type: "aij" (monolithic, the only type this engine assembles).
    :arg pmat_type: as mat_type; ``Jp`` is not supported by the engine path.
    :arg appctx / pre_*/post_* callbacks / options_prefix: as in the reference.
    :arg assembler: "b200" (default) or an object with ``ifunction`` /
        ``ijacobian`` / ``pattern`` methods (the tests plug the CPU oracle in here
        to produce reference trajectories through the very same callbacks).
    """

    def __init__(self, problem, mat_type=None, pmat_type=None, appctx=None,
                 pre_jacobian_callback=None, pre_function_callback=None,
                 post_jacobian_callback=None, post_function_callback=None,
                 options_prefix=None, transfer_manager=None, assembler="b200", comm=None):
        if mat_type not in (None, "aij", "aijcusparse"):
            raise ValueError(f"mat_type {mat_type!r}: the B200 engine assembles monolithic AIJ only")
        if problem.Jp is not None:
            raise NotImplementedError("a separate preconditioner form Jp is outside the engine's scope")
        if problem.G is not None:
            raise NotImplementedError("the IMEX G(t,u) branch is outside the engine's scope (SURVEY.md 2.1 row 2)")
        self._problem = problem
        self.mat_type, self.pmat_type = mat_type or "aij", pmat_type or "aij"
        self.appctx = appctx or {}
        self.options_prefix = options_prefix
        self.transfer_manager = transfer_manager
        self._pre_jacobian_callback = pre_jacobian_callback
        self._pre_function_callback = pre_function_callback
        self._post_jacobian_callback = post_jacobian_callback
        self._post_function_callback = post_function_callback
        self.comm = comm
        # state the callbacks update (reference :92-96 and base class)
        self._x = problem.u_restrict
        self._xdot = problem.udot
        self._shift = problem.shift
        self._time = problem.time
        self.F, self.J, self.Jp = problem.F, problem.J, problem.Jp
        self.G, self.dGdu = None, None
        self.bcs_F = tuple(problem.dirichlet_bcs())
        self.fcp = problem.form_compiler_parameters
        self._jacobian_assembled = False
        self._nullspace = self._nullspace_T = self._near_nullspace = None
        self._splits = {}
        V = self._x.function_space()
        self.is_mixed = V.layout == "blocked" and V.ncomp > 1
        dev = self._x.data.device

        if assembler == "b200":
            from .engine import Engine

            self.engine = Engine(problem.F, self.bcs_F, device=dev if dev.type == "cuda" else None,
                                 scatter=self.appctx.get("scatter"))
            self._device = self.engine.device
            # device-resident state: the Functions the callbacks update live on the engine's GPU
            self._x.data = self._x.data.to(self._device)
            self._xdot.data = self._xdot.data.to(self._device)
            rowptr, colidx = self.engine.csr()
            values = self.engine.new_values()
        else:
            self.engine = assembler
            self._device = dev
            rp, ci = assembler.pattern()
            rowptr, colidx = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
            values = torch.zeros(colidx.numel(), dtype=torch.float64, device=dev)
        self._halo = None
        if V.mesh.partition.size > 1:
            from .halo import HaloExchange

            self._halo = HaloExchange(V, self.engine if assembler == "b200" else None, group=comm,
                                      device=self._device)
        # residual Cofunction and Jacobian Matrix stand-ins (base class: _F, _jac, _pjac)
        self._F = Vec(torch.zeros(V.num_owned_dofs, dtype=torch.float64, device=self._device), comm)
        self._jac = _Assembled(Mat(rowptr, colidx, values, V.num_dofs, halo=self._halo))
        self._pjac = self._jac
        self._assemble_residual = self._b200_assemble_residual
        self._assemble_jac = self._b200_assemble_jac

    # -- the two assemblers that replace get_assembler(...).assemble ----------------
    def _b200_assemble_residual(self, tensor=None):
        tensor = tensor if tensor is not None else self._F
        out = self.engine.ifunction(float(self._time), self._x.data, self._xdot.data, out=tensor.array)
        if out is not tensor.array:
            tensor.array.copy_(torch.as_tensor(out))
        return tensor

    def _b200_assemble_jac(self, jac):
        mat = jac.petscmat
        out = self.engine.ijacobian(float(self._time), self._x.data, self._xdot.data, float(self._shift),
                                    out=mat.values)
        if out is not mat.values:
            mat.values.copy_(torch.as_tensor(out))
        return jac

    def _refresh_halo(self):
        if self._halo is not None:
            self._halo.forward(self._x.data)
            self._halo.forward(self._xdot.data)

    # -- registration (reference :115-122) ----------------------------------------------
    def set_ifunction(self, ts):
        r"""Set the function to compute F(t,U,U_t) where F() = 0 is the DAE to be solved."""
        ts.setIFunction(self.form_function, self._F)

    def set_ijacobian(self, ts):
        r"""Set the function to compute the matrix dF/dU + a*dF/dU_t."""
        ts.setIJacobian(self.form_jacobian, J=self._jac.petscmat, P=self._pjac.petscmat)

    def set_rhs_function(self, ts):
        pass  # G is None on the engine path

    def set_rhs_jacobian(self, ts):
        pass

    def set_nullspace(self, nullspace, ises=None, transpose=False, near=False):
        if nullspace is None:
            return
        self._jac.petscmat._nullspace = nullspace

    # -- callbacks (reference :236-317) ------------------------------------------------------
    @staticmethod
    def form_function(ts, t, X, Xdot, F):
        r"""Form the residual for this problem

        :arg ts: a TS object
        :arg t: the time at step/stage being solved
        :arg X: state vector
        :arg Xdot: time derivative of state vector
        :arg F: function vector
        """
        dm = ts.getDM()
        ctx = dmhooks.get_appctx(dm)
        # X may not be the same vector as the vec behind self._x, so copy guess in from X.
        ctx._x.set_owned(X.array)
        ctx._xdot.set_owned(Xdot.array)
        ctx._refresh_halo()
        ctx._time.assign(t)

        if ctx._pre_function_callback is not None:
            ctx._pre_function_callback(X, Xdot)

        ctx._assemble_residual(tensor=ctx._F)

        if ctx._post_function_callback is not None:
            ctx._post_function_callback(X, Xdot, ctx._F)

        # F may not be the same vector as self._F, so copy residual out to F.
        ctx._F.copy(F)

    @staticmethod
    def form_jacobian(ts, t, X, Xdot, shift, J, P):
        r"""Form the Jacobian for this problem

        :arg ts: a TS object
        :arg t: the time at step/stage being solved
        :arg X: state vector
        :arg Xdot: time derivative of state vector
        :arg J: the Jacobian (a Mat)
        :arg P: the preconditioner matrix (a Mat)
        """
        dm = ts.getDM()
        ctx = dmhooks.get_appctx(dm)
        problem = ctx._problem

        assert J.handle == ctx._jac.petscmat.handle
        if problem._constant_jacobian and ctx._jacobian_assembled:
            # Don't need to do any work with a constant jacobian that's already assembled
            return
        ctx._jacobian_assembled = True

        ctx._x.set_owned(X.array)
        ctx._xdot.set_owned(Xdot.array)
        ctx._refresh_halo()
        ctx._time.assign(t)

        if ctx._pre_jacobian_callback is not None:
            ctx._pre_jacobian_callback(X, Xdot)

        ctx._shift.assign(shift)
        ctx._assemble_jac(ctx._jac)

        if ctx._post_jacobian_callback is not None:
            ctx._post_jacobian_callback(X, Xdot, J)

        ctx.set_nullspace(ctx._nullspace, None, transpose=False, near=False)
        ctx.set_nullspace(ctx._nullspace_T, None, transpose=True, near=False)
        ctx.set_nullspace(ctx._near_nullspace, None, transpose=False, near=True)
