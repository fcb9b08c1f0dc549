"""Minimal PETSc-TS stand-in: the *caller* of the hot path.

PETSc (TS/SNES/KSP) is what invokes ``form_function`` / ``form_jacobian`` in the
reference (firedrake_ts/ts_solver.py:233-272, :364) and stays in place in a real
deployment (out of scope).  PETSc is absent in this image,
so this module provides just enough of its surface -- ``Vec``, ``Mat``, ``DM``,
``SNES``/``KSP`` settings and ``TS`` with the BEULER / THETA integrators of
SURVEY.md A.5 -- for the drop-in ``DAESolver`` to run standalone and for the
trajectory tests.  Callback signatures are PETSc's:
``IFunction(ts, t, X, Xdot, F)`` and ``IJacobian(ts, t, X, Xdot, shift, J, P)``.

Everything here works on torch tensors, on the GPU or (for the oracle-driven
tests) on the CPU.
"""
from __future__ import annotations

import math

import torch


class Vec:
    """PETSc.Vec look-alike over one torch tensor (the rank's owned dofs)."""

    def __init__(self, array: torch.Tensor, comm=None):
        self.array = array
        self.comm = comm

    @property
    def handle(self):
        return id(self)

    def copy(self, result=None):
        if result is None:
            return Vec(self.array.clone(), self.comm)
        if result.array.data_ptr() != self.array.data_ptr():
            result.array.copy_(self.array)
        return result

    def duplicate(self):
        return Vec(torch.zeros_like(self.array), self.comm)

    def getSize(self):
        return self.array.numel()

    def set(self, v):
        self.array.fill_(v)

    def axpy(self, alpha, x):
        self.array.add_(x.array, alpha=alpha)

    def dot(self, other):
        return _allreduce_sum(torch.dot(self.array, other.array), self.comm)

    def norm(self):
        return math.sqrt(float(self.dot(self)))


def _allreduce_sum(t, comm):
    if comm is not None:
        import torch.distributed as dist

        t = t.clone()
        dist.all_reduce(t, group=comm)
    return t


class Mat:
    """PETSc.Mat (AIJ) look-alike: device-resident CSR whose value array the
    engine fills in place (``assert J.handle == ctx._jac.petscmat.handle``,
    firedrake_ts/solving_utils.py:284)."""

    def __init__(self, rowptr, colidx, values, ncols, halo=None, space=None):
        self.rowptr, self.colidx, self.values = rowptr, colidx, values
        self.nrows = rowptr.numel() - 1
        self.ncols = ncols
        self.halo = halo  # forward exchange for the off-process columns (None on one rank)
        # dof layout of the rows / columns (local column of owned dof (m, i) is m*num_nodes + i for a
        # blocked space on a partition): needed to find the diagonal
        self.ncomp = 1 if space is None else space.ncomp
        self.layout = 0 if space is None or space.layout == "interleaved" else 1
        self.num_owned_nodes = self.nrows // self.ncomp if space is None else space.num_owned_nodes
        self.num_nodes = self.num_owned_nodes if space is None else space.num_nodes
        self._nullspace = self._transpose_nullspace = self._near_nullspace = None
        self._lu = None

    @property
    def handle(self):
        return id(self)

    def zeroEntries(self):
        self.values.zero_()

    def getValuesCSR(self):
        return self.rowptr, self.colidx, self.values

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.values.cpu().numpy(), self.colidx.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=(self.nrows, self.ncols))

    def mult(self, x_local: torch.Tensor) -> torch.Tensor:
        """y = A x for a *local* (owned+ghost) x; returns owned rows."""
        from . import linalg

        return linalg.csr_spmv(self.rowptr, self.colidx, self.values, x_local)

    def diagonal(self):
        from . import linalg

        return linalg.csr_diagonal(self.rowptr, self.colidx, self.values, self.num_owned_nodes, self.num_nodes,
                                   self.ncomp, self.layout)


class DM:
    """DMShell stand-in that carries the application context (the reference looks
    its ``_TSContext`` up through ``dmhooks.get_appctx(ts.getDM())``,
    firedrake_ts/solving_utils.py:245-246)."""

    def __init__(self, comm=None):
        self.comm = comm
        self._appctx = []


class SNES:
    def __init__(self):
        self.rtol, self.atol, self.stol, self.max_it = 1e-8, 1e-50, 1e-8, 50
        self.ksp = KSP()
        self.its = 0
        self.reason = 0


class KSP:
    def __init__(self):
        self.type, self.pc_type = "gmres", "jacobi"
        self.rtol, self.atol, self.max_it, self.restart = 1e-5, 1e-50, 10000, 30
        self.its = 0
        self.options = {}       # the full option dictionary (fieldsplit_* sub-options are read from it)
        self.prec = None        # pc_type fieldsplit: the block preconditioner built by the TS for this solve


class ConvergedReason:
    CONVERGED_ITERATING = 0
    CONVERGED_TIME = 1
    CONVERGED_ITS = 2
    DIVERGED_NONLINEAR_SOLVE = -1
    DIVERGED_STEP_REJECTED = -2


def ksp_solve(A, rhs: torch.Tensor, k, comm=None, P=None) -> torch.Tensor:
    """Solve A x = rhs with the Krylov method / preconditioner named by the KSP ``k``
    (cg | gmres | preonly, jacobi | none | lu).  The preconditioner is built from ``P`` when the
    caller registered a separate preconditioning matrix (``Jp``, reference solving_utils.py:120-122,
    310-312), otherwise from ``A``.  Shared by the Newton step and by the right-hand-side projection
    solver (reference solving_utils.py:429-458)."""
    from . import linalg

    if P is None:
        P = A

    def to_local(v):
        return v if A.halo is None else A.halo.owned_to_local(v)

    def matvec(v):
        return A.mult(to_local(v))

    def dot(a, b):
        return float(_allreduce_sum(torch.dot(a, b), comm))

    if k.pc_type == "lu":
        assert comm is None, "pc_type lu is a single-rank test facility"
        import numpy as np
        import scipy.sparse.linalg as spla

        lu = spla.splu(P.to_scipy().tocsc())

        def lu_apply(r):
            return torch.from_numpy(np.ascontiguousarray(lu.solve(r.cpu().numpy()))).to(r.device)

        if P is A or k.type == "preonly":  # direct solve (with P != A: one application of P^-1, as PETSc's preonly)
            k.its = 1
            return lu_apply(rhs)
        x, k.its = linalg.gmres(matvec, rhs, lu_apply, dot, k.rtol, k.atol, k.max_it, k.restart)
        return x
    if k.type not in ("cg", "gmres", "preonly"):
        raise NotImplementedError(f"ksp_type {k.type!r}: the stand-in KSP knows cg, gmres (and preonly with pc_type lu)")
    if k.type == "preonly":
        raise NotImplementedError("ksp_type preonly needs pc_type lu")
    if any(ns is not None for ns in (A._nullspace, A._transpose_nullspace)):
        raise NotImplementedError("nullspace projection is not part of the stand-in KSP (PETSc's KSP applies it in "
                                  "the real stack); attach no nullspace or solve with the Firedrake adapter")
    if k.pc_type == "jacobi":
        d = P.diagonal()
        if bool((d == 0).any()):
            raise ZeroDivisionError("pc_type jacobi: the matrix has a zero diagonal entry")
        dinv = 1.0 / d
        prec = lambda r: r * dinv  # noqa: E731
    elif k.pc_type == "none":
        prec = lambda r: r  # noqa: E731
    elif k.pc_type == "fieldsplit":
        if k.prec is None:
            raise NotImplementedError("pc_type fieldsplit needs the field sub-contexts of a mixed problem "
                                      "(_TSContext.split); it is set up by TS for the Newton systems only")
        prec = k.prec
    else:
        raise NotImplementedError(f"pc_type {k.pc_type!r}: the stand-in KSP knows jacobi, none, lu and fieldsplit")
    if k.type == "cg" and rhs.is_cuda:
        red = None if comm is None else (lambda t: _allreduce_sum(t, comm))
        x, k.its = linalg.pcg_device(matvec, rhs, prec, k.rtol, k.atol, k.max_it, allreduce=red)
    elif k.type == "cg":
        x, k.its = linalg.pcg(matvec, rhs, prec, dot, k.rtol, k.atol, k.max_it)
    else:
        x, k.its = linalg.gmres(matvec, rhs, prec, dot, k.rtol, k.atol, k.max_it, k.restart)
    return x


# IMEX Runge-Kutta tableaus (At: implicit part, A: explicit part) in PETSc's TSARKIMEX naming.
# Only schemes whose first stage is explicit and carries no implicit weight are listed, so the
# implicit slope of the first stage is never needed.
ARKIMEX_TABLEAUS = {
    # first-order IMEX Euler (forward Euler on G, backward Euler on F)
    "1": dict(At=[[0.0, 0.0], [0.0, 1.0]], A=[[0.0, 0.0], [1.0, 0.0]], bt=[0.0, 1.0], b=[1.0, 0.0]),
    # Ascher-Ruuth-Spiteri (1,2,2): implicit-explicit midpoint, second order
    "ars122": dict(At=[[0.0, 0.0], [0.0, 0.5]], A=[[0.0, 0.0], [0.5, 0.0]], bt=[0.0, 1.0], b=[0.0, 1.0]),
}


class TS:
    """BEULER / THETA time stepper with Newton + Krylov, PETSc callback protocol."""

    ConvergedReason = ConvergedReason

    class EquationType:
        IMPLICIT = 1

    TYPES = ("beuler", "theta", "arkimex")

    def __init__(self, comm=None):
        self.comm = comm
        self.dm = None
        self.snes = SNES()
        self.type, self.theta = "beuler", 1.0
        self.time, self.max_time, self.dt = 0.0, 1.0, 0.1  # PETSc's default step is 0.1
        self.max_steps = 5000
        self.exact_final_time = "stepover"
        self.steps = 0
        self.reason = 0
        self._ifunction = self._ijacobian = None
        self._F = self._J = self._P = None
        self._rhsfunction = self._rhsjacobian = None
        self._G = self._Jrhs = self._Prhs = None
        self.arkimex_type = "ars122"
        self.rhs_evals = 0
        self._monitor = None
        self.equation_type = None
        self.newton_its = 0
        self.linear_its = 0

    def _reached_end(self):
        """end-time test with an absolute-plus-relative tolerance tied to the step (works for max_time <= 0)"""
        return self.time >= self.max_time - 1e-12 * max(abs(self.dt), 1e-300) - 1e-14 * abs(self.max_time)

    # -- PETSc-like setters / getters -------------------------------------------
    def create(self, comm=None):
        self.comm = comm
        return self

    def getSNES(self):
        return self.snes

    def setDM(self, dm):
        self.dm = dm

    def getDM(self):
        return self.dm

    def setMonitor(self, monitor):
        self._monitor = monitor

    def setTime(self, t):
        self.time = float(t)

    def getTime(self):
        return self.time

    def setMaxTime(self, t):
        self.max_time = float(t)

    def setTimeStep(self, dt):
        self.dt = float(dt)

    def getTimeStep(self):
        return self.dt

    def setMaxSteps(self, n):
        self.max_steps = int(n)

    def getStepNumber(self):
        return self.steps

    def setEquationType(self, et):
        self.equation_type = et

    def setType(self, name):
        assert name in self.TYPES, f"TS type {name!r} is not provided by the stand-in"
        self.type = name
        if name == "beuler":
            self.theta = 1.0

    def setTheta(self, theta):
        self.theta = float(theta)

    def setIFunction(self, f, vec=None):
        self._ifunction, self._F = f, vec

    def setIJacobian(self, f, J=None, P=None):
        self._ijacobian, self._J, self._P = f, J, P

    def setRHSFunction(self, f, vec=None):
        self._rhsfunction, self._G = f, vec

    def setRHSJacobian(self, f, J=None, P=None):
        self._rhsjacobian, self._Jrhs, self._Prhs = f, J, P

    def setARKIMEXType(self, name):
        assert name in ARKIMEX_TABLEAUS, f"ARKIMEX type {name!r}: the stand-in provides {sorted(ARKIMEX_TABLEAUS)}"
        self.arkimex_type = name

    def getConvergedReason(self):
        return self.reason

    def setFromOptions(self, opts: dict):
        o = dict(opts or {})
        if "ts_type" in o:
            self.setType(o["ts_type"])
        if "ts_arkimex_type" in o:
            self.setARKIMEXType(o["ts_arkimex_type"])
        if "ts_theta_theta" in o:
            self.setTheta(o["ts_theta_theta"])
        if "ts_dt" in o:
            self.setTimeStep(o["ts_dt"])
        if "ts_max_steps" in o:
            self.setMaxSteps(o["ts_max_steps"])
        if "ts_exact_final_time" in o:
            self.exact_final_time = o["ts_exact_final_time"]
        s, k = self.snes, self.snes.ksp
        s.rtol = float(o.get("snes_rtol", s.rtol))
        s.atol = float(o.get("snes_atol", s.atol))
        s.stol = float(o.get("snes_stol", s.stol))
        s.max_it = int(o.get("snes_max_it", s.max_it))
        k.type = o.get("ksp_type", k.type)
        k.pc_type = o.get("pc_type", k.pc_type)
        k.rtol = float(o.get("ksp_rtol", k.rtol))
        k.atol = float(o.get("ksp_atol", k.atol))
        k.max_it = int(o.get("ksp_max_it", k.max_it))
        k.restart = int(o.get("ksp_gmres_restart", k.restart))
        k.options = o

    # -- linear solve ---------------------------------------------------------------
    def _linear_solve(self, rhs: torch.Tensor) -> torch.Tensor:
        k = self.snes.ksp
        if k.pc_type == "fieldsplit":
            k.prec = self._fieldsplit_preconditioner()
        return ksp_solve(self._J, rhs, k, self.comm, P=self._P)

    def _fieldsplit_preconditioner(self):
        """PCFIELDSPLIT over the fields of a mixed space, one split per field: the diagonal blocks come from
        the context's field sub-contexts (``_TSContext.split``, reference solving_utils.py:145-233, which
        PETSc reaches through the DM's field decomposition).  ``pc_fieldsplit_type`` additive (block Jacobi) or
        multiplicative (block Gauss-Seidel, PETSc's default); ``fieldsplit_<i>_pc_type`` jacobi | lu per block
        (one application each, i.e. ``fieldsplit_<i>_ksp_type preonly``).  Schur factorisations, which need a
        user Schur preconditioner as in examples/cahn-hilliard.py:73-112, are not part of the stand-in."""
        from . import dmhooks

        ctx = dmhooks.get_appctx(self.dm)
        k = self.snes.ksp
        kind = k.options.get("pc_fieldsplit_type", "multiplicative")
        if kind not in ("additive", "multiplicative"):
            raise NotImplementedError(f"pc_fieldsplit_type {kind!r}: additive and multiplicative only")
        V = ctx._x.function_space()
        nf = V.ncomp if getattr(ctx, "is_mixed", False) else 1
        splits = ctx.split(tuple((i,) for i in range(nf)))
        applies = []
        for i, sc in enumerate(splits):
            B = sc.extract_jacobian_block().petscmat
            pct = k.options.get(f"fieldsplit_{i}_pc_type", "jacobi")
            if pct == "jacobi":
                d = B.diagonal()
                if bool((d == 0).any()):
                    raise ZeroDivisionError(f"fieldsplit_{i}_pc_type jacobi: zero diagonal entry in block {i}")
                applies.append(lambda r, dinv=1.0 / d: r * dinv)
            elif pct == "lu":
                assert self.comm is None, "fieldsplit_<i>_pc_type lu is a single-rank test facility"
                import numpy as np
                import scipy.sparse.linalg as spla

                n = B.nrows
                lu = spla.splu(B.to_scipy()[:, :n].tocsc())
                applies.append(lambda r, lu=lu: torch.from_numpy(np.ascontiguousarray(lu.solve(r.cpu().numpy()))).to(r.device))
            else:
                raise NotImplementedError(f"fieldsplit_{i}_pc_type {pct!r}: jacobi and lu only")
        A = self._J
        rows = [sc.rows for sc in splits]

        def prec(r):
            z = torch.zeros_like(r)
            res = r
            for i, ap in enumerate(applies):
                z[rows[i]] = ap(res[rows[i]])
                if kind == "multiplicative" and i + 1 < len(applies):
                    zl = z if A.halo is None else A.halo.owned_to_local(z)
                    res = r - A.mult(zl)
            return z

        return prec

    # -- the integrator ---------------------------------------------------------------
    def _stage_solve(self, X: Vec, X0: torch.Tensor, ts_time: float, shift: float):
        s = self.snes
        Xdot = Vec(torch.empty_like(X.array), self.comm)
        F = self._F if self._F is not None else X.duplicate()
        fnorm0 = None
        for it in range(s.max_it + 1):
            torch.sub(X.array, X0, out=Xdot.array)
            Xdot.array.mul_(shift)
            self._ifunction(self, ts_time, X, Xdot, F)
            fnorm = F.norm()
            if fnorm0 is None:
                fnorm0 = fnorm
            if fnorm <= s.atol or fnorm <= s.rtol * fnorm0:
                s.its, s.reason = it, 1
                return Xdot
            if it == s.max_it:
                break
            self._ijacobian(self, ts_time, X, Xdot, shift, self._J, self._P)
            dx = self._linear_solve(-F.array)
            self.linear_its += s.ksp.its
            X.array.add_(dx)
            self.newton_its += 1
            xnorm = X.norm()
            if float(_allreduce_sum(torch.dot(dx, dx), self.comm)) ** 0.5 <= s.stol * xnorm:
                # step tolerance: re-evaluate once so that Xdot matches the final X
                torch.sub(X.array, X0, out=Xdot.array)
                Xdot.array.mul_(shift)
                s.its, s.reason = it + 1, 2
                return Xdot
        s.reason = -1
        return None

    def _solve_arkimex(self, u: Vec):
        """Fixed-step additive Runge-Kutta IMEX integrator for F(t, u, udot) = G(t, u) in PETSc's
        formulation (TSARKIMEX): every implicit stage solves F(t, Y, shift (Y - Z)) = 0 with the
        IFunction / IJacobian callbacks only, Z carrying the earlier implicit slopes and the
        explicit slopes YdotRHS_j = RHSFunction(t_j, Y_j).  As in PETSc, the RHSFunction value is
        used as a slope, which is why the reference projects G with the mass matrix first
        (solving_utils.py:447-458)."""
        tab = ARKIMEX_TABLEAUS[self.arkimex_type]
        At, A, bt, b = tab["At"], tab["A"], tab["bt"], tab["b"]
        ns = len(b)
        c = [sum(row) for row in A]
        ct = [sum(row) for row in At]
        Y = Vec(u.array.clone(), self.comm)
        G = self._G if self._G is not None else Y.duplicate()
        while not self._reached_end() and self.steps < self.max_steps:
            h = self.dt
            if self.exact_final_time == "matchstep" and self.time + h > self.max_time:
                h = self.max_time - self.time
            YdotI = [None] * ns
            YdotRHS = [None] * ns
            for i in range(ns):
                Z = u.array.clone()
                for j in range(i):
                    if At[i][j] != 0.0:
                        Z.add_(YdotI[j], alpha=h * At[i][j])
                    if A[i][j] != 0.0:
                        Z.add_(YdotRHS[j], alpha=h * A[i][j])
                if At[i][i] == 0.0:
                    Y.array.copy_(Z)  # explicit stage
                    YdotI[i] = None
                else:
                    Y.array.copy_(Z)
                    shift = 1.0 / (h * At[i][i])
                    Ydot = self._stage_solve(Y, Z, self.time + ct[i] * h, shift)
                    if Ydot is None:
                        self.reason = ConvergedReason.DIVERGED_NONLINEAR_SOLVE
                        return
                    YdotI[i] = Ydot.array.clone()
                if any(A[k][i] != 0.0 for k in range(i + 1, ns)) or b[i] != 0.0:
                    self._rhsfunction(self, self.time + c[i] * h, Y, G)
                    self.rhs_evals += 1
                    YdotRHS[i] = G.array.clone()
            for i in range(ns):
                if bt[i] != 0.0:
                    u.array.add_(YdotI[i], alpha=h * bt[i])
                if b[i] != 0.0:
                    u.array.add_(YdotRHS[i], alpha=h * b[i])
            self.time += h
            self.steps += 1
            if self._monitor:
                self._monitor(self, self.steps, self.time, u)

    def solve(self, u: Vec):
        assert self._ifunction is not None and self._ijacobian is not None
        self.steps, self.reason = 0, 0
        if self._rhsfunction is not None and self.type != "arkimex":
            raise NotImplementedError(f"TS type {self.type!r} with a right-hand-side function G: use ts_type arkimex")
        if self.type == "arkimex":
            assert self._rhsfunction is not None, "ts_type arkimex needs G (setRHSFunction)"
            if self._monitor:
                self._monitor(self, self.steps, self.time, u)
            self._solve_arkimex(u)
            if self.reason == 0:
                self.reason = (ConvergedReason.CONVERGED_TIME if self._reached_end()
                               else ConvergedReason.CONVERGED_ITS)
            return
        theta = self.theta
        if self._monitor:
            self._monitor(self, self.steps, self.time, u)
        X = Vec(u.array.clone(), self.comm)
        while not self._reached_end() and self.steps < self.max_steps:
            dt = self.dt
            if self.exact_final_time == "matchstep" and self.time + dt > self.max_time:
                dt = self.max_time - self.time
            X0 = u.array.clone()
            X.array.copy_(X0)
            shift = 1.0 / (theta * dt)
            Xdot = self._stage_solve(X, X0, self.time + theta * dt, shift)
            if Xdot is None:
                self.reason = ConvergedReason.DIVERGED_NONLINEAR_SOLVE
                return
            if theta == 1.0:
                u.array.copy_(X.array)
            else:
                u.array.add_(Xdot.array, alpha=dt)
            self.time += dt
            self.steps += 1
            if self._monitor:
                self._monitor(self, self.steps, self.time, u)
        self.reason = ConvergedReason.CONVERGED_TIME if self._reached_end() else ConvergedReason.CONVERGED_ITS
