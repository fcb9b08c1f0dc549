"""_TSContext: the callback/context layer (reference firedrake_ts/solving_utils.py).

Same names, argument meaning and error behaviour as the reference's
``_TSContext`` (solving_utils.py:42-317) for the IFunction / IJacobian path:
``set_ifunction``, ``set_ijacobian``, ``form_function(ts, t, X, Xdot, F)``,
``form_jacobian(ts, t, X, Xdot, shift, J, P)``, the four user callbacks, the
``J.handle`` assertion, the ``_jacobian_assembled`` flag and the
``_time``/``_shift`` Constants.  What changes is what sits behind
``_assemble_residual`` / ``_assemble_jac``: the B200 engine (engine.Engine)
writing straight into the device-resident Vec / AIJ value arrays, plus a halo
exchange in place of PyOP2's.  The IMEX RHS branch (solving_utils.py:320-384) and ``split``
(:145-233, field sub-contexts of mixed spaces for fieldsplit preconditioners) are here as well;
``compute_operators`` is out of scope.
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
                 options_prefix=None, transfer_manager=None, assembler="b200", comm=None,
                 project_rhs=True, rhs_projection_parameters=None, rhs_assembler=None, mass_assembler=None,
                 jac_du_assembler=None, pjac_assembler=None):
        if mat_type not in (None, "aij", "aijcusparse"):
            raise ValueError(f"mat_type {mat_type!r}: the B200 engine assembles monolithic AIJ only")
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
        self.G, self.dGdu = problem.G, problem.dGdu
        self.bcs_F = tuple(problem.dirichlet_bcs())
        self.bcs_G = self.bcs_dGdu = self.bcs_F
        self.fcp = problem.form_compiler_parameters
        self._jacobian_assembled = False
        self._nullspace = self._nullspace_T = self._near_nullspace = None
        self._splits = {}
        V = self._x.function_space()
        F_ = problem.F
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
        self._native_halo = False
        self._ghosts_fresh = True
        if V.mesh.partition.size > 1:
            from .halo import HaloExchange

            self._halo = HaloExchange(V, self.engine if assembler == "b200" else None, group=comm,
                                      device=self._device)
            # B200 engine on a CUDA device with NCCL: the exchange runs inside the library (its own communicator),
            # fused with the residual; the torch.distributed path stays for CPU/gloo runs and for mat-vecs
            self._native_halo = False
            if assembler == "b200" and self._device.type == "cuda" and self.appctx.get("native_halo", True):
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized() and dist.get_backend(comm) == "nccl":
                    self.engine.halo_setup(comm)
                    self._native_halo = True
        # residual Cofunction and Jacobian Matrix stand-ins (base class: _F, _jac, _pjac)
        self._F = Vec(torch.zeros(V.num_owned_dofs, dtype=torch.float64, device=self._device), comm)
        self._jac = _Assembled(Mat(rowptr, colidx, values, V.num_dofs, halo=self._halo, space=V))
        self._pjac = self._jac
        self._assemble_residual = self._b200_assemble_residual
        self._assemble_jac = self._b200_assemble_jac

        def side_engine(form, test_assembler, what):
            """an engine of its own for a form that is not F (user J, Jp): same space, same sparsity,
            same Dirichlet rows"""
            if form.space is not V:
                raise ValueError(f"{what} must live on the function space of F")
            if assembler == "b200":
                from .engine import Engine

                return Engine(forms.Form(form.family, self._x, self._xdot, form.params), self.bcs_F,
                              device=self._device, scatter=self.appctx.get("scatter"))
            if test_assembler is None:
                raise ValueError(f"a test assembler for F needs an assembler for {what} as well")
            return test_assembler

        # user-provided J: J = shift * dF/dudot + J_du  (reference ts_solver.py:108)
        self._jac_du_engine = self._mass_engine = None
        Ju = getattr(problem, "J_du", None)
        if Ju is not None and not (Ju.family == F_.family and tuple(Ju.params) == tuple(F_.params)):
            self._jac_du_engine = side_engine(Ju, jac_du_assembler, "J")
            self._mass_engine = side_engine(forms.mass_form(F_), mass_assembler, "derivative(F, udot)")
            self._jac_scratch = torch.zeros_like(values)
            self._bc_diag_slots = self._find_bc_diagonal_slots(rowptr, colidx, V)
        # preconditioning form Jp: a Mat of its own, assembled after J (reference :310-312)
        self._pjac_engine = None
        if self.Jp is not None:
            self._pjac_engine = side_engine(self.Jp, pjac_assembler, "Jp")
            self._pjac = _Assembled(Mat(rowptr, colidx, torch.zeros_like(values), V.num_dofs, halo=self._halo, space=V))

        # IMEX branch (reference :104-111): G is assembled by an engine of its own on the same
        # space, so it shares the sparsity and the Dirichlet rows of F
        if self.G is not None:
            if assembler == "b200":
                from .engine import Engine

                self.rhs_engine = Engine(problem.G, self.bcs_G, device=self._device,
                                         scatter=self.appctx.get("scatter"))
            else:
                if rhs_assembler is None:
                    raise ValueError("a test assembler for F needs rhs_assembler= for G")
                self.rhs_engine = rhs_assembler
            self._mass_assembler = mass_assembler
            self.rhs_projection_parameters = rhs_projection_parameters
            self.project_rhs = project_rhs
            self._G_vec = Vec(torch.zeros(V.num_owned_dofs, dtype=torch.float64, device=self._device), comm)
            self._projected_G = Vec(torch.zeros_like(self._G_vec.array), comm)
            self._G_or_projected_G = self._projected_G if self.project_rhs else self._G_vec
            self._rhs_jac = _Assembled(Mat(rowptr, colidx, torch.zeros_like(values), V.num_dofs, halo=self._halo, space=V))
            self._rhs_pjac = self._rhs_jac
            self._dGdu = self._rhs_jac
            self._rhs_jacobian_assembled = False
            self._rhs_projection_solver_cache = None

    # -- the two assemblers that replace get_assembler(...).assemble ----------------
    def _b200_assemble_residual(self, tensor=None):
        tensor = tensor if tensor is not None else self._F
        if not self._ghosts_fresh:
            # exchange deferred by form_function: one library call packs, sends/receives on the communication
            # stream, assembles the interior cells meanwhile, unpacks and assembles the remaining cells
            self._ghosts_fresh = True
            self.engine.ifunction_exchange(float(self._time), self._x.data, self._xdot.data, out=tensor.array)
            return tensor
        out = self.engine.ifunction(float(self._time), self._x.data, self._xdot.data, out=tensor.array)
        if out is not tensor.array:
            tensor.array.copy_(torch.as_tensor(out))
        return tensor

    def _b200_assemble_jac(self, jac):
        mat = jac.petscmat
        if self._jac_du_engine is not None:
            # shift * derivative(F, udot) + J_du: mass matrix of F at the current shift plus the user's
            # bilinear form (its own udot term switched off); both assemblies put 1 on the Dirichlet diagonal
            t, x, xd = float(self._time), self._x.data, self._xdot.data
            self._fill(self._mass_engine.ijacobian(t, x, xd, float(self._shift), out=mat.values), mat.values)
            self._fill(self._jac_du_engine.ijacobian(t, x, xd, 0.0, out=self._jac_scratch), self._jac_scratch)
            mat.values.add_(self._jac_scratch)
            if self._bc_diag_slots.numel():
                mat.values[self._bc_diag_slots] = 1.0
            return jac
        out = self.engine.ijacobian(float(self._time), self._x.data, self._xdot.data, float(self._shift),
                                    out=mat.values)
        if out is not mat.values:
            mat.values.copy_(torch.as_tensor(out))
        return jac

    @staticmethod
    def _fill(out, target):
        if out is not target:
            target.copy_(torch.as_tensor(out))

    def _assemble_pjac(self, pjac):
        """assemble(problem.Jp) into the preconditioning matrix (reference :310-312): the form is taken
        verbatim, its own shift constant (problem.shift when the user built it from that) applies"""
        mat = pjac.petscmat
        sh = self.Jp.shift if self.Jp.shift is not None else self._shift
        self._fill(self._pjac_engine.ijacobian(float(self._time), self._x.data, self._xdot.data, float(sh),
                                               out=mat.values), mat.values)
        return pjac

    def _find_bc_diagonal_slots(self, rowptr, colidx, V):
        """CSR slots of the diagonal entries of the Dirichlet rows (owned dofs)"""
        nodes = [bc.nodes.to(torch.int64) for bc in self.bcs_F]
        if not nodes:
            return torch.zeros(0, dtype=torch.int64, device=rowptr.device)
        nodes = torch.unique(torch.cat(nodes))
        nodes = nodes[nodes < V.num_owned_nodes].to(rowptr.device)
        slots = []
        for m in range(V.ncomp):
            rows = V.dof_index(nodes, m, V.num_owned_nodes)
            cols = V.dof_index(nodes, m)
            for r, c in zip(rows.tolist(), cols.tolist()):
                seg = colidx[int(rowptr[r]):int(rowptr[r + 1])]
                k = int(torch.nonzero(seg == c)[0])
                slots.append(int(rowptr[r]) + k)
        return torch.tensor(slots, dtype=torch.int64, device=rowptr.device)

    def _refresh_halo(self, defer=False):
        """ghost update of u and udot (PyOP2 does it when the reference leaves dat.vec_wo, :249-252).  With the
        library's own exchange and ``defer`` the update is folded into the residual assembly that follows."""
        if self._halo is None:
            return
        if self._native_halo:
            if defer:
                self._ghosts_fresh = False
            else:
                self.engine.halo_exchange(self._x.data, self._xdot.data)
            return
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
        r"""Set the function to compute G(t,u) where F() = G() is the equation to be solved
        (reference :124-128)."""
        if self.G is not None:
            ts.setRHSFunction(self.form_rhs_function, self._G_or_projected_G)

    def set_rhs_jacobian(self, ts):
        r"""Set the function to compute the Jacobian of G (reference :130-133)."""
        if self.G is not None:
            ts.setRHSJacobian(self.form_rhs_jacobian, J=self._rhs_jac.petscmat, P=self._rhs_pjac.petscmat)

    # -- IMEX assemblers (reference :429-458, :462-485) ---------------------------------------
    def _assemble_rhs_residual(self, tensor):
        out = self.rhs_engine.ifunction(float(self._time), self._x.data, self._xdot.data, out=tensor.array)
        if out is not tensor.array:
            tensor.array.copy_(torch.as_tensor(out))
        return tensor

    def _assemble_rhs_jac(self, jac):
        mat = jac.petscmat
        out = self.rhs_engine.ijacobian(float(self._time), self._x.data, self._xdot.data, 0.0, out=mat.values)
        if out is not mat.values:
            mat.values.copy_(torch.as_tensor(out))
        return jac

    @property
    def _rhs_projection_solver(self):
        r"""mass matrix = derivative(F, udot) assembled once with bcs_F, and the KSP that solves
        with it (reference :429-445).  Default parameters: Jacobi-preconditioned CG to 1e-12 on
        the device CSR (the reference's default is a direct solve; ``pc_type: lu`` gives that on
        one rank)."""
        if self.G is None:
            return None
        if self._rhs_projection_solver_cache is None:
            from . import forms
            from .ts import KSP, ksp_solve

            F = self._problem.F
            is_pure_mass = F.family == forms.DIFFUSION and F.params[2] == 0.0
            if is_pure_mass:
                eng = self.engine
            elif self._mass_assembler is not None:
                eng = self._mass_assembler
            else:
                from .engine import Engine

                eng = Engine(forms.mass_form(F), self.bcs_F, device=self._device)
            base = self._jac.petscmat
            vals = torch.zeros_like(base.values)
            out = eng.ijacobian(0.0, self._x.data, self._xdot.data, 1.0, out=vals)
            if out is not vals:
                vals.copy_(torch.as_tensor(out))
            M = Mat(base.rowptr, base.colidx, vals, base.ncols, halo=self._halo, space=self._x.function_space())
            k = KSP()
            k.type, k.pc_type, k.rtol, k.atol = "cg", "jacobi", 1e-12, 1e-50
            o = dict(self.rhs_projection_parameters or {})
            k.type = o.get("ksp_type", k.type)
            k.pc_type = o.get("pc_type", k.pc_type)
            k.rtol = float(o.get("ksp_rtol", k.rtol))
            k.atol = float(o.get("ksp_atol", k.atol))
            k.max_it = int(o.get("ksp_max_it", k.max_it))
            comm = self.comm

            class _Solver:
                mass_matrix, ksp = M, k

                @staticmethod
                def solve(x, bvec):
                    x.array.copy_(ksp_solve(M, bvec.array, k, comm))

            self._rhs_projection_solver_cache = _Solver
        return self._rhs_projection_solver_cache

    def _assemble_projected_rhs_residual(self):
        """solve(mass_matrix_form == self.G, self._projected_G, self.bcs_G)  (reference :447-458)"""
        if self.G is not None:
            self._assemble_rhs_residual(self._G_vec)
            if self.project_rhs:
                self._rhs_projection_solver.solve(self._projected_G, self._G_vec)

    # -- field sub-contexts (reference :145-233) ---------------------------------------------------
    def split(self, fields):
        r"""Contexts of the sub-problems on ``fields`` (an iterable of field-index tuples) of a mixed space:
        each one assembles the rows ``field`` of F and the diagonal block (field, field) of J, with the other
        fields of the solution entering as coefficients (their data are shared with this context's
        Functions, as the reference shares the Dats).  Cached per ``fields`` like the reference's
        ``self._splits``.  The engine assembles the monolithic system once and the sub-context reads its
        rows / its block out of it, so no second set of kernels is needed."""
        fields = tuple(tuple(int(i) for i in f) for f in fields)
        splits = self._splits.get(fields)
        if splits is not None:
            return splits
        V = self._x.function_space()
        nfields = V.ncomp if self.is_mixed else 1
        for f in fields:
            if not f or any(i < 0 or i >= nfields for i in f) or len(set(f)) != len(f):
                raise ValueError(f"split: field tuple {f} is not a subset of the {nfields} fields of the space")
        splits = [_SplitContext(self, f) for f in fields]
        return self._splits.setdefault(fields, splits)

    def set_nullspace(self, nullspace, ises=None, transpose=False, near=False):
        # three separate attachments, as MatSetNullSpace / MatSetTransposeNullSpace / MatSetNearNullSpace
        # keep them (reference solving_utils.py:314-317 re-attaches all three after every assembly)
        if nullspace is None:
            return
        attr = "_near_nullspace" if near else ("_transpose_nullspace" if transpose else "_nullspace")
        for mat in {id(self._jac.petscmat): self._jac.petscmat, id(self._pjac.petscmat): self._pjac.petscmat}.values():
            setattr(mat, attr, nullspace)

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
        # a user pre-callback may read the ghosts: exchange first in that case, otherwise inside the assembly
        ctx._refresh_halo(defer=ctx._pre_function_callback is None)
        ctx._time.assign(t)

        if ctx._pre_function_callback is not None:
            ctx._pre_function_callback(X, Xdot)

        ctx._assemble_residual(tensor=ctx._F)

        if ctx._post_function_callback is not None:
            ctx._post_function_callback(X, Xdot, ctx._F)

        # F may not be the same vector as self._F, so copy residual out to F.
        ctx._F.copy(F)

    @staticmethod
    def form_rhs_function(ts, t, X, G):
        r"""Form the right-hand side G(t, u) for this problem (reference :320-347)

        :arg ts: a TS object
        :arg t: the time at step/stage being solved
        :arg X: state vector
        :arg G: function vector
        """
        dm = ts.getDM()
        ctx = dmhooks.get_appctx(dm)
        # X may not be the same vector as the vec behind self._x, so copy guess in from X.
        ctx._x.set_owned(X.array)
        if ctx._halo is not None:
            ctx._halo.forward(ctx._x.data)
        ctx._time.assign(t)

        ctx._assemble_projected_rhs_residual()

        # G may not be the same vector as self._projected_G, so copy residual out to G.
        ctx._G_or_projected_G.copy(G)

    @staticmethod
    def form_rhs_jacobian(ts, t, X, J, P):
        r"""Form the Jacobian dG/du for this problem (reference :349-384)

        :arg ts: a TS object
        :arg t: the time at step/stage being solved
        :arg X: state vector
        :arg J: the Jacobian (a Mat)
        :arg P: the preconditioner matrix (a Mat)
        """
        dm = ts.getDM()
        ctx = dmhooks.get_appctx(dm)
        problem = ctx._problem

        assert J.handle == ctx._rhs_jac.petscmat.handle
        if problem._constant_rhs_jacobian and ctx._rhs_jacobian_assembled:
            # Don't need to do any work with a constant jacobian that's already assembled
            return
        ctx._rhs_jacobian_assembled = True

        ctx._x.set_owned(X.array)
        if ctx._halo is not None:
            ctx._halo.forward(ctx._x.data)
        ctx._time.assign(t)

        ctx._assemble_rhs_jac(ctx._dGdu)

        ctx.set_nullspace(ctx._nullspace, None, transpose=False, near=False)
        ctx.set_nullspace(ctx._nullspace_T, None, transpose=True, near=False)
        ctx.set_nullspace(ctx._near_nullspace, None, transpose=False, near=True)

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

        if ctx.Jp is not None:
            assert P.handle == ctx._pjac.petscmat.handle
            ctx._assemble_pjac(ctx._pjac)

        ctx.set_nullspace(ctx._nullspace, None, transpose=False, near=False)
        ctx.set_nullspace(ctx._nullspace_T, None, transpose=True, near=False)
        ctx.set_nullspace(ctx._near_nullspace, None, transpose=False, near=True)


class _SplitContext:
    """Sub-problem of a mixed problem on the fields ``field`` (reference ``_TSContext.split``,
    solving_utils.py:145-233): same callback protocol as ``_TSContext``; ``_x`` / ``_xdot`` are Functions on the
    sub-space whose values are written through to the parent's Functions; ``_jac`` holds the diagonal block
    (field, field) of the parent's AIJ pattern."""

    def __init__(self, parent, field):
        from . import mesh as M
        from .function import Function

        self.parent, self.field = parent, tuple(field)
        self._problem = parent._problem
        self.appctx, self.comm = parent.appctx, parent.comm
        self.mat_type, self.pmat_type = parent.mat_type, parent.pmat_type
        V = parent._x.function_space()
        nf = len(self.field)
        dev = parent._device
        self._pre_jacobian_callback = self._pre_function_callback = None
        self._post_jacobian_callback = self._post_function_callback = None
        self._jacobian_assembled = False
        self._time, self._shift = parent._time, parent._shift
        self.Jp = None
        if not parent.is_mixed:  # scalar / vector space: the only field is the whole problem
            self.space = V
            self.rows = torch.arange(V.num_owned_dofs, device=dev)
            self.cols_local = torch.arange(V.num_dofs, device=dev)
        else:
            cache = V.__dict__.setdefault("_sub_spaces", {})
            if nf not in cache:
                cache[nf] = M.FunctionSpace(V.mesh, V.degree, ncomp=nf, layout="blocked" if nf > 1 else "interleaved",
                                            variant=V.variant, tile=V.tile, tile_offset=V.tile_offset)
            self.space = cache[nf]
            nown, nn = V.num_owned_nodes, V.num_nodes
            self.rows = torch.cat([torch.arange(m * nown, (m + 1) * nown, device=dev) for m in self.field])
            self.cols_local = torch.cat([torch.arange(m * nn, (m + 1) * nn, device=dev) for m in self.field])
        Vs = self.space
        self._x, self._xdot = Function(Vs, device=dev), Function(Vs, device=dev)
        self._pull()
        # diagonal block of the parent's pattern: entries in the rows of `field` whose column lies in `field`
        pm = parent._jac.petscmat
        rp, ci = pm.rowptr, pm.colidx.to(torch.int64)
        newcol = torch.full((V.num_dofs,), -1, dtype=torch.int64, device=ci.device)
        newcol[self.cols_local.to(ci.device)] = torch.arange(self.cols_local.numel(), device=ci.device)
        rows = self.rows.to(rp.device)
        lens = rp[rows + 1] - rp[rows]
        start = torch.repeat_interleave(rp[rows], lens)
        within = torch.arange(int(lens.sum()), device=rp.device) - torch.repeat_interleave(torch.cumsum(lens, 0) - lens, lens)
        entry = start + within  # parent CSR slots of all entries in the selected rows, row by row
        keep = newcol[ci[entry]] >= 0
        rowid = torch.repeat_interleave(torch.arange(rows.numel(), device=rp.device), lens)[keep]
        self.slots = entry[keep]
        sub_rp = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rp.device)
        sub_rp[1:] = torch.cumsum(torch.bincount(rowid, minlength=rows.numel()), 0)
        sub_ci = newcol[ci[self.slots]].to(torch.int32)
        values = torch.zeros(self.slots.numel(), dtype=torch.float64, device=pm.values.device)
        halo = None
        if parent._halo is not None:
            from .halo import HaloExchange

            halo = HaloExchange(Vs, None, group=parent.comm, device=dev)
        self._halo = halo
        self._F = Vec(torch.zeros(rows.numel(), dtype=torch.float64, device=dev), parent.comm)
        self._jac = _Assembled(Mat(sub_rp, sub_ci, values, Vs.num_dofs, halo=halo, space=Vs))
        self._pjac = self._jac

    # values of the sub-Functions <-> the parent's Functions (local vectors, ghosts included)
    def _pull(self):
        self._x.data.copy_(self.parent._x.data[self.cols_local])
        self._xdot.data.copy_(self.parent._xdot.data[self.cols_local])

    def _push(self):
        self.parent._x.data[self.cols_local] = self._x.data
        self.parent._xdot.data[self.cols_local] = self._xdot.data

    def extract_jacobian_block(self):
        """copy the block out of the parent's assembled Jacobian (no assembly)"""
        self._jac.petscmat.values.copy_(self.parent._jac.petscmat.values[self.slots])
        return self._jac

    def set_ifunction(self, ts):
        ts.setIFunction(self.form_function, self._F)

    def set_ijacobian(self, ts):
        ts.setIJacobian(self.form_jacobian, J=self._jac.petscmat, P=self._pjac.petscmat)

    def _assemble_parent(self, t, X, Xdot):
        p = self.parent
        self._x.set_owned(X.array)
        self._xdot.set_owned(Xdot.array)
        self._push()
        p._refresh_halo()
        p._time.assign(t)

    @staticmethod
    def form_function(ts, t, X, Xdot, F):
        ctx = dmhooks.get_appctx(ts.getDM())
        p = ctx.parent
        ctx._assemble_parent(t, X, Xdot)
        p._assemble_residual(tensor=p._F)
        ctx._F.array.copy_(p._F.array[ctx.rows])
        ctx._F.copy(F)

    @staticmethod
    def form_jacobian(ts, t, X, Xdot, shift, J, P):
        ctx = dmhooks.get_appctx(ts.getDM())
        p = ctx.parent
        assert J.handle == ctx._jac.petscmat.handle
        ctx._jacobian_assembled = True
        ctx._assemble_parent(t, X, Xdot)
        p._shift.assign(shift)
        p._assemble_jac(p._jac)
        ctx.extract_jacobian_block()
