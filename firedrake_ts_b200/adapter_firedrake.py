"""Drop-in adapter for a real Firedrake + PETSc(CUDA) installation.

IMPORTABLE BUT NOT EXECUTABLE IN THIS IMAGE: firedrake, pyop2, ufl and petsc4py are absent
(SURVEY.md section 0 item 4), so everything below is written against the accessor names listed
in SURVEY.md Appendix A.2 and must be verified the first time a real install is available.
Nothing in tests/, bench.py or the standalone API imports this module.

What it does
  * ``B200TSContext`` subclasses the reference's ``firedrake_ts.solving_utils._TSContext`` and
    overrides the two static callbacks ``form_function`` / ``form_jacobian``
    (firedrake_ts/solving_utils.py:236-317).  Everything else of the context -- registration
    (``set_ifunction`` / ``set_ijacobian``), nullspaces, ``split``, the G branch -- is inherited.
  * the engine is built from the live problem: cell->node map, coordinates, DirichletBC nodes,
    halo star forest (Appendix A.2) and is handed the *device* arrays of the PETSc objects
    (``Vec.getCUDAHandle`` / MatSeqAIJCUSPARSE CSR values), so the residual and the Jacobian
    values are written in place with no host round trip.
  * ``use_b200(solver)`` swaps the context of an already constructed ``firedrake_ts.DAESolver``
    and switches its work Vec / Mat types (``ts_solver.py:239``, ``solving_utils.py:122``).
"""
from __future__ import annotations

try:  # pragma: no cover - needs a Firedrake install
    import firedrake  # noqa: F401
    from firedrake import dmhooks
    from firedrake_ts.solving_utils import _TSContext as _RefTSContext
    HAVE_FIREDRAKE = True
except Exception:  # pragma: no cover
    HAVE_FIREDRAKE = False
    _RefTSContext = object
    dmhooks = None

from . import fem, forms


def classify_form(F, u, udot):  # pragma: no cover - needs UFL
    """Recognise which catalogue family a UFL residual belongs to.  The engine has no form
    compiler: unsupported forms keep Firedrake's own assembler (the callbacks fall through to
    the reference implementation)."""
    import ufl
    from ufl.algorithms import expand_derivatives

    sig = expand_derivatives(F).signature()
    V = u.function_space()
    table = {}
    v = ufl.TestFunction(V)
    candidates = {
        forms.HEAT: lambda: ufl.inner(udot, v) * ufl.dx + ufl.inner(ufl.grad(u), ufl.grad(v)) * ufl.dx - 1.0 * v * ufl.dx,
        forms.BBM: lambda: (ufl.inner(udot, v) + ufl.inner(u.dx(0), v) + ufl.inner(u * u.dx(0), v)
                            + ufl.inner(udot.dx(0), v.dx(0))) * ufl.dx,
    }
    for fam, make in candidates.items():
        try:
            table[expand_derivatives(make()).signature()] = fam
        except Exception:
            pass
    return table.get(sig)


class _LiveSpace:  # pragma: no cover - needs a Firedrake install
    """Adapts a Firedrake FunctionSpace to the attributes engine.Engine reads from
    firedrake_ts_b200.mesh.FunctionSpace."""

    def __init__(self, V):
        import numpy as np
        import torch

        mesh = V.mesh()
        self.dim = mesh.geometric_dimension()
        self.degree = V.ufl_element().degree()
        self.variant = "gll" if self.degree >= 3 else "equispaced"
        self.ncomp = V.value_size if len(V) == 1 else len(V)
        self.layout = "interleaved" if len(V) == 1 else "blocked"
        self.nd = fem.num_nodes(self.dim, self.degree)
        scalar = V if len(V) == 1 else V.sub(0)
        self.cell_node_map = torch.from_numpy(np.ascontiguousarray(scalar.cell_node_map().values_with_halo))
        self.num_owned_nodes = scalar.dof_dset.size
        self.num_nodes = scalar.dof_dset.total_size
        self.num_ghost_nodes = self.num_nodes - self.num_owned_nodes
        self.num_dofs = self.num_nodes * self.ncomp
        self.num_owned_dofs = self.num_owned_nodes * self.ncomp

        class _M:
            pass

        m = _M()
        cs = mesh.coordinates.function_space()
        m.coords = torch.from_numpy(np.ascontiguousarray(mesh.coordinates.dat.data_ro_with_halos.reshape(-1, self.dim)))
        m.cell_coord_map = torch.from_numpy(np.ascontiguousarray(cs.cell_node_map().values_with_halo))
        m.num_cells = m.cell_coord_map.shape[0]
        m.num_interior_cells = mesh.cell_set.core_size
        m.num_coord_nodes = m.coords.shape[0]
        m.device = torch.device("cuda")
        self.mesh = m

    def row_block_starts(self):
        return None  # Firedrake's numbering: uniform row blocks (engine default)


class B200TSContext(_RefTSContext):  # pragma: no cover - needs a Firedrake install
    """``_TSContext`` whose IFunction / IJacobian callbacks run on the B200 engine."""

    def __init__(self, problem, *args, **kwargs):
        super().__init__(problem, *args, **kwargs)
        from .engine import Engine

        family = classify_form(problem.F, problem.u_restrict, problem.udot)
        self._b200 = None
        if family is not None and problem.Jp is None:
            space = _LiveSpace(problem.u_restrict.function_space())

            class _F:
                pass

            f = _F()
            f.family, f.space, f.params = family, space, (1.0, 0.0, 0.0, 0.0)
            f.qdegree = forms.quadrature_degree(family, space.degree)

            class _BC:
                def __init__(self, nodes):
                    import torch
                    self.nodes = torch.as_tensor(nodes)

            bcs = [_BC(bc.nodes) for bc in problem.dirichlet_bcs()]
            self._b200 = Engine(f, bcs)

    @staticmethod
    def form_function(ts, t, X, Xdot, F):
        ctx = dmhooks.get_appctx(ts.getDM())
        if ctx._b200 is None:
            return _RefTSContext.form_function(ts, t, X, Xdot, F)
        ctx._time.assign(t)
        if ctx._pre_function_callback is not None:
            ctx._pre_function_callback(X, Xdot)
        # device pointers straight from the PETSc Vecs (VECCUDA): no host copies
        import ctypes as C
        from .engine import _check, _stream
        hx, hxd, hf = X.getCUDAHandle("r"), Xdot.getCUDAHandle("r"), F.getCUDAHandle("w")
        _check(ctx._b200._lib.fdts_ifunction(ctx._b200._h, float(t), C.c_void_p(hx), C.c_void_p(hxd), C.c_void_p(hf),
                                             _stream()), "fdts_ifunction")
        X.restoreCUDAHandle(hx, "r"); Xdot.restoreCUDAHandle(hxd, "r"); F.restoreCUDAHandle(hf, "w")
        if ctx._post_function_callback is not None:
            ctx._post_function_callback(X, Xdot, F)

    @staticmethod
    def form_jacobian(ts, t, X, Xdot, shift, J, P):
        ctx = dmhooks.get_appctx(ts.getDM())
        if ctx._b200 is None:
            return _RefTSContext.form_jacobian(ts, t, X, Xdot, shift, J, P)
        assert J.handle == ctx._jac.petscmat.handle
        ctx._jacobian_assembled = True
        ctx._time.assign(t)
        if ctx._pre_jacobian_callback is not None:
            ctx._pre_jacobian_callback(X, Xdot)
        ctx._shift.assign(shift)
        import ctypes as C
        from .engine import _check, _stream
        hx, hxd = X.getCUDAHandle("r"), Xdot.getCUDAHandle("r")
        vals = _mat_values_device_pointer(J)  # MatSeqAIJGetCSRAndMemType through PETSc's C API
        _check(ctx._b200._lib.fdts_ijacobian(ctx._b200._h, float(t), C.c_void_p(hx), C.c_void_p(hxd), float(shift),
                                             C.c_void_p(vals), _stream()), "fdts_ijacobian")
        X.restoreCUDAHandle(hx, "r"); Xdot.restoreCUDAHandle(hxd, "r")
        J.assemble()
        if ctx._post_jacobian_callback is not None:
            ctx._post_jacobian_callback(X, Xdot, J)
        ises = ctx._problem.J.arguments()[0].function_space()._ises
        ctx.set_nullspace(ctx._nullspace, ises, transpose=False, near=False)
        ctx.set_nullspace(ctx._nullspace_T, ises, transpose=True, near=False)
        ctx.set_nullspace(ctx._near_nullspace, ises, transpose=False, near=True)


def _mat_values_device_pointer(mat):  # pragma: no cover - needs libpetsc
    """device address of the AIJCUSPARSE value array (INTEGRATION.md shows the C stub)."""
    import ctypes as C
    lib = C.CDLL("libpetsc.so")  # the PETSc the running petsc4py is linked against
    i, j, a, mtype = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
    lib.MatSeqAIJGetCSRAndMemType(C.c_void_p(mat.handle), C.byref(i), C.byref(j), C.byref(a), C.byref(mtype))
    return a.value


def use_b200(solver):  # pragma: no cover - needs a Firedrake install
    """Switch an existing firedrake_ts.DAESolver to the B200 engine (same problem, same options)."""
    ctx = solver._ctx
    new = B200TSContext(ctx._problem, mat_type="aij", pmat_type="aij", appctx=ctx.appctx,
                        pre_jacobian_callback=ctx._pre_jacobian_callback,
                        pre_function_callback=ctx._pre_function_callback,
                        post_jacobian_callback=ctx._post_jacobian_callback,
                        post_function_callback=ctx._post_function_callback,
                        options_prefix=ctx.options_prefix)
    new._jac.petscmat.setType("aijcusparse")
    solver._ctx = new
    solver._work.setType("cuda")
    return solver
