"""Drop-in adapter: the B200 engine behind the reference's own ``_TSContext`` callbacks.

firedrake, pyop2, ufl and petsc4py are absent from this image, so the adapter
is written against *accessor names* only and never imports them at module level
beyond an optional probe.  ``tests/test_adapter_fake.py`` executes every function below on the GPU with
PETSc-/Firedrake-shaped stand-ins (``tests/fake_petsc.py``: a Vec with ``getCUDAHandle``, a Mat with
``handle`` / ``getRowIJ`` / ``assemble``, Functions with ``dat.vec_wo``, a stub ``_TSContext`` base carrying
the reference's attribute names) and compares the results with the oracle; only ``classify_form`` (needs
UFL) and ``mat_values_device_pointer`` (needs libpetsc) stay unexecuted here.

What it does
  * ``make_context_class(base)`` derives from the reference's ``firedrake_ts.solving_utils._TSContext`` a
    class whose static callbacks ``form_function`` / ``form_jacobian`` (firedrake_ts/solving_utils.py:236-317)
    run on the engine.  Everything else of the context -- registration (``set_ifunction`` / ``set_ijacobian``),
    nullspaces, ``split``, the G branch, ``Jp`` assembly -- is inherited.
  * the engine is built from the live problem: cell->node map, coordinates, DirichletBC nodes and the halo
    star forest, and is handed the *device* arrays of the PETSc objects (``Vec.getCUDAHandle``,
    MatSeqAIJCUSPARSE values), so residual and Jacobian values are written in place.
  * PETSc's TS passes GLOBAL vectors (owned dofs); the kernels read LOCAL vectors (owned + ghost).  The
    callbacks copy the owned part into persistent device buffers (``fdts_owned_to_local``) and, on more than
    one rank, refresh the ghosts over NCCL (``fdts_halo_exchange``), which replaces the PyOP2 halo exchange
    the reference triggers at solving_utils.py:249-252.
  * ``use_b200(solver)`` swaps the context of an already constructed ``firedrake_ts.DAESolver`` and switches
    its work Vec / Mat types (``ts_solver.py:239``, ``solving_utils.py:122``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

try:  # pragma: no cover - needs a Firedrake install
    import firedrake  # noqa: F401
    from firedrake import dmhooks
    from firedrake_ts.solving_utils import _TSContext as _RefTSContext
    HAVE_FIREDRAKE = True
except Exception:
    HAVE_FIREDRAKE = False
    _RefTSContext = None
    dmhooks = None  # tests install a stand-in with get_appctx(dm)

from . import fem, forms


# ------------------------------------------------------------------------------------------------ forms
def classify_form(F, u, udot):  # pragma: no cover - needs UFL
    """Recognise which catalogue family a UFL residual belongs to: returns (family, params) or None.
    The engine has no form compiler: unsupported forms keep Firedrake's own assembler (the callbacks fall
    through to the reference implementation).  Matching is by UFL signature after expand_derivatives, against
    the examples' forms rebuilt on the same arguments; constants are read off the candidate that matches."""
    import ufl
    from ufl.algorithms import expand_derivatives

    sig = expand_derivatives(F).signature()
    V = u.function_space()
    inner, grad, dx, dot = ufl.inner, ufl.grad, ufl.dx, ufl.dot

    def heat(src):  # examples/heat.py:11
        v = ufl.TestFunction(V)
        return inner(udot, v) * dx + inner(grad(u), grad(v)) * dx - src * v * dx

    def bbm():  # examples/bbm.py:35-40
        v = ufl.TestFunction(V)
        return (inner(udot, v) + inner(u.dx(0), v) + inner(u * u.dx(0), v) + inner(udot.dx(0), v.dx(0))) * dx

    def burgers(nu):  # examples/burgers.py:21-25
        v = ufl.TestFunction(V)
        return (inner(udot, v) + inner(dot(u, ufl.nabla_grad(u)), v) + nu * inner(grad(u), grad(v))) * dx

    def cahn_hilliard(lmbda):  # examples/cahn-hilliard.py:26-38
        q, v = ufl.TestFunctions(V)
        c, mu = ufl.split(u)
        c_t, _ = ufl.split(udot)
        c = ufl.variable(c)
        f = 100 * c ** 2 * (1 - c) ** 2
        dfdc = ufl.diff(f, c)
        return (c_t * q * dx + dot(grad(mu), grad(q)) * dx + mu * v * dx - dfdc * v * dx
                - lmbda * dot(grad(c), grad(v)) * dx)

    def mass_only():  # examples/heat-explicit.py:13 (the implicit side of the IMEX split)
        v = ufl.TestFunction(V)
        return inner(udot, v) * dx

    def minus_diffusion(src):  # examples/heat-explicit.py:14 (the explicit side, G)
        v = ufl.TestFunction(V)
        return -(inner(grad(u), grad(v)) * dx - src * v * dx)

    candidates = [
        (forms.HEAT, (1.0, 0.0, 0.0, 0.0), lambda: heat(1.0)),
        (forms.BBM, (0.0, 0.0, 0.0, 0.0), bbm),
        (forms.BURGERS, (1.0e-4, 0.0, 0.0, 0.0), lambda: burgers(1.0e-4)),
        (forms.CAHN_HILLIARD, (1.0e-2, 0.0, 0.0, 0.0), lambda: cahn_hilliard(1.0e-2)),
        (forms.DIFFUSION, (0.0, 1.0, 0.0, 0.0), mass_only),
        (forms.DIFFUSION, (-1.0, 0.0, -1.0, 0.0), lambda: minus_diffusion(1.0)),
    ]
    for fam, params, make in candidates:
        try:
            if expand_derivatives(make()).signature() == sig:
                return fam, params
        except Exception:
            continue
    return None


class CatalogueForm:
    """what engine.Engine reads from a firedrake_ts_b200.forms.Form"""

    def __init__(self, family, params, space):
        self.family, self.params, self.space = int(family), tuple(float(p) for p in params), space
        self.qdegree = forms.quadrature_degree(self.family, space.degree)


# ------------------------------------------------------------------------------------------------ spaces
class LiveSpace:
    """Adapts a Firedrake FunctionSpace (or a stand-in with the same accessors) to the attributes
    engine.Engine reads from firedrake_ts_b200.mesh.FunctionSpace:
    V.mesh(), V.ufl_element().degree(), len(V), V.value_size, V.sub(0), V.cell_node_map().values_with_halo,
    V.dof_dset.size / .total_size, mesh.coordinates.dat.data_ro_with_halos,
    mesh.coordinates.function_space().cell_node_map().values_with_halo, mesh.cell_set.core_size."""

    def __init__(self, V, device=None):
        import torch

        if device is None:  # the rank's current CUDA device (one rank per GPU)
            device = torch.device("cuda", torch.cuda.current_device())

        mesh = V.mesh()
        self.dim = int(mesh.geometric_dimension())
        self.degree = int(V.ufl_element().degree()) if len(V) == 1 else int(V.sub(0).ufl_element().degree())
        self.variant = "gll" if self.degree >= 3 else "equispaced"
        self.ncomp = int(V.value_size) if len(V) == 1 else len(V)
        self.layout = "interleaved" if len(V) == 1 else "blocked"
        cellname = str(mesh.ufl_cell().cellname())
        cell = "quadrilateral" if cellname == "quadrilateral" else "simplex"
        self.nd = fem.num_nodes(self.dim, self.degree, cell)
        scalar = V if len(V) == 1 else V.sub(0)
        self.firedrake_space = scalar
        cn = np.ascontiguousarray(scalar.cell_node_map().values_with_halo, dtype=np.int32)
        assert cn.shape[1] == self.nd, "cell->node map does not match the catalogue element"
        self.cell_node_map = torch.from_numpy(cn)
        self.num_owned_nodes = int(scalar.dof_dset.size)
        self.num_nodes = int(scalar.dof_dset.total_size)
        self.num_ghost_nodes = self.num_nodes - self.num_owned_nodes
        self.num_dofs = self.num_nodes * self.ncomp
        self.num_owned_dofs = self.num_owned_nodes * self.ncomp

        class _Mesh:
            pass

        m = _Mesh()
        cs = mesh.coordinates.function_space()
        m.cell = cell
        m.coords = torch.from_numpy(np.ascontiguousarray(
            np.asarray(mesh.coordinates.dat.data_ro_with_halos, dtype=np.float64).reshape(-1, self.dim)))
        m.cell_coord_map = torch.from_numpy(np.ascontiguousarray(cs.cell_node_map().values_with_halo, dtype=np.int32))
        m.num_cells = int(m.cell_coord_map.shape[0])
        # PyOP2 lists core cells (no ghost dependence) first: they can run before the ghosts arrive
        m.num_interior_cells = int(getattr(mesh.cell_set, "core_size", m.num_cells if self.num_ghost_nodes == 0 else 0))
        m.num_coord_nodes = int(m.coords.shape[0])
        m.device = torch.device(device)
        self.mesh = m

    def row_block_starts(self):
        return None  # Firedrake's numbering is not tiled: uniform row blocks (engine default)


def halo_lists_from_sf(rank, nranks, num_owned, ilocal, iremote, alltoall):
    """Neighbour send / receive lists of one rank from the point star forest of a PyOP2 halo
    (``halo.sf.getGraph()`` -> nroots, ilocal, iremote): leaf ``ilocal[k]`` (a ghost node, >= num_owned) is
    owned by rank ``iremote[k][0]`` under the owner's local index ``iremote[k][1]``.

    Returns (neighbours, send_lists, recv_lists): for every neighbour r, recv_lists[r] are my ghost node ids
    in the order r packs them, send_lists[r] my owned node ids that r reads.  ``alltoall(list_per_rank)`` is
    the communicator's object all-to-all (mpi4py ``comm.alltoall``): the owners learn who reads what."""
    ilocal = np.asarray(ilocal, dtype=np.int64).reshape(-1)
    iremote = np.asarray(iremote, dtype=np.int64).reshape(-1, 2)
    ghost = ilocal >= num_owned
    ilocal, iremote = ilocal[ghost], iremote[ghost]
    assert (iremote[:, 0] != rank).all(), "a ghost node cannot be owned by its own rank"
    recv_lists, wanted = {}, [np.zeros(0, dtype=np.int64) for _ in range(nranks)]
    for r in np.unique(iremote[:, 0]).tolist():
        sel = iremote[:, 0] == r
        order = np.argsort(iremote[sel, 1], kind="stable")  # the owner packs in ascending owned index
        recv_lists[r] = ilocal[sel][order].astype(np.int32)
        wanted[r] = iremote[sel, 1][order]
    asked = alltoall(wanted)  # asked[r]: my owned indices that rank r reads
    send_lists = {}
    for r, idx in enumerate(asked):
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        if idx.size:
            assert idx.min() >= 0 and idx.max() < num_owned
            send_lists[r] = idx.astype(np.int32)
    neighbours = sorted(set(recv_lists) | set(send_lists))
    return neighbours, send_lists, recv_lists


def mat_values_device_pointer(mat):  # pragma: no cover - needs libpetsc
    """device address of the AIJCUSPARSE value array (INTEGRATION.md shows the C stub)."""
    lib = C.CDLL("libpetsc.so")  # the PETSc the running petsc4py is linked against
    i, j, a, mtype = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
    lib.MatSeqAIJGetCSRAndMemType(C.c_void_p(mat.handle), C.byref(i), C.byref(j), C.byref(a), C.byref(mtype))
    return a.value


class _B200State:
    """engine + persistent local (owned + ghost) device buffers of one context"""

    def __init__(self, family, params, V, bc_nodes, comm=None, appctx=None):
        import torch

        from .engine import Engine

        appctx = appctx or {}
        self.space = LiveSpace(V)
        self.form = CatalogueForm(family, params, self.space)

        class _BC:
            def __init__(self, nodes):
                self.nodes = torch.as_tensor(np.asarray(nodes, dtype=np.int64))

        self.engine = Engine(self.form, [_BC(n) for n in bc_nodes], scatter=appctx.get("scatter"))
        dev = self.engine.device
        self.x = torch.zeros(self.space.num_dofs, dtype=torch.float64, device=dev)
        self.xdot = torch.zeros(self.space.num_dofs, dtype=torch.float64, device=dev)
        self.nranks = 1 if comm is None else int(comm.size)
        self.pattern_checked = False
        if self.nranks > 1:
            rank = int(comm.rank)
            halo = self.space.firedrake_space.dof_dset.halo
            _, ilocal, iremote = halo.sf.getGraph()
            nb, send, recv = halo_lists_from_sf(rank, self.nranks, self.space.num_owned_nodes, ilocal, iremote,
                                                comm.alltoall)
            uid = self.engine.nccl_unique_id() if rank == 0 else None
            uid = comm.bcast(uid, root=0)
            self.engine.halo_setup_raw(rank, self.nranks, nb, send, {}, uid, recv_lists=recv)

    def refresh(self, X, Xdot):
        """global PETSc Vecs -> local device vectors with fresh ghosts"""
        from .engine import _check, _stream

        e = self.engine
        hx, hxd = X.getCUDAHandle("r"), Xdot.getCUDAHandle("r")
        try:
            _check(e._lib.fdts_owned_to_local(e._h, C.c_void_p(hx), C.c_void_p(self.x.data_ptr()), _stream()),
                   "fdts_owned_to_local")
            _check(e._lib.fdts_owned_to_local(e._h, C.c_void_p(hxd), C.c_void_p(self.xdot.data_ptr()), _stream()),
                   "fdts_owned_to_local")
        finally:
            X.restoreCUDAHandle(hx, "r")
            Xdot.restoreCUDAHandle(hxd, "r")
        if self.nranks > 1:
            e.halo_exchange(self.x, self.xdot)

    def check_pattern(self, mat):
        """the values are written in the engine's CSR order: the Mat must have exactly that pattern"""
        if self.pattern_checked:
            return
        rp, ci = self.engine.csr()
        mi, mj = mat.getRowIJ()
        if not (np.array_equal(np.asarray(mi, dtype=np.int64), rp.cpu().numpy())
                and np.array_equal(np.asarray(mj, dtype=np.int64), ci.cpu().numpy().astype(np.int64))):
            raise RuntimeError("B200 adapter: the PETSc Mat's CSR pattern differs from the engine's "
                               "(unexpected preallocation or column order); refusing to write values in place")
        self.pattern_checked = True


def make_context_class(base, hooks=None, values_pointer=None):
    """``base``: the reference's _TSContext (or a stand-in with its attribute names).  ``hooks``: the module
    providing get_appctx(dm) (firedrake.dmhooks).  ``values_pointer(mat)``: device address of the Mat values."""
    hooks = hooks or dmhooks
    values_pointer = values_pointer or mat_values_device_pointer

    class B200TSContext(base):
        """``_TSContext`` whose IFunction / IJacobian callbacks run on the B200 engine."""

        def __init__(self, problem, *args, **kwargs):
            super().__init__(problem, *args, **kwargs)
            appctx = getattr(self, "appctx", None) or {}
            spec = appctx.get("b200_form")  # explicit (family, params); else recognise the UFL form
            if spec is None:
                spec = classify_form(problem.F, problem.u_restrict, problem.udot)
            self._b200 = None
            if spec is not None:
                family, params = spec
                V = problem.u_restrict.function_space()
                comm = getattr(V.mesh(), "comm", None)
                bc_nodes = [bc.nodes for bc in problem.dirichlet_bcs()]
                self._b200 = _B200State(family, params, V, bc_nodes, comm=comm, appctx=appctx)

        @staticmethod
        def form_function(ts, t, X, Xdot, F):
            from .engine import _check, _stream

            ctx = hooks.get_appctx(ts.getDM())
            st = ctx._b200
            if st is None:
                return base.form_function(ts, t, X, Xdot, F)
            # as the reference: the context's Functions follow the iterate (pre/post callbacks and the
            # inherited G branch read them)
            with ctx._x.dat.vec_wo as v:
                X.copy(v)
            with ctx._xdot.dat.vec_wo as v:
                Xdot.copy(v)
            ctx._time.assign(t)
            if ctx._pre_function_callback is not None:
                ctx._pre_function_callback(X, Xdot)
            st.refresh(X, Xdot)
            e = st.engine
            hf = F.getCUDAHandle("w")
            try:
                _check(e._lib.fdts_ifunction(e._h, float(t), C.c_void_p(st.x.data_ptr()), C.c_void_p(st.xdot.data_ptr()),
                                             C.c_void_p(hf), _stream()), "fdts_ifunction")
            finally:
                F.restoreCUDAHandle(hf, "w")
            if ctx._post_function_callback is not None:
                ctx._post_function_callback(X, Xdot, F)

        @staticmethod
        def form_jacobian(ts, t, X, Xdot, shift, J, P):
            from .engine import _check, _stream

            ctx = hooks.get_appctx(ts.getDM())
            st = ctx._b200
            if st is None:
                return base.form_jacobian(ts, t, X, Xdot, shift, J, P)
            problem = ctx._problem
            assert J.handle == ctx._jac.petscmat.handle
            if getattr(problem, "_constant_jacobian", False) and ctx._jacobian_assembled:
                return
            ctx._jacobian_assembled = True
            with ctx._x.dat.vec_wo as v:
                X.copy(v)
            with ctx._xdot.dat.vec_wo as v:
                Xdot.copy(v)
            ctx._time.assign(t)
            if ctx._pre_jacobian_callback is not None:
                ctx._pre_jacobian_callback(X, Xdot)
            ctx._shift.assign(shift)
            st.check_pattern(J)
            st.refresh(X, Xdot)
            e = st.engine
            vals = values_pointer(J)
            _check(e._lib.fdts_ijacobian(e._h, float(t), C.c_void_p(st.x.data_ptr()), C.c_void_p(st.xdot.data_ptr()),
                                         float(shift), C.c_void_p(vals), _stream()), "fdts_ijacobian")
            J.assemble()
            if ctx._post_jacobian_callback is not None:
                ctx._post_jacobian_callback(X, Xdot, J)
            if ctx.Jp is not None:  # the preconditioner form keeps the reference's own assembler (solving_utils.py:310-312)
                assert P.handle == ctx._pjac.petscmat.handle
                ctx._assemble_pjac(ctx._pjac)
            ises = problem.J.arguments()[0].function_space()._ises
            ctx.set_nullspace(ctx._nullspace, ises, transpose=False, near=False)
            ctx.set_nullspace(ctx._nullspace_T, ises, transpose=True, near=False)
            ctx.set_nullspace(ctx._near_nullspace, ises, transpose=False, near=True)

    return B200TSContext


B200TSContext = make_context_class(_RefTSContext) if HAVE_FIREDRAKE else None  # pragma: no cover


def use_b200(solver, context_class=None):
    """Switch an existing firedrake_ts.DAESolver to the B200 engine (same problem, same options):
    a new context of the B200 class replaces solver._ctx, the Jacobian Mat becomes MATAIJCUSPARSE
    (reference solving_utils.py:122 preallocates MATAIJ) and the work Vec VECCUDA (ts_solver.py:239)."""
    cls = context_class or B200TSContext
    if cls is None:
        raise RuntimeError("use_b200 needs firedrake and firedrake_ts (or an explicit context_class)")
    ctx = solver._ctx
    new = cls(ctx._problem, mat_type="aij", pmat_type="aij", appctx=ctx.appctx,
              pre_jacobian_callback=ctx._pre_jacobian_callback,
              pre_function_callback=ctx._pre_function_callback,
              post_jacobian_callback=ctx._post_jacobian_callback,
              post_function_callback=ctx._post_function_callback,
              options_prefix=ctx.options_prefix)
    new._jac.petscmat.setType("aijcusparse")
    solver._ctx = new
    solver._work.setType("cuda")
    return solver
