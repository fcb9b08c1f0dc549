"""PETSc-/Firedrake-shaped stand-ins for executing firedrake_ts_b200.adapter_firedrake without either
package (tests only).  They carry the accessor names of SURVEY.md Appendix A.2 and the attribute names the
reference's _TSContext reads and writes (firedrake_ts/solving_utils.py:77-111, 236-317); nothing here
computes anything on behalf of the product."""
import contextlib

import numpy as np
import torch


class FakeVec:
    """petsc4py.PETSc.Vec of type VECCUDA: global vector (owned dofs) living on the device."""

    def __init__(self, n=None, array=None, device="cuda:0"):
        self.array = torch.zeros(n, dtype=torch.float64, device=device) if array is None else array
        self.handles_out = 0

    def getCUDAHandle(self, mode="rw"):
        self.handles_out += 1
        return self.array.data_ptr()

    def restoreCUDAHandle(self, handle, mode="rw"):
        assert handle == self.array.data_ptr()
        self.handles_out -= 1

    def copy(self, dest):
        dest.array.copy_(self.array)

    def getLocalSize(self):
        return self.array.numel()

    def setType(self, t):
        self.type = t


class FakeMat:
    """MATAIJCUSPARSE: CSR pattern on the host (getRowIJ), values on the device."""

    _next = [1000]

    def __init__(self, rowptr, colidx, device="cuda:0"):
        self.rowptr, self.colidx = np.asarray(rowptr), np.asarray(colidx)
        self.values = torch.zeros(len(self.colidx), dtype=torch.float64, device=device)
        FakeMat._next[0] += 1
        self.handle = FakeMat._next[0]
        self.assembled = 0
        self.type = "aij"

    def getRowIJ(self):
        return self.rowptr, self.colidx

    def assemble(self):
        self.assembled += 1

    def setType(self, t):
        self.type = t


class FakeConstant:
    def __init__(self, v=0.0):
        self.value = v

    def assign(self, v):
        self.value = float(v)


class _Dat:
    def __init__(self, vec):
        self._vec = vec

    @property
    @contextlib.contextmanager
    def vec_wo(self):
        yield self._vec

    vec = vec_ro = vec_wo


class FakeFunction:
    """firedrake.Function: .dat.vec_wo yields the PETSc Vec view of its (owned) values."""

    def __init__(self, V, n_owned_dofs, device="cuda:0"):
        self._V = V
        self.dat = _Dat(FakeVec(n_owned_dofs, device=device))

    def function_space(self):
        return self._V


class _Map:
    def __init__(self, values):
        self.values_with_halo = values


class _Element:
    def __init__(self, degree):
        self._degree = degree

    def degree(self):
        return self._degree


class _Cell:
    def __init__(self, name):
        self._name = name

    def cellname(self):
        return self._name


class _SF:
    def __init__(self, graph):
        self._graph = graph

    def getGraph(self):
        return self._graph


class FakeFiredrakeMesh:
    def __init__(self, m, comm=None, coord_space=None):
        self._m = m
        self.comm = comm

        class _C:
            pass

        self.coordinates = _C()
        self.coordinates.dat = _C()
        self.coordinates.dat.data_ro_with_halos = m.coords.cpu().numpy()
        self.coordinates.function_space = lambda: coord_space
        self.cell_set = _C()
        self.cell_set.core_size = m.num_interior_cells

    def geometric_dimension(self):
        return self._m.dim

    def ufl_cell(self):
        names = {1: "interval", 2: "triangle", 3: "tetrahedron"}
        return _Cell("quadrilateral" if self._m.cell == "quadrilateral" else names[self._m.dim])


class FakeFiredrakeSpace:
    """firedrake FunctionSpace / MixedFunctionSpace built from the synthetic generator's arrays."""

    def __init__(self, V, comm=None, sf_graph=None, _scalar_of=None):
        self._V = V
        self._scalar_of = _scalar_of
        mesh = V.mesh

        class _CoordSpace:
            def cell_node_map(self_inner):
                return _Map(mesh.cell_coord_map.cpu().numpy())

        self._mesh = FakeFiredrakeMesh(mesh, comm=comm, coord_space=_CoordSpace())
        self.value_size = V.ncomp if (V.layout == "interleaved" and _scalar_of is None) else 1
        self._nfields = V.ncomp if (V.layout == "blocked" and _scalar_of is None and V.ncomp > 1) else 1

        class _DSet:
            pass

        self.dof_dset = _DSet()
        self.dof_dset.size = V.num_owned_nodes
        self.dof_dset.total_size = V.num_nodes
        if sf_graph is not None:
            self.dof_dset.halo = type("H", (), {"sf": _SF(sf_graph)})()
        self._sf_graph = sf_graph
        self._comm = comm
        self._ises = None

    def mesh(self):
        return self._mesh

    def ufl_element(self):
        return _Element(self._V.degree)

    def __len__(self):
        return self._nfields

    def sub(self, i):
        return FakeFiredrakeSpace(self._V, comm=self._comm, sf_graph=self._sf_graph, _scalar_of=self)

    def cell_node_map(self):
        return _Map(self._V.cell_node_map.cpu().numpy())


class FakeDM:
    def __init__(self, ctx):
        self.ctx = ctx


class FakeTS:
    def __init__(self, ctx):
        self._dm = FakeDM(ctx)

    def getDM(self):
        return self._dm


class fake_dmhooks:
    @staticmethod
    def get_appctx(dm):
        return dm.ctx


class FakeBC:
    def __init__(self, nodes):
        self.nodes = np.asarray(nodes)


class FakeProblem:
    """what _TSContext reads from firedrake_ts.DAEProblem"""

    def __init__(self, V_fd, n_owned_dofs, bcs=(), Jp=None, device="cuda:0"):
        self.u_restrict = FakeFunction(V_fd, n_owned_dofs, device=device)
        self.udot = FakeFunction(V_fd, n_owned_dofs, device=device)
        self.F = object()
        self.Jp = Jp
        self.shift, self.time = FakeConstant(), FakeConstant()
        self._constant_jacobian = False
        self._bcs = list(bcs)

        class _J:
            @staticmethod
            def arguments():
                return (type("A", (), {"function_space": lambda self_: V_fd})(),)

        self.J = _J()

    def dirichlet_bcs(self):
        return self._bcs


class StubRefTSContext:
    """stand-in for firedrake_ts.solving_utils._TSContext: the attributes its callbacks use, and a
    reference-path form_function / form_jacobian that record that they were reached."""

    def __init__(self, problem, mat_type="aij", pmat_type="aij", appctx=None, pre_jacobian_callback=None,
                 pre_function_callback=None, post_jacobian_callback=None, post_function_callback=None,
                 options_prefix=None, jac=None, pjac=None):
        self._problem = problem
        self.appctx = appctx or {}
        self._x, self._xdot = problem.u_restrict, problem.udot
        self._shift, self._time = problem.shift, problem.time
        self._pre_jacobian_callback, self._pre_function_callback = pre_jacobian_callback, pre_function_callback
        self._post_jacobian_callback, self._post_function_callback = post_jacobian_callback, post_function_callback
        self.options_prefix = options_prefix
        self.Jp = problem.Jp
        self._jacobian_assembled = False
        self._nullspace = self._nullspace_T = self._near_nullspace = None
        self.nullspace_calls = []
        self.pjac_assembled = 0
        self._jac = type("M", (), {"petscmat": jac})()
        self._pjac = type("M", (), {"petscmat": pjac if pjac is not None else jac})()
        self.reference_path = []

    def set_nullspace(self, nullspace, ises=None, transpose=False, near=False):
        self.nullspace_calls.append((transpose, near))

    def _assemble_pjac(self, pjac):
        self.pjac_assembled += 1

    @staticmethod
    def form_function(ts, t, X, Xdot, F):
        fake_dmhooks.get_appctx(ts.getDM()).reference_path.append("F")

    @staticmethod
    def form_jacobian(ts, t, X, Xdot, shift, J, P):
        fake_dmhooks.get_appctx(ts.getDM()).reference_path.append("J")


def sf_graph_of(V):
    """PetscSF graph (nroots, ilocal, iremote) of the ghost nodes of a partitioned synthetic space: leaf =
    local ghost node id, root = (owner rank, owner's local id)."""
    nown, nn = V.num_owned_nodes, V.num_nodes
    if nn == nown:
        return nown, np.zeros(0, dtype=np.int32), np.zeros((0, 2), dtype=np.int32)
    gt = V._ghost_table(V.mesh.partition.rank)
    pts, owner = gt["points"], gt["owner"]
    ilocal = np.arange(nown, nn, dtype=np.int32)
    iremote = np.zeros((nn - nown, 2), dtype=np.int32)
    iremote[:, 0] = owner.cpu().numpy()
    for r in torch.unique(owner).tolist():
        sel = owner == r
        iremote[sel.cpu().numpy(), 1] = V._local_ids(pts[sel], rank=r).cpu().numpy()
    return nown, ilocal, iremote
