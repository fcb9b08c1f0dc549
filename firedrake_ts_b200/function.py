"""Function / Constant / DirichletBC: the small slice of Firedrake's object
model that the reference's callbacks touch (firedrake_ts/solving_utils.py:92-96,
249-253, 294-304; firedrake_ts/ts_solver.py:78-104, 121-123, 342-343), backed
by device-resident torch tensors instead of PyOP2 Dats."""
from __future__ import annotations

import torch


class Constant:
    """run-time scalar (``problem.time`` / ``problem.shift``, ts_solver.py:102-104)."""

    def __init__(self, value=0.0):
        self._value = float(value)

    def assign(self, value):
        self._value = float(value._value if isinstance(value, Constant) else value)
        return self

    def __float__(self):
        return self._value

    def __repr__(self):
        return f"Constant({self._value})"


class Function:
    """Degree-of-freedom vector on a FunctionSpace.  ``.data`` is the *local*
    vector (owned + ghost dofs, space layout); ``.owned`` views the owned part
    in the layout of the PETSc Vec the reference would hand out
    (``dat.vec_ro/vec_wo``, solving_utils.py:249-252)."""

    def __init__(self, V, val=None, name=None, device=None):
        self._V = V
        self.name = name
        dev = torch.device(device) if device is not None else V.mesh.device
        if val is None:
            self.data = torch.zeros(V.num_dofs, dtype=torch.float64, device=dev)
        else:
            self.data = val.to(device=dev, dtype=torch.float64).contiguous()
            assert self.data.numel() == V.num_dofs

    def function_space(self):
        return self._V

    @property
    def device(self):
        return self.data.device

    def owned(self):
        """owned dofs as one tensor in Vec layout (a view when possible)."""
        V = self._V
        if V.num_ghost_nodes == 0:
            return self.data
        if V.layout == "interleaved":
            return self.data[: V.num_owned_dofs]
        return self.data.view(V.ncomp, V.num_nodes)[:, : V.num_owned_nodes].reshape(-1)

    def set_owned(self, x):
        V = self._V
        if V.num_ghost_nodes == 0 or V.layout == "interleaved":
            self.data[: V.num_owned_dofs].copy_(x)
        else:
            self.data.view(V.ncomp, V.num_nodes)[:, : V.num_owned_nodes].copy_(x.view(V.ncomp, V.num_owned_nodes))

    def assign(self, other):
        if isinstance(other, Function):
            self.data.copy_(other.data)
        else:
            self.data.fill_(float(other))
        return self

    def interpolate(self, fn):
        self.data.copy_(self._V.interpolate(fn).to(self.data.device))
        return self

    def copy(self):
        return Function(self._V, self.data.clone(), name=self.name)

    def sub(self, i):
        """component view (mixed spaces: ``u.subfunctions``)."""
        V = self._V
        if V.layout == "blocked":
            return self.data.view(V.ncomp, V.num_nodes)[i]
        return self.data.view(V.num_nodes, V.ncomp)[:, i]


class DirichletBC:
    """Strong boundary condition on every component of ``V`` at ``sub_domain``
    ("on_boundary" or an explicit node tensor)."""

    def __init__(self, V, value, sub_domain="on_boundary"):
        self._V = V
        self.value = value
        if isinstance(sub_domain, str):
            assert sub_domain == "on_boundary"
            self.nodes = V.boundary_nodes
        else:
            self.nodes = torch.as_tensor(sub_domain, dtype=torch.int32, device=V.mesh.device)

    def function_space(self):
        return self._V

    def dirichlet_bcs(self):
        yield self

    def apply(self, u: Function):
        """set the boundary values on ``u`` (ts_solver.py:342-343)."""
        V = self._V
        val = float(self.value)
        nodes = self.nodes.to(torch.int64).to(u.data.device)
        for m in range(V.ncomp):
            u.data[V.dof_index(nodes, m)] = val
