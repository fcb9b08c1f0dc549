"""ctypes front-end of the CPU oracle (oracle/fdts_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
firedrake_ts_b200 never imports this module.  PARITY UNPINNED with respect to
the reference's own tests (they pin no numbers); pinned
instead against exact known answers in tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libfdts_oracle.so")
    src = os.path.join(_HERE, "fdts_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfdts_oracle.so"], stdout=subprocess.DEVNULL)
    return so


class _Problem(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("nd", C.c_int32), ("ncomp", C.c_int32), ("layout", C.c_int32),
        ("family", C.c_int32), ("nq", C.c_int32),
        ("num_cells", C.c_int64), ("num_nodes", C.c_int64), ("num_owned_nodes", C.c_int64),
        ("num_coord_nodes", C.c_int64),
        ("coords", C.c_void_p), ("cell_coord", C.c_void_p), ("cell_node", C.c_void_p),
        ("B", C.c_void_p), ("w", C.c_void_p),
        ("params", C.c_double * 4),
        ("num_bc_nodes", C.c_int64), ("bc_nodes", C.c_void_p),
        ("cell_type", C.c_int32),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for name in ("oracle_ifunction", "oracle_node_pattern", "oracle_dof_pattern", "oracle_ijacobian",
                     "oracle_element"):
            getattr(_LIB, name).restype = C.c_int
    return _LIB


def _np(x, dtype):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


class Oracle:
    """CPU restatement of one assembly problem (same inputs as engine.Engine)."""

    def __init__(self, *, dim, nd, ncomp, layout, family, params, coords, cell_coord, cell_node, B, w,
                 num_nodes, num_owned_nodes, bc_nodes=None, cell_type=0):
        self._keep = dict(
            coords=_np(coords, np.float64), cell_coord=_np(cell_coord, np.int32), cell_node=_np(cell_node, np.int32),
            B=_np(B, np.float64), w=_np(w, np.float64),
            bc=_np(bc_nodes if bc_nodes is not None else np.zeros(0), np.int32))
        k = self._keep
        p = _Problem()
        p.dim, p.nd, p.ncomp = dim, nd, ncomp
        p.layout = 0 if layout == "interleaved" else 1
        p.family, p.nq = family, len(k["w"])
        p.num_cells = k["cell_node"].shape[0]
        p.num_nodes, p.num_owned_nodes = num_nodes, num_owned_nodes
        p.num_coord_nodes = k["coords"].shape[0]
        for name, key in (("coords", "coords"), ("cell_coord", "cell_coord"), ("cell_node", "cell_node"),
                          ("B", "B"), ("w", "w"), ("bc_nodes", "bc")):
            setattr(p, name, k[key].ctypes.data)
        for i, v in enumerate(list(params)[:4]):
            p.params[i] = float(v)
        p.num_bc_nodes = len(k["bc"])
        p.cell_type = int(cell_type)
        assert k["cell_coord"].shape[1] == (4 if cell_type == 1 else dim + 1)
        self.p = p
        self.ncomp, self.nd = ncomp, nd
        self.num_nodes, self.num_owned_nodes = num_nodes, num_owned_nodes
        self._pattern = None

    @classmethod
    def from_form(cls, form, bcs=()):
        """build from a firedrake_ts_b200 catalogue Form (tests' convenience)."""
        from firedrake_ts_b200 import fem
        V = form.space
        m = V.mesh
        et = fem.element_tables(V.dim, V.degree, form.qdegree, V.variant, m.cell)
        bc_nodes = None
        if bcs:
            import torch
            bc_nodes = torch.unique(torch.cat([bc.nodes.to(torch.int64) for bc in bcs])).to(torch.int32)
        return cls(dim=V.dim, nd=V.nd, ncomp=V.ncomp, layout=V.layout, family=form.family, params=form.params,
                   coords=m.coords, cell_coord=m.cell_coord_map, cell_node=V.cell_node_map, B=et.B, w=et.w,
                   num_nodes=V.num_nodes, num_owned_nodes=V.num_owned_nodes, bc_nodes=bc_nodes,
                   cell_type=1 if m.cell == "quadrilateral" else 0)

    # -- residual ---------------------------------------------------------
    def ifunction(self, t, u, udot):
        u, udot = _np(u, np.float64), _np(udot, np.float64)
        assert u.size == self.num_nodes * self.ncomp == udot.size
        F = np.zeros(self.num_owned_nodes * self.ncomp)
        rc = _lib().oracle_ifunction(C.byref(self.p), C.c_double(t), u.ctypes.data_as(C.c_void_p),
                                     udot.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return F

    # -- pattern ----------------------------------------------------------
    def node_pattern(self):
        nown = self.num_owned_nodes
        rowptr = np.zeros(nown + 1, dtype=np.int64)
        _lib().oracle_node_pattern(C.byref(self.p), rowptr.ctypes.data_as(C.c_void_p), None)
        colidx = np.zeros(rowptr[-1], dtype=np.int32)
        _lib().oracle_node_pattern(C.byref(self.p), rowptr.ctypes.data_as(C.c_void_p),
                                   colidx.ctypes.data_as(C.c_void_p))
        return rowptr, colidx

    def pattern(self):
        """dof-level CSR (rowptr int64, colidx int32)."""
        if self._pattern is None:
            nrp, nci = self.node_pattern()
            nrows = self.num_owned_nodes * self.ncomp
            rowptr = np.zeros(nrows + 1, dtype=np.int64)
            colidx = np.zeros(int(nrp[-1]) * self.ncomp * self.ncomp, dtype=np.int32)
            _lib().oracle_dof_pattern(C.byref(self.p), nrp.ctypes.data_as(C.c_void_p),
                                      nci.ctypes.data_as(C.c_void_p), rowptr.ctypes.data_as(C.c_void_p),
                                      colidx.ctypes.data_as(C.c_void_p))
            self._pattern = (rowptr, colidx)
        return self._pattern

    # -- jacobian ---------------------------------------------------------
    def ijacobian(self, t, u, udot, shift):
        u, udot = _np(u, np.float64), _np(udot, np.float64)
        rowptr, colidx = self.pattern()
        vals = np.zeros(rowptr[-1])
        rc = _lib().oracle_ijacobian(C.byref(self.p), C.c_double(t), u.ctypes.data_as(C.c_void_p),
                                     udot.ctypes.data_as(C.c_void_p), C.c_double(shift),
                                     rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p),
                                     vals.ctypes.data_as(C.c_void_p))
        assert rc == 0, "entry outside the sparsity pattern"
        return vals

    def element(self, cell, u, udot, shift):
        u, udot = _np(u, np.float64), _np(udot, np.float64)
        ne = self.nd * self.ncomp
        Fe, Ae = np.zeros(ne), np.zeros((ne, ne))
        _lib().oracle_element(C.byref(self.p), C.c_int64(cell), u.ctypes.data_as(C.c_void_p),
                              udot.ctypes.data_as(C.c_void_p), C.c_double(shift), Fe.ctypes.data_as(C.c_void_p),
                              Ae.ctypes.data_as(C.c_void_p))
        return Fe, Ae
