"""shared builders for the tests: the synthetic configurations of SURVEY.md section 8(d)."""
import math

import numpy as np
import torch

from firedrake_ts_b200 import forms, mesh as M
from firedrake_ts_b200.function import DirichletBC, Function


def make_problem(family, dim, degree, n, *, tile=0, partition=None, device="cpu", variant="gll", seed=0,
                 ftile=None, foffset=0):
    """returns (form, bcs, u, udot) with the seeded states of SURVEY.md 8(d)."""
    shape = (n,) * dim
    msh = M.SimplexMesh(shape, partition=partition, tile=tile, device=device)
    if family in ("heat", "bbm"):
        V = M.FunctionSpace(msh, degree, variant=variant, tile=ftile, tile_offset=foffset)
    elif family == "burgers":
        V = M.FunctionSpace(msh, degree, ncomp=dim, layout="interleaved", variant=variant)
    elif family == "cahn_hilliard":
        V = M.FunctionSpace(msh, degree, ncomp=2, layout="blocked", variant=variant)
    else:
        raise ValueError(family)
    u, udot = Function(V), Function(V)
    g = torch.Generator().manual_seed(seed)
    gid = V.global_node_ids.cpu()
    nglob = V.global_num_nodes

    def rnd(ncomp):
        # rank-independent random field: draw for all global nodes, pick the local ones
        full = torch.randn((nglob, ncomp), generator=g, dtype=torch.float64)
        loc = full[gid]
        if V.layout == "interleaved":
            return loc.reshape(-1)
        return loc.t().reshape(-1)

    x = V.node_coords.cpu()
    if family in ("heat", "bbm"):
        u.data.copy_(torch.prod(torch.sin(math.pi * x), dim=1))
        udot.data.copy_(rnd(1))
        form = forms.heat(u, udot) if family == "heat" else forms.bbm(u, udot)
        bcs = [DirichletBC(V, 0.0, "on_boundary")] if family == "heat" else []
    elif family == "burgers":
        v = torch.zeros((V.num_nodes, dim), dtype=torch.float64)
        v[:, 0] = torch.sin(math.pi * x[:, 0])
        v += 0.1 * rnd(dim).reshape(V.num_nodes, dim)
        u.data.copy_(v.reshape(-1))
        udot.data.copy_(rnd(dim))
        form, bcs = forms.burgers(u, udot, nu=1e-4), []
    else:
        full = torch.rand((nglob,), generator=g, dtype=torch.float64)
        c = 0.63 + 0.2 * (0.5 - full[gid])
        mu = 0.1 * torch.randn((nglob,), generator=g, dtype=torch.float64)[gid]
        u.data.copy_(torch.cat([c, mu]))
        udot.data.copy_(rnd(2))
        form, bcs = forms.cahn_hilliard(u, udot, lmbda=1e-2), []
    return form, bcs, u, udot


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


class OracleAssembler:
    """plugs the CPU oracle behind _TSContext's assembler interface (tests only): reference
    trajectories are produced through the very same callbacks and time stepper."""

    def __init__(self, form, bcs):
        from oracle import Oracle

        self.o = Oracle.from_form(form, bcs)
        self.device = torch.device("cpu")

    def pattern(self):
        return self.o.pattern()

    def ifunction(self, t, x, xdot, out=None):
        F = torch.from_numpy(self.o.ifunction(t, x, xdot))
        if out is not None:
            out.copy_(F)
            return out
        return F

    def ijacobian(self, t, x, xdot, shift, out=None):
        J = torch.from_numpy(self.o.ijacobian(t, x, xdot, shift))
        if out is not None:
            out.copy_(J)
            return out
        return J
