"""Synthetic workloads C1-C5 of SURVEY.md section 8(d): seeded meshes, states and boundary
conditions used by the parity tests, bench.py and the smoke test."""
import math

import numpy as np
import torch

from . import forms, mesh as M
from .function import DirichletBC, Function

# name -> (family, dim, degree, full-size n, function-space tile, tile offset, mesh tile)
CONFIGS = {
    "C1": ("heat", 2, 1, 4096, 14, 0, 14),  # 14 x 14 row blocks: 15^2 = 225 vertices per block (limit 254)
    "C2": ("burgers", 2, 2, 2048, 0, 0, 8),
    "C3": ("cahn_hilliard", 2, 1, 4096, 0, 0, 16),
    # the literal mesh of examples/cahn-hilliard.py:14 (quadrilateral Q1 x Q1) at the size of C3
    "C3q": ("cahn_hilliard", 2, 1, 4096, 0, 0, 16),
    "C4": ("heat", 3, 2, 184, 6, 0, 3),
    "C5": ("bbm", 3, 3, 96, 0, 0, 2),
}


def make_problem(family, dim, degree, n, *, tile=0, partition=None, device="cpu", variant="gll", seed=0,
                 ftile=None, foffset=0, quadrilateral=False, warp=0.0, numbering=None, block_rows=192):
    """returns (form, bcs, u, udot) with the seeded states of SURVEY.md 8(d).

    quadrilateral: Q1 cells (examples/cahn-hilliard.py:14); warp > 0 displaces the vertices smoothly
    (x += warp sin(2 pi x) sin(2 pi y) ...) so that quadrilaterals are no longer parallelograms and the
    non-affine geometry path is exercised; boundary vertices stay on the boundary."""
    shape = (n,) * dim
    msh = M.SimplexMesh(shape, partition=partition, tile=tile, device=device, quadrilateral=quadrilateral,
                        numbering=numbering)
    if warp:
        X = msh.coords.clone()
        bump = torch.ones(X.shape[0], dtype=X.dtype, device=X.device)
        for a in range(dim):
            bump = bump * torch.sin(2 * math.pi * X[:, a])
        for a in range(dim):
            msh.coords[:, a] = X[:, a] + warp * (1.0 + 0.5 * a) * bump
    if family in ("heat", "bbm", "diffusion"):
        V = M.FunctionSpace(msh, degree, variant=variant, tile=ftile, tile_offset=foffset, block_rows=block_rows)
    elif family == "burgers":
        V = M.FunctionSpace(msh, degree, ncomp=dim, layout="interleaved", variant=variant)
    elif family == "cahn_hilliard":
        V = M.FunctionSpace(msh, degree, ncomp=2, layout="blocked", variant=variant)
    else:
        raise ValueError(family)
    u, udot = Function(V), Function(V)
    g = torch.Generator().manual_seed(seed)
    # rank- and numbering-independent key of a node: its lexicographic index on the refined lattice
    gid = V._lex_key(V.node_lattice).cpu()
    nglob = V.global_num_nodes

    def rnd(ncomp):
        # rank-independent random field: draw for all global nodes, pick the local ones
        full = torch.randn((nglob, ncomp), generator=g, dtype=torch.float64)
        loc = full[gid]
        if V.layout == "interleaved":
            return loc.reshape(-1)
        return loc.t().reshape(-1)

    x = V.node_coords.cpu()
    if family in ("heat", "bbm", "diffusion"):
        u.data.copy_(torch.prod(torch.sin(math.pi * x), dim=1))
        udot.data.copy_(rnd(1))
        if family == "diffusion":  # generic coefficients; the IMEX split uses (1, 0, 0) and (0, -1, -1)
            form = forms.diffusion(u, udot, mass=0.75, diffusivity=-1.5, source=0.3)
        else:
            form = forms.heat(u, udot) if family == "heat" else forms.bbm(u, udot)
        bcs = [DirichletBC(V, 0.0, "on_boundary")] if family != "bbm" else []
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


