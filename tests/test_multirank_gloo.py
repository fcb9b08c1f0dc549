"""N > 1 path on CPU (gloo, world_size 2 and 4): partitioned meshes, halo exchange, owner-only row
assembly and the distributed TS solve reproduce the single-rank results."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from firedrake_ts_b200.workloads import make_problem, rel_err
from helpers import OracleAssembler
from mr_worker import run


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, *args):
    outdir = tempfile.mkdtemp()
    port = free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=run, args=(r, world, port, outdir) + args) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return [torch.load(os.path.join(outdir, f"rank{r}.pt"), weights_only=False) for r in range(world)]


def global_matrix(res, nglob):
    import scipy.sparse as sp

    rows, cols, vals = [], [], []
    for o in res:
        nc, nown, nn = o["ncomp"], o["nown"], o["nnodes"]
        gid = o["gid"].numpy()
        rp, ci, v = o["rowptr"].numpy(), o["colidx"].numpy(), o["J"].numpy()
        r_local = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        if o["layout"] == "interleaved":
            grow = gid[r_local // nc] * nc + r_local % nc
            gcol = gid[ci // nc] * nc + ci % nc
        else:
            grow = (r_local // nown) * nglob + gid[r_local % nown]
            gcol = (ci // nn) * nglob + gid[ci % nn]
        rows.append(grow); cols.append(gcol); vals.append(v)
    n = nglob * res[0]["ncomp"]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


@pytest.mark.parametrize("family,dim,degree,n,grid,solve", [
    ("heat", 2, 2, 6, (2, 1), True),
    ("heat", 2, 2, 6, (2, 1), "imex"),
    ("heat", 3, 2, 4, (2, 2, 1), False),
    ("burgers", 2, 2, 4, (2, 1), False),
    ("cahn_hilliard", 2, 1, 6, (1, 2), False),
    ("cahn_hilliard", 2, 1, 6, (2, 1), "ch"),
])
def test_partitioned_equals_single_rank(family, dim, degree, n, grid, solve):
    import scipy.sparse as sp

    world = int(np.prod(grid))
    res = launch(world, family, dim, degree, n, list(grid), solve)
    assert all(o["halo_ok"] for o in res)
    form, bcs, u, udot = make_problem(family, dim, degree, n, tile=2)
    V = form.space
    asm = OracleAssembler(form, bcs)
    F1 = asm.ifunction(0.0, u.data, udot.data).numpy()
    rp, ci = asm.pattern()
    J1 = sp.csr_matrix((asm.ijacobian(0.0, u.data, udot.data, 4.0).numpy(), ci, rp), shape=(V.num_dofs, V.num_dofs))
    # single-rank numbering -> global ids of the partitioned numbering, via node coordinates
    key = lambda X: [tuple(np.round(x, 9)) for x in X]
    # residual: compare by global id through coordinates
    nglob = V.global_num_nodes
    coord_to_single = {k: i for i, k in enumerate(key(V.node_coords.numpy()))}
    from firedrake_ts_b200 import mesh as M
    Fg = np.zeros(V.num_dofs)
    perm = np.zeros(nglob, dtype=np.int64)  # partition-global id -> single-rank node id
    for r, o in enumerate(res):
        Vr = M.FunctionSpace(M.SimplexMesh((n,) * dim, partition=M.Partition(tuple(grid), r), tile=2), degree,
                             ncomp=V.ncomp, layout=V.layout)
        single = np.array([coord_to_single[k] for k in key(Vr.node_coords.numpy())])
        perm[o["gid"].numpy()] = single
        nown, nc = o["nown"], o["ncomp"]
        Fr = o["F"].numpy()
        for m in range(nc):
            if V.layout == "interleaved":
                Fg[single[:nown] * nc + m] = Fr[np.arange(nown) * nc + m]
            else:
                Fg[m * V.num_nodes + single[:nown]] = Fr[m * nown + np.arange(nown)]
    assert rel_err(Fg, F1) < 1e-12
    Jg = global_matrix(res, nglob).tocoo()
    nc = V.ncomp
    if V.layout == "interleaved":
        rr = perm[Jg.row // nc] * nc + Jg.row % nc
        cc = perm[Jg.col // nc] * nc + Jg.col % nc
    else:
        rr = (Jg.row // nglob) * V.num_nodes + perm[Jg.row % nglob]
        cc = (Jg.col // nglob) * V.num_nodes + perm[Jg.col % nglob]
    Jp = sp.csr_matrix((Jg.data, (rr, cc)), shape=J1.shape)
    diff = (Jp - J1)
    assert abs(diff).max() < 1e-12 * abs(J1).max()
    assert Jp.nnz == J1.nnz  # same sparsity pattern, every row assembled by exactly one owner
    if solve:
        import firedrake_ts_b200 as fts
        if solve == "imex":
            from firedrake_ts_b200 import forms
            from mr_worker import IMEX_DT

            F, G = forms.heat_explicit(u, udot)
            problem = fts.DAEProblem(F, u, udot, (0.0, 4 * IMEX_DT), bcs=bcs, G=G)
            s1 = fts.DAESolver(problem, assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs),
                               solver_parameters={"pc_type": "lu"}, rhs_projection_parameters={"pc_type": "lu"})
            s1.ts.setTimeStep(IMEX_DT)
            u0 = u.data.clone()
            s1.solve()
            assert s1.ts.getStepNumber() == 4 and all(o["steps"] == 4 for o in res)
            assert (u.data - u0).abs().max() > 1e-3
        elif solve == "ch":
            from mr_worker import CH_DT

            problem = fts.DAEProblem(form, u, udot, (0.0, 2 * CH_DT))
            s1 = fts.DAESolver(problem, assembler=asm, solver_parameters={"pc_type": "lu", "snes_rtol": 1e-10,
                                                                           "snes_atol": 1e-11, "ts_dt": CH_DT})
            s1.solve()
            assert s1.ts.getStepNumber() == 2 and all(o["steps"] == 2 for o in res)
        else:
            problem = fts.DAEProblem(form, u, udot, (0.0, 0.2), bcs=bcs or None)
            fts.DAESolver(problem, assembler=asm, solver_parameters={"pc_type": "lu", "snes_rtol": 1e-12}).solve()
        ug = np.zeros(V.num_dofs)
        for r, o in enumerate(res):
            nown = o["nown"]
            gid = o["gid"].numpy()[:nown]
            uf = o["u_final"].numpy()
            if V.layout == "interleaved":
                for m in range(V.ncomp):
                    ug[perm[gid] * V.ncomp + m] = uf[np.arange(nown) * V.ncomp + m]
            else:
                for m in range(V.ncomp):
                    ug[m * V.num_nodes + perm[gid]] = uf[m * nown:(m + 1) * nown]
        assert np.abs(ug - u.data.numpy()).max() < 1e-9 * max(1.0, np.abs(u.data.numpy()).max())
