"""worker of tests/test_multirank_gloo.py: one rank of a world_size-N gloo job (CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


IMEX_DT = 2.0e-4
CH_DT = 5.0e-6


def run(rank, world, port, outdir, family, dim, degree, n, grid, solve):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import firedrake_ts_b200 as fts
    from firedrake_ts_b200 import mesh as M
    from firedrake_ts_b200.halo import HaloExchange
    from firedrake_ts_b200.workloads import make_problem
    from helpers import OracleAssembler

    part = M.Partition(tuple(grid), rank)
    form, bcs, u, udot = make_problem(family, dim, degree, n, partition=part, tile=2)
    V = form.space
    halo = HaloExchange(V, None, group=None, device="cpu")
    # scramble the ghost tails, then let the exchange restore them from the owners
    ref_u, ref_ud = u.data.clone(), udot.data.clone()
    if V.num_ghost_nodes:
        for m in range(V.ncomp):
            idx = V.dof_index(torch.arange(V.num_owned_nodes, V.num_nodes), m)
            u.data[idx] = -777.0
            udot.data[idx] = 555.0
    halo.forward(u.data)
    halo.forward(udot.data)
    halo_ok = bool(torch.equal(u.data, ref_u) and torch.equal(udot.data, ref_ud))
    asm = OracleAssembler(form, bcs)
    F = asm.ifunction(0.0, u.data, udot.data)
    rp, ci = asm.pattern()
    Jv = asm.ijacobian(0.0, u.data, udot.data, 4.0)
    out = {"halo_ok": halo_ok, "F": F, "rowptr": torch.from_numpy(rp), "colidx": torch.from_numpy(ci), "J": Jv,
           "gid": V.global_node_ids.clone(), "nown": V.num_owned_nodes, "nnodes": V.num_nodes, "ncomp": V.ncomp,
           "layout": V.layout}
    if solve == "imex":
        # the IMEX split on the partitioned space: G assembled per owner, projected with the distributed
        # mass matrix (CG + Jacobi, dots all-reduced), ARKIMEX stepping
        from firedrake_ts_b200 import forms

        F, G = forms.heat_explicit(u, udot)
        problem = fts.DAEProblem(F, u, udot, (0.0, 4 * IMEX_DT), bcs=bcs, G=G)
        solver = fts.DAESolver(problem, assembler=OracleAssembler(F, bcs), rhs_assembler=OracleAssembler(G, bcs),
                               comm=dist.group.WORLD,
                               solver_parameters={"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-13},
                               rhs_projection_parameters={"ksp_rtol": 1e-14})
        solver.ts.setTimeStep(IMEX_DT)
        solver.solve()
        out["u_final"] = u.owned().clone()
        out["steps"] = solver.ts.getStepNumber()
    elif solve == "ch":
        # mixed (field-blocked) space on a partition with the default-style Jacobi preconditioner: the diagonal of
        # dof (m, i) sits at local column m*num_nodes + i, not m*num_owned + i
        problem = fts.DAEProblem(form, u, udot, (0.0, 2 * CH_DT), bcs=None)
        solver = fts.DAESolver(problem, assembler=asm, comm=dist.group.WORLD,
                               solver_parameters={"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-13,
                                                  "snes_rtol": 1e-10, "snes_atol": 1e-11, "ksp_max_it": 4000,
                                                  "ksp_gmres_restart": 200, "ts_dt": CH_DT})
        solver.solve()
        out["u_final"] = u.owned().clone()
        out["steps"] = solver.ts.getStepNumber()
    elif solve:
        problem = fts.DAEProblem(form, u, udot, (0.0, 0.2), bcs=bcs or None)
        solver = fts.DAESolver(problem, assembler=asm, comm=dist.group.WORLD,
                               solver_parameters={"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-13,
                                                  "snes_rtol": 1e-12, "ksp_max_it": 2000})
        solver.solve()
        out["u_final"] = u.owned().clone()
    torch.save(out, os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()
