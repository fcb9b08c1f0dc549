"""Multi-GPU check of the library's own ghost exchange (one rank per GPU, NCCL), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py

Per rank: (1) fdts_halo_exchange restores scrambled ghosts bit for bit (compared with the values the owners
hold); (2) fdts_ifunction_exchange == the torch.distributed-driven exchange + plain residual; (3) a DAESolver run
on the partitioned mesh (form_function -> fused exchange, form_jacobian -> fdts_halo_exchange, Krylov dots
all-reduced) reproduces the single-rank GPU trajectory to 1e-10.  Prints MGPU_CHECK_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import firedrake_ts_b200 as fts
    from firedrake_ts_b200 import mesh as M
    from firedrake_ts_b200.engine import Engine
    from firedrake_ts_b200.halo import HaloExchange
    from firedrake_ts_b200.workloads import make_problem, rel_err

    for family, dim, degree, n in (("heat", 3, 2, 8), ("cahn_hilliard", 2, 1, 12), ("burgers", 2, 2, 8)):
        part = M.Partition(M.default_grid(world, dim), rank)
        form, bcs, u, udot = make_problem(family, dim, degree, n, partition=part, tile=2, device=dev)
        V = form.space
        eng = Engine(form, bcs, device=dev)
        eng.halo_setup()
        halo = HaloExchange(V, eng, group=None, device=dev)
        x_ref, xd_ref = u.data.clone(), udot.data.clone()  # make_problem fills the ghosts with the owners' values
        F_ref = eng.ifunction(0.0, x_ref, xd_ref)

        def scrambled():
            x, xd = x_ref.clone(), xd_ref.clone()
            for m in range(V.ncomp):
                idx = V.dof_index(torch.arange(V.num_owned_nodes, V.num_nodes), m).to(dev)
                x[idx] = -777.0
                xd[idx] = 555.0
            return x, xd

        for _ in range(3):
            x, xd = scrambled()
            eng.halo_exchange(x, xd)
            assert torch.equal(x, x_ref) and torch.equal(xd, xd_ref), (family, rank, "halo_exchange")
            x, xd = scrambled()
            F = eng.ifunction_exchange(0.0, x, xd)
            assert torch.equal(x, x_ref) and torch.equal(xd, xd_ref), (family, rank, "ifunction_exchange ghosts")
            assert rel_err(F.cpu().numpy(), F_ref.cpu().numpy()) < 1e-14, (family, rank)
            x, xd = scrambled()
            F2 = eng.ifunction_overlapped(0.0, x, xd, halo)
            assert rel_err(F2.cpu().numpy(), F_ref.cpu().numpy()) < 1e-14, (family, rank, "python path")
        torch.cuda.synchronize()
        dist.barrier()

    # the Firedrake adapter on the partition: PETSc-shaped global Vecs (owned dofs) in, ghosts refreshed through
    # the star-forest neighbour lists (explicit ghost node lists) and the library's NCCL exchange
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from firedrake_ts_b200 import adapter_firedrake as A
    from fake_petsc import (FakeBC, FakeFiredrakeSpace, FakeMat, FakeProblem, FakeTS, FakeVec, StubRefTSContext,
                            fake_dmhooks, sf_graph_of)

    class Comm:  # mpi4py-shaped communicator over torch.distributed
        rank, size = rank, world

        @staticmethod
        def alltoall(lists):
            box = [None] * world
            dist.all_gather_object(box, lists)
            return [box[o][rank] for o in range(world)]

        @staticmethod
        def bcast(obj, root=0):
            b = [obj]
            dist.broadcast_object_list(b, src=root)
            return b[0]

    for family, dim, degree, n in (("heat", 3, 2, 6), ("cahn_hilliard", 2, 1, 10)):
        part = M.Partition(M.default_grid(world, dim), rank)
        form, bcs, u, udot = make_problem(family, dim, degree, n, partition=part, tile=2, device=dev)
        V = form.space
        eng = Engine(form, bcs, device=dev)
        rp, ci = eng.csr()
        Vfd = FakeFiredrakeSpace(V, comm=Comm, sf_graph=sf_graph_of(V))
        problem = FakeProblem(Vfd, V.num_owned_dofs, bcs=[FakeBC(bc.nodes.cpu().numpy()) for bc in bcs], device=dev)
        jac = FakeMat(rp.cpu().numpy(), ci.cpu().numpy(), device=dev)
        cls = A.make_context_class(StubRefTSContext, hooks=fake_dmhooks, values_pointer=lambda m: m.values.data_ptr())
        ctx = cls(problem, appctx={"b200_form": (form.family, form.params)}, jac=jac)
        assert ctx._b200.nranks == world and ctx._b200.engine.has_native_halo

        def owned(vec):  # global (owned-dof) view of a local vector
            if V.layout == "interleaved":
                return vec[: V.num_owned_dofs].clone()
            return torch.cat([vec[m * V.num_nodes: m * V.num_nodes + V.num_owned_nodes] for m in range(V.ncomp)])

        X, Xd, F = FakeVec(array=owned(u.data)), FakeVec(array=owned(udot.data)), FakeVec(V.num_owned_dofs, device=dev)
        ts = FakeTS(ctx)
        cls.form_function(ts, 0.0, X, Xd, F)
        assert rel_err(F.array.cpu().numpy(), eng.ifunction(0.0, u.data, udot.data).cpu().numpy()) < 1e-14, (family, rank)
        assert torch.equal(ctx._b200.x, u.data) and torch.equal(ctx._b200.xdot, udot.data), (family, rank, "ghosts")
        cls.form_jacobian(ts, 0.0, X, Xd, 10.0, jac, jac)
        assert rel_err(jac.values.cpu().numpy(), eng.ijacobian(0.0, u.data, udot.data, 10.0).cpu().numpy()) < 1e-14
        torch.cuda.synchronize()
        dist.barrier()

    # DAESolver on the partition vs single rank
    n = 6
    part = M.Partition(M.default_grid(world, 2), rank)
    form, bcs, u, udot = make_problem("heat", 2, 2, n, partition=part, tile=2, device=dev)
    udot.data.zero_()
    params = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-13, "snes_rtol": 1e-12}
    solver = fts.DAESolver(fts.DAEProblem(form, u, udot, (0.0, 0.3), bcs=bcs), comm=dist.group.WORLD,
                           solver_parameters=params)
    assert solver._ctx._native_halo
    solver.solve()
    V = form.space
    key = V._lex_key(V.node_lattice)[: V.num_owned_nodes]
    form1, bcs1, u1, udot1 = make_problem("heat", 2, 2, n, tile=2, device=dev)
    udot1.data.zero_()
    fts.DAESolver(fts.DAEProblem(form1, u1, udot1, (0.0, 0.3), bcs=bcs1), solver_parameters=params).solve()
    V1 = form1.space
    full = torch.zeros(V1.global_num_nodes, dtype=torch.float64, device=dev)
    full[V1._lex_key(V1.node_lattice)] = u1.data
    err = float((u.owned() - full[key]).abs().max())
    assert err < 1e-10, err
    assert solver.ts.getStepNumber() == 3
    dist.barrier()
    if rank == 0:
        print("MGPU_CHECK_OK world", world, "trajectory err", err, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
