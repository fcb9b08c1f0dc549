"""Dumps assemble(F) / assemble(J) of the REAL reference stack (Firedrake + firedrake_ts) for the
configurations C1-C5 of SURVEY.md section 8(d) at small n, as tests/golden/firedrake/<name>.npz.

CANNOT RUN IN THE BUILD IMAGE (no ufl / firedrake / pyop2 / petsc4py); it is
committed so that parity can be pinned to the reference itself the day an install exists:

    python tests/golden/make_firedrake_fixtures.py            # inside a Firedrake venv, one rank
    python -m pytest tests/test_firedrake_fixtures.py         # oracle (and, with a GPU, the engine) vs the dump

Every fixture holds the INPUTS the engine consumes (coordinates, cell->vertex and cell->node maps, Dirichlet
nodes, state vectors, shift, constants) and the reference OUTPUTS at the boundary the engine replaces:
the residual assembled by ctx._assemble_residual(tensor=ctx._F) (firedrake_ts/solving_utils.py:258) and the
AIJ matrix assembled by ctx._assemble_jac(ctx._jac) (:304-305) as CSR (rowptr, colidx, values).  The loader
feeds the maps to the oracle, so "sparsity pattern and DOF/CSR index maps bit-exact" is checked on
Firedrake's own numbering.

Accessor names are upstream Firedrake API as of 2024; verify on first use.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "firedrake")


def build(name, n):
    import firedrake as fd
    import firedrake_ts

    if name == "C1":  # examples/heat.py: P1 heat on UnitSquareMesh
        mesh = fd.UnitSquareMesh(n, n)
        V = fd.FunctionSpace(mesh, "CG", 1)
    elif name == "C2":  # examples/burgers.py: vector CG2
        mesh = fd.UnitSquareMesh(n, n)
        V = fd.VectorFunctionSpace(mesh, "CG", 2)
    elif name == "C3":  # examples/cahn-hilliard.py on triangles (the literal example uses quadrilateral=True)
        mesh = fd.UnitSquareMesh(n, n)
        W = fd.FunctionSpace(mesh, "CG", 1)
        V = W * W
    elif name == "C3Q":  # the literal example: Q1 x Q1 on quadrilaterals
        mesh = fd.UnitSquareMesh(n, n, quadrilateral=True)
        W = fd.FunctionSpace(mesh, "Lagrange", 1)
        V = W * W
    elif name == "C4":  # headline: 3-D heat, P2 tetrahedra
        mesh = fd.UnitCubeMesh(n, n, n)
        V = fd.FunctionSpace(mesh, "CG", 2)
    elif name == "C5":  # examples/bbm.py form on P3 tetrahedra
        mesh = fd.UnitCubeMesh(n, n, n)
        V = fd.FunctionSpace(mesh, "CG", 3)
    else:
        raise ValueError(name)
    u, udot = fd.Function(V), fd.Function(V)
    rng = np.random.default_rng(0)
    x = fd.SpatialCoordinate(mesh)
    params, bcs, family = [0.0] * 4, [], None
    if name in ("C1", "C4"):
        v = fd.TestFunction(V)
        F = fd.inner(udot, v) * fd.dx + fd.inner(fd.grad(u), fd.grad(v)) * fd.dx - 1.0 * v * fd.dx  # heat.py:11
        bcs = [fd.DirichletBC(V, 0.0, "on_boundary")]
        u.interpolate(fd.sin(fd.pi * x[0]) * fd.sin(fd.pi * x[1]) * (fd.sin(fd.pi * x[2]) if name == "C4" else 1.0))
        family, params[0] = "heat", 1.0
    elif name == "C2":
        v = fd.TestFunction(V)
        nu = 1.0e-4
        F = (fd.inner(udot, v) + fd.inner(fd.dot(u, fd.nabla_grad(u)), v) + nu * fd.inner(fd.grad(u), fd.grad(v))) * fd.dx
        u.interpolate(fd.as_vector([fd.sin(fd.pi * x[0]), 0.0]))  # burgers.py:17
        family, params[0] = "burgers", nu
    elif name in ("C3", "C3Q"):
        q, v = fd.TestFunctions(V)
        c, mu = fd.split(u)
        c_t, _mu_t = fd.split(udot)
        lmbda = 1.0e-2
        c_ = fd.variable(c)
        f = 100 * c_ ** 2 * (1 - c_) ** 2
        dfdc = fd.diff(f, c_)
        F0 = c_t * q * fd.dx + fd.dot(fd.grad(mu), fd.grad(q)) * fd.dx
        F1 = mu * v * fd.dx - dfdc * v * fd.dx - lmbda * fd.dot(fd.grad(c), fd.grad(v)) * fd.dx
        F = F0 + F1  # cahn-hilliard.py:26-38
        u.sub(0).dat.data[:] = 0.63 + 0.2 * (0.5 - np.random.default_rng(11).random(u.sub(0).dat.data.shape[0]))
        family, params[0] = "cahn_hilliard", lmbda
    elif name == "C5":
        v = fd.TestFunction(V)
        F = (fd.inner(udot, v) + fd.inner(u.dx(0), v) + fd.inner(u * u.dx(0), v) + fd.inner(udot.dx(0), v.dx(0))) * fd.dx
        u.interpolate(fd.sin(fd.pi * x[0]) * fd.sin(fd.pi * x[1]) * fd.sin(fd.pi * x[2]))
        family = "bbm"
    with udot.dat.vec_wo as vec:
        vec.setArray(rng.standard_normal(vec.getLocalSize()))
    problem = firedrake_ts.DAEProblem(F, u, udot, (0.0, 1.0), bcs=bcs)
    solver = firedrake_ts.DAESolver(problem)
    return mesh, V, u, udot, bcs, problem, solver, family, params


def dump(name, n, shift=10.0):
    mesh, V, u, udot, bcs, problem, solver, family, params = build(name, n)
    ctx = solver._ctx
    ctx._time.assign(0.0)
    ctx._shift.assign(shift)
    ctx._assemble_residual(tensor=ctx._F)   # firedrake_ts/solving_utils.py:258
    ctx._assemble_jac(ctx._jac)             # firedrake_ts/solving_utils.py:304-305
    mat = ctx._jac.petscmat
    if mat.getType() == "nest":
        raise SystemExit("use mat_type aij (the default) for the dump")
    rowptr, colidx, vals = mat.getValuesCSR()
    coords = mesh.coordinates
    out = dict(
        family=family, params=np.array(params), shift=shift, dim=mesh.geometric_dimension(),
        degree=V.sub(0).ufl_element().degree() if len(V) > 1 else V.ufl_element().degree(),
        cell=str(mesh.ufl_cell()), value_size=V.value_size if len(V) == 1 else len(V),
        layout="blocked" if len(V) > 1 else "interleaved",
        coords=np.array(coords.dat.data_ro), cell_coord=np.array(coords.cell_node_map().values),
        cell_node=np.array((V.sub(0) if len(V) > 1 else V).cell_node_map().values),
        bc_nodes=np.unique(np.concatenate([bc.nodes for bc in bcs])) if bcs else np.zeros(0, dtype=np.int32),
        F=np.concatenate([d.data_ro.reshape(-1) for d in ctx._F.dat]),
        rowptr=np.array(rowptr), colidx=np.array(colidx), values=np.array(vals),
        u=np.concatenate([d.data_ro.reshape(-1) for d in u.dat]),
        udot=np.concatenate([d.data_ro.reshape(-1) for d in udot.dat]),
    )
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"{name}_n{n}.npz"), **out)
    print("wrote", name, n, "dofs", out["u"].size, "nnz", out["values"].size)


if __name__ == "__main__":
    try:
        import firedrake  # noqa: F401
    except ImportError:
        sys.exit("firedrake is not importable here: run this inside a Firedrake + firedrake_ts environment")
    for name, n in (("C1", 8), ("C2", 4), ("C3", 6), ("C3Q", 6), ("C4", 3), ("C5", 2)):
        dump(name, n)
