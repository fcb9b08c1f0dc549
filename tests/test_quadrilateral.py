"""Q1 x Q1 quadrilateral cells (reference examples/cahn-hilliard.py:14-16, UnitSquareMesh(96, 96,
quadrilateral=True) with "Lagrange" 1): element tables, oracle known answers, finite-difference
consistency of the Jacobian, and the literal example's first steps through DAESolver (CPU, oracle-assembled)."""
import math

import mpmath
import numpy as np
import pytest
import torch

from firedrake_ts_b200 import fem, forms, mesh as M
from firedrake_ts_b200.function import Function
from helpers import make_problem, rel_err
from oracle import Oracle


def test_q1_tables():
    et = fem.element_tables(2, 1, 4, "gll", "quadrilateral")
    assert et.nd == 4 and et.nq == 9
    assert abs(et.w.sum() - 1.0) < 1e-15
    # Kronecker property at the reference vertices, in the tensor-product order (second factor fastest)
    T = fem.tabulate(2, 1, np.array(fem.QUAD_VERTICES, dtype=float), cell="quadrilateral")
    assert np.allclose(T[0], np.eye(4), atol=1e-15)
    # partition of unity, derivatives sum to zero
    assert np.allclose(et.B[:, 0, :].sum(axis=1), 1.0)
    assert np.allclose(et.B[:, 1:, :].sum(axis=2), 0.0)
    # exact Q1 mass matrix on the unit square: (1/36) [[4,2,2,1],[2,4,1,2],[2,1,4,2],[1,2,2,4]]
    Mref = np.array([[4, 2, 2, 1], [2, 4, 1, 2], [2, 1, 4, 2], [1, 2, 2, 4]]) / 36.0
    assert np.allclose(et.R[0, 0], Mref, atol=1e-15)


def _one_cell(verts, family, ncomp, layout, params, qdegree=4):
    et = fem.element_tables(2, 1, qdegree, "gll", "quadrilateral")
    coords = np.array(verts, dtype=np.float64)
    return Oracle(dim=2, nd=4, ncomp=ncomp, layout=layout, family=family, params=params, coords=coords,
                  cell_coord=np.arange(4, dtype=np.int32)[None, :], cell_node=np.arange(4, dtype=np.int32)[None, :],
                  B=et.B, w=et.w, num_nodes=4, num_owned_nodes=4, cell_type=1), et


def test_parallelogram_known_answer():
    """On a parallelogram the map is affine: mass = |det| * reference mass, stiffness = |det| sum_ab G_ab R_ab
    (exact rationals); vertex order (0,0),(0,1),(1,0),(1,1) of the reference square."""
    a, b, o = np.array([2.0, 0.5]), np.array([0.25, 1.5]), np.array([0.3, -0.2])
    verts = [o, o + b, o + a, o + a + b]  # xi along a, eta along b
    orc, et = _one_cell(verts, forms.HEAT, 1, "interleaved", (1.0, 0, 0, 0), qdegree=2)
    shift = 3.0
    Fe, Ae = orc.element(0, np.zeros(4), np.zeros(4), shift)
    J = np.stack([a, b], axis=1)  # dx/dxi
    det = abs(np.linalg.det(J))
    Ji = np.linalg.inv(J)  # dxi/dx
    G = Ji @ Ji.T
    Mref = np.array([[4, 2, 2, 1], [2, 4, 1, 2], [2, 1, 4, 2], [1, 2, 2, 4]]) / 36.0
    # reference stiffness tensors of Q1 (exact): R_ab[i][j] = int d_a phi_i d_b phi_j
    R = et.R
    K = sum(G[p, q] * R[1 + p, 1 + q] for p in range(2) for q in range(2))
    assert np.allclose(Ae, det * (shift * Mref + K), rtol=1e-13, atol=1e-14)
    assert np.allclose(Fe, -det * 0.25, rtol=1e-13)


def test_general_quadrilateral_against_mpmath():
    """Non-parallelogram cell: the oracle's 3x3 Gauss sums against a 50-digit restatement of the same sums
    (bilinear map, Jacobian at the point, physical gradients) written independently in mpmath; and the
    rule-independent facts: sum of the mass matrix = area, constants in the stiffness null space."""
    mpmath.mp.dps = 50
    verts = [(0.0, 0.0), (0.1, 1.2), (1.3, -0.1), (1.9, 1.4)]
    lam = 0.01
    orc, et = _one_cell(verts, forms.CAHN_HILLIARD, 2, "blocked", (lam, 0, 0, 0))
    rng = np.random.default_rng(5)
    u = np.concatenate([0.63 + 0.2 * (0.5 - rng.random(4)), 0.1 * rng.standard_normal(4)])
    ud = rng.standard_normal(8)
    shift = 7.0
    Fe, Ae = orc.element(0, u, ud, shift)  # element order: (node i, field m) -> i*2 + m

    g = [mpmath.mpf(1) / 2 - mpmath.sqrt(mpmath.mpf(3) / 5) / 2, mpmath.mpf(1) / 2, mpmath.mpf(1) / 2 + mpmath.sqrt(mpmath.mpf(3) / 5) / 2]
    gw = [mpmath.mpf(5) / 18, mpmath.mpf(8) / 18, mpmath.mpf(5) / 18]
    X = [[mpmath.mpf(v[0]), mpmath.mpf(v[1])] for v in verts]
    c = [mpmath.mpf(float(x)) for x in u[:4]]
    mu = [mpmath.mpf(float(x)) for x in u[4:]]
    ct = [mpmath.mpf(float(x)) for x in ud[:4]]
    F = [[mpmath.mpf(0)] * 2 for _ in range(4)]
    A = [[mpmath.mpf(0)] * 8 for _ in range(8)]
    area = mpmath.mpf(0)
    for xi, wx in zip(g, gw):
        for eta, wy in zip(g, gw):
            ax, ay = (1 - xi, xi), (1 - eta, eta)
            phi, dxi, deta = [], [], []
            for (ix, iy) in fem.QUAD_VERTICES:
                phi.append(ax[ix] * ay[iy])
                dxi.append((-1 if ix == 0 else 1) * ay[iy])
                deta.append(ax[ix] * (-1 if iy == 0 else 1))
            J = [[sum(X[v][k] * (dxi[v], deta[v])[a] for v in range(4)) for a in range(2)] for k in range(2)]
            det = J[0][0] * J[1][1] - J[0][1] * J[1][0]
            Ji = [[J[1][1] / det, -J[0][1] / det], [-J[1][0] / det, J[0][0] / det]]  # Ji[a][k] = dxi_a/dx_k
            G = [[Ji[0][k] * dxi[v] + Ji[1][k] * deta[v] for k in range(2)] for v in range(4)]
            w = wx * wy * abs(det)
            area += w
            cq = sum(phi[v] * c[v] for v in range(4))
            muq = sum(phi[v] * mu[v] for v in range(4))
            ctq = sum(phi[v] * ct[v] for v in range(4))
            gc = [sum(G[v][k] * c[v] for v in range(4)) for k in range(2)]
            gmu = [sum(G[v][k] * mu[v] for v in range(4)) for k in range(2)]
            dfdc = 200 * cq * (1 - cq) * (1 - 2 * cq)
            d2f = 200 * (1 - 6 * cq + 6 * cq * cq)
            for i in range(4):
                F[i][0] += w * (ctq * phi[i] + gmu[0] * G[i][0] + gmu[1] * G[i][1])
                F[i][1] += w * (muq * phi[i] - dfdc * phi[i] - lam * (gc[0] * G[i][0] + gc[1] * G[i][1]))
                for j in range(4):
                    gg = G[i][0] * G[j][0] + G[i][1] * G[j][1]
                    A[2 * i][2 * j] += w * shift * phi[i] * phi[j]
                    A[2 * i][2 * j + 1] += w * gg
                    A[2 * i + 1][2 * j] += w * (-d2f * phi[i] * phi[j] - lam * gg)
                    A[2 * i + 1][2 * j + 1] += w * phi[i] * phi[j]
    Fm = np.array([[float(x) for x in r] for r in F]).reshape(-1)
    Am = np.array([[float(x) for x in r] for r in A])
    assert rel_err(Fe, Fm) < 1e-13
    assert rel_err(Ae, Am) < 1e-13
    # shoelace area of the quadrilateral (perimeter order v0, v2, v3, v1): the bilinear Jacobian is linear in
    # (xi, eta), so the rule integrates it exactly
    P = [verts[0], verts[2], verts[3], verts[1]]
    shoelace = 0.5 * abs(sum(P[i][0] * P[(i + 1) % 4][1] - P[(i + 1) % 4][0] * P[i][1] for i in range(4)))
    assert abs(float(area) - shoelace) < 1e-14
    mass = Ae[1::2, 1::2]
    assert abs(mass.sum() - shoelace) < 1e-13
    assert np.abs(Ae[0::2, 1::2].sum(axis=1)).max() < 1e-12  # stiffness annihilates constants


@pytest.mark.parametrize("family", ["heat", "cahn_hilliard", "diffusion"])
@pytest.mark.parametrize("warp", [0.0, 0.03])
def test_jacobian_is_derivative_of_residual(family, warp):
    form, bcs, u, udot = make_problem(family, 2, 1, 5, quadrilateral=True, warp=warp)
    orc = Oracle.from_form(form, bcs)
    rp, ci = orc.pattern()
    V = form.space
    assert V.nd == 4 and V.mesh.num_cells == 25 and V.num_nodes == 36
    # interior Q1 rows couple to 9 nodes
    nrp, _ = orc.node_pattern()
    assert int(np.diff(nrp).max()) == 9
    shift = 2.5
    x, xd = u.data.numpy().copy(), udot.data.numpy().copy()
    J = orc.ijacobian(0.0, x, xd, shift)
    rng = np.random.default_rng(1)
    d = rng.standard_normal(x.size)
    if bcs:
        bn = torch.unique(torch.cat([bc.nodes for bc in bcs])).numpy()
        d[bn] = 0.0
    eps = 1e-6
    Fp = orc.ifunction(0.0, x + eps * d, xd + eps * shift * d)
    Fm = orc.ifunction(0.0, x - eps * d, xd - eps * shift * d)
    Jd = np.array([J[rp[r]:rp[r + 1]] @ d[ci[rp[r]:rp[r + 1]]] for r in range(len(rp) - 1)])
    fd = (Fp - Fm) / (2 * eps)
    if bcs:
        rows = np.setdiff1d(np.arange(len(rp) - 1), bn)
        Jd, fd = Jd[rows], fd[rows]
    assert rel_err(Jd, fd) < 1e-6


def test_heat_on_quads_equals_triangles_in_the_limit_facts():
    """global known answers on the (unwarped and warped) quadrilateral mesh: the mass matrix sums to |Omega| = 1
    and the residual of u = const, udot = 0, source 0 vanishes."""
    for warp in (0.0, 0.04):
        form, _, u, udot = make_problem("diffusion", 2, 1, 6, quadrilateral=True, warp=warp)
        u.data.fill_(1.7)
        udot.data.zero_()
        f = forms.diffusion(u, udot, mass=1.0, diffusivity=1.0, source=0.0)
        orc = Oracle.from_form(f, [])
        assert np.abs(orc.ifunction(0.0, u.data, udot.data)).max() < 1e-13
        f = forms.diffusion(u, udot, mass=1.0, diffusivity=0.0, source=0.0)
        orc = Oracle.from_form(f, [])
        assert abs(orc.ijacobian(0.0, u.data, udot.data, 1.0).sum() - 1.0) < 1e-13


def test_literal_cahn_hilliard_example_runs():
    """examples/cahn-hilliard.py as written (quadrilateral Q1 x Q1, lmbda 1e-2, dt 5e-6, two steps, seeded
    c = 0.63 + 0.2 (0.5 - rand)) through DAEProblem/DAESolver with the oracle behind the callbacks; mass of c
    is conserved by the scheme (q = 1 in F0)."""
    from firedrake_ts_b200 import DAEProblem, DAESolver
    from helpers import OracleAssembler

    n = 12  # the example's 96 at a CPU-test size
    msh = M.UnitSquareMesh(n, n, quadrilateral=True)
    V = M.FunctionSpace(msh, 1, ncomp=2, layout="blocked")
    u, u_t = Function(V), Function(V)
    rng = np.random.default_rng(11)
    nn = V.num_nodes
    u.data[:nn] = torch.from_numpy(0.63 + 0.2 * (0.5 - rng.random(nn)))
    F = forms.cahn_hilliard(u, u_t, lmbda=1.0e-2)
    dt = 5.0e-6
    params = {"ts_type": "beuler", "ts_dt": dt, "snes_rtol": 1e-9, "snes_atol": 1e-10, "ksp_type": "gmres",
              "ksp_rtol": 1e-10, "pc_type": "jacobi"}
    problem = DAEProblem(F, u, u_t, (0.0, 2 * dt))
    solver = DAESolver(problem, solver_parameters=params, assembler=OracleAssembler(F, []))
    mass_form = forms.diffusion(Function(M.FunctionSpace(msh, 1)), Function(M.FunctionSpace(msh, 1)), mass=1.0,
                                diffusivity=0.0)
    Mo = Oracle.from_form(mass_form, [])
    rp, ci = Mo.pattern()
    mv = Mo.ijacobian(0.0, np.zeros(nn), np.zeros(nn), 1.0)

    def total_c():
        c = u.data[:nn].numpy()
        return sum(mv[rp[r]:rp[r + 1]] @ c[ci[rp[r]:rp[r + 1]]] for r in range(nn))

    m0 = total_c()
    solver.solve()
    assert solver.ts.getStepNumber() == 2
    assert abs(total_c() - m0) < 1e-9 * abs(m0)
    assert torch.isfinite(u.data).all()
