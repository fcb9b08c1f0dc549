"""Generates tests/golden/independent_kats.json: known answers for the oracle that share NOTHING with the
product package (no import from firedrake_ts_b200 at all).

What is restated here, from the conventions written down in SURVEY.md section 8(c) item 4 (FIAT / UFC as
Firedrake uses them), is only the local node ORDER of the Lagrange elements:

    vertices 0..d, then edges in the order   triangle (1,2) (0,2) (0,1)
                                              tetrahedron (2,3) (1,3) (1,2) (0,3) (0,2) (0,1)
    (interval: the cell interior), nodes on an edge from the lower to the higher local vertex,
    then tetrahedron faces (1,2,3) (0,2,3) (0,1,3) (0,1,2), then the cell interior;
    P3 edge nodes at 1/3, 2/3 ("equispaced") or at the Gauss-Lobatto points (1 -+ 1/sqrt 5)/2 ("gll",
    Firedrake's default variant), face / interior nodes at the centroid.

tests/test_oracle_independent.py checks the product's tables (fem.lattice_weights / fem.nodes_barycentric)
against NODE_TABLE below, so a wrong local order or a wrong GLL position in the product can no longer hide
behind an oracle that was fed the same tables.

Everything else is exact polynomial algebra in 50-digit arithmetic (mpmath): Lagrange bases by inverting the
Vandermonde matrix of the node set, products / derivatives of polynomials as coefficient dictionaries, and
the closed-form monomial integral over the reference simplex  int x^a = prod a_i! / (d + |a|)!.  Because all
cells are affine, every integrand of the reference's example forms is a polynomial, so the element residuals
and Jacobians below are exact (to 50 digits) for

    heat   (examples/heat.py:11)              GLL-P3 mass / stiffness / load, dim 1-3
    bbm    (examples/bbm.py:35-40)            P2 triangle, GLL-P3 tetrahedron, non-zero state
    burgers (examples/burgers.py:21-25)       vector P2 triangle, vector P1 tetrahedron, non-zero state
    cahn-hilliard (examples/cahn-hilliard.py:26-38)  mixed P1xP1 and P2xP2 triangle, P1xP1 tetrahedron

with J = shift*dF/dudot + dF/du (firedrake_ts/ts_solver.py:108).  The file also stores the monomial
coefficients of every basis function so that the test can tabulate B at quadrature points of ITS OWN rule
(Gauss-Legendre-Duffy from numpy) and feed the oracle tables that never touched fem.py.

Run: python tests/golden/make_golden_independent.py
"""
import itertools
import json
import os
from fractions import Fraction

import mpmath as mp

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))

EDGES = {1: [(0, 1)], 2: [(1, 2), (0, 2), (0, 1)], 3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]}
FACES = {3: [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]}


def M(x):
    """exact conversion of ints / Fractions to mpf"""
    f = Fraction(x)
    return mp.mpf(f.numerator) / mp.mpf(f.denominator)


def edge_params(degree, variant):
    if degree == 2:
        return [mp.mpf(1) / 2]
    if degree == 3:
        if variant == "gll":
            s = 1 / mp.sqrt(5)
            return [(1 - s) / 2, (1 + s) / 2]
        return [mp.mpf(1) / 3, mp.mpf(2) / 3]
    return []


def node_table(dim, degree, variant):
    """barycentric coordinates of the nodes in FIAT order (list of (dim+1)-tuples of mpf)"""
    nv = dim + 1
    nodes = []
    for a in range(nv):
        nodes.append(tuple(mp.mpf(1 if b == a else 0) for b in range(nv)))
    par = edge_params(degree, variant)
    for (a, b) in EDGES[dim]:
        for t in par:  # from the lower local vertex a to the higher one b
            p = [mp.mpf(0)] * nv
            p[a], p[b] = 1 - t, t
            nodes.append(tuple(p))
    if degree == 3 and dim == 2:
        nodes.append((mp.mpf(1) / 3,) * 3)
    if degree == 3 and dim == 3:
        for f in FACES[3]:
            p = [mp.mpf(0)] * nv
            for a in f:
                p[a] = mp.mpf(1) / 3
            nodes.append(tuple(p))
    return nodes


# ---- polynomials in xi_1..xi_d as {exponent tuple: mpf}
def padd(p, q, s=1):
    r = dict(p)
    for e, c in q.items():
        r[e] = r.get(e, mp.mpf(0)) + s * c
    return r


def pmul(p, q):
    r = {}
    for e1, c1 in p.items():
        for e2, c2 in q.items():
            e = tuple(a + b for a, b in zip(e1, e2))
            r[e] = r.get(e, mp.mpf(0)) + c1 * c2
    return r


def pscale(p, s):
    return {e: c * s for e, c in p.items()}


def pdiff(p, a):
    r = {}
    for e, c in p.items():
        if e[a] > 0:
            e2 = list(e)
            e2[a] -= 1
            r[tuple(e2)] = r.get(tuple(e2), mp.mpf(0)) + c * e[a]
    return r


def pint(p, dim):
    """integral over the reference simplex {xi >= 0, sum xi <= 1}"""
    tot = mp.mpf(0)
    for e, c in p.items():
        num = mp.mpf(1)
        for k in e:
            num *= mp.factorial(k)
        tot += c * num / mp.factorial(dim + sum(e))
    return tot


def lagrange_basis(dim, degree, variant):
    nodes = node_table(dim, degree, variant)
    mons = [e for e in itertools.product(range(degree + 1), repeat=dim) if sum(e) <= degree]
    assert len(mons) == len(nodes)
    V = mp.matrix(len(nodes), len(mons))
    for i, b in enumerate(nodes):
        xi = b[1:]  # xi_a = lambda_a, a = 1..d  (v0 = origin)
        for m, e in enumerate(mons):
            v = mp.mpf(1)
            for a in range(dim):
                v *= xi[a] ** e[a]
            V[i, m] = v
    C = V ** -1  # column i = coefficients of phi_i
    return [{mons[m]: C[m, i] for m in range(len(mons))} for i in range(len(nodes))], mons


def lincomb(coeffs, polys):
    r = {}
    for c, p in zip(coeffs, polys):
        r = padd(r, pscale(p, M(c)))
    return r


def geometry(verts, dim):
    J = mp.matrix(dim, dim)  # J[k][a] = d x_k / d xi_a
    for k in range(dim):
        for a in range(dim):
            J[k, a] = M(verts[a + 1][k]) - M(verts[0][k])
    return abs(mp.det(J)), J ** -1  # Jinv[a][k] = d xi_a / d x_k


def phys_grad(p, Jinv, dim):
    """[d p / d x_k for k] of a reference polynomial"""
    ref = [pdiff(p, a) for a in range(dim)]
    out = []
    for k in range(dim):
        g = {}
        for a in range(dim):
            g = padd(g, pscale(ref[a], Jinv[a, k]))
        out.append(g)
    return out


def pdot(ga, gb):
    r = {}
    for x, y in zip(ga, gb):
        r = padd(r, pmul(x, y))
    return r


def element(family, dim, degree, variant, verts, ncomp, u, ud, shift, params):
    """exact element residual Fe[i*ncomp+m] and Jacobian Ae[(i,m),(j,n)] (same local layout as the
    oracle's oracle_element: node-major, component-minor)"""
    phi, _ = lagrange_basis(dim, degree, variant)
    nd = len(phi)
    adet, Jinv = geometry(verts, dim)
    G = [phys_grad(p, Jinv, dim) for p in phi]
    I = lambda p: pint(p, dim) * adet  # noqa: E731
    one = {(0,) * dim: mp.mpf(1)}
    uc = [lincomb([u[i * ncomp + m] for i in range(nd)], phi) for m in range(ncomp)]
    udc = [lincomb([ud[i * ncomp + m] for i in range(nd)], phi) for m in range(ncomp)]
    gu = [phys_grad(p, Jinv, dim) for p in uc]
    gud = [phys_grad(p, Jinv, dim) for p in udc]
    ne = nd * ncomp
    Fe = [mp.mpf(0)] * ne
    Ae = [[mp.mpf(0)] * ne for _ in range(ne)]
    sh = M(shift)
    if family == "heat":
        src = M(params[0])
        for i in range(nd):
            Fe[i] = I(padd(padd(pmul(udc[0], phi[i]), pdot(gu[0], G[i])), pscale(phi[i], src), -1))
            for j in range(nd):
                Ae[i][j] = I(padd(pscale(pmul(phi[j], phi[i]), sh), pdot(G[j], G[i])))
    elif family == "bbm":
        ux = gu[0][0]
        for i in range(nd):
            t = padd(pmul(udc[0], phi[i]), pmul(ux, phi[i]))
            t = padd(t, pmul(pmul(uc[0], ux), phi[i]))
            t = padd(t, pmul(gud[0][0], G[i][0]))
            Fe[i] = I(t)
            for j in range(nd):
                t = pscale(padd(pmul(phi[j], phi[i]), pmul(G[j][0], G[i][0])), sh)
                t = padd(t, pmul(G[j][0], phi[i]))
                t = padd(t, pmul(padd(pmul(ux, phi[j]), pmul(uc[0], G[j][0])), phi[i]))
                Ae[i][j] = I(t)
    elif family == "burgers":
        nu = M(params[0])
        for i in range(nd):
            for m in range(ncomp):
                adv = {}
                for k in range(dim):
                    adv = padd(adv, pmul(uc[k], gu[m][k]))
                t = padd(pmul(udc[m], phi[i]), pmul(adv, phi[i]))
                t = padd(t, pscale(pdot(gu[m], G[i]), nu))
                Fe[i * ncomp + m] = I(t)
                for j in range(nd):
                    ugj = {}
                    for k in range(dim):
                        ugj = padd(ugj, pmul(uc[k], G[j][k]))
                    for n in range(ncomp):
                        t = pmul(pmul(phi[j], gu[m][n]), phi[i])  # (du . grad) u
                        if m == n:
                            t = padd(t, pscale(pmul(phi[j], phi[i]), sh))
                            t = padd(t, pmul(ugj, phi[i]))
                            t = padd(t, pscale(pdot(G[j], G[i]), nu))
                        Ae[i * ncomp + m][j * ncomp + n] = I(t)
    elif family == "cahn_hilliard":
        lam = M(params[0])
        c, mu = uc[0], uc[1]
        omc = padd(one, c, -1)
        om2c = padd(one, pscale(c, 2), -1)
        dfdc = pscale(pmul(pmul(c, omc), om2c), 200)                       # f'  = 200 c (1-c)(1-2c)
        d2f = pscale(padd(padd(one, pscale(c, 6), -1), pscale(pmul(c, c), 6)), 200)  # f'' = 200 (1 - 6c + 6c^2)
        for i in range(nd):
            Fe[i * 2 + 0] = I(padd(pmul(udc[0], phi[i]), pdot(gu[1], G[i])))
            t = padd(pmul(mu, phi[i]), pmul(dfdc, phi[i]), -1)
            Fe[i * 2 + 1] = I(padd(t, pscale(pdot(gu[0], G[i]), lam), -1))
            for j in range(nd):
                Mij, Kij = pmul(phi[j], phi[i]), pdot(G[j], G[i])
                Ae[i * 2 + 0][j * 2 + 0] = I(pscale(Mij, sh))
                Ae[i * 2 + 0][j * 2 + 1] = I(Kij)
                Ae[i * 2 + 1][j * 2 + 0] = I(padd(pscale(pmul(d2f, Mij), -1), pscale(Kij, lam), -1))
                Ae[i * 2 + 1][j * 2 + 1] = I(Mij)
    else:
        raise ValueError(family)
    return Fe, Ae


def rat(x):
    return Fraction(x).limit_denominator(10 ** 6)


def make_state(n, seed):
    """small rationals, deterministic, non-zero"""
    u = [Fraction(((seed * 7 + 3 * k * k + 5 * k) % 17) - 8, 9) for k in range(n)]
    ud = [Fraction(((seed * 5 + 2 * k * k + 3 * k) % 13) - 6, 7) for k in range(n)]
    return u, ud


CASES = [
    # family, dim, degree, variant, ncomp, vertices, shift, params
    ("heat", 1, 3, "gll", 1, [[0], [Fraction(3, 4)]], Fraction(5, 2), [1]),
    ("heat", 2, 3, "gll", 1, [[0, 0], [1, Fraction(1, 4)], [Fraction(1, 2), 2]], Fraction(5, 2), [1]),
    ("heat", 3, 3, "gll", 1, [[0, 0, 0], [1, Fraction(1, 5), 0], [Fraction(1, 3), 1, Fraction(1, 7)],
                              [0, Fraction(1, 2), 2]], Fraction(5, 2), [1]),
    ("heat", 3, 3, "gll", 1, [[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]], 3, [1]),  # a Kuhn tetrahedron
    ("heat", 3, 2, "gll", 1, [[0, 0, 0], [1, Fraction(1, 5), 0], [Fraction(1, 3), 1, Fraction(1, 7)],
                              [0, Fraction(1, 2), 2]], 10, [1]),
    ("bbm", 1, 1, "gll", 1, [[Fraction(1, 5)], [Fraction(6, 5)]], Fraction(7, 3), []),
    ("bbm", 2, 2, "gll", 1, [[0, 0], [2, Fraction(1, 2)], [Fraction(1, 3), 1]], Fraction(7, 3), []),
    ("bbm", 3, 3, "gll", 1, [[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]], Fraction(25, 2), []),
    ("bbm", 3, 3, "equispaced", 1, [[0, 0, 0], [1, Fraction(1, 5), 0], [Fraction(1, 3), 1, Fraction(1, 7)],
                                    [0, Fraction(1, 2), 2]], 2, []),
    ("burgers", 2, 2, "gll", 2, [[0, 0], [2, Fraction(1, 2)], [Fraction(1, 3), 1]], Fraction(3, 2), [Fraction(1, 10000)]),
    ("burgers", 2, 2, "gll", 2, [[1, 1], [0, 0], [2, 0]], 4, [Fraction(1, 8)]),  # negatively oriented
    ("burgers", 3, 1, "gll", 3, [[0, 0, 0], [1, Fraction(1, 5), 0], [Fraction(1, 3), 1, Fraction(1, 7)],
                                 [0, Fraction(1, 2), 2]], Fraction(3, 2), [Fraction(1, 100)]),
    ("cahn_hilliard", 2, 1, "gll", 2, [[0, 0], [Fraction(1, 96), 0], [Fraction(1, 96), Fraction(1, 96)]], 200000,
     [Fraction(1, 100)]),
    ("cahn_hilliard", 2, 2, "gll", 2, [[0, 0], [2, Fraction(1, 2)], [Fraction(1, 3), 1]], Fraction(9, 2), [Fraction(1, 100)]),
    ("cahn_hilliard", 3, 1, "gll", 2, [[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]], 7, [Fraction(1, 100)]),
]


def s(x):
    return mp.nstr(x, 25)


def main():
    out = {"generator": "tests/golden/make_golden_independent.py (mpmath %s, 50 digits)" % mp.__version__,
           "node_tables": {}, "bases": {}, "cases": []}
    for dim in (1, 2, 3):
        for degree in (1, 2, 3):
            for variant in (("equispaced", "gll") if degree == 3 else ("gll",)):
                key = f"{dim}_{degree}_{variant}"
                out["node_tables"][key] = [[s(c) for c in b] for b in node_table(dim, degree, variant)]
                phi, mons = lagrange_basis(dim, degree, variant)
                out["bases"][key] = {"monomials": [list(e) for e in mons],
                                     "coefficients": [[s(p[e]) for e in mons] for p in phi]}
    for k, (family, dim, degree, variant, ncomp, verts, shift, params) in enumerate(CASES):
        nd = len(node_table(dim, degree, variant))
        u, ud = make_state(nd * ncomp, k)
        if family == "cahn_hilliard":  # concentration near the example's 0.63 +- 0.1, small chemical potential
            u = [Fraction(63, 100) + x / 40 if (i % 2 == 0) else x / 10 for i, x in enumerate(u)]
        Fe, Ae = element(family, dim, degree, variant, verts, ncomp, u, ud, shift, params)
        out["cases"].append({
            "family": family, "dim": dim, "degree": degree, "variant": variant, "ncomp": ncomp,
            "vertices": [[str(Fraction(v)) for v in row] for row in verts],
            "vertices_float": [[float(Fraction(v)) for v in row] for row in verts],
            "shift": float(Fraction(shift)), "params": [float(Fraction(p)) for p in params],
            "u": [float(x) for x in u], "udot": [float(x) for x in ud],
            "u_exact": [str(x) for x in u], "udot_exact": [str(x) for x in ud],
            "F": [s(v) for v in Fe], "A": [[s(v) for v in row] for row in Ae],
        })
        print("case", k, family, dim, degree, variant, "done", flush=True)
    json.dump(out, open(os.path.join(HERE, "independent_kats.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
