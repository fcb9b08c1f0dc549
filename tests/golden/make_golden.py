"""Generates tests/golden/element_kats.json: exact (rational arithmetic, sympy) element mass and
stiffness matrices and load vectors of Lagrange P1-P3 (equispaced nodes) on affine simplices with
rational vertices.  Independent of firedrake_ts_b200.fem (symbolic Lagrange interpolation +
closed-form monomial integrals), so it pins both the tabulation and the oracle.

The reference ships no golden vectors for this path, and its Python stack
cannot be imported here, so these known answers are what the oracle is pinned against.
Run: python tests/golden/make_golden.py
"""
import itertools
import json
import os
import sys

import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from firedrake_ts_b200 import fem  # only for the node ordering convention (integer lattice weights)

CASES = [
    (1, 1, [[0], [1]]), (1, 2, [[sp.Rational(1, 3)], [2]]), (1, 3, [[0], [sp.Rational(1, 2)]]),
    (2, 1, [[0, 0], [1, 0], [0, 1]]), (2, 2, [[0, 0], [2, sp.Rational(1, 2)], [sp.Rational(1, 3), 1]]),
    (2, 2, [[1, 1], [0, 0], [2, 0]]),  # negatively oriented
    (2, 3, [[0, 0], [1, 0], [0, 1]]), (2, 3, [[0, 0], [1, sp.Rational(1, 4)], [sp.Rational(1, 2), 2]]),
    (3, 1, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]),
    (3, 2, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]),
    (3, 2, [[0, 0, 0], [1, sp.Rational(1, 5), 0], [sp.Rational(1, 3), 1, sp.Rational(1, 7)], [0, sp.Rational(1, 2), 2]]),
    (3, 3, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]),
    (3, 3, [[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 1, 1]]),  # a Kuhn tetrahedron of the unit cube
]


def integrate_ref(poly, xs):
    """exact integral of a polynomial over the reference simplex."""
    dim = len(xs)
    total = sp.Integer(0)
    for mon, coeff in sp.Poly(sp.expand(poly), *xs).terms():
        num = sp.Integer(1)
        for e in mon:
            num *= sp.factorial(e)
        total += coeff * num / sp.factorial(dim + sum(mon))
    return total


def basis(dim, degree, xs):
    W = fem.lattice_weights(dim, degree)
    nodes = [[sp.Rational(int(w[a + 1]), degree) for a in range(dim)] for w in W]
    mons = [e for e in itertools.product(range(degree + 1), repeat=dim) if sum(e) <= degree]
    V = sp.Matrix([[sp.prod([p[a] ** e[a] for a in range(dim)]) for e in mons] for p in nodes])
    C = V.inv()
    return [sum(C[m, i] * sp.prod([xs[a] ** mons[m][a] for a in range(dim)]) for m in range(len(mons)))
            for i in range(len(nodes))]


def main():
    out = []
    for dim, degree, verts in CASES:
        xs = sp.symbols(f"x0:{dim}")
        phi = basis(dim, degree, xs)
        Vt = sp.Matrix(verts)
        J = sp.Matrix([[Vt[a + 1, k] - Vt[0, k] for a in range(dim)] for k in range(dim)])  # J[k][a]
        det = J.det()
        Jinv = J.inv()  # Jinv[a][k] = d xi_a / d x_k
        adet = abs(det)
        nd = len(phi)
        grads = [[sum(Jinv[a, k] * sp.diff(p, xs[a]) for a in range(dim)) for k in range(dim)] for p in phi]
        M = [[integrate_ref(phi[i] * phi[j], xs) * adet for j in range(nd)] for i in range(nd)]
        K = [[integrate_ref(sum(grads[i][k] * grads[j][k] for k in range(dim)), xs) * adet for j in range(nd)]
             for i in range(nd)]
        b = [integrate_ref(phi[i], xs) * adet for i in range(nd)]
        # advection-type matrix used by the BBM form: C_ij = int (d phi_j / d x_0) phi_i
        Cx = [[integrate_ref(grads[j][0] * phi[i], xs) * adet for j in range(nd)] for i in range(nd)]
        out.append({
            "dim": dim, "degree": degree, "vertices": [[str(sp.nsimplify(v)) for v in row] for row in verts],
            "vertices_float": [[float(v) for v in row] for row in verts],
            "mass": [[sp.N(v, 25).__str__() for v in row] for row in M],
            "stiffness": [[sp.N(v, 25).__str__() for v in row] for row in K],
            "load": [sp.N(v, 25).__str__() for v in b],
            "ddx0": [[sp.N(v, 25).__str__() for v in row] for row in Cx],
        })
        print("case", dim, degree, "done")
    json.dump({"generator": "tests/golden/make_golden.py (sympy %s)" % sp.__version__, "cases": out},
              open(os.path.join(HERE, "element_kats.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
