"""Reference-element tables for Lagrange P1-P3 on simplices (dim 1-3).

These tables are *inputs* to the assembly kernels: node sets,
basis/gradient tabulations at quadrature points, and the constant-coefficient
reference tensors used by the closed-form heat path.  They restate what the
reference obtains implicitly through TSFC/FInAT/FIAT when it first calls
``get_assembler(...).assemble`` (reference call sites
firedrake_ts/solving_utils.py:105-107,258,305).

Conventions:
  * reference simplex: v0 = origin, v_a = unit vector e_a; barycentric
    lambda_0 = 1 - sum(xi), lambda_a = xi_a;
  * local node order = vertices, edges, faces, interior, with UFC/FIAT entity
    numbering: triangle edges (1,2),(0,2),(0,1); tetrahedron edges
    (2,3),(1,3),(1,2),(0,3),(0,2),(0,1); tetrahedron faces
    (1,2,3),(0,2,3),(0,1,3),(0,1,2);
  * nodes on an edge run from the lower local vertex to the higher one; cell
    vertices are sorted by global vertex number, so this is globally consistent;
  * P3 ``variant``: "equispaced" (edge nodes at 1/3, 2/3) or "gll" (edge nodes
    at (1 -/+ 1/sqrt(5))/2, Firedrake's default "spectral" variant, SURVEY.md
    hard part 3).  P1/P2 are variant independent.
"""
from __future__ import annotations

import functools
import itertools
import math

import mpmath
import numpy as np

mpmath.mp.dps = 50

EDGES = {
    1: [(0, 1)],
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}
FACES = {3: [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]}


def num_nodes(dim: int, degree: int, cell: str = "simplex") -> int:
    if cell == "quadrilateral":
        return (degree + 1) ** dim
    return math.comb(degree + dim, dim)


# ---------------------------------------------------------------------------
# quadrilateral Q1 (reference examples/cahn-hilliard.py:14-16: UnitSquareMesh(96, 96, quadrilateral=True),
# "Lagrange" 1).  Firedrake builds Q1 as the tensor product of two interval P1 elements, so the local node
# order is the tensor-product order with the SECOND factor fastest: node l = 2*ix + iy sits at the
# reference vertex (ix, iy) of [0,1]^2, and phi_l = a_ix(xi) a_iy(eta) with a_0(t) = 1 - t, a_1(t) = t.
# The coordinate field is the same Q1 element (isoparametric bilinear map).
# ---------------------------------------------------------------------------
QUAD_VERTICES = [(0, 0), (0, 1), (1, 0), (1, 1)]


def _check_quad(dim, degree):
    if dim != 2 or degree != 1:
        raise NotImplementedError("quadrilateral cells: Q1 in two dimensions only")


def _tabulate_quad(points: np.ndarray) -> np.ndarray:
    points = np.atleast_2d(np.asarray(points, dtype=float))
    out = np.zeros((3, points.shape[0], 4))
    for q, (x, y) in enumerate(points):
        ax, ay = (1.0 - x, x), (1.0 - y, y)
        dx, dy = (-1.0, 1.0), (-1.0, 1.0)
        for l, (ix, iy) in enumerate(QUAD_VERTICES):
            out[0, q, l] = ax[ix] * ay[iy]
            out[1, q, l] = dx[ix] * ay[iy]
            out[2, q, l] = ax[ix] * dy[iy]
    return out


def _quadrature_quad(degree: int):
    """tensor Gauss-Legendre rule on [0,1]^2, exact for degree ``degree`` in each variable."""
    n = max(int(degree), 1) // 2 + 1
    x, w = _gauss_jacobi_01(n, 0)
    pts = np.array([(x[i], x[j]) for i in range(n) for j in range(n)])
    wts = np.array([w[i] * w[j] for i in range(n) for j in range(n)])
    return pts, wts


@functools.lru_cache(maxsize=None)
def lattice_weights(dim: int, degree: int, cell: str = "simplex") -> np.ndarray:
    if cell == "quadrilateral":
        _check_quad(dim, degree)
        return np.eye(4, dtype=np.int64)  # Q1: the nodes are the cell's vertices
    return _lattice_weights_simplex(dim, degree)


@functools.lru_cache(maxsize=None)
def _lattice_weights_simplex(dim: int, degree: int) -> np.ndarray:
    """Integer matrix W (nd x (dim+1)): ``degree`` times the barycentric
    coordinates of the *equispaced* Lagrange nodes in local (FIAT) order.

    ``W @ vertex_lattice_coords`` is the node's position on the ``degree``-times
    refined vertex lattice; the synthetic mesh generator uses it to identify
    shared nodes between cells (mesh.py).  The identification is the same for
    the GLL variant (only the basis functions differ)."""
    k = degree
    nv = dim + 1
    rows = []
    for a in range(nv):
        w = [0] * nv
        w[a] = k
        rows.append(w)
    if dim == 1:
        # the single edge is the cell interior: nodes from v0 to v1
        for m in range(1, k):
            w = [0] * nv
            w[0] = k - m
            w[1] = m
            rows.append(w)
    else:
        for (a, b) in EDGES[dim]:
            for m in range(1, k):
                w = [0] * nv
                w[a] = k - m
                w[b] = m
                rows.append(w)
        if k >= 3:
            if dim == 2:
                assert k == 3
                rows.append([1, 1, 1])
            else:
                assert k == 3
                for (a, b, c) in FACES[3]:
                    w = [0] * nv
                    w[a] = w[b] = w[c] = 1
                    rows.append(w)
    W = np.array(rows, dtype=np.int64)
    assert W.shape == (num_nodes(dim, degree), nv), (W.shape, dim, degree)
    assert (W.sum(axis=1) == k).all()
    return W


def _edge_param(degree: int, variant: str):
    """Parameters in (0,1) of the interior edge nodes, as mpmath numbers."""
    k = degree
    if k <= 2 or variant == "equispaced":
        return [mpmath.mpf(m) / k for m in range(1, k)]
    if variant == "gll":
        assert k == 3
        s = 1 / mpmath.sqrt(5)
        return [(1 - s) / 2, (1 + s) / 2]
    raise ValueError(f"unknown variant {variant!r}")


@functools.lru_cache(maxsize=None)
def _nodes_bary_mp(dim: int, degree: int, variant: str):
    """Barycentric coordinates (mpmath) of the nodes in local order."""
    W = lattice_weights(dim, degree)
    k = degree
    par = _edge_param(degree, variant)
    out = []
    for w in W:
        nz = [a for a in range(dim + 1) if w[a] != 0]
        b = [mpmath.mpf(0)] * (dim + 1)
        if len(nz) == 1:
            b[nz[0]] = mpmath.mpf(1)
        elif len(nz) == 2:
            a0, a1 = nz  # a0 < a1: runs from the lower to the higher local vertex
            m = int(w[a1])  # m-th interior node counted from a0
            t = par[m - 1]
            b[a0] = 1 - t
            b[a1] = t
        else:
            for a in nz:
                b[a] = mpmath.mpf(int(w[a])) / k
        out.append(b)
    return out


def nodes_barycentric(dim: int, degree: int, variant: str = "gll", cell: str = "simplex") -> np.ndarray:
    if cell == "quadrilateral":
        _check_quad(dim, degree)
        return np.eye(4)
    return np.array([[float(x) for x in b] for b in _nodes_bary_mp(dim, degree, variant)])


def _monomials(dim: int, degree: int):
    return [e for e in itertools.product(range(degree + 1), repeat=dim) if sum(e) <= degree]


@functools.lru_cache(maxsize=None)
def _basis_coeffs(dim: int, degree: int, variant: str):
    """Coefficient matrix C (nd x nmono, mpmath) with phi_i = sum_m C[i,m] xi^m."""
    mons = _monomials(dim, degree)
    nodes = _nodes_bary_mp(dim, degree, variant)
    nd = len(nodes)
    assert nd == len(mons)
    V = mpmath.matrix(nd, nd)
    for r, b in enumerate(nodes):
        xi = b[1:]
        for c, e in enumerate(mons):
            v = mpmath.mpf(1)
            for a in range(dim):
                v *= xi[a] ** e[a]
            V[r, c] = v
    Vinv = V ** -1  # columns: coefficients of the dual basis
    return Vinv.T, mons


def tabulate(dim: int, degree: int, points: np.ndarray, variant: str = "gll", cell: str = "simplex") -> np.ndarray:
    """Return B[a, q, i]: a = 0 value of phi_i at point q, a = 1..dim the
    derivative with respect to reference coordinate xi_{a-1}."""
    if cell == "quadrilateral":
        _check_quad(dim, degree)
        return _tabulate_quad(points)
    C, mons = _basis_coeffs(dim, degree, variant)
    points = np.atleast_2d(np.asarray(points, dtype=float))
    nq = points.shape[0]
    nd = len(mons)
    out = np.zeros((dim + 1, nq, nd))
    for q in range(nq):
        xi = [mpmath.mpf(float(x)) for x in points[q]]
        mv = []
        for e in mons:
            row = []
            for a in range(-1, dim):
                v = mpmath.mpf(1)
                for b in range(dim):
                    p = e[b]
                    if b == a:
                        v = v * p * (xi[b] ** (p - 1)) if p > 0 else mpmath.mpf(0)
                    else:
                        v = v * xi[b] ** p
                row.append(v)
            mv.append(row)
        for i in range(nd):
            for a in range(dim + 1):
                s = mpmath.mpf(0)
                for m in range(nd):
                    s += C[i, m] * mv[m][a]
                out[a, q, i] = float(s)
    return out


# ---------------------------------------------------------------------------
# quadrature
# ---------------------------------------------------------------------------
def _gauss_jacobi_01(n: int, alpha: int):
    """n-point Gauss-Jacobi rule on [0,1] for weight (1-x)^alpha."""
    from scipy.special import roots_jacobi

    x, w = roots_jacobi(n, alpha, 0)
    return (x + 1) / 2, w / 2 ** (alpha + 1)


# A few economical rules with positive weights; each is verified for exactness
# in tests/test_fem.py.  (points in reference coordinates, weights sum to the
# reference-simplex volume)
def _sym_tri(groups):
    pts, wts = [], []
    for kind, w, *p in groups:
        if kind == 1:
            pts.append((1 / 3, 1 / 3)); wts.append(w)
        elif kind == 3:
            a = p[0]; b = 1 - 2 * a
            for t in ((a, a), (a, b), (b, a)):
                pts.append(t); wts.append(w)
        elif kind == 6:
            a, b = p; c = 1 - a - b
            for t in itertools.permutations((a, b, c), 3):
                pts.append(t[:2]); wts.append(w)
    return np.array(pts), np.array(wts) / 2.0


def _sym_tet(groups):
    pts, wts = [], []
    for kind, w, *p in groups:
        if kind == 1:
            pts.append((0.25, 0.25, 0.25)); wts.append(w)
        elif kind == 4:
            a = p[0]; b = 1 - 3 * a
            for t in set(itertools.permutations((a, a, a, b), 4)):
                pts.append(t[:3]); wts.append(w)
        elif kind == 6:
            a = p[0]; b = 0.5 - a
            for t in set(itertools.permutations((a, a, b, b), 4)):
                pts.append(t[:3]); wts.append(w)
    return np.array(pts), np.array(wts) / 6.0


_TRI_RULES = {
    1: [(1, 1.0)],
    2: [(3, 1 / 3, 1 / 6)],
    4: [(3, 0.223381589678011, 0.445948490915965), (3, 0.109951743655322, 0.091576213509771)],
    # Radon's 7-point rule (degree 5: vector-P2 Burgers, configuration C2) and the 12- / 16-point rules of
    # degree 6 / 8 (Dunavant 1985) instead of the 9 / 16 / 25 points of the collapsed Gauss rules
    5: [(1, 0.225), (3, 0.125939180544827, 0.101286507323456), (3, 0.132394152788506, 0.470142064105115)],
    6: [(3, 0.116786275726379, 0.249286745170910), (3, 0.050844906370207, 0.063089014491502),
        (6, 0.082851075618374, 0.053145049844817, 0.310352451033784)],
    8: [(1, 0.144315607677787), (3, 0.095091634267285, 0.459292588292723), (3, 0.103217370534718, 0.170569307751760),
        (3, 0.032458497623198, 0.050547228317031), (6, 0.027230314174435, 0.008394777409958, 0.263112829634638)],
}
_TET_RULES = {
    1: [(1, 1.0)],
    2: [(4, 0.25, 0.1381966011250105)],
}


def _grundmann_moeller(dim: int, s: int):
    """Grundmann-Moeller rule of index s on the reference simplex: exact for total degree 2s+1 with
    C(s + dim + 1, dim + 1) points (SIAM J. Numer. Anal. 15 (1978)).  Some weights are negative
    (sum |w| / sum w = 12 for s = 3, 25 for s = 4 in 3-D), harmless at the 1e-12 parity bar."""
    d = dim
    pts, wts = [], []
    for i in range(s + 1):
        m = d + 2 * s - 2 * i + 1
        w = ((-1) ** i) * 2.0 ** (-2 * s) * float(m) ** (2 * s + 1) / (math.factorial(i) * math.factorial(d + 2 * s + 1 - i))
        for beta in itertools.product(range(s - i + 1), repeat=d + 1):
            if sum(beta) == s - i:
                pts.append([(2 * b + 1) / m for b in beta[1:]])
                wts.append(w)  # times d! / d! : weights sum to the reference volume 1/d!
    return np.array(pts), np.array(wts)


@functools.lru_cache(maxsize=None)
def quadrature(dim: int, degree: int, cell: str = "simplex"):
    if cell == "quadrilateral":
        return _quadrature_quad(degree)
    return _quadrature_simplex(dim, degree)


@functools.lru_cache(maxsize=None)
def _quadrature_simplex(dim: int, degree: int):
    """Rule exact for polynomials of total degree <= ``degree`` on the reference
    simplex.  Returns (points[nq, dim], weights[nq]).  On affine cells every
    integrand of the supported forms is polynomial, so any exact rule gives the
    same result up to rounding."""
    degree = max(int(degree), 1)
    if dim == 2:
        for d in sorted(_TRI_RULES):
            if d >= degree:
                return _sym_tri(_TRI_RULES[d])
    if dim == 3:
        for d in sorted(_TET_RULES):
            if d >= degree:
                return _sym_tet(_TET_RULES[d])
    n = degree // 2 + 1
    if dim == 3 and degree >= 3:
        # tetrahedra beyond the 4-point rule: Grundmann-Moeller rules with 5 / 15 / 35 / 70 points (degree 3 / 5 /
        # 7 / 9) instead of the 8 / 27 / 64 / 125 of the collapsed rule below -- the quadrature kernels are
        # compute bound there (P2 heat and Burgers: degree 4-5, P3 heat: 6, P3 BBM / P2 Cahn-Hilliard: 8)
        return _grundmann_moeller(3, degree // 2)
    if dim == 1:
        x, w = _gauss_jacobi_01(n, 0)
        return x[:, None].copy(), w.copy()
    # collapsed (Stroud conical product) Gauss-Jacobi rule
    rules = [_gauss_jacobi_01(n, dim - 1 - a) for a in range(dim)]
    pts, wts = [], []
    for idx in itertools.product(range(n), repeat=dim):
        t = [rules[a][0][idx[a]] for a in range(dim)]
        w = math.prod(rules[a][1][idx[a]] for a in range(dim))
        # xi_0 = t0, xi_1 = t1 (1-t0), xi_2 = t2 (1-t0)(1-t1)
        xi, scale = [], 1.0
        for a in range(dim):
            xi.append(t[a] * scale)
            scale *= 1 - t[a]
        pts.append(xi)
        wts.append(w)
    return np.array(pts), np.array(wts)


# ---------------------------------------------------------------------------
# element tables handed to the engines
# ---------------------------------------------------------------------------
class ElementTables:
    """Everything the oracle and the CUDA engine need about one scalar Lagrange
    element and one quadrature degree.

    B   : (nq, dim+1, nd)  B[q, 0, i] = phi_i(x_q), B[q, 1+a, i] = d phi_i / d xi_a
    w   : (nq,)            quadrature weights (sum = reference volume)
    R   : (dim+1, dim+1, nd, nd)  R[a, b, i, j] = sum_q w_q B[q,a,i] B[q,b,j]
                           (exact reference tensors when the rule is exact for
                           degree 2k; used by the constant-coefficient path)
    M0  : (nd,)            sum_q w_q B[q,0,i]   (integral of phi_i)
    """

    def __init__(self, dim: int, degree: int, qdegree: int, variant: str = "gll", cell: str = "simplex"):
        self.dim, self.degree, self.qdegree, self.variant, self.cell = dim, degree, qdegree, variant, cell
        self.nd = num_nodes(dim, degree, cell)
        pts, w = quadrature(dim, qdegree, cell)
        self.points, self.w = pts, np.ascontiguousarray(w)
        self.nq = len(w)
        T = tabulate(dim, degree, pts, variant, cell)  # (dim+1, nq, nd)
        self.B = np.ascontiguousarray(np.transpose(T, (1, 0, 2)))
        pr, wr = quadrature(dim, max(2 * degree, 1), cell)
        Tr = tabulate(dim, degree, pr, variant, cell)
        self.R = np.ascontiguousarray(np.einsum("q,aqi,bqj->abij", wr, Tr, Tr))
        self.M0 = np.ascontiguousarray(np.einsum("q,qi->i", wr, Tr[0]))
        self.nodes_bary = nodes_barycentric(dim, degree, variant, cell)
        self.W = lattice_weights(dim, degree, cell)


@functools.lru_cache(maxsize=None)
def element_tables(dim: int, degree: int, qdegree: int, variant: str = "gll", cell: str = "simplex") -> ElementTables:
    return ElementTables(dim, degree, qdegree, variant, cell)
