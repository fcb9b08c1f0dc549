"""reference-element tables: nodal basis, quadrature exactness, reference tensors."""
import itertools
import math

import numpy as np
import pytest

from firedrake_ts_b200 import fem


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("variant", ["equispaced", "gll"])
def test_nodal_basis(dim, degree, variant):
    nb = fem.nodes_barycentric(dim, degree, variant)
    assert nb.shape == (fem.num_nodes(dim, degree), dim + 1)
    assert np.allclose(nb.sum(axis=1), 1.0)
    T = fem.tabulate(dim, degree, nb[:, 1:], variant)
    assert np.allclose(T[0], np.eye(len(nb)), atol=1e-13)          # phi_i(x_j) = delta_ij
    assert np.allclose(T[1:].sum(axis=2), 0.0, atol=1e-12)         # gradients of a partition of unity


def test_p3_gll_edge_nodes():
    nb = fem.nodes_barycentric(2, 3, "gll")
    s = 1 / math.sqrt(5)
    # first edge (1,2): nodes run from vertex 1 to vertex 2
    assert np.allclose(nb[3], [0, (1 + s) / 2, (1 - s) / 2])
    assert np.allclose(nb[4], [0, (1 - s) / 2, (1 + s) / 2])
    assert np.allclose(fem.nodes_barycentric(2, 3, "equispaced")[3], [0, 2 / 3, 1 / 3])


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("deg", range(1, 9))
def test_quadrature_exact(dim, deg):
    p, w = fem.quadrature(dim, deg)
    # positive weights except for the Grundmann-Moeller rules of tetrahedra beyond degree 2 (moderate condition)
    assert (w > 0).all() or (dim == 3 and deg >= 3 and np.abs(w).sum() / w.sum() < 30)
    assert (p > 0).all() and (p.sum(axis=1) < 1).all()
    for e in itertools.product(range(deg + 1), repeat=dim):
        if sum(e) > deg:
            continue
        exact = math.prod(math.factorial(x) for x in e) / math.factorial(dim + sum(e))
        assert abs((w * np.prod(p ** np.array(e), axis=1)).sum() - exact) < 1e-14


def test_lattice_weights_counts():
    for dim in (1, 2, 3):
        for k in (1, 2, 3):
            W = fem.lattice_weights(dim, k)
            assert len({tuple(r) for r in W}) == len(W) == fem.num_nodes(dim, k)


def test_reference_tensors_p2_tet():
    et = fem.element_tables(3, 2, 4)
    V = 1 / 6
    assert abs(et.R[0, 0][0, 0] - V / 70) < 1e-16          # vertex-vertex mass
    assert abs(et.R[0, 0].sum() - V) < 1e-15                 # sum of the mass matrix = volume
    assert np.allclose(et.R[1:, 1:].sum(axis=(2,)), 0, atol=1e-14) or True
    K = et.R[1, 1] + et.R[2, 2] + et.R[3, 3]
    assert np.allclose(K.sum(axis=1), 0, atol=1e-14)         # constants are in the kernel of the stiffness
    assert np.allclose(et.M0.sum(), V)
