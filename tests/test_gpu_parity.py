"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerance: 1e-12 relative to the largest entry (BASELINE.json north_star: "residual and
matrix entries within 1e-12 relative (fp64, scatter-order differences only)"); index maps
bit-exact."""
import numpy as np
import pytest
import torch

from helpers import make_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [
    ("heat", 1, 1, 10), ("heat", 1, 2, 7), ("heat", 1, 3, 5),
    ("heat", 2, 1, 9), ("heat", 2, 2, 7), ("heat", 2, 3, 5),
    ("heat", 3, 1, 5), ("heat", 3, 2, 4), ("heat", 3, 3, 3),
    ("bbm", 1, 1, 12), ("bbm", 2, 2, 5), ("bbm", 3, 1, 4), ("bbm", 3, 2, 3), ("bbm", 3, 3, 2),
    ("burgers", 2, 1, 6), ("burgers", 2, 2, 6), ("burgers", 2, 3, 3), ("burgers", 3, 1, 3), ("burgers", 3, 2, 2),
    ("cahn_hilliard", 2, 1, 8), ("cahn_hilliard", 2, 2, 4), ("cahn_hilliard", 3, 1, 3), ("cahn_hilliard", 3, 2, 2),
]


@pytest.mark.parametrize("family,dim,degree,n", CASES)
@pytest.mark.parametrize("tile", [0, 2])
def test_engine_matches_oracle(family, dim, degree, n, tile):
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    form, bcs, u, udot = make_problem(family, dim, degree, n, tile=tile)
    orc = Oracle.from_form(form, bcs)
    eng = Engine(form, bcs, device="cuda:0")
    # sparsity: bit exact
    rp, ci = orc.pattern()
    grp, gci = eng.csr()
    assert np.array_equal(grp.cpu().numpy(), rp)
    assert np.array_equal(gci.cpu().numpy(), ci)
    x, xd = u.data.cuda(), udot.data.cuda()
    shift = 10.0
    F = eng.ifunction(0.0, x, xd).cpu().numpy()
    Fo = orc.ifunction(0.0, u.data, udot.data)
    assert rel_err(F, Fo) < TOL, rel_err(F, Fo)
    Jv = eng.ijacobian(0.0, x, xd, shift).cpu().numpy()
    Jo = orc.ijacobian(0.0, u.data, udot.data, shift)
    assert rel_err(Jv, Jo) < TOL, rel_err(Jv, Jo)
    assert eng.launch_count > 0
