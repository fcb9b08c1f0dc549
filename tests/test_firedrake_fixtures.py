"""Parity against dumps of the real reference stack (tests/golden/firedrake/*.npz, written by
tests/golden/make_firedrake_fixtures.py inside a Firedrake environment).  No such dump can be produced in
the build image (Firedrake/PETSc are not installable), so without fixtures these
tests are skipped and parity stays "unpinned with respect to the reference itself" (DESIGN.md section 5)."""
import glob
import os

import numpy as np
import pytest

from firedrake_ts_b200 import fem, forms

FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "firedrake", "*.npz")))
FAMILY = {"heat": forms.HEAT, "bbm": forms.BBM, "burgers": forms.BURGERS, "cahn_hilliard": forms.CAHN_HILLIARD}


def test_fixture_generator_is_committed():
    here = os.path.dirname(__file__)
    src = open(os.path.join(here, "golden", "make_firedrake_fixtures.py")).read()
    for needle in ("_assemble_residual(tensor=ctx._F)", "_assemble_jac(ctx._jac)", "getValuesCSR", "cell_node_map"):
        assert needle in src


@pytest.mark.skipif(not FIX, reason="no Firedrake dumps (the reference stack is not installable in this image)")
@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p) for p in FIX])
def test_oracle_matches_firedrake_dump(path):
    from oracle import Oracle

    d = np.load(path, allow_pickle=False)
    if "quadrilateral" in str(d["cell"]):
        pytest.skip("quadrilateral dump: compared by the tensor-product element tests")
    dim, k = int(d["dim"]), int(d["degree"])
    fam = FAMILY[str(d["family"])]
    et = fem.element_tables(dim, k, forms.quadrature_degree(fam, k), "gll")
    nn = int(d["cell_node"].max()) + 1
    o = Oracle(dim=dim, nd=et.nd, ncomp=int(d["value_size"]), layout=str(d["layout"]), family=fam, params=d["params"],
               coords=d["coords"], cell_coord=d["cell_coord"], cell_node=d["cell_node"], B=et.B, w=et.w,
               num_nodes=nn, num_owned_nodes=nn, bc_nodes=d["bc_nodes"])
    rp, ci = o.pattern()
    assert np.array_equal(rp, d["rowptr"]) and np.array_equal(ci, d["colidx"])  # bit-exact index maps
    F = o.ifunction(0.0, d["u"], d["udot"])
    J = o.ijacobian(0.0, d["u"], d["udot"], float(d["shift"]))
    assert np.abs(F - d["F"]).max() < 1e-12 * np.abs(d["F"]).max()
    assert np.abs(J - d["values"]).max() < 1e-12 * np.abs(d["values"]).max()
