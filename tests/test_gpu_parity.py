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
    ("bbm", 1, 1, 12), ("bbm", 2, 2, 5), ("bbm", 2, 3, 4), ("bbm", 3, 1, 4), ("bbm", 3, 2, 3), ("bbm", 3, 3, 2),
    ("burgers", 2, 1, 6), ("burgers", 2, 2, 6), ("burgers", 2, 3, 3), ("burgers", 3, 1, 3), ("burgers", 3, 2, 2),
    ("diffusion", 1, 1, 10), ("diffusion", 1, 3, 5), ("diffusion", 2, 2, 7), ("diffusion", 3, 1, 5),
    ("diffusion", 3, 2, 4), ("diffusion", 3, 3, 3),
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
    paths = [0, 1, 2] if family in ("heat", "diffusion") else [0]
    # sparsity: bit exact
    rp, ci = orc.pattern()
    grp, gci = eng.csr()
    assert np.array_equal(grp.cpu().numpy(), rp)
    assert np.array_equal(gci.cpu().numpy(), ci)
    x, xd = u.data.cuda(), udot.data.cuda()
    shift = 10.0
    Fo = orc.ifunction(0.0, u.data, udot.data)
    Jo = orc.ijacobian(0.0, u.data, udot.data, shift)
    # 0: generic quadrature, 1: closed form + atomics, 2: closed form + owner-computes gather (plan 1 / plan 2)
    for path, plan in [(p, 0) for p in paths if p < 2] + ([(2, 1), (2, 2)] if 2 in paths else []):
        if path == 2:
            starts = form.space.row_block_starts()
            if starts is None:
                starts = torch.unique(torch.arange(0, eng.num_rows + 50, 50).clamp(max=eng.num_rows))
            eng.set_option("plan_version", plan)
            eng.set_row_blocks(starts)
            assert eng.get_option("plan_version") == plan
            eng.set_option("gather_residual", 1)
        eng.set_option("scatter", path)
        F = eng.ifunction(0.0, x, xd).cpu().numpy()
        assert rel_err(F, Fo) < TOL, (path, plan, rel_err(F, Fo))
        Jv = eng.ijacobian(0.0, x, xd, shift).cpu().numpy()
        assert rel_err(Jv, Jo) < TOL, (path, plan, rel_err(Jv, Jo))
    assert eng.launch_count > 0


@pytest.mark.parametrize("plan", [1, 2])
@pytest.mark.parametrize("dim,degree,n,ftile,foffset", [(3, 2, 6, 5, 1), (3, 2, 5, 6, 0), (2, 2, 9, 5, 1), (3, 1, 7, 3, 1),
                                                        (3, 3, 3, 4, 1), (2, 3, 5, 7, 2), (1, 2, 20, 5, 1),
                                                        (3, 2, 8, (8, 6, 4), 0), (3, 2, 7, (4, 6, 8), 3),
                                                        (2, 2, 12, (8, 4), 0)])
def test_gather_path_tiled_numbering(dim, degree, n, ftile, foffset, plan):
    """owner-computes gather (write-once, deterministic) on tile-numbered spaces with offset and
    anisotropic tiles, both plan versions; also checks run-to-run bitwise reproducibility."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    form, bcs, u, udot = make_problem("heat", dim, degree, n, tile=2, ftile=ftile, foffset=foffset)
    orc = Oracle.from_form(form, bcs)
    eng = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=plan)
    eng.set_option("gather_residual", 1)
    assert eng.get_option("scatter") == 2 and eng.get_option("plan_blocks") > 0
    assert eng.get_option("plan_version") == plan
    x, xd = u.data.cuda(), udot.data.cuda()
    J1 = eng.ijacobian(0.0, x, xd, 7.5)
    J2 = eng.ijacobian(0.0, x, xd, 7.5)
    assert torch.equal(J1, J2)
    F1 = eng.ifunction(0.0, x, xd)
    F2 = eng.ifunction(0.0, x, xd)
    if plan == 1:  # plan 2 has no residual lists: the residual stays on the atomic path (order not fixed)
        assert torch.equal(F1, F2)
    assert rel_err(J1.cpu().numpy(), orc.ijacobian(0.0, u.data, udot.data, 7.5)) < TOL
    assert rel_err(F1.cpu().numpy(), orc.ifunction(0.0, u.data, udot.data)) < TOL


@pytest.mark.parametrize("n,ftile", [(12, (8, 6, 4)), (13, 6)])
def test_pattern_dictionary(n, ftile):
    """plan 2: blocks with identical index lists share one copy; sharing must not change any value
    (bitwise) and must actually find the repeated interior blocks of a structured mesh."""
    from firedrake_ts_b200.engine import Engine

    mt = tuple(t // 2 for t in ftile) if isinstance(ftile, tuple) else ftile // 2
    form, bcs, u, udot = make_problem("heat", 3, 2, n, tile=mt, ftile=ftile)
    x, xd = u.data.cuda(), udot.data.cuda()
    e1 = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=1, options={"optimize": 0, "sector": 4})
    e0 = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=0, options={"sector": 4, "optimize": 0})
    nb = e1.get_option("plan_blocks")
    assert e0.get_option("plan_patterns") == nb
    assert e1.get_option("plan_patterns") < nb
    assert e1.get_option("plan_shared_contributions") < e1.get_option("plan_contributions")
    J0 = e0.ijacobian(0.0, x, xd, 10.0)
    assert torch.equal(e1.ijacobian(0.0, x, xd, 10.0), J0)
    # bank-conflict-aware column order, 16/8-byte store groups, CTA sizes and the warp-specialised
    # (overlapped) kernel only change the summation order / the schedule, never a value beyond rounding
    for opts in ({"optimize": 1}, {"optimize": 1, "sector": 2}, {"optimize": 0, "sector": 2}, {"sector": 1},
                 {"threads": 1024}, {"overlap": 1}, {"overlap": 1, "sector": 2}):
        e2 = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=1, options=opts)
        J2 = e2.ijacobian(0.0, x, xd, 10.0)
        assert rel_err(J2.cpu().numpy(), J0.cpu().numpy()) < TOL
        assert torch.equal(J2, e2.ijacobian(0.0, x, xd, 10.0))


@pytest.mark.parametrize("dim,degree,n,ftile", [(3, 2, 8, (6, 4, 4)), (3, 2, 6, 4), (2, 2, 12, (8, 4)), (3, 1, 9, 4),
                                                (2, 3, 6, 4), (1, 2, 30, 6)])
def test_overlapped_kernel_bitwise_equal(dim, degree, n, ftile):
    """warp-specialised Jacobian kernel (producer / consumer / copy warps, mbarrier hand-over) against the
    two-phase kernel on the same plan: identical summation order => identical bits; plus the oracle."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    form, bcs, u, udot = make_problem("heat", dim, degree, n, tile=2, ftile=ftile)
    x, xd = u.data.cuda(), udot.data.cuda()
    # cache_metric 0: both kernels form the geometry inline (the cached metrics come out of another kernel)
    e0 = Engine(form, bcs, device="cuda:0", scatter=2, options={"overlap": 0, "cache_metric": 0})
    e1 = Engine(form, bcs, device="cuda:0", scatter=2, options={"overlap": 1})
    J0 = e0.ijacobian(0.0, x, xd, 3.0)
    assert e0.get_option("plan_metric_cached") == 0
    for _ in range(3):  # repeated launches: the hand-over barriers must not depend on timing
        assert torch.equal(e1.ijacobian(0.0, x, xd, 3.0), J0)
    assert rel_err(J0.cpu().numpy(), Oracle.from_form(form, bcs).ijacobian(0.0, u.data, udot.data, 3.0)) < TOL


@pytest.mark.parametrize("dim,degree,n,tile", [(3, 2, 7, 3), (3, 2, 6, 0), (2, 2, 12, 4), (3, 1, 9, 3), (2, 3, 6, 2),
                                               (1, 2, 30, 6), (3, 3, 3, 1)])
def test_residual_tile_kernel(dim, degree, n, tile):
    """residual tile kernel (coalesced staging of node values, per-tile pre-reduction, one atomic per
    (tile, node)) against the oracle and against the per-cell kernel, incl. tile-aligned cell ranges."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    form, bcs, u, udot = make_problem("heat", dim, degree, n, tile=tile, ftile=2 * tile if tile else None)
    x, xd = u.data.cuda(), udot.data.cuda()
    eng = Engine(form, bcs, device="cuda:0", scatter=1)
    V = form.space
    starts = V.mesh.cell_tile_starts(max_pairs=4096 // V.nd)
    eng.set_cell_tiles(starts)
    assert eng.get_option("residual_tiles") == starts.numel() - 1
    Fo = Oracle.from_form(form, bcs).ifunction(0.0, u.data, udot.data)
    F1 = eng.ifunction(0.0, x, xd)
    assert rel_err(F1.cpu().numpy(), Fo) < TOL
    eng.set_option("residual_tiles", 2)  # variant: cells gather their inputs directly, tile pre-reduction only
    assert rel_err(eng.ifunction(0.0, x, xd).cpu().numpy(), Fo) < TOL
    eng.set_option("residual_tiles", 3)  # the same with the tile's reduction tables prefetched into shared memory
    F3 = eng.ifunction(0.0, x, xd).cpu().numpy()
    assert rel_err(F3, Fo) < TOL
    eng.set_option("residual_tiles", 0)
    F0 = eng.ifunction(0.0, x, xd)
    eng.set_option("residual_tiles", 1)
    assert rel_err(F1.cpu().numpy(), F0.cpu().numpy()) < TOL
    # two tile-aligned ranges add up to the full residual
    mid = int(starts[(starts.numel() - 1) // 2])
    F2 = torch.empty_like(F1)
    eng.ifunction_range(0.0, x, xd, F2, 0, mid, 1)
    eng.ifunction_range(0.0, x, xd, F2, mid, eng.num_cells, 0)
    assert rel_err(F2.cpu().numpy(), Fo) < TOL


@pytest.mark.parametrize("with_bc", [False, True])
@pytest.mark.parametrize("dim,degree,n", [(3, 3, 3), (3, 2, 4), (2, 3, 6), (3, 3, 2)])
def test_bbm_dmma_jacobian(dim, degree, n, with_bc):
    """fp64 tensor-core (DMMA) BBM Jacobians against the oracle and against the generic quadrature kernel:
    option dmma = 2 (default) contracts the reference tensor with the cell's coefficient vector (one small
    GEMM per 16 cells), dmma = 1 is the quadrature formulation, dmma = 0 the generic kernel.  State-dependent,
    with and without Dirichlet rows / columns; cell counts that do not fill the last group of 16."""
    from firedrake_ts_b200.engine import Engine
    from firedrake_ts_b200.function import DirichletBC
    from oracle import Oracle

    form, bcs, u, udot = make_problem("bbm", dim, degree, n, tile=2)
    if with_bc:
        bcs = [DirichletBC(form.space, 0.0, "on_boundary")]
    x, xd = u.data.cuda(), udot.data.cuda()
    eng = Engine(form, bcs, device="cuda:0")
    assert eng.get_option("dmma") == 2
    Jo = Oracle.from_form(form, bcs).ijacobian(0.0, u.data, udot.data, 12.5)
    J2 = eng.ijacobian(0.0, x, xd, 12.5)
    assert rel_err(J2.cpu().numpy(), Jo) < TOL
    eng.set_option("dmma", 1)
    J1 = eng.ijacobian(0.0, x, xd, 12.5)
    assert rel_err(J1.cpu().numpy(), Jo) < TOL
    eng.set_option("dmma", 0)
    J0 = eng.ijacobian(0.0, x, xd, 12.5)
    assert rel_err(J2.cpu().numpy(), J0.cpu().numpy()) < TOL
    # cell ranges (the halo split of the scaling path) add up
    eng.set_option("dmma", 2)
    mid = (eng.num_cells // 3) | 1
    Js = torch.empty_like(J2)
    eng.ijacobian_range(0.0, x, xd, 12.5, Js, 0, mid, 1)
    eng.ijacobian_range(0.0, x, xd, 12.5, Js, mid, eng.num_cells, 2)
    assert rel_err(Js.cpu().numpy(), Jo) < TOL


def test_pattern_dictionary_compaction():
    """with enough repeated blocks only the tables of the unique patterns are kept (the duplicates are
    freed after hashing); values must not change."""
    from firedrake_ts_b200.engine import Engine

    form, bcs, u, udot = make_problem("heat", 3, 2, 24, tile=3, ftile=6, device="cuda:0")
    x, xd = u.data, udot.data
    e1 = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=1, options={"optimize": 0})
    e0 = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=0, options={"sector": 4, "optimize": 0})
    # store-group width left to the engine: 32-byte groups while the dictionary pays, 16-byte groups for a
    # mesh without repeated blocks (its index tables stream from HBM; less SELL padding)
    ea = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=0, options={"optimize": 0})
    eb = Engine(form, bcs, device="cuda:0", scatter=2, plan_version=2, dedup=0)  # default: GPU lane orders, 32-byte groups
    assert e1.get_option("plan_sector") == 4 and e0.get_option("plan_sector") == 4 and ea.get_option("plan_sector") == 2
    assert eb.get_option("plan_sector") == 4 and eb.get_option("plan_optimized") == 2
    assert rel_err(eb.ijacobian(0.0, x, xd, 10.0).cpu().numpy(), e0.ijacobian(0.0, x, xd, 10.0).cpu().numpy()) < TOL
    assert rel_err(ea.ijacobian(0.0, x, xd, 10.0).cpu().numpy(), e0.ijacobian(0.0, x, xd, 10.0).cpu().numpy()) < TOL
    assert e1.get_option("plan_compacted") == 1 and e0.get_option("plan_compacted") == 0
    assert e1.get_option("plan_patterns") * 4 <= e1.get_option("plan_blocks")
    assert torch.equal(e1.ijacobian(0.0, x, xd, 10.0), e0.ijacobian(0.0, x, xd, 10.0))


@pytest.mark.parametrize("dim,degree,n", [(1, 1, 10), (2, 2, 6), (3, 2, 4)])
def test_imex_split_forms(dim, degree, n):
    """the two sides of examples/heat-explicit.py:13-14 on the engine: F = (u_t, v) and
    G = -((grad u, grad v) - (1, v)) against the oracle, and F - G against the closed-form heat
    kernels (a different kernel family, so this is not a tautology)."""
    from firedrake_ts_b200 import forms
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    heat, bcs, u, udot = make_problem("heat", dim, degree, n, tile=2)
    F, G = forms.heat_explicit(u, udot)
    x, xd = u.data.cuda(), udot.data.cuda()
    out = {}
    for name, form in (("F", F), ("G", G), ("heat", heat)):
        orc, eng = Oracle.from_form(form, bcs), Engine(form, bcs, device="cuda:0")
        shift = 0.0 if name == "G" else 4.0
        r = eng.ifunction(0.0, x, xd).cpu().numpy()
        Jv = eng.ijacobian(0.0, x, xd, shift).cpu().numpy()
        assert rel_err(r, orc.ifunction(0.0, u.data, udot.data)) < TOL
        assert rel_err(Jv, orc.ijacobian(0.0, u.data, udot.data, shift)) < TOL
        out[name] = (r, Jv)
    assert rel_err(out["F"][0] - out["G"][0], out["heat"][0]) < 1e-11
    isbc_diag = np.abs(out["F"][1] - out["G"][1] - out["heat"][1])  # Dirichlet rows: 1 - 1 vs 1
    assert np.sort(isbc_diag)[-1] <= 1.0 + 1e-12 and np.count_nonzero(isbc_diag > 1e-10) == sum(
        int(b.nodes.numel()) for b in bcs)


@pytest.mark.parametrize("family", ["heat", "diffusion", "cahn_hilliard"])
@pytest.mark.parametrize("warp", [0.0, 0.03])
@pytest.mark.parametrize("tile", [0, 2])
def test_quadrilateral_engine_matches_oracle(family, warp, tile):
    """Q1 (x Q1) on quadrilaterals, affine and non-affine cells (reference examples/cahn-hilliard.py:14-16):
    pattern bit-exact, values 1e-12."""
    from firedrake_ts_b200.engine import Engine, EngineError
    from oracle import Oracle

    form, bcs, u, udot = make_problem(family, 2, 1, 9, tile=tile, quadrilateral=True, warp=warp)
    orc = Oracle.from_form(form, bcs)
    eng = Engine(form, bcs, device="cuda:0")
    rp, ci = orc.pattern()
    grp, gci = eng.csr()
    assert np.array_equal(grp.cpu().numpy(), rp)
    assert np.array_equal(gci.cpu().numpy(), ci)
    x, xd = u.data.cuda(), udot.data.cuda()
    F = eng.ifunction(0.0, x, xd).cpu().numpy()
    assert rel_err(F, orc.ifunction(0.0, u.data, udot.data)) < TOL
    Jv = eng.ijacobian(0.0, x, xd, 10.0).cpu().numpy()
    assert rel_err(Jv, orc.ijacobian(0.0, u.data, udot.data, 10.0)) < TOL
    # the simplex-only fast paths refuse loudly
    with pytest.raises(EngineError):
        eng.set_option("scatter", 2)
    with pytest.raises(EngineError):
        eng.set_row_blocks(torch.tensor([0, eng.num_rows]))


@pytest.mark.parametrize("dim,degree,n,ftile", [(3, 2, 8, (6, 4, 4)), (3, 2, 7, 6), (2, 2, 12, (8, 4)), (3, 1, 9, 4),
                                                (2, 3, 6, 4), (1, 2, 30, 6), (2, 1, 20, 14)])
def test_cached_metric_gather(dim, degree, n, ftile):
    """plan 2 with the per-block cell metrics precomputed (default) against the kernel that forms the geometry
    from the vertex table, against the oracle, run-to-run bitwise, and across a change of shift."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    mt = tuple(max(t // degree, 1) for t in ftile) if isinstance(ftile, tuple) else max(ftile // degree, 1)
    form, bcs, u, udot = make_problem("heat", dim, degree, n, tile=mt, ftile=ftile)
    x, xd = u.data.cuda(), udot.data.cuda()
    orc = Oracle.from_form(form, bcs)
    e1 = Engine(form, bcs, device="cuda:0", scatter=2)
    e0 = Engine(form, bcs, device="cuda:0", scatter=2, options={"cache_metric": 0})
    for shift in (10.0, 0.0, 1234.5):
        J1 = e1.ijacobian(0.0, x, xd, shift)
        assert e1.get_option("plan_metric_cached") == 1 and e0.get_option("plan_metric_cached") == 0
        assert torch.equal(J1, e1.ijacobian(0.0, x, xd, shift))
        assert rel_err(J1.cpu().numpy(), e0.ijacobian(0.0, x, xd, shift).cpu().numpy()) < TOL
        assert rel_err(J1.cpu().numpy(), orc.ijacobian(0.0, u.data, udot.data, shift)) < TOL


def test_host_jacobian_pipelined_readback():
    """fdts_ijacobian_host on the gather path assembles the row blocks in 8 ranges and reads every range back
    while the next one is assembled: bitwise equal to the device call, for repeated calls and changing shift."""
    from firedrake_ts_b200.engine import Engine

    form, bcs, u, udot = make_problem("heat", 3, 2, 24, tile=3, ftile=6, device="cuda:0")
    eng = Engine(form, bcs, device="cuda:0", scatter=2, options={"pipelined_readback": 1})
    assert eng.get_option("plan_blocks") >= 512
    x, xd = u.data, udot.data
    xh, xdh = x.cpu().pin_memory(), xd.cpu().pin_memory()
    Jh = torch.empty(eng.nnz, dtype=torch.float64).pin_memory()
    for shift in (10.0, 0.5, 10.0):
        Jh.fill_(float("nan"))
        eng.ijacobian_host(0.0, xh, xdh, shift, Jh)
        assert torch.equal(Jh, eng.ijacobian(0.0, x, xd, shift).cpu())
    moved = eng.get_option("d2h_bytes")
    assert moved == 3 * 8 * eng.nnz and eng.get_option("h2d_bytes") == 0


@pytest.mark.parametrize("dim,degree,n,ftile,sector", [(3, 2, 8, (6, 4, 4), 4), (3, 2, 7, 6, 2), (2, 2, 12, (8, 4), 4),
                                                       (3, 1, 9, 4, 1), (2, 3, 6, 4, 2), (3, 3, 3, 3, 4)])
def test_gpu_lane_order_optimiser(dim, degree, n, ftile, sector):
    """plans without a dictionary get their lane orders refined on the GPU (one warp per slice): only the
    summation order changes -- values against the unoptimised plan and the oracle at 1e-12, run-to-run bitwise."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    mt = tuple(max(t // degree, 1) for t in ftile) if isinstance(ftile, tuple) else max(ftile // degree, 1)
    form, bcs, u, udot = make_problem("heat", dim, degree, n, tile=mt, ftile=ftile)
    x, xd = u.data.cuda(), udot.data.cuda()
    e1 = Engine(form, bcs, device="cuda:0", scatter=2, dedup=0, options={"sector": sector, "optimize": 1})
    e0 = Engine(form, bcs, device="cuda:0", scatter=2, dedup=0, options={"sector": sector, "optimize": 0})
    assert e1.get_option("plan_optimized") == 2 and e0.get_option("plan_optimized") == 0
    J1 = e1.ijacobian(0.0, x, xd, 10.0)
    assert torch.equal(J1, e1.ijacobian(0.0, x, xd, 10.0))
    assert rel_err(J1.cpu().numpy(), e0.ijacobian(0.0, x, xd, 10.0).cpu().numpy()) < TOL
    assert rel_err(J1.cpu().numpy(), Oracle.from_form(form, bcs).ijacobian(0.0, u.data, udot.data, 10.0)) < TOL


@pytest.mark.parametrize("dim,degree,n,block_rows", [(3, 2, 7, 120), (3, 2, 6, 60), (2, 2, 14, 90), (3, 1, 10, 40), (2, 1, 24, 150)])
def test_gather_on_morton_numbering(dim, degree, n, block_rows):
    """a numbering without any tile structure (Morton order: compact but irregular row blocks, hardly any
    repeated pattern): gather path with uniform row blocks against the oracle; the dictionary finds (almost)
    nothing, so the tables are per block and their lane orders come from the GPU optimiser when there are many."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    form, bcs, u, udot = make_problem("heat", dim, degree, n, numbering="morton", block_rows=block_rows)
    orc = Oracle.from_form(form, bcs)
    eng = Engine(form, bcs, device="cuda:0", scatter=2)
    rp, ci = orc.pattern()
    grp, gci = eng.csr()
    assert np.array_equal(grp.cpu().numpy(), rp) and np.array_equal(gci.cpu().numpy(), ci)
    x, xd = u.data.cuda(), udot.data.cuda()
    J = eng.ijacobian(0.0, x, xd, 10.0)
    assert torch.equal(J, eng.ijacobian(0.0, x, xd, 10.0))
    assert rel_err(J.cpu().numpy(), orc.ijacobian(0.0, u.data, udot.data, 10.0)) < TOL
    assert rel_err(eng.ifunction(0.0, x, xd).cpu().numpy(), orc.ifunction(0.0, u.data, udot.data)) < TOL


def test_untiled_numbering_shrinks_row_blocks():
    """a numbering the engine knows nothing about (row_block_starts() is None, as for a live Firedrake space): uniform
    chunks of rows are tried from 160 downwards until the plan fits; values against the oracle."""
    from firedrake_ts_b200.engine import Engine
    from oracle import Oracle

    form, bcs, u, udot = make_problem("heat", 3, 2, 10, numbering="morton")
    form.space.row_block_starts = lambda: None
    eng = Engine(form, bcs, device="cuda:0", scatter=2)
    nb = eng.get_option("plan_blocks")
    assert nb >= form.space.num_owned_nodes // 160
    x, xd = u.data.cuda(), udot.data.cuda()
    J = eng.ijacobian(0.0, x, xd, 10.0)
    assert rel_err(J.cpu().numpy(), Oracle.from_form(form, bcs).ijacobian(0.0, u.data, udot.data, 10.0)) < TOL
    # lexicographic numbering of a 3-D mesh: 160 consecutive rows form a pencil that touches far too many cells
    form, bcs, u, udot = make_problem("heat", 3, 2, 24, tile=0)
    eng = Engine(form, bcs, device="cuda:0", scatter=2)
    assert eng.get_option("plan_blocks") > form.space.num_owned_nodes // 160
    J = eng.ijacobian(0.0, u.data.cuda(), udot.data.cuda(), 10.0)
    assert rel_err(J.cpu().numpy(), Oracle.from_form(form, bcs).ijacobian(0.0, u.data, udot.data, 10.0)) < TOL
