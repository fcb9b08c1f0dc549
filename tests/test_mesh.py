"""synthetic meshes: entity counts of SURVEY.md section 8, numbering bijectivity, partition + halo lists."""
import numpy as np
import pytest
import torch

from firedrake_ts_b200 import mesh as M
from firedrake_ts_b200.mesh import _tiled_index

NNZ = {
    (2, 1): lambda n: 7 * n * n + 6 * n + 1, (2, 2): lambda n: 46 * n * n + 16 * n + 1,
    (2, 3): lambda n: 153 * n * n + 30 * n + 1, (3, 1): lambda n: 15 * n ** 3 + 21 * n * n + 9 * n + 1,
    (3, 2): lambda n: 230 * n ** 3 + 138 * n * n + 24 * n + 1, (3, 3): lambda n: 1311 * n ** 3 + 459 * n * n + 45 * n + 1,
}
MAXROW = {(2, 1): 7, (2, 2): 19, (2, 3): 37, (3, 1): 15, (3, 2): 65, (3, 3): 175}


@pytest.mark.parametrize("dim,degree", list(NNZ))
@pytest.mark.parametrize("tile", [0, 2])
def test_counts_match_survey_table(dim, degree, tile):
    from firedrake_ts_b200 import forms
    from firedrake_ts_b200.function import Function
    from oracle import Oracle

    n = 5 if dim == 2 else 3
    m = M.SimplexMesh((n,) * dim, tile=tile)
    V = M.FunctionSpace(m, degree)
    assert m.num_cells == (2 * n * n if dim == 2 else 6 * n ** 3)
    assert V.num_nodes == (degree * n + 1) ** dim
    assert len(torch.unique(V.cell_node_map)) == V.num_nodes
    rp, ci = Oracle.from_form(forms.heat(Function(V), Function(V))).node_pattern()
    assert rp[-1] == NNZ[(dim, degree)](n)
    assert np.diff(rp).max() == MAXROW[(dim, degree)]
    # sorted, duplicate-free columns (what PETSc AIJ stores)
    for r in range(len(rp) - 1):
        seg = ci[rp[r]:rp[r + 1]]
        assert (np.diff(seg) > 0).all()


def test_headline_formulas():
    n = 184
    assert 6 * n ** 3 == 37_377_024 and (2 * n + 1) ** 3 == 50_243_409 and NNZ[(3, 2)](n) == 1_437_462_465
    n = 96
    assert 6 * n ** 3 == 5_308_416 and (3 * n + 1) ** 3 == 24_137_569 and NNZ[(3, 3)](n) == 1_164_123_361


@pytest.mark.parametrize("dims", [(7,), (9, 5), (6, 7, 5)])
@pytest.mark.parametrize("T,o", [(3, 0), (3, 1), (5, 1), (4, 2), (5, 4), (6, 0)])
def test_tiled_index_bijective(dims, T, o):
    g = torch.meshgrid(*[torch.arange(d) for d in dims], indexing="ij")
    p = torch.stack([x.reshape(-1) for x in g], -1)
    idx = _tiled_index(p, list(dims), T, o)
    assert sorted(idx.tolist()) == list(range(int(np.prod(dims))))


def test_row_blocks_are_tiles():
    m = M.SimplexMesh((4, 4, 4), tile=2)
    V = M.FunctionSpace(m, 2, tile=6, tile_offset=0)
    st = V.row_block_starts()
    for b in range(len(st) - 1):
        sub = V.node_lattice[st[b]:st[b + 1]]
        ext = sub.max(0).values - sub.min(0).values + 1
        assert int(ext.prod()) == int(st[b + 1] - st[b])


def test_periodic_interval():
    m = M.PeriodicIntervalMesh(10, 100.0)
    V = M.FunctionSpace(m, 1)
    assert V.num_nodes == 10 and int(V.cell_node_map[-1, 1]) == 0
    assert V.boundary_nodes.numel() == 0
    assert abs(float(m.coords[-1, 0]) - 100.0) < 1e-12


@pytest.mark.parametrize("dim,grid,n,k,tile", [(2, (2, 2), 5, 2, 0), (3, (2, 2, 2), 3, 2, 2), (3, (2, 1, 1), 4, 3, 0),
                                               (2, (2, 1), 4, 1, 3), (1, (2,), 9, 2, 0)])
def test_partition_and_halo_lists(dim, grid, n, k, tile):
    P = int(np.prod(grid))
    Vs = [M.FunctionSpace(M.SimplexMesh((n,) * dim, partition=M.Partition(grid, r), tile=tile), k) for r in range(P)]
    tot = sum(V.num_owned_nodes for V in Vs)
    assert tot == (k * n + 1) ** dim
    allg = torch.cat([V.global_node_ids[:V.num_owned_nodes] for V in Vs])
    assert len(torch.unique(allg)) == tot                       # every node owned exactly once
    for r in range(P):
        for s, (st, cnt) in Vs[r].recv_slices.items():
            sent = Vs[s].send_lists[r]
            assert len(sent) == cnt
            assert torch.equal(Vs[r].global_node_ids[st:st + cnt], Vs[s].global_node_ids[sent.long()])
        assert sum(c for _, c in Vs[r].recv_slices.values()) == Vs[r].num_ghost_nodes
        ni = Vs[r].mesh.num_interior_cells
        assert (Vs[r].cell_node_map[:ni] < Vs[r].num_owned_nodes).all()   # interior cells touch no ghost
    # every global cell that touches an owned node of rank r is local to r (overlap-1 exec halo)
    total_cells = sum(V.mesh.num_cells for V in Vs)
    assert total_cells >= Vs[0].mesh.global_num_cells
