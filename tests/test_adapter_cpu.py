"""Host logic of the Firedrake adapter that needs no GPU: neighbour lists from a PetscSF graph against the
synthetic partitioner's own send/recv lists; the module imports without firedrake."""
import numpy as np
import pytest

from firedrake_ts_b200 import adapter_firedrake as A
from firedrake_ts_b200 import mesh as M
from fake_petsc import sf_graph_of


def test_module_imports_without_firedrake():
    assert A.HAVE_FIREDRAKE is False and A.B200TSContext is None
    with pytest.raises(RuntimeError):
        A.use_b200(object())


@pytest.mark.parametrize("dim,degree,n,grid", [(2, 1, 6, (2, 1)), (2, 2, 5, (2, 2)), (3, 2, 4, (2, 2, 2)), (3, 1, 5, (2, 1, 1))])
def test_halo_lists_from_sf_match_partitioner(dim, degree, n, grid):
    """all ranks in one process: the all-to-all is simulated by collecting every rank's request lists."""
    size = int(np.prod(grid))
    spaces = []
    for r in range(size):
        msh = M.SimplexMesh((n,) * dim, partition=M.Partition(grid, r))
        spaces.append(M.FunctionSpace(msh, degree))
    graphs = [sf_graph_of(V) for V in spaces]
    # pass 1: every rank's "wanted" lists (what it sends into the all-to-all)
    wanted = []
    for r, V in enumerate(spaces):
        box = {}
        A.halo_lists_from_sf(r, size, V.num_owned_nodes, graphs[r][1], graphs[r][2],
                             lambda w, box=box: box.setdefault("w", w) and [np.zeros(0, dtype=np.int64)] * size)
        wanted.append(box["w"])
    # pass 2: with the transposed lists
    for r, V in enumerate(spaces):
        nb, send, recv = A.halo_lists_from_sf(r, size, V.num_owned_nodes, graphs[r][1], graphs[r][2],
                                              lambda w: [wanted[o][r] for o in range(size)])
        assert nb == sorted(set(V.recv_slices) | set(V.send_lists))
        for o, (first, count) in V.recv_slices.items():
            assert sorted(recv[o].tolist()) == list(range(first, first + count))
        for o, lst in V.send_lists.items():
            assert sorted(send[o].tolist()) == sorted(lst.tolist())
        # consistency of the pairing: the k-th node rank o packs for me is the k-th ghost I unpack
        for o in recv:
            Vo = spaces[o]
            _, sendo, _ = A.halo_lists_from_sf(o, size, Vo.num_owned_nodes, graphs[o][1], graphs[o][2],
                                               lambda w: [wanted[q][o] for q in range(size)])
            assert len(sendo[r]) == len(recv[o])
            assert np.array_equal(Vo.global_node_ids[sendo[r].astype(np.int64)].numpy(),
                                  V.global_node_ids[recv[o].astype(np.int64)].numpy())
