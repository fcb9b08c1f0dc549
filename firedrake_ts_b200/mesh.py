"""Synthetic structured simplex meshes, Lagrange function spaces and the
box partitioner.

This module produces exactly the arrays the engine *consumes* from a live
Firedrake problem: coordinates, cell->coordinate map,
cell->node map, boundary-node list, owned/ghost split and neighbour send/recv
lists.  It mimics Firedrake's ``UnitIntervalMesh`` / ``UnitSquareMesh``
(diagonal "left") / ``UnitCubeMesh`` (six tetrahedra per cube around the
v0-v7 diagonal) / ``PeriodicIntervalMesh`` cell lists.  Real Firedrake
renumbers, so global ids are a property of this generator, not of the
reference; counts (cells, DOFs, nnz) are numbering independent and are checked
against SURVEY.md section 8's table in tests/test_mesh.py.

Everything is closed-form torch arithmetic so that the 37 M-cell headline mesh
can be generated directly on the GPU in well under a second.

Node identification: a P_k node is the integer point ``W[l] @ vertex_lattice``
of the k-times refined vertex lattice (fem.lattice_weights).  Cell vertices are
listed in ascending lexicographic lattice order (a rank-independent total
order), which fixes edge orientation globally (SURVEY.md hard part 4).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np
import torch

from . import fem

# cells of one lattice box, as offsets of their (sorted) vertices
_CELLS = {
    1: [[(0,), (1,)]],
    # quad corners around the perimeter v0=(0,0) v1=(1,0) v2=(1,1) v3=(0,1);
    # triangles (v0,v1,v3) and (v1,v2,v3), vertices sorted lexicographically
    2: [[(0, 0), (1, 0), (0, 1)], [(1, 0), (0, 1), (1, 1)]],
    # Kuhn paths 0 -> 7; listed in the order of SURVEY.md A.3
    3: [],
}
for perm in [(0, 1, 2), (0, 2, 1), (2, 0, 1), (1, 0, 2), (2, 1, 0), (1, 2, 0)]:
    v = [0, 0, 0]
    path = [tuple(v)]
    for ax in perm:
        v[ax] = 1
        path.append(tuple(v))
    _CELLS[3].append(path)


def _tile_sizes(n, tile, offset):
    """sizes of the 1-D tiles of range(n): the first tile has ``offset`` points
    (if offset > 0), then full tiles of ``tile`` points, then the remainder."""
    out = []
    if offset > 0:
        out.append(min(offset, n))
    left = n - sum(out)
    while left > 0:
        out.append(min(tile, left))
        left -= out[-1]
    return out


def _tile_tuple(tile, dim):
    """normalise a tile spec (None, int or per-dimension sequence) to a tuple of ints (0 = untiled)."""
    if tile is None:
        return (0,) * dim
    if isinstance(tile, (int, np.integer)):
        return (max(int(tile), 0),) * dim
    t = tuple(max(int(x), 0) for x in tile)
    assert len(t) == dim, "tile must be an int or one entry per dimension"
    if any(x == 0 for x in t):
        return (0,) * dim
    return t


def _tiled_index(p, dims, tile, offset=0):
    """Bijective numbering of the integer box prod(range(dims[a])) that walks
    tile by tile (tile edges ``tile[a]``; tile boundaries at offset + k*tile[a]; first
    and last tiles ragged), x fastest inside a tile.  ``p``: (..., dim) int64
    tensor.  tile <= 0 means lexicographic."""
    dim = len(dims)
    tt = _tile_tuple(tile, dim)
    if tt[0] <= 0:
        idx = p[..., dim - 1]
        for a in range(dim - 2, -1, -1):
            idx = idx * dims[a] + p[..., a]
        return idx
    start, l, s = [], [], []
    for a in range(dim):
        T = tt[a]
        sh = (T - int(offset) % T) % T  # shift so that tiles are uniform on x' = x + sh
        xs = p[..., a] + sh
        t = torch.div(xs, T, rounding_mode="floor")
        st = torch.clamp(t * T, min=sh)
        en = torch.clamp((t + 1) * T, max=dims[a] + sh)
        start.append(st - sh)
        l.append(xs - st)
        s.append(en - st)
    # offset of the tile: full slabs of higher dimensions first
    off = torch.zeros_like(p[..., 0])
    # dims above `a` are restricted to the current tile extent s[b], dims below are full
    for a in range(dim - 1, -1, -1):
        block = start[a]
        for b in range(dim):
            if b < a:
                block = block * dims[b]
            elif b > a:
                block = block * s[b]
        off = off + block
    loc = l[dim - 1]
    for a in range(dim - 2, -1, -1):
        loc = loc * s[a] + l[a]
    return off + loc


_MORTON_CACHE = {}


def _morton_rank(dims, device):
    """rank of every point of the integer box prod(range(dims[a])) (lexicographic index, first coordinate
    fastest) along the Morton (Z-order) curve: a numbering without any tile structure -- consecutive ids are
    spatially compact but no two row blocks look alike, which is what a mesh numbered by a generic
    cache-friendly traversal (DMPlex / RCM) gives the engine.  Single rank only."""
    key = (tuple(int(d) for d in dims), str(device))
    if key not in _MORTON_CACHE:
        dim = len(dims)
        n = int(np.prod(dims))
        lex = torch.arange(n, device=device, dtype=torch.int64)
        coords, rem = [], lex
        for a in range(dim):
            coords.append(rem % dims[a])
            rem = torch.div(rem, dims[a], rounding_mode="floor")
        code = torch.zeros(n, dtype=torch.int64, device=device)
        for b in range(max(int(d - 1).bit_length() for d in dims)):
            for a in range(dim):
                code |= ((coords[a] >> b) & 1) << (b * dim + a)
        order = torch.argsort(code)
        rank = torch.empty(n, dtype=torch.int64, device=device)
        rank[order] = torch.arange(n, device=device, dtype=torch.int64)
        _MORTON_CACHE[key] = rank
    return _MORTON_CACHE[key]


def _numbering_index(p, dims, tile, offset=0, numbering=None):
    """_tiled_index, or the Morton rank when numbering == "morton" """
    if numbering == "morton":
        return _morton_rank(dims, p.device)[_tiled_index(p, dims, 0)]
    return _tiled_index(p, dims, tile, offset)


@dataclass
class Partition:
    """Box partition of the cube lattice: ``grid`` ranks per dimension."""
    grid: tuple
    rank: int = 0

    @property
    def size(self):
        return int(np.prod(self.grid))

    def coords_of(self, rank):
        c = []
        for g in self.grid:
            c.append(rank % g)
            rank //= g
        return tuple(c)

    def rank_of(self, coords):
        r, m = 0, 1
        for c, g in zip(coords, self.grid):
            r += c * m
            m *= g
        return r


def default_grid(size: int, dim: int):
    """2/4/8 ranks -> 2x1x1 / 2x2x1 / 2x2x2."""
    grid = [1] * dim
    a = 0
    while size > 1:
        assert size % 2 == 0, "rank count must be a power of two"
        grid[a % dim] *= 2
        size //= 2
        a += 1
    return tuple(grid)


def _split(n, parts, i):
    """cube range [lo, hi) of part i of n cubes."""
    base, rem = divmod(n, parts)
    lo = i * base + min(i, rem)
    return lo, lo + base + (1 if i < rem else 0)


class SimplexMesh:
    """Structured simplex mesh of the box prod([0, L_a]) with ``shape`` cubes.

    With a :class:`Partition` the object holds the *local* part of one rank:
    all cells of the cubes [c0-1, c1) (clipped), i.e. every cell that touches
    an owned node (overlap-1 exec halo); interior cells
    (no ghost nodes) are listed first.
    """

    def __init__(self, shape, lengths=None, periodic=False, partition: Partition | None = None,
                 tile=0, device="cpu", quadrilateral=False, numbering=None):
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        # quadrilateral=True: one Q1 cell per lattice square (reference examples/cahn-hilliard.py:14,
        # UnitSquareMesh(96, 96, quadrilateral=True)); vertices in the tensor-product order of fem.QUAD_VERTICES
        self.cell = "quadrilateral" if quadrilateral else "simplex"
        assert not quadrilateral or self.dim == 2, "quadrilateral cells: two dimensions only"
        self.lengths = tuple(float(x) for x in (lengths or [1.0] * self.dim))
        self.periodic = bool(periodic)
        assert not self.periodic or self.dim == 1, "periodic meshes: 1-D only"
        self.partition = partition or Partition((1,) * self.dim, 0)
        assert not (self.periodic and self.partition.size > 1)
        self.tile = _tile_tuple(tile, len(self.shape))
        self.numbering = numbering  # None (tiled / lexicographic) or "morton" (no tile structure; single rank)
        assert numbering in (None, "morton")
        self.device = torch.device(device)
        dim = self.dim
        pc = self.partition.coords_of(self.partition.rank)
        self.own_lo, self.own_hi = [], []
        for a in range(dim):
            lo, hi = _split(self.shape[a], self.partition.grid[a], pc[a])
            self.own_lo.append(lo)
            self.own_hi.append(hi)
        # local cube box (with the low-side halo layer)
        self.box_lo = [max(self.own_lo[a] - 1, 0) for a in range(dim)]
        self.box_hi = list(self.own_hi)
        self._build_cells()

    # -- cells -------------------------------------------------------------
    def _cube_list(self):
        """Local cubes, interior cubes first; each group in tiled order."""
        dim, dev = self.dim, self.device
        dims = [self.box_hi[a] - self.box_lo[a] for a in range(dim)]
        axes = [torch.arange(dims[a], device=dev, dtype=torch.int64) for a in range(dim)]
        grids = torch.meshgrid(*axes, indexing="ij")
        p = torch.stack([g.reshape(-1) for g in grids], dim=-1)  # local cube coords
        key = _numbering_index(p, dims, self.tile, 0, self.numbering)
        p = p[torch.argsort(key)]
        cube = p + torch.tensor(self.box_lo, device=dev, dtype=torch.int64)
        # a cube touches ghost nodes iff it lies in the halo layer or at the
        # upper owned face next to another rank
        pc = self.partition.coords_of(self.partition.rank)
        halo = torch.zeros(cube.shape[0], dtype=torch.bool, device=dev)
        for a in range(dim):
            if self.box_lo[a] < self.own_lo[a]:
                halo |= cube[:, a] < self.own_lo[a]
            if pc[a] + 1 < self.partition.grid[a]:
                halo |= cube[:, a] == self.own_hi[a] - 1
        order = torch.argsort(halo.to(torch.int8), stable=True)
        self.num_interior_cubes = int((~halo).sum())
        return cube[order]

    def _build_cells(self):
        dim, dev = self.dim, self.device
        cube = self._cube_list()  # (ncube, dim) global cube coords
        cells = [fem.QUAD_VERTICES] if self.cell == "quadrilateral" else _CELLS[dim]
        tmpl = torch.tensor(cells, device=dev, dtype=torch.int64)  # (cpc, nverts, dim)
        self.cells_per_cube = tmpl.shape[0]
        self.nverts = tmpl.shape[1]  # coordinate nodes per cell: dim+1 (simplex) or 4 (quadrilateral)
        # (ncube, cpc, nverts, dim) -> (nc, nverts, dim): global vertex-lattice coordinates
        self.cell_vlat = (cube[:, None, None, :] + tmpl[None]).reshape(-1, self.nverts, dim)
        self.num_cells = self.cell_vlat.shape[0]
        self.num_interior_cells = self.num_interior_cubes * self.cells_per_cube
        # coordinate field: P1 (DG-like copy of the last vertex for periodic meshes)
        lo = torch.tensor(self.box_lo, device=dev, dtype=torch.int64)
        vdims = [self.box_hi[a] - self.box_lo[a] + 1 for a in range(dim)]
        self._vdims = vdims
        self.cell_coord_map = _numbering_index(self.cell_vlat - lo, vdims, self.tile, 0,
                                               self.numbering).to(torch.int32).contiguous()
        nv = int(np.prod(vdims))
        self.num_coord_nodes = nv
        coords = torch.zeros((nv, dim), dtype=torch.float64, device=dev)
        h = torch.tensor([self.lengths[a] / self.shape[a] for a in range(dim)], dtype=torch.float64, device=dev)
        flat = self.cell_coord_map.reshape(-1).to(torch.int64)
        coords[flat] = self.cell_vlat.reshape(-1, dim).to(torch.float64) * h
        self.coords = coords

    def cell_tile_starts(self, max_pairs=409):
        """cell offsets of contiguous cell tiles for the residual tile kernel: the cubes of one numbering
        tile (single rank, exact tile sizes) or uniform chunks (partitioned meshes / untiled numbering);
        the interior/halo boundary is always a tile boundary; at most ``max_pairs`` cells per tile."""
        cpc = self.cells_per_cube
        dims = [self.box_hi[a] - self.box_lo[a] for a in range(self.dim)]
        sizes = None
        if self.partition.size == 1 and self.tile[0] > 0:
            out = torch.ones(1, dtype=torch.int64)
            for a in range(self.dim - 1, -1, -1):
                out = (out[:, None] * torch.tensor(_tile_sizes(dims[a], self.tile[a], 0), dtype=torch.int64)[None, :]).reshape(-1)
            sizes = out * cpc
            if int(sizes.max()) > max_pairs:
                sizes = None
        if sizes is None:
            chunk = max(cpc, (min(max_pairs, 27 * cpc) // cpc) * cpc)
            parts = []
            for lo, hi in ((0, self.num_interior_cells), (self.num_interior_cells, self.num_cells)):
                n = hi - lo
                if n > 0:
                    full, rem = divmod(n, chunk)
                    parts += [chunk] * full + ([rem] if rem else [])
            sizes = torch.tensor(parts, dtype=torch.int64)
        starts = torch.zeros(sizes.numel() + 1, dtype=torch.int64)
        starts[1:] = torch.cumsum(sizes, 0)
        assert int(starts[-1]) == self.num_cells
        return starts

    # -- counts ------------------------------------------------------------
    @property
    def global_num_cells(self):
        return int(np.prod(self.shape)) * self.cells_per_cube


class FunctionSpace:
    """Lagrange P_k space (``ncomp`` components) on a :class:`SimplexMesh`.

    layout "interleaved": dof = node*ncomp + comp  (VectorFunctionSpace)
    layout "blocked":     dof = comp*num_nodes + node (mixed space V*V, one
                          field block after the other)

    Local node numbering: owned nodes first (tiled order inside the owned box),
    then ghost nodes grouped by owner rank (ascending), lexicographic inside a
    group; so each neighbour's halo data lands in one contiguous tail slice.
    """

    def __init__(self, mesh: SimplexMesh, degree: int, ncomp: int = 1, layout: str = "interleaved",
                 variant: str = "gll", tile=None, tile_offset=0, block_rows=192):
        assert degree in (1, 2, 3)
        assert layout in ("interleaved", "blocked")
        self.mesh, self.degree, self.ncomp, self.layout, self.variant = mesh, degree, ncomp, layout, variant
        self.dim = mesh.dim
        self.tile = (tuple(t * degree for t in mesh.tile) if tile is None else _tile_tuple(tile, mesh.dim))
        self.tile_offset = int(tile_offset)
        self.numbering = getattr(mesh, "numbering", None)
        self.block_rows = int(block_rows)  # rows per gather block for numberings without tiles
        assert self.numbering is None or mesh.partition.size == 1, "Morton numbering: single rank only"
        self.nd = fem.num_nodes(self.dim, degree, mesh.cell)
        self._build()

    # lattice helpers --------------------------------------------------------
    def _gdims(self):
        m, k = self.mesh, self.degree
        return [k * m.shape[a] + (0 if m.periodic else 1) for a in range(m.dim)]

    def _owner_coords(self, p):
        """rank coordinates of the owner of refined-lattice points p (..., dim)."""
        m, k = self.mesh, self.degree
        out = []
        for a in range(m.dim):
            g = m.partition.grid[a]
            bounds = [k * _split(m.shape[a], g, i)[1] for i in range(g)]  # exclusive upper ends
            bounds[-1] += 1  # the last rank owns the closing boundary plane
            b = torch.tensor(bounds, device=p.device, dtype=torch.int64)
            out.append(torch.bucketize(p[..., a].contiguous(), b, right=True))
        return torch.stack(out, dim=-1)

    def _owned_box(self, rank):
        m, k = self.mesh, self.degree
        pc = m.partition.coords_of(rank)
        lo, hi = [], []
        for a in range(m.dim):
            l, h = _split(m.shape[a], m.partition.grid[a], pc[a])
            lo.append(k * l)
            hi.append(k * h + (1 if (pc[a] + 1 == m.partition.grid[a] and not m.periodic) else 0))
        return lo, hi

    def _local_ids(self, p, rank=None):
        """local node id of refined-lattice points ``p`` on ``rank`` (own rank by
        default); -1 where the point is not a local node of that rank."""
        m = self.mesh
        part = m.partition
        rank = part.rank if rank is None else rank
        dev = p.device
        lo, hi = self._owned_box(rank)
        lo_t = torch.tensor(lo, device=dev, dtype=torch.int64)
        dims = [hi[a] - lo[a] for a in range(m.dim)]
        nown = int(np.prod(dims))
        dims_t = torch.tensor(dims, device=dev, dtype=torch.int64)
        q = p - lo_t
        owned = ((q >= 0) & (q < dims_t)).all(dim=-1)
        ids = torch.full(p.shape[:-1], -1, dtype=torch.int64, device=dev)
        qo = torch.where(owned[..., None], q, torch.zeros_like(q))
        ids_owned = _numbering_index(qo, dims, self.tile, self.tile_offset, self.numbering)
        ids = torch.where(owned, ids_owned, ids)
        if part.size > 1:
            gl = self._ghost_table(rank)
            if gl["keys"].numel():
                key = self._lex_key(p)
                pos = torch.searchsorted(gl["keys"], key)
                pos = pos.clamp(max=gl["keys"].numel() - 1)
                hit = (gl["keys"][pos] == key) & ~owned
                ids = torch.where(hit, nown + gl["rank_of_key"][pos], ids)
        return ids

    def _lex_key(self, p):
        gd = self._gdims()
        return _tiled_index(p, gd, 0)

    def _ghost_table(self, rank):
        """ghost nodes of ``rank``: sorted lex keys plus their ghost ordinal."""
        cache = self.__dict__.setdefault("_ghost_cache", {})
        if rank in cache:
            return cache[rank]
        m, k = self.mesh, self.degree
        part = m.partition
        dev = m.device
        pc = part.coords_of(rank)
        lo, hi = self._owned_box(rank)
        # closure box of the rank's local cells on the refined lattice
        blo, bhi = [], []
        for a in range(m.dim):
            l, h = _split(m.shape[a], part.grid[a], pc[a])
            blo.append(k * max(l - 1, 0))
            bhi.append(k * h + 1)
        axes = [torch.arange(blo[a], bhi[a], device=dev, dtype=torch.int64) for a in range(m.dim)]
        # enumerate only the shell: all points of the closure box that are not owned
        grids = torch.meshgrid(*axes, indexing="ij")
        p = torch.stack([g.reshape(-1) for g in grids], dim=-1)
        lo_t = torch.tensor(lo, device=dev)
        hi_t = torch.tensor(hi, device=dev)
        owned = ((p >= lo_t) & (p < hi_t)).all(dim=-1)
        p = p[~owned]
        oc = self._owner_coords(p)
        orank = torch.zeros(p.shape[0], dtype=torch.int64, device=dev)
        mul = 1
        for a in range(m.dim):
            orank += oc[:, a] * mul
            mul *= part.grid[a]
        key = self._lex_key(p)
        # ghost order: by owner rank, then lexicographic
        order = torch.argsort(orank * (int(key.max()) + 1 if key.numel() else 1) + key) if key.numel() else key
        p, orank, key = p[order], orank[order], key[order]
        ordinal = torch.arange(p.shape[0], device=dev, dtype=torch.int64)
        ks = torch.argsort(key)
        tab = {"points": p, "owner": orank, "keys": key[ks], "rank_of_key": ordinal[ks]}
        cache[rank] = tab
        return tab

    # build ---------------------------------------------------------------
    def _build(self):
        m, k, dev = self.mesh, self.degree, self.mesh.device
        dim = m.dim
        W = torch.tensor(fem.lattice_weights(dim, k, m.cell), device=dev, dtype=torch.int64)  # (nd, nverts)
        gd = torch.tensor(self._gdims(), device=dev, dtype=torch.int64)
        cols = []
        for l in range(self.nd):
            p = (W[l][None, :, None] * m.cell_vlat).sum(dim=1)  # (nc, dim)
            if m.periodic:
                p = p % gd
            cols.append(self._local_ids(p).to(torch.int32))
        self.cell_node_map = torch.stack(cols, dim=1).contiguous()
        assert int(self.cell_node_map.min()) >= 0
        lo, hi = self._owned_box(m.partition.rank)
        self.owned_lo, self.owned_hi = lo, hi
        self.num_owned_nodes = int(np.prod([hi[a] - lo[a] for a in range(dim)]))
        if m.partition.size > 1:
            gt = self._ghost_table(m.partition.rank)
            self.num_ghost_nodes = int(gt["points"].shape[0])
        else:
            self.num_ghost_nodes = 0
        self.num_nodes = self.num_owned_nodes + self.num_ghost_nodes
        self.global_num_nodes = int(np.prod(self._gdims()))
        self._build_node_info()
        self._build_halo()

    def _build_node_info(self):
        """node coordinates, global ids and boundary flags for local nodes."""
        m, k, dev = self.mesh, self.degree, self.mesh.device
        dim = m.dim
        bary = torch.tensor(fem.nodes_barycentric(dim, k, self.variant, m.cell), device=dev, dtype=torch.float64)
        W = torch.tensor(fem.lattice_weights(dim, k, m.cell), device=dev, dtype=torch.int64)
        h = torch.tensor([m.lengths[a] / m.shape[a] for a in range(dim)], dtype=torch.float64, device=dev)
        nn = self.num_nodes
        self.node_coords = torch.zeros((nn, dim), dtype=torch.float64, device=dev)
        lat = torch.zeros((nn, dim), dtype=torch.int64, device=dev)
        vx = m.cell_vlat.to(torch.float64) * h  # (nc, nverts, dim)
        gd = torch.tensor(self._gdims(), device=dev, dtype=torch.int64)
        for l in range(self.nd):
            idx = self.cell_node_map[:, l].to(torch.int64)
            x = (bary[l][None, :, None] * vx).sum(dim=1)
            p = (W[l][None, :, None] * m.cell_vlat).sum(dim=1)
            if m.periodic:
                p = p % gd
                x = x % m.lengths[0]
            self.node_coords[idx] = x
            lat[idx] = p
        self.node_lattice = lat
        self.global_node_ids = self._global_ids(lat)
        if m.periodic:
            self.boundary_nodes = torch.zeros(0, dtype=torch.int32, device=dev)
        else:
            onb = ((lat == 0) | (lat == gd - 1)).any(dim=-1)
            self.boundary_nodes = torch.nonzero(onb).reshape(-1).to(torch.int32)

    def _global_ids(self, lat):
        """rank-independent global node numbering: ranks in order, owned-local order inside."""
        m = self.mesh
        part = m.partition
        if part.size == 1:
            return self._local_ids(lat)
        oc = self._owner_coords(lat)
        orank = torch.zeros(lat.shape[0], dtype=torch.int64, device=lat.device)
        mul = 1
        for a in range(m.dim):
            orank += oc[:, a] * mul
            mul *= part.grid[a]
        out = torch.zeros(lat.shape[0], dtype=torch.int64, device=lat.device)
        base = 0
        for r in range(part.size):
            lo, hi = self._owned_box(r)
            dims = [hi[a] - lo[a] for a in range(m.dim)]
            sel = orank == r
            if sel.any():
                q = lat[sel] - torch.tensor(lo, device=lat.device)
                out[sel] = base + _tiled_index(q, dims, self.tile, self.tile_offset)
            base += int(np.prod(dims))
        return out

    def _build_halo(self):
        """neighbour lists: recv[(rank)] = (start, count) slice of the ghost tail;
        send[(rank)] = local owned node ids to pack, in the receiver's ghost order."""
        m = self.mesh
        part = m.partition
        self.recv_slices, self.send_lists = {}, {}
        if part.size == 1:
            return
        me = part.rank
        gt = self._ghost_table(me)
        owners = gt["owner"]
        if owners.numel():
            uniq, counts = torch.unique_consecutive(owners, return_counts=True)
            start = 0
            for r, c in zip(uniq.tolist(), counts.tolist()):
                self.recv_slices[r] = (self.num_owned_nodes + start, c)
                start += c
        # what do I send?  every rank whose ghost table names me as owner
        pc = part.coords_of(me)
        for delta in itertools.product((-1, 0, 1), repeat=m.dim):
            if not any(delta):
                continue
            nc = tuple(pc[a] + delta[a] for a in range(m.dim))
            if any(c < 0 or c >= part.grid[a] for a, c in enumerate(nc)):
                continue
            r = part.rank_of(nc)
            gtr = self._ghost_table(r)
            sel = gtr["owner"] == me
            if sel.any():
                ids = self._local_ids(gtr["points"][sel])
                assert int(ids.min()) >= 0 and int(ids.max()) < self.num_owned_nodes
                self.send_lists[r] = ids.to(torch.int32).contiguous()

    def row_block_starts(self):
        """owned-node row ranges of the numbering tiles (each tile is contiguous
        in the numbering): int64 tensor of nblocks+1 offsets.  The gather kernels
        use them as row blocks (one CTA per tile)."""
        dims = [self.owned_hi[a] - self.owned_lo[a] for a in range(self.dim)]
        if self.numbering == "morton":  # no tiles: uniform chunks of consecutive (spatially compact) rows
            starts = torch.arange(0, self.num_owned_nodes + self.block_rows, self.block_rows, dtype=torch.int64)
            return torch.unique(starts.clamp(max=self.num_owned_nodes))
        if self.tile[0] <= 0:
            return None
        per_dim = [_tile_sizes(dims[a], self.tile[a], self.tile_offset) for a in range(self.dim)]
        sizes = torch.tensor(per_dim[self.dim - 1], dtype=torch.int64)
        # order: slowest dimension outermost; tile (t_{d-1}, ..., t_0) has prod of sizes
        out = torch.ones(1, dtype=torch.int64)
        for a in range(self.dim - 1, -1, -1):
            out = (out[:, None] * torch.tensor(per_dim[a], dtype=torch.int64)[None, :]).reshape(-1)
        starts = torch.zeros(out.numel() + 1, dtype=torch.int64)
        starts[1:] = torch.cumsum(out, 0)
        assert int(starts[-1]) == self.num_owned_nodes
        return starts

    # dof-level helpers --------------------------------------------------------
    @property
    def num_owned_dofs(self):
        return self.num_owned_nodes * self.ncomp

    @property
    def num_dofs(self):
        return self.num_nodes * self.ncomp

    @property
    def global_num_dofs(self):
        return self.global_num_nodes * self.ncomp

    def dof_index(self, node, comp, n_nodes=None):
        n_nodes = self.num_nodes if n_nodes is None else n_nodes
        if self.layout == "interleaved":
            return node * self.ncomp + comp
        return comp * n_nodes + node

    def interpolate(self, fn):
        """sample ``fn(x) -> (n,) or (n, ncomp)`` at the local nodes; returns the
        local dof vector (owned + ghost) in this space's layout."""
        v = fn(self.node_coords)
        if v.ndim == 1:
            v = v[:, None]
        v = v.to(torch.float64).expand(self.num_nodes, self.ncomp)
        if self.layout == "interleaved":
            return v.reshape(-1).contiguous()
        return v.t().reshape(-1).contiguous()


def UnitIntervalMesh(n, **kw):
    return SimplexMesh((n,), **kw)


def PeriodicIntervalMesh(n, length, **kw):
    return SimplexMesh((n,), lengths=(length,), periodic=True, **kw)


def UnitSquareMesh(nx, ny=None, quadrilateral=False, **kw):
    return SimplexMesh((nx, ny or nx), quadrilateral=quadrilateral, **kw)


def UnitCubeMesh(nx, ny=None, nz=None, **kw):
    return SimplexMesh((nx, ny or nx, nz or nx), **kw)
