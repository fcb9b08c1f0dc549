"""ctypes shim over libfdts_b200.so (the C ABI declared in include/fdts.h).

PyTorch is only used for device buffers and streams (``tensor.data_ptr()``,
``torch.cuda.current_stream()``); every kernel on the assembly path lives in
the shared library.  There is no CPU fallback: importing works without a GPU
(so that host-side logic can be tested), but constructing an :class:`Engine`
raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import fem

_HERE = os.path.dirname(os.path.abspath(__file__))
# FDTS_B200_LIB: development override (tools/ load the build with the timing switches compiled in)
LIB_PATH = os.environ.get("FDTS_B200_LIB") or os.path.join(_HERE, "csrc", "libfdts_b200.so")
_LIB = None

SYMBOLS = [
    "fdts_last_error", "fdts_version", "fdts_create", "fdts_destroy", "fdts_sizes", "fdts_get_csr",
    "fdts_get_node_csr", "fdts_set_param", "fdts_set_option", "fdts_get_option", "fdts_ifunction", "fdts_ijacobian",
    "fdts_ifunction_range", "fdts_ijacobian_range", "fdts_ifunction_host", "fdts_ijacobian_host", "fdts_halo_pack",
    "fdts_halo_unpack", "fdts_launch_count", "fdts_measure_fp64_peak", "fdts_csr_spmv", "fdts_csr_spmv_warp", "fdts_csr_diagonal", "fdts_set_row_blocks",
    "fdts_set_cell_tiles", "fdts_nccl_unique_id", "fdts_halo_setup", "fdts_halo_exchange", "fdts_ifunction_exchange",
    "fdts_owned_to_local",
]


class EngineError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("degree", C.c_int32), ("nd", C.c_int32), ("ncomp", C.c_int32), ("layout", C.c_int32),
        ("family", C.c_int32), ("nq", C.c_int32), ("memspace", C.c_int32), ("device", C.c_int32),
        ("cell_type", C.c_int32),
        ("num_cells", C.c_int64), ("num_interior_cells", C.c_int64), ("num_nodes", C.c_int64),
        ("num_owned_nodes", C.c_int64), ("num_coord_nodes", C.c_int64),
        ("coords", C.c_void_p), ("cell_coord", C.c_void_p), ("cell_node", C.c_void_p),
        ("B", C.c_void_p), ("w", C.c_void_p), ("R", C.c_void_p), ("M0", C.c_void_p),
        ("params", C.c_double * 4),
        ("num_bc_nodes", C.c_int64), ("bc_nodes", C.c_void_p),
    ]


class _Sizes(C.Structure):
    _fields_ = [("num_rows", C.c_int64), ("num_cols", C.c_int64), ("nnz", C.c_int64), ("node_nnz", C.c_int64),
                ("max_row_nodes", C.c_int32), ("reserved", C.c_int32)]


class _HaloConfig(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("num_neighbours", C.c_int32), ("reserved", C.c_int32),
                ("neighbour_rank", C.c_void_p), ("send_offset", C.c_void_p), ("send_nodes", C.c_void_p),
                ("recv_first", C.c_void_p), ("recv_count", C.c_void_p), ("nccl_library", C.c_char_p),
                ("nccl_unique_id", C.c_void_p), ("recv_offset", C.c_void_p), ("recv_nodes", C.c_void_p)]


def nccl_library_path():
    """the libnccl.so.2 this process already uses (torch's bundled one once torch.distributed/NCCL is loaded),
    else the wheel's copy; None lets the dynamic loader decide"""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia.nccl as _n

        cand = os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except ImportError:
        pass
    return None


def load_library():
    """dlopen the in-tree library; raises EngineError when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.fdts_last_error.restype = C.c_char_p
        lib.fdts_version.restype = C.c_char_p
        lib.fdts_launch_count.restype = C.c_int64
        lib.fdts_launch_count.argtypes = [C.c_void_p]
        lib.fdts_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
        lib.fdts_destroy.argtypes = [C.c_void_p]
        lib.fdts_sizes.argtypes = [C.c_void_p, C.POINTER(_Sizes)]
        lib.fdts_get_csr.argtypes = [C.c_void_p] * 4
        lib.fdts_get_node_csr.argtypes = [C.c_void_p] * 4
        lib.fdts_set_param.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        lib.fdts_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        lib.fdts_get_option.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
        lib.fdts_ifunction.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fdts_ijacobian.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                       C.c_void_p]
        lib.fdts_ifunction_range.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_int64, C.c_int32, C.c_void_p]
        lib.fdts_ijacobian_range.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                             C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
        lib.fdts_ifunction_host.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fdts_ijacobian_host.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        lib.fdts_halo_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.fdts_halo_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        lib.fdts_set_row_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        lib.fdts_set_cell_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        lib.fdts_csr_spmv.argtypes = [C.c_int64] + [C.c_void_p] * 6
        lib.fdts_csr_spmv_warp.argtypes = [C.c_int64] + [C.c_void_p] * 6
        lib.fdts_csr_diagonal.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                          C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.fdts_measure_fp64_peak.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
        lib.fdts_nccl_unique_id.argtypes = [C.c_char_p, C.c_void_p]
        lib.fdts_halo_setup.argtypes = [C.c_void_p, C.POINTER(_HaloConfig)]
        lib.fdts_halo_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fdts_owned_to_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fdts_ifunction_exchange.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


def _check(rc, what):
    if rc != 0:
        raise EngineError(f"{what} failed (code {rc}): {load_library().fdts_last_error().decode()}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def measure_fp64_peak(device=0, dmma=False) -> float:
    out = C.c_double(0.0)
    _check(load_library().fdts_measure_fp64_peak(device, 1 if dmma else 0, C.byref(out)), "fdts_measure_fp64_peak")
    return out.value


class Engine:
    """One assembly problem resident on one GPU (replaces ``_assemble_residual``
    / ``_assemble_jac`` of the reference's ``_TSContext``,
    firedrake_ts/solving_utils.py:258,305)."""

    def __init__(self, form, bcs=(), device=None, scatter=None, plan_version=None, dedup=None, options=None):
        if not torch.cuda.is_available():
            raise EngineError("no CUDA device: the B200 assembly engine has no CPU fallback")
        lib = load_library()
        V = form.space
        m = V.mesh
        dev = torch.device(device if device is not None else (m.device if m.device.type == "cuda" else "cuda:0"))
        self.device = dev
        self.form, self.V = form, V
        cell = getattr(m, "cell", "simplex")
        et = fem.element_tables(V.dim, V.degree, form.qdegree, V.variant, cell)
        bc_nodes = torch.zeros(0, dtype=torch.int32)
        if bcs:
            bc_nodes = torch.unique(torch.cat([bc.nodes.to(torch.int64).cpu() for bc in bcs])).to(torch.int32)

        def dv(x, dtype):
            if not torch.is_tensor(x):
                x = torch.from_numpy(np.ascontiguousarray(x))
            return x.to(device=dev, dtype=dtype).contiguous()

        keep = dict(coords=dv(m.coords, torch.float64), cc=dv(m.cell_coord_map, torch.int32),
                    cn=dv(V.cell_node_map, torch.int32), B=dv(et.B, torch.float64), w=dv(et.w, torch.float64),
                    R=dv(et.R, torch.float64), M0=dv(et.M0, torch.float64), bc=dv(bc_nodes, torch.int32))
        cfg = _Config()
        cfg.dim, cfg.degree, cfg.nd, cfg.ncomp = V.dim, V.degree, V.nd, V.ncomp
        cfg.layout = 0 if V.layout == "interleaved" else 1
        cfg.family, cfg.nq, cfg.memspace = form.family, et.nq, 1
        cfg.device = dev.index or 0
        cfg.cell_type = 1 if cell == "quadrilateral" else 0
        cfg.num_cells, cfg.num_interior_cells = m.num_cells, m.num_interior_cells
        cfg.num_nodes, cfg.num_owned_nodes = V.num_nodes, V.num_owned_nodes
        cfg.num_coord_nodes = m.num_coord_nodes
        cfg.coords, cfg.cell_coord, cfg.cell_node = keep["coords"].data_ptr(), keep["cc"].data_ptr(), keep["cn"].data_ptr()
        cfg.B, cfg.w, cfg.R, cfg.M0 = keep["B"].data_ptr(), keep["w"].data_ptr(), keep["R"].data_ptr(), keep["M0"].data_ptr()
        for i, v in enumerate(form.params):
            cfg.params[i] = v
        cfg.num_bc_nodes, cfg.bc_nodes = bc_nodes.numel(), keep["bc"].data_ptr()
        self._h = C.c_void_p()
        torch.cuda.synchronize(dev)
        with torch.cuda.device(dev):
            _check(lib.fdts_create(C.byref(cfg), C.byref(self._h)), "fdts_create")
        del keep  # the library holds its own copies
        s = _Sizes()
        _check(lib.fdts_sizes(self._h, C.byref(s)), "fdts_sizes")
        self.num_rows, self.num_cols, self.nnz, self.node_nnz = s.num_rows, s.num_cols, s.nnz, s.node_nnz
        self.max_row_nodes, self.max_degree = s.max_row_nodes, s.reserved
        self.num_cells, self.num_interior_cells = m.num_cells, m.num_interior_cells
        self._lib = lib
        self.has_native_halo = False
        if plan_version is not None:
            self.set_option("plan_version", plan_version)
        if dedup is not None:
            self.set_option("dedup", dedup)
        for k, v in (options or {}).items():  # plan options must be set before the plan is built
            self.set_option(k, v)
        if scatter is not None:
            if int(scatter) == 2:
                starts = V.row_block_starts()
                if starts is None:
                    # untiled numbering (e.g. a live Firedrake space): uniform chunks of consecutive rows; how many
                    # cells / vertices a chunk touches depends on the numbering, so shrink until the plan fits
                    err = None
                    for rows in (160, 128, 96, 64, 32):
                        starts = torch.unique(torch.arange(0, V.num_owned_nodes + rows, rows).clamp(max=V.num_owned_nodes))
                        try:
                            self.set_row_blocks(starts)
                            err = None
                            break
                        except EngineError as exc:
                            err = exc
                    if err is not None:
                        raise err
                else:
                    self.set_row_blocks(starts)
                if form.family == 0 and V.ncomp == 1 and (options or {}).get("residual_tiles", 0):
                    self.set_cell_tiles(m.cell_tile_starts(max_pairs=4096 // V.nd))
            self.set_option("scatter", scatter)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.fdts_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # -- options / params -------------------------------------------------
    def set_option(self, key, value):
        _check(self._lib.fdts_set_option(self._h, key.encode(), int(value)), "fdts_set_option")

    def get_option(self, key):
        v = C.c_int64(0)
        _check(self._lib.fdts_get_option(self._h, key.encode(), C.byref(v)), "fdts_get_option")
        return v.value

    def set_row_blocks(self, starts):
        """row blocks (owned-node row offsets) of the owner-computes gather path; builds the plan."""
        st = torch.as_tensor(starts, dtype=torch.int64).cpu().contiguous()
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_set_row_blocks(self._h, st.data_ptr(), st.numel() - 1), "fdts_set_row_blocks")

    def set_cell_tiles(self, starts):
        """cell tiles (cell offsets) of the residual tile kernel; builds the tile plan."""
        st = torch.as_tensor(starts, dtype=torch.int64).cpu().contiguous()
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_set_cell_tiles(self._h, st.data_ptr(), st.numel() - 1), "fdts_set_cell_tiles")

    def set_param(self, index, value):
        _check(self._lib.fdts_set_param(self._h, index, float(value)), "fdts_set_param")

    @property
    def launch_count(self):
        return self._lib.fdts_launch_count(self._h)

    # -- sparsity -----------------------------------------------------------
    def csr(self):
        """dof-level (rowptr int64, colidx int32) device tensors of the AIJ pattern."""
        rowptr = torch.empty(self.num_rows + 1, dtype=torch.int64, device=self.device)
        colidx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=self.device)[: self.nnz]
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_get_csr(self._h, rowptr.data_ptr(), colidx.data_ptr(), _stream()), "fdts_get_csr")
        return rowptr, colidx

    def node_csr(self):
        nown = self.V.num_owned_nodes
        rowptr = torch.empty(nown + 1, dtype=torch.int64, device=self.device)
        colidx = torch.empty(max(self.node_nnz, 1), dtype=torch.int32, device=self.device)[: self.node_nnz]
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_get_node_csr(self._h, rowptr.data_ptr(), colidx.data_ptr(), _stream()),
                   "fdts_get_node_csr")
        return rowptr, colidx

    def new_values(self):
        return torch.empty(max(self.nnz, 1), dtype=torch.float64, device=self.device)[: self.nnz]

    # -- the hot path ---------------------------------------------------------
    def _chk_vec(self, x, n, name):
        if not (x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.numel() == n):
            raise EngineError(f"{name}: expected a contiguous float64 CUDA tensor of {n} entries, got "
                              f"{tuple(x.shape)} {x.dtype} {x.device}")

    def ifunction(self, t, x, xdot, out=None):
        """F(t, x, xdot) for the owned dofs; x/xdot are local (owned+ghost) vectors."""
        self._chk_vec(x, self.num_cols, "x")
        self._chk_vec(xdot, self.num_cols, "xdot")
        if out is None:
            out = torch.empty(self.num_rows, dtype=torch.float64, device=self.device)
        self._chk_vec(out, self.num_rows, "F")
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_ifunction(self._h, float(t), x.data_ptr(), xdot.data_ptr(), out.data_ptr(),
                                            _stream()), "fdts_ifunction")
        return out

    def ijacobian(self, t, x, xdot, shift, out=None):
        """CSR values of shift*dF/dxdot + dF/dx at (t, x, xdot)."""
        self._chk_vec(x, self.num_cols, "x")
        self._chk_vec(xdot, self.num_cols, "xdot")
        if out is None:
            out = self.new_values()
        self._chk_vec(out, self.nnz, "J values")
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_ijacobian(self._h, float(t), x.data_ptr(), xdot.data_ptr(), float(shift),
                                            out.data_ptr(), _stream()), "fdts_ijacobian")
        return out

    def ifunction_range(self, t, x, xdot, out, c0, c1, flags):
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_ifunction_range(self._h, float(t), x.data_ptr(), xdot.data_ptr(), out.data_ptr(),
                                                  c0, c1, flags, _stream()), "fdts_ifunction_range")

    def ifunction_overlapped(self, t, x, xdot, halo, out=None):
        """IFunction with the ghost exchange of x / xdot hidden behind the interior cells (the cells
        that touch no ghost node are listed first): pack + NCCL send/recv are posted, the interior
        range is assembled meanwhile, the remaining cells after the receives completed.  Same split as
        PyOP2's core / owned+exec-halo parloop halves around ``global_to_local_begin/end``."""
        if out is None:
            out = torch.empty(self.num_rows, dtype=torch.float64, device=self.device)
        r1 = halo.start(x, 0)
        r2 = halo.start(xdot, 1)
        ni = self.num_interior_cells
        self.ifunction_range(t, x, xdot, out, 0, ni, 1)
        halo.finish(x, r1, 0)
        halo.finish(xdot, r2, 1)
        if ni < self.num_cells:
            self.ifunction_range(t, x, xdot, out, ni, self.num_cells, 0)
        return out

    def ijacobian_range(self, t, x, xdot, shift, out, c0, c1, flags):
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_ijacobian_range(self._h, float(t), x.data_ptr(), xdot.data_ptr(), float(shift),
                                                  out.data_ptr(), c0, c1, flags, _stream()), "fdts_ijacobian_range")

    # host-buffer variants (reference-facing call with host Vec/Mat storage)
    def ifunction_host(self, t, x, xdot, out):
        for a in (x, xdot, out):
            assert not a.is_cuda and a.dtype == torch.float64 and a.is_contiguous()
        _check(self._lib.fdts_ifunction_host(self._h, float(t), x.data_ptr(), xdot.data_ptr(), out.data_ptr()),
               "fdts_ifunction_host")
        return out

    def ijacobian_host(self, t, x, xdot, shift, out):
        for a in (x, xdot, out):
            assert not a.is_cuda and a.dtype == torch.float64 and a.is_contiguous()
        _check(self._lib.fdts_ijacobian_host(self._h, float(t), x.data_ptr(), xdot.data_ptr(), float(shift),
                                             out.data_ptr()), "fdts_ijacobian_host")
        return out

    # halo: exchange inside the library (NCCL send/recv grouped per call, fused with the residual) -----------
    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        path = nccl_library_path()
        _check(load_library().fdts_nccl_unique_id(path.encode() if path else None, buf), "fdts_nccl_unique_id")
        return buf.raw

    def halo_setup_raw(self, rank, nranks, neighbours, send_lists, recv_slices, unique_id, recv_lists=None):
        """neighbours: ranks; send_lists[r]: owned node ids to pack for r (the receiver's ghost order);
        recv_slices[r] = (first local node, count), or recv_lists[r] = explicit ghost node ids (numberings
        whose ghosts are not grouped by owner).  Collective over the ranks (ncclCommInitRank)."""
        nb = np.asarray(list(neighbours), dtype=np.int32)
        off, nodes, first, count = [0], [], [], []
        for r in nb.tolist():
            lst = np.asarray(send_lists.get(r, np.zeros(0)), dtype=np.int32).reshape(-1)
            nodes.append(lst)
            off.append(off[-1] + lst.size)
            f, c = recv_slices.get(r, (self.V.num_owned_nodes, 0))
            first.append(int(f))
            count.append(int(c))
        off = np.asarray(off, dtype=np.int64)
        nodes = np.ascontiguousarray(np.concatenate(nodes) if nodes else np.zeros(0), dtype=np.int32)
        first, count = np.asarray(first, dtype=np.int64), np.asarray(count, dtype=np.int64)
        idbuf = C.create_string_buffer(bytes(unique_id), 128)
        path = nccl_library_path()
        cfg = _HaloConfig()
        cfg.rank, cfg.nranks, cfg.num_neighbours = int(rank), int(nranks), int(nb.size)
        cfg.neighbour_rank, cfg.send_offset = nb.ctypes.data, off.ctypes.data
        cfg.send_nodes, cfg.recv_first, cfg.recv_count = nodes.ctypes.data, first.ctypes.data, count.ctypes.data
        cfg.nccl_library = path.encode() if path else None
        cfg.nccl_unique_id = C.cast(idbuf, C.c_void_p)
        if recv_lists is not None:
            rl = [np.asarray(recv_lists.get(r, np.zeros(0)), dtype=np.int32).reshape(-1) for r in nb.tolist()]
            roff = np.concatenate([[0], np.cumsum([x.size for x in rl])]).astype(np.int64)
            rnodes = np.ascontiguousarray(np.concatenate(rl) if rl else np.zeros(0), dtype=np.int32)
            cfg.recv_offset, cfg.recv_nodes = roff.ctypes.data, rnodes.ctypes.data
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_halo_setup(self._h, C.byref(cfg)), "fdts_halo_setup")
        self.has_native_halo = True

    def halo_setup(self, group=None):
        """open the library's own NCCL communicator over the ranks of ``group`` (default: WORLD) for the
        ghost exchange of this engine's space; collective."""
        import torch.distributed as dist

        V = self.V
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        assert rank == V.mesh.partition.rank and world == V.mesh.partition.size
        box = [self.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        nbs = sorted(set(V.recv_slices) | set(V.send_lists))
        self.halo_setup_raw(rank, world, nbs, {r: v.cpu().numpy() for r, v in V.send_lists.items()},
                            dict(V.recv_slices), box[0])

    def halo_exchange(self, x, xdot):
        """refresh the ghost entries of the local vectors x and xdot (stream-ordered)"""
        self._chk_vec(x, self.num_cols, "x")
        self._chk_vec(xdot, self.num_cols, "xdot")
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_halo_exchange(self._h, x.data_ptr(), xdot.data_ptr(), _stream()), "fdts_halo_exchange")

    def ifunction_exchange(self, t, x, xdot, out=None):
        """IFunction including the ghost exchange of x and xdot, hidden behind the interior cells; one
        library call (fdts_ifunction_exchange)"""
        self._chk_vec(x, self.num_cols, "x")
        self._chk_vec(xdot, self.num_cols, "xdot")
        if out is None:
            out = torch.empty(self.num_rows, dtype=torch.float64, device=self.device)
        self._chk_vec(out, self.num_rows, "F")
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_ifunction_exchange(self._h, float(t), x.data_ptr(), xdot.data_ptr(), out.data_ptr(),
                                                     _stream()), "fdts_ifunction_exchange")
        return out

    # halo helpers ----------------------------------------------------------------
    def halo_pack(self, x, nodes, buf):
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_halo_pack(self._h, x.data_ptr(), nodes.data_ptr(), nodes.numel(), buf.data_ptr(),
                                            _stream()), "fdts_halo_pack")

    def halo_unpack(self, x, first, count, buf):
        with torch.cuda.device(self.device):
            _check(self._lib.fdts_halo_unpack(self._h, x.data_ptr(), first, count, buf.data_ptr(), _stream()),
                   "fdts_halo_unpack")
