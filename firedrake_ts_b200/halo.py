"""Ghost-value (forward) halo exchange over torch.distributed send/recv.

Replaces PyOP2's ``global_to_local`` PetscSF broadcast that the reference
triggers when it leaves ``dat.vec_wo`` (firedrake_ts/solving_utils.py:249-252;
SURVEY.md section 5 "Distributed communication backend").  Owned values that are
ghosts elsewhere are packed by a gather kernel (fdts_halo_pack) into one
contiguous buffer per neighbour, sent with NCCL ``isend``/``irecv`` (gloo in
the CPU tests), and land in the contiguous per-neighbour slices of the ghost
tail (scalar spaces: received in place, no unpack kernel).  No reverse
exchange exists: every rank computes all cells that touch its owned rows and
drops ghost rows.
"""
from __future__ import annotations

import torch


class HaloExchange:
    def __init__(self, V, engine=None, group=None, device=None):
        import torch.distributed as dist

        self.V, self.engine, self.group = V, engine, group
        self.rank = V.mesh.partition.rank
        self.dev = torch.device(device) if device is not None else (engine.device if engine is not None else torch.device("cpu"))
        self.nc = V.ncomp
        self.neighbours = sorted(set(V.recv_slices) | set(V.send_lists))
        self.send_idx = {r: V.send_lists[r].to(self.dev) for r in V.send_lists}
        # two buffer sets ("slots") so that the exchanges of u and udot can be in flight together
        self.send_buf = {(k, r): torch.empty(len(v) * self.nc, dtype=torch.float64, device=self.dev)
                         for r, v in self.send_idx.items() for k in (0, 1)}
        self.recv_buf = {}
        if self.nc > 1:
            self.recv_buf = {(k, r): torch.empty(c * self.nc, dtype=torch.float64, device=self.dev)
                             for r, (s, c) in V.recv_slices.items() for k in (0, 1)}
        self._dist = dist
        self.bytes_per_exchange = 8 * self.nc * sum(len(v) for v in self.send_idx.values())

    def _global_rank(self, r):
        return self._dist.get_global_rank(self.group, r) if self.group is not None else r

    # -- pieces (so that the caller can overlap interior work) -----------------
    def start(self, x_local: torch.Tensor, slot: int = 0):
        """pack + post sends/recvs; returns the list of pending requests.  Exchanges that are in
        flight at the same time must use different buffer ``slot``s (0 or 1)."""
        dist, V, nc = self._dist, self.V, self.nc
        ops = []
        for r in self.neighbours:
            if r in self.send_idx:
                idx, buf = self.send_idx[r], self.send_buf[(slot, r)]
                if x_local.is_cuda and self.engine is not None:
                    self.engine.halo_pack(x_local, idx, buf)
                else:  # CPU tensors: host-side logic tests under gloo
                    li = idx.to(torch.int64)
                    if V.layout == "interleaved":
                        buf.view(nc, -1).copy_(x_local.view(-1, nc)[li].t())
                    else:
                        buf.view(nc, -1).copy_(x_local.view(nc, -1)[:, li])
                ops.append(dist.P2POp(dist.isend, buf, self._global_rank(r), group=self.group))
            if r in V.recv_slices:
                s, c = V.recv_slices[r]
                tgt = x_local[s:s + c] if nc == 1 else self.recv_buf[(slot, r)]
                ops.append(dist.P2POp(dist.irecv, tgt, self._global_rank(r), group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def finish(self, x_local: torch.Tensor, reqs, slot: int = 0):
        for q in reqs:
            q.wait()
        V, nc = self.V, self.nc
        if nc > 1:
            for r, (s, c) in V.recv_slices.items():
                buf = self.recv_buf[(slot, r)]
                if x_local.is_cuda and self.engine is not None:
                    self.engine.halo_unpack(x_local, s, c, buf)
                elif V.layout == "interleaved":
                    x_local.view(-1, nc)[s:s + c].copy_(buf.view(nc, c).t())
                else:
                    x_local.view(nc, -1)[:, s:s + c].copy_(buf.view(nc, c))

    def forward(self, x_local: torch.Tensor):
        self.finish(x_local, self.start(x_local))

    def owned_to_local(self, x_owned: torch.Tensor):
        """local (owned+ghost) copy of an owned vector with fresh ghost values."""
        V = self.V
        loc = torch.empty(V.num_dofs, dtype=torch.float64, device=x_owned.device)
        if V.layout == "interleaved":
            loc[: V.num_owned_dofs].copy_(x_owned)
        else:
            loc.view(V.ncomp, V.num_nodes)[:, : V.num_owned_nodes].copy_(x_owned.view(V.ncomp, V.num_owned_nodes))
        self.forward(loc)
        return loc
