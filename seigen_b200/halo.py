"""One-layer DG halo exchange between the ranks of one node (torch.distributed: NCCL on GPUs, gloo in CPU tests).

This is the explicit form of what PyOP2 does implicitly around every par_loop that reads a Dat through a map
(SURVEY.md sections 2a, 5): after each of the six passes of a time step the cells next to a partition cut send
the field that pass produced to the ranks that see them across a facet.  Whole cells travel (K = comps*nd
doubles per cell), packed in the order fixed by ``layout.build_rank_plan``.

The exchanger is independent of where the field lives: ``pack(which) -> send tensor`` and
``unpack(which, recv tensor)`` are supplied by the caller (device kernels via the C ABI in ``elastic.py``;
NumPy indexing in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["HaloExchanger", "connect_peers"]


def connect_peers(dev, plan, group=None):
    """Wire the peer-memory exchange of ``include/seigen_b200.h`` (sg_ipc_export / sg_peer_connect): every rank
    publishes the CUDA-IPC handles of its four fields and its control words together with where each neighbour's
    cells land in its halo; torch.distributed is only the bootstrap channel (any backend)."""
    import ctypes as C

    from . import capi
    from .capi import check, lib

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = (C.c_ubyte * (5 * 64))()
    check(lib.sg_ipc_export(dev.handle, raw))
    mine = {"handles": bytes(raw), "n_owned": int(plan.n_owned), "recv": {int(q): tuple(v) for q, v in plan.recv.items()},
            "peers": sorted(int(q) for q in plan.recv)}
    infos = [None] * world
    dist.all_gather_object(infos, mine, group=group)
    tile = lib.sg_tile_cells(dev.dim, dev.degree)
    peers = mine["peers"]
    descs = (capi.PeerDesc * max(len(peers), 1))()
    for i, q in enumerate(peers):
        qi = infos[q]
        first, count = qi["recv"][rank]
        so, sn = plan.send_offsets[q]
        if sn != count:
            raise capi.SgError(f"halo plan mismatch between ranks {rank} and {q}: send {sn} cells, peer expects {count}")
        d = descs[i]
        d.rank, d.flag_slot = q, qi["peers"].index(rank)
        d.send_offset, d.send_count = so, sn
        d.remote_first_cell = -(-qi["n_owned"] // tile) * tile + first
        C.memmove(d.handles, qi["handles"], 5 * 64)
    check(lib.sg_peer_connect(dev.handle, len(peers), descs))
    dist.barrier(group=group)


class HaloExchanger:
    def __init__(self, plan, max_k: int, device, group=None):
        self.plan = plan
        self.group = group
        self.device = torch.device(device)
        self.peers = sorted(plan.recv)
        nsend = int(len(plan.send_cells)) if plan.send_cells is not None else 0
        self.nsend = nsend
        self.nrecv = plan.n_halo
        self.sendbuf = torch.empty(max(nsend, 1) * max_k, dtype=torch.float64, device=self.device)
        self.recvbuf = torch.empty(max(self.nrecv, 1) * max_k, dtype=torch.float64, device=self.device)

    @property
    def active(self):
        return len(self.peers) > 0

    def views(self, K):
        """Per-peer (send view, recv view) of the flat buffers for a field with K doubles per cell."""
        out = []
        for q in self.peers:
            so, sn = self.plan.send_offsets[q]
            ro, rn = self.plan.recv[q]
            out.append((q, self.sendbuf[so * K:(so + sn) * K], self.recvbuf[ro * K:(ro + rn) * K]))
        return out

    def exchange(self, K):
        """Post all sends/receives for the packed buffers; returns after the transfers are ordered on the current
        stream (NCCL) or complete (gloo)."""
        if not self.active:
            return
        ops = []
        for q, sv, rv in self.views(K):
            ops.append(dist.P2POp(dist.isend, sv, q, group=self.group))
            ops.append(dist.P2POp(dist.irecv, rv, q, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
