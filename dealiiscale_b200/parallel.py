"""Host-side bookkeeping of the multi-GPU run: which macro DoFs a rank owns and how the per-point
micro contributions are exchanged.

The reference has no distributed layer; its macro solver reads the micro solutions through raw
pointers (src/elliptic-elliptic/manager.cpp:25-26).  With the macro DoFs sharded over ranks that
hand-off becomes, once per fixed-point iteration,

1. an allgather of the coupling scalars (N doubles; bio: 2N) — every rank owns the contiguous block
   ``[r*B, min((r+1)*B, N))`` with ``B = ceil(N / R)``, blocks are padded to ``B`` so that the
   in-place gather has equal counts;
2. a broadcast of the macro solution from the rank that solved the macro system.

On the GPU both are issued by libmsgpu on the context's stream (``msgpu_coupling``: the fused peer-memory kernel or
``ncclAllGather``; ``msgpu_broadcast_macro``).  This module holds the same layout logic in Python for the hosts that
drive libmsgpu from Python: ``bench.py`` takes its partition from it and checks, after the timed region, that what
every rank received for EVERY macro point equals what the owner computed (``exchange_check``: an independent allgather
through ``torch.distributed``); the ``gloo`` test runs the same code on CPU tensors.
"""
from __future__ import annotations

import numpy as np

from .capi import partition


class MacroSharding:
    def __init__(self, n_total: int, rank: int, world_size: int):
        self.n_total, self.rank, self.world_size = n_total, rank, world_size
        self.block = (n_total + world_size - 1) // world_size
        self.k0, self.k1 = partition(n_total, world_size, rank)

    @property
    def n_local(self) -> int:
        return self.k1 - self.k0

    def local(self, values):
        """This rank's block of a global per-macro-DoF array."""
        return np.asarray(values)[self.k0:self.k1]

    def allgather(self, local_values, group=None, device=None):
        """Gather per-point scalars from all ranks (torch.distributed; ``device``: where the tensors of the
        collective live - "cuda" for the NCCL backend, None / "cpu" for gloo)."""
        import torch
        import torch.distributed as dist

        local_values = np.asarray(local_values, dtype=np.float64)
        assert local_values.shape[0] == self.n_local
        padded = torch.zeros(self.block * self.world_size, dtype=torch.float64, device=device)
        mine = padded[self.rank * self.block:(self.rank + 1) * self.block]
        mine[: self.n_local] = torch.from_numpy(local_values).to(padded.device)
        if self.world_size > 1:
            dist.all_gather_into_tensor(padded, mine.clone(), group=group)
        return padded[: self.n_total].cpu().numpy().copy()

    def check_exchange(self, *received, local, group=None, device=None):
        """``received[f]``: the n_total values this rank holds for field f after the exchange; ``local[f]``: the
        values it computed for its own block.  Returns the number of entries, summed over all ranks, where what a
        rank received differs from what the owner computed (0 = the exchange delivered every value to everybody)."""
        import torch
        import torch.distributed as dist

        bad = 0
        for got, mine in zip(received, local):
            truth = self.allgather(mine, group=group, device=device)
            bad += int((np.asarray(got) != truth).sum())
        if self.world_size > 1:
            t = torch.tensor([bad], dtype=torch.int64, device=device)
            dist.all_reduce(t, group=group)
            bad = int(t.item())
        return bad

    def broadcast(self, values, root: int = 0, group=None):
        """Broadcast the macro solution from the rank that solved the macro system."""
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64).copy())
        if self.world_size > 1:
            dist.broadcast(t, src=root, group=group)
        return t.numpy()
