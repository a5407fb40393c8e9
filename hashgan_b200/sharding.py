"""Multi-GPU evaluation; the reference is single-GPU (main.py:263 only sets
CUDA_VISIBLE_DEVICES), so this layer is new.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch; gloo on CPU for the tests):
  1. the database is row-sharded: rank g holds rows [start_g, start_g + n_g) and packs them locally;
  2. ONE all-gather of the packed rows (code words + label words in one tensor) -- rank order is the
     global database row order, which preserves the (distance, row) tie rule;
  3. queries are row-sharded: each rank ranks its own queries against the full packed database;
  4. the per-query APs are all-gathered so every rank computes the same mean in the same order
     (lib/metric.py:24).
The host logic here is device-agnostic: the packers / ranker are injected, so the gloo tests can drive it
with CPU tensors.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["row_shard", "shard_bounds", "gather_rows", "gather_vector", "ShardedMAPs"]


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row ranges: the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def row_shard(n: int, rank: int, world: int) -> Tuple[int, int]:
    return shard_bounds(n, world)[rank]


def _dist():
    import torch.distributed as dist

    return dist


def gather_rows(local, group=None):
    """All-gather of row blocks with possibly different row counts.  local: [n_r, ...] tensor (any device the
    group's backend supports).  Returns the [sum n_r, ...] concatenation in rank order plus the row counts."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    if world == 1:
        return local, [int(local.shape[0])]
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts_t = torch.empty((world,), dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts_t, n_local, group=group)
    counts = [int(x) for x in counts_t.cpu().tolist()]
    n_max = max(counts)
    tail = tuple(local.shape[1:])
    if all(c == n_max for c in counts):
        out = torch.empty((world * n_max,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out, counts
    padded = torch.zeros((n_max,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world, n_max) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view((world * n_max,) + tail), padded, group=group)
    return torch.cat([buf[r, : counts[r]] for r in range(world)], 0), counts


def gather_vector(local, group=None):
    out, _ = gather_rows(local.reshape(-1, 1), group)
    return out.reshape(-1)


class ShardedMAPs:
    """mAP@R over a process group.  ``get_maps_by_feature(database_shard, query_shard)`` takes THIS rank's
    contiguous row blocks (rank order == global row order) and returns the global mAP on every rank."""

    def __init__(self, r: int, group=None, *, device=None, flags: int = 0,
                 pack_rows: Optional[Callable] = None, rank_fn: Optional[Callable] = None):
        self.R = r
        self.group = group
        self.device = device
        self.flags = flags
        self._pack_rows = pack_rows
        self._rank_fn = rank_fn
        self.last_counts: Sequence[int] = ()

    def _hooks(self):
        from . import metric

        pack = self._pack_rows or (lambda out, lab: metric.pack_rows(out, lab, self.device))

        def rank(q_rows, db_rows, b, L, R):
            ap, _, _, _ = metric.hamming_map_device(q_rows, db_rows, b, L, R, flags=self.flags)
            return ap

        return pack, (self._rank_fn or rank)

    def per_query_ap_device(self, database, query):
        """Global per-query AP vector (rank order) as a tensor on the compute device."""
        pack, rank = self._hooks()
        b = int(database.output.shape[1])
        L = int(database.label.shape[1])
        db_rows_local = pack(database.output, database.label)
        q_rows = pack(query.output, query.label)
        db_rows, counts = gather_rows(db_rows_local, self.group)   # the ONE exchange step: packed code + label words
        self.last_counts = counts
        ndb = int(db_rows.shape[0])
        if self.R > ndb:
            raise ValueError(f"operands could not be broadcast together: R={self.R} exceeds the database size {ndb}")
        ap_local = rank(q_rows, db_rows, b, L, int(self.R))
        return gather_vector(ap_local, self.group)

    def get_maps_by_feature(self, database, query):
        ap = self.per_query_ap_device(database, query).cpu().numpy()
        kept = ap[~np.isnan(ap)]
        return np.mean(kept)  # lib/metric.py:24, identical on every rank
