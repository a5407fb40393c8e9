"""Multi-GPU evaluation; the reference is single-GPU (main.py:263 only sets
CUDA_VISIBLE_DEVICES), so this layer is new.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch; gloo on CPU for the tests):
  1. the database is row-sharded: rank g holds rows [start_g, start_g + n_g) and packs them locally;
  2. ONE exchange of the packed rows (code words + label words in one tensor) -- rank order is the global database
     row order, which preserves the (distance, row) tie rule.  On CUDA the exchange is fused into the pack kernel
     (SymmetricRows: peer-memory stores into every rank's buffer + one barrier); otherwise one all-gather;
  3. queries are row-sharded: each rank ranks its own queries against the full packed database;
  4. the per-query APs are all-gathered so every rank computes the same mean in the same order
     (lib/metric.py:24).
The host logic here is device-agnostic: the packers / ranker are injected, so the gloo tests can drive it
with CPU tensors.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["row_shard", "shard_bounds", "gather_rows", "gather_vector", "ShardedMAPs", "SymmetricRows"]


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


def gather_rows(local, group=None, counts: Optional[Sequence[int]] = None):
    """All-gather of row blocks with possibly different row counts.  local: [n_r, ...] tensor (any device the
    group's backend supports).  Returns the [sum n_r, ...] concatenation in rank order plus the row counts.
    `counts` (rows of every rank, e.g. from shard_bounds) skips the count exchange and its host synchronisation,
    which leaves ONE collective on the stream and lets the host run ahead of the GPU."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    if world == 1:
        return local, [int(local.shape[0])]
    if counts is not None:
        counts = [int(c) for c in counts]
        if len(counts) != world or counts[dist.get_rank(group)] != int(local.shape[0]):
            raise ValueError(f"counts {counts} do not describe this rank's block of {int(local.shape[0])} rows")
    else:
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


def gather_vector(local, group=None, counts: Optional[Sequence[int]] = None):
    out, _ = gather_rows(local.reshape(-1, 1), group, counts)
    return out.reshape(-1)


class SymmetricRows:
    """The exchange step fused into the pack kernel (C ABI: hg_pack_rows_push): every rank owns a full-size packed
    database buffer in SYMMETRIC memory (torch.distributed._symmetric_memory: same allocation on every GPU, peer
    pointers over NVLink / NVSwitch); the pack kernel stores each packed row into all of them, one signal-pad barrier
    publishes the rows.  No NCCL collective, no extra pass over the packed rows.

    Two buffers alternate: a rank may already pack step t+1 into its peers while they still rank step t.  A rank's
    writes into buffer i of step t+2 happen after barrier t+1, which every peer enters only after its step-t ranking
    (the last reader of buffer i) has finished on its stream."""

    def __init__(self, ndb: int, b: int, L: int, device, group=None):
        import torch
        import torch.distributed._symmetric_memory as symm

        from . import _native

        dist = _dist()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.ndb, self.b, self.L, self.Wr = int(ndb), int(b), int(L), _native.row_words(b, L)
        self.device = torch.device(device)
        self.lib = _native.lib()
        self.bufs, self.handles = [], []
        with torch.cuda.device(self.device):
            for _ in range(2):
                t = symm.empty((self.ndb, self.Wr), dtype=torch.int32, device=self.device)
                self.handles.append(symm.rendezvous(t, self.group))
                self.bufs.append(t)
        self.step = 0

    def pack(self, feat, lab, row_lo: int, bad_flag=None, expect_rows: Optional[int] = None):
        """Packs this rank's rows [row_lo, row_lo + n) into every rank's buffer; returns this rank's full buffer,
        valid on the current stream once the call returns (kernel + barrier are enqueued).  `expect_rows`: the row count
        the other ranks assume for this rank (db_counts[rank]) -- a different block would leave rows of the symmetric
        buffer unwritten or overlapping a neighbour's, so it raises instead."""
        import ctypes as C

        import torch

        from . import _native, metric

        i = self.step & 1
        self.step += 1
        h = self.handles[i]
        with torch.cuda.device(self.device):
            f = metric._features_to_device(torch, feat, self.device)
            t, nbytes = metric._labels_to_device(torch, lab, self.device)
            n = int(f.shape[0])
            if expect_rows is not None and n != int(expect_rows):
                raise ValueError(f"rank {self.rank} holds {n} database rows but db_counts says {int(expect_rows)}")
            if int(f.shape[1]) != self.b or int(t.shape[1]) != self.L or t.shape[0] != n or row_lo + n > self.ndb:
                raise ValueError("block does not fit the symmetric database buffer")
            off = int(row_lo) * self.Wr * 4
            ptrs = (C.c_void_p * self.world)(*[int(p) + off for p in h.buffer_ptrs])
            stream = torch.cuda.current_stream(self.device)
            _native.check(self.lib.hg_pack_rows_push(f.data_ptr(), self.b, t.data_ptr(), nbytes, n, self.b, self.L, ptrs, self.world,
                                                     bad_flag.data_ptr() if bad_flag is not None else None, int(stream.cuda_stream)))
            f.record_stream(stream)
            t.record_stream(stream)
            h.barrier(channel=0)
        return self.bufs[i]


class ShardedMAPs:
    """mAP@R over a process group.  ``get_maps_by_feature(database_shard, query_shard)`` takes THIS rank's
    contiguous row blocks (rank order == global row order) and returns the global mAP on every rank.
    ``binarize=False`` ranks the raw features by inner product (lib/metric.py:13-14 literally; EVAL.BINARIZE False): the
    database features are all-gathered next to the packed label rows and every rank runs hg_ip_map on its queries."""

    def __init__(self, r: int, group=None, *, device=None, flags: int = 0,
                 pack_rows: Optional[Callable] = None, rank_fn: Optional[Callable] = None,
                 db_counts: Optional[Sequence[int]] = None, query_counts: Optional[Sequence[int]] = None, symmetric: bool = False,
                 binarize: bool = True, rank_real_fn: Optional[Callable] = None):
        # symmetric: fuse the exchange into the pack kernel (SymmetricRows; needs db_counts and the CUDA packers)
        self.symmetric = bool(symmetric)
        self._sym = None
        # db_counts / query_counts: rows per rank when the caller knows them (shard_bounds): no count exchange, no host sync
        self.db_counts, self.query_counts = db_counts, query_counts
        self.R = r
        self.group = group
        self.device = device
        self.flags = flags
        self.binarize = bool(binarize)
        self._pack_rows = pack_rows
        self._rank_fn = rank_fn
        self._rank_real_fn = rank_real_fn
        self.last_counts: Sequence[int] = ()
        self.exchange = "none"

    def _hooks(self):
        from . import metric

        pack = self._pack_rows or (lambda out, lab, bad=None: metric.pack_rows(out, lab, self.device, bad))

        def rank(q_rows, db_rows, b, L, R):
            ap, _, _, _ = metric.hamming_map_device(q_rows, db_rows, b, L, R, flags=self.flags)
            return ap

        def rank_real(q_feat, q_rows, db_feat, db_rows, b, L, R):
            ap, _, _, _ = metric.ip_map_device(q_feat, q_rows, db_feat, db_rows, b, L, R)
            return ap

        return pack, (self._rank_fn or rank), (self._rank_real_fn or rank_real)

    def per_query_ap_device(self, database, query):
        """Global per-query AP vector (rank order) as a tensor on the compute device."""
        pack, rank, rank_real = self._hooks()
        dist = _dist()
        b = int(database.output.shape[1])
        L = int(database.label.shape[1])
        injected = self._pack_rows is not None
        if injected:
            q_rows = pack(query.output, query.label)
            bad = None
        else:
            from . import metric

            torch = metric._torch()
            device = metric._require_cuda(torch, self.device)
            bad = torch.zeros((1,), dtype=torch.int32, device=device)
            q_rows = pack(query.output, query.label, bad)
        my_rank = dist.get_rank(self.group)
        db_rows = None
        if self.symmetric and self.db_counts is not None and not injected and b % 32 == 0:
            counts = [int(c) for c in self.db_counts]
            ndb_all = sum(counts)
            try:
                if self._sym is None or (self._sym.ndb, self._sym.b, self._sym.L) != (ndb_all, b, L):
                    self._sym = SymmetricRows(ndb_all, b, L, q_rows.device, self.group)
            except (ImportError, RuntimeError, AttributeError) as exc:  # no symmetric memory / no P2P on this box: all-gather instead
                import warnings

                warnings.warn(f"symmetric memory unavailable ({exc}); exchanging the packed rows with an all-gather")
                self.symmetric, self._sym = False, None
            if self._sym is not None:
                db_rows = self._sym.pack(database.output, database.label, sum(counts[:my_rank]), bad, expect_rows=counts[my_rank])  # pack + exchange in one kernel
                self.exchange = "push"
        if db_rows is None:
            db_rows_local = pack(database.output, database.label) if injected else pack(database.output, database.label, bad)
            db_rows, counts = gather_rows(db_rows_local, self.group, self.db_counts)   # the ONE exchange step: packed code + label words
            self.exchange = "all-gather"
        self.last_counts = counts
        ndb = int(db_rows.shape[0])
        if self.R > ndb:
            raise ValueError(f"operands could not be broadcast together: R={self.R} exceeds the database size {ndb}")
        if self.binarize:
            ap_local = rank(q_rows, db_rows, b, L, int(self.R))
        else:
            if injected:
                import torch as _t

                as_t = lambda x: x if isinstance(x, _t.Tensor) else _t.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))  # noqa: E731
                db_feat_local, q_feat = as_t(database.output), as_t(query.output)
            else:
                db_feat_local = metric._features_to_device(torch, database.output, device)
                q_feat = metric._features_to_device(torch, query.output, device)
            db_feat, _ = gather_rows(db_feat_local, self.group, counts)   # the raw features travel too (ndb x b float32)
            ap_local = rank_real(q_feat, q_rows, db_feat, db_rows, b, L, int(self.R))
        out = gather_vector(ap_local, self.group, self.query_counts)
        if bad is not None:
            # a label that is not 0/1 on ANY rank fails the call on EVERY rank (MAPs raises ValueError for the same input)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
            if int(bad.item()) != 0:
                raise ValueError("labels must be 0/1 integers (lib/metric.py:17-19 is only defined for 0/1 labels)")
        return out

    def get_maps_by_feature(self, database, query):
        ap = self.per_query_ap_device(database, query).cpu().numpy()
        kept = ap[~np.isnan(ap)]
        return np.mean(kept)  # lib/metric.py:24, identical on every rank
