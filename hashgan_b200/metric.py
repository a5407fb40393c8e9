"""mAP@R by Hamming ranking -- the drop-in for the reference's ``lib/metric.py``.

Reference surface kept (thuml/HashGAN):
    lib/metric.py:4-6    MAPs.__init__(self, r) -> self.R
    lib/metric.py:8-10   MAPs.distance(a, b)            (dead code in the reference; kept)
    lib/metric.py:12-24  MAPs.get_maps_by_feature(database, query) -> numpy.float64
    main.py:164          MAPs(cfg.DATA.MAP_R).get_maps_by_feature(db, test)

``database`` / ``query`` are any objects with ``.output`` ([N, b] features) and ``.label`` ([N, L] 0/1
integers), e.g. the EasyDict of main.py:157 or a SimpleNamespace.  Arrays may be NumPy arrays or
torch tensors (host or CUDA).  The work is done by hand-written sm_100a kernels behind the C ABI of
``include/hashgan_b200.h``; torch only owns device memory and streams.  There is no CPU path.

Semantics relative to the reference:
  * features are binarised by sign (bit = x > 0); on {-1,+1} inputs the reference's inner-product order
    (lib/metric.py:13-14) is exactly the Hamming order computed here (ip = b - 2 d_H);
  * ties are broken by database row (== ``np.argsort(kind='stable')``); the reference's default argsort
    leaves tie order to the NumPy build;
  * queries with no relevant row in their top-R are skipped (lib/metric.py:22-23), all skipped -> nan;
  * R > Ndb raises ValueError (the reference fails in the broadcast of lib/metric.py:21).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _native

__all__ = ["MAPs", "MAPs_CQ", "pack_rows", "hamming_map_device", "ip_map_device", "relevant_totals_device"]

# One call of hg_hamming_map handles a query chunk whose workspace stays under this many bytes.
DEFAULT_WORKSPACE_LIMIT = 24 << 30


def _torch():
    import torch  # torch is the device-memory / stream / process-group plumbing

    return torch


def _require_cuda(torch, device):
    if not torch.cuda.is_available():
        raise _native.NativeLibraryError(
            "hashgan_b200 needs a CUDA device (sm_100a); there is no CPU fallback for the metric path")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("device must be a CUDA device")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _stream_ptr(torch, device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _features_to_device(torch, x, device):
    """[N, b] features -> contiguous float32 CUDA tensor (non-blocking copy when the source is pinned)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dim() != 2:
        raise ValueError(f"features must be 2-D [N, b], got shape {tuple(t.shape)}")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _labels_to_device(torch, x, device):
    """[N, L] 0/1 labels -> contiguous CUDA tensor of int64 / int32 / int8 (as given) and its item size."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype == np.bool_:
            a = a.astype(np.int8)
        elif a.dtype.kind not in "iu":
            raise TypeError(f"labels must be an integer 0/1 matrix, got dtype {a.dtype}")
        elif a.dtype.itemsize == 2:
            a = a.astype(np.int32)
        elif a.dtype.kind == "u" and a.dtype.itemsize in (4, 8):
            a = a.astype(np.int64)
        elif a.dtype == np.uint8:
            a = a.view(np.int8)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dim() != 2:
        raise ValueError(f"labels must be 2-D [N, L], got shape {tuple(t.shape)}")
    if t.dtype == torch.bool:
        t = t.to(torch.int8)
    elif t.dtype == torch.uint8:
        t = t.view(torch.int8)
    elif t.dtype == torch.int16:
        t = t.to(torch.int32)
    elif t.dtype not in (torch.int64, torch.int32, torch.int8):
        raise TypeError(f"labels must be an integer 0/1 matrix, got dtype {t.dtype}")
    if t.device != device:
        t = t.to(device, non_blocking=True)
    t = t.contiguous()
    return t, t.element_size()


def pack_rows(feat, lab=None, device=None, bad_flag=None, L: Optional[int] = None):
    """sign + bit-pack [N, b] features and (optionally) [N, L] 0/1 labels on the GPU into packed rows
    [ code words | label words | pad ] -- an int32 CUDA tensor [N, hg_row_words(b, L)] (C ABI: hg_pack_rows).
    ``bad_flag`` (int32 CUDA tensor [1]) is OR-ed with 1 when a label is not 0/1."""
    torch = _torch()
    device = _require_cuda(torch, device)
    lib = _native.lib()
    with torch.cuda.device(device):
        f = _features_to_device(torch, feat, device)
        n, b = f.shape
        if lab is not None:
            t, nbytes = _labels_to_device(torch, lab, device)
            if t.shape[0] != n:
                raise ValueError(f"output has {n} rows but label has {t.shape[0]}")
            L = int(t.shape[1])
        else:
            t, nbytes, L = None, 8, int(L or 1)
        Wr = _native.row_words(b, L)
        rows = torch.empty((n, Wr), dtype=torch.int32, device=device)
        stream = torch.cuda.current_stream(device)
        _native.check(lib.hg_pack_rows(f.data_ptr(), b, t.data_ptr() if t is not None else None, nbytes, n, b, L, rows.data_ptr(),
                                       bad_flag.data_ptr() if bad_flag is not None else None, int(stream.cuda_stream)))
        f.record_stream(stream)
        if t is not None:
            t.record_stream(stream)
    return rows


def _query_chunk(nq: int, ndb: int, b: int, L: int, R: int, limit: int) -> int:
    """Largest query count whose hg_hamming_map workspace fits in `limit` bytes."""
    lib = _native.lib()
    def fits(n):  # 0 = the plan refuses the batch (list area beyond 32-bit offsets): treat as "too big"
        need = lib.hg_hamming_map_workspace_bytes(n, ndb, b, L, R)
        return 0 < need <= limit

    if lib.hg_hamming_map_workspace_bytes(1, ndb, b, L, R) == 0:
        raise ValueError(f"sizes out of range for the Hamming kernel: ndb={ndb} b={b} L={L} R={R}")
    if fits(nq):
        return nq
    lo, hi = 1, nq
    while lo < hi:  # the workspace grows with nq
        mid = (lo + hi + 1) // 2
        if fits(mid):
            lo = mid
        else:
            hi = mid - 1
    if not fits(lo):
        raise MemoryError(f"workspace limit {limit} B is too small for a single query against ndb={ndb}")
    return lo


def hamming_map_device(q_rows, db_rows, b: int, L: int, R: int, *, flags: int = 0,
                       want_ids: bool = False, want_rel: bool = False,
                       workspace_limit: int = DEFAULT_WORKSPACE_LIMIT, stats: Optional[dict] = None):
    """Per-query AP@R from packed-row device tensors (C ABI: hg_hamming_map).  Returns (ap, ids, dist, rel);
    ids/dist/rel are None unless requested.  Everything stays on the device and on the current stream."""
    torch = _torch()
    lib = _native.lib()
    device = db_rows.device
    nq, ndb = int(q_rows.shape[0]), int(db_rows.shape[0])
    Wr = _native.row_words(b, L)
    if q_rows.shape[1] != Wr or db_rows.shape[1] != Wr or not q_rows.is_contiguous() or not db_rows.is_contiguous():
        raise ValueError(f"packed rows must be contiguous [N, {Wr}] for b={b}, L={L}")
    if R > ndb:
        raise ValueError(f"operands could not be broadcast together: R={R} exceeds the database size {ndb}")
    if R <= 0:
        raise ValueError("R must be positive")
    with torch.cuda.device(device):
        ap = torch.empty((nq,), dtype=torch.float64, device=device)
        ids = torch.empty((nq, R), dtype=torch.int32, device=device) if want_ids else None
        dist = torch.empty((nq, R), dtype=torch.int16, device=device) if want_ids else None
        rel = torch.empty((nq,), dtype=torch.int32, device=device) if want_rel else None
        if nq == 0:
            return ap, ids, dist, rel
        chunk = _query_chunk(nq, ndb, b, L, R, workspace_limit)
        ws_bytes = lib.hg_hamming_map_workspace_bytes(chunk, ndb, b, L, R)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=device)
        stream = _stream_ptr(torch, device)
        for s in range(0, nq, chunk):
            n = min(chunk, nq - s)
            _native.check(lib.hg_hamming_map(
                q_rows[s:s + n].data_ptr(), n, db_rows.data_ptr(), ndb, b, L, R, flags,
                ap[s:s + n].data_ptr(),
                ids[s:s + n].data_ptr() if ids is not None else None,
                dist[s:s + n].data_ptr() if dist is not None else None,
                rel[s:s + n].data_ptr() if rel is not None else None,
                ws.data_ptr(), ws_bytes, stream))
            if stats is not None:
                out = (C.c_int64 * 8)()
                _native.check(lib.hg_hamming_map_stats(ws.data_ptr(), ws_bytes, n, ndb, b, L, R, out, stream))
                stats.setdefault("chunks", []).append(
                    dict(exact_queries=int(out[0]), splits=int(out[1]), rows_per_split=int(out[2]), bin_entries=int(out[3]),
                         queries_per_cta=int(out[4]), sample_rows=int(out[5]), exact_failures=int(out[6]),
                         wide_queries=int(out[7]), nq=n))
    return ap, ids, dist, rel


def ip_map_device(q_feat, q_rows, db_feat, db_rows, b: int, L: int, R: int, *, want_ids: bool = False, want_rel: bool = False,
                  workspace_limit: int = DEFAULT_WORKSPACE_LIMIT):
    """Per-query AP@R by REAL-VALUED inner-product ranking (C ABI: hg_ip_map; lib/metric.py:13-23 without binarisation).
    q_feat / db_feat: float32 CUDA tensors [N, b]; q_rows / db_rows: their packed rows (label words).  Returns
    (ap, ids, ips, rel); ids / ips / rel are None unless requested."""
    torch = _torch()
    lib = _native.lib()
    device = db_feat.device
    nq, ndb = int(q_feat.shape[0]), int(db_feat.shape[0])
    if R > ndb:
        raise ValueError(f"operands could not be broadcast together: R={R} exceeds the database size {ndb}")
    if R <= 0:
        raise ValueError("R must be positive")
    with torch.cuda.device(device):
        ap = torch.empty((nq,), dtype=torch.float64, device=device)
        ids = torch.empty((nq, R), dtype=torch.int32, device=device) if want_ids else None
        ips = torch.empty((nq, R), dtype=torch.float32, device=device) if want_ids else None
        rel = torch.empty((nq,), dtype=torch.int32, device=device) if want_rel else None
        if nq == 0:
            return ap, ids, ips, rel
        need = lib.hg_ip_map_workspace_bytes(nq, ndb, b, L, R)
        if need == 0:
            raise ValueError(f"sizes out of range for the inner-product ranking: ndb={ndb} b={b} L={L} R={R}")
        one = lib.hg_ip_map_workspace_bytes(1, ndb, b, L, R)
        ws_bytes = max(one, min(need, workspace_limit))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=device)
        _native.check(lib.hg_ip_map(q_feat.data_ptr(), q_rows.data_ptr(), nq, db_feat.data_ptr(), db_rows.data_ptr(), ndb, b, L, R,
                                    ap.data_ptr(), ids.data_ptr() if ids is not None else None, ips.data_ptr() if ips is not None else None,
                                    rel.data_ptr() if rel is not None else None, ws.data_ptr(), ws_bytes, _stream_ptr(torch, device)))
    return ap, ids, ips, rel


def relevant_totals_device(q_rows, db_rows, b: int, L: int):
    """Relevant rows of the whole database per query (C ABI: hg_relevant_totals): int32 CUDA tensor [Nq]."""
    torch = _torch()
    lib = _native.lib()
    device = db_rows.device
    with torch.cuda.device(device):
        total = torch.empty((int(q_rows.shape[0]),), dtype=torch.int32, device=device)
        _native.check(lib.hg_relevant_totals(q_rows.data_ptr(), int(q_rows.shape[0]), db_rows.data_ptr(), int(db_rows.shape[0]), b, L,
                                             total.data_ptr(), _stream_ptr(torch, device)))
    return total


class _Record:
    __slots__ = ("output", "label")

    def __init__(self, output, label):
        self.output, self.label = output, label


def _as_record(x):
    """Accepts the reference's EasyDict / any object with .output and .label; list-like fields become arrays."""
    out, lab = x.output, x.label
    if not hasattr(out, "shape"):
        out = np.asarray(out)
    if not hasattr(lab, "shape"):
        lab = np.asarray(lab)
    if len(out.shape) != 2 or len(lab.shape) != 2:
        raise ValueError("output must be [N, b] and label [N, L]")
    return _Record(out, lab)


class MAPs:
    """Drop-in for ``lib.metric.MAPs`` (lib/metric.py:4-24)."""

    def __init__(self, r, *, device=None, flags: int = 0, workspace_limit: int = DEFAULT_WORKSPACE_LIMIT, binarize: bool = True):
        self.R = r
        self.device = device
        # binarize=False: rank by the real-valued inner products exactly as lib/metric.py:13-14 does (config: EVAL.BINARIZE)
        self.binarize = binarize
        self.flags = flags
        self.workspace_limit = workspace_limit
        self.collect_stats = False
        self.last_stats = None

    @staticmethod
    def distance(a, b):
        # lib/metric.py:8-10 (unused by the reference itself)
        return np.dot(a, b)

    def _pack_all(self, database, query):
        torch = _torch()
        device = _require_cuda(torch, self.device)
        with torch.cuda.device(device):
            bad = torch.zeros((1,), dtype=torch.int32, device=device)
            b, bq = int(database.output.shape[1]), int(query.output.shape[1])
            L, Lq = int(database.label.shape[1]), int(query.label.shape[1])
            if b != bq:
                raise ValueError(f"shapes {tuple(query.output.shape)} and {tuple(database.output.shape)} not aligned: hash lengths differ")
            if L != Lq:
                raise ValueError(f"label widths differ: database {L}, query {Lq}")
            db_rows = pack_rows(database.output, database.label, device, bad)
            q_rows = pack_rows(query.output, query.label, device, bad)
        return device, bad, db_rows, q_rows, b, L

    def _host_call(self, database, query):
        """Host arrays in -> per-query AP out through ONE C call (hg_maps_by_feature_host): chunked H2D on a copy
        stream overlapped with packing and ranking.  Returns None when the inputs are not plain host buffers (CUDA
        tensors) or the batch needs query chunking; the caller then takes the device path."""
        torch = _torch()
        device = _require_cuda(torch, self.device)

        def host_view(x, kinds):
            if isinstance(x, torch.Tensor):
                if x.device.type != "cpu" or not x.is_contiguous():
                    return None
                x = x.numpy()  # shares memory (pinned stays pinned)
            a = np.asarray(x)
            if "f4" not in kinds and a.dtype.kind not in "iub":
                raise TypeError(f"labels must be an integer 0/1 matrix, got dtype {a.dtype}")
            if a.dtype == np.bool_:
                a = a.view(np.int8)
            elif a.dtype == np.uint8 and "u1" in kinds:
                a = a.view(np.int8)
            if a.dtype.str[1:] not in kinds:
                a = a.astype(np.float32 if "f4" in kinds else np.int64)
            return np.ascontiguousarray(a)

        arrs = [host_view(database.output, ("f4",)), host_view(database.label, ("i8", "i4", "i1", "u1")),
                host_view(query.output, ("f4",)), host_view(query.label, ("i8", "i4", "i1", "u1"))]
        if any(a is None for a in arrs):
            return None
        db_f, db_l, q_f, q_l = arrs
        if db_f.ndim != 2 or q_f.ndim != 2 or db_l.ndim != 2 or q_l.ndim != 2:
            raise ValueError("output must be [N, b] and label [N, L]")
        if db_f.shape[1] != q_f.shape[1]:
            raise ValueError(f"shapes {q_f.shape} and {db_f.shape} not aligned: hash lengths differ")
        if db_l.shape[1] != q_l.shape[1]:
            raise ValueError(f"label widths differ: database {db_l.shape[1]}, query {q_l.shape[1]}")
        if db_f.shape[0] != db_l.shape[0] or q_f.shape[0] != q_l.shape[0]:
            raise ValueError("output and label row counts differ")
        if q_l.dtype != db_l.dtype:
            q_l, db_l = q_l.astype(np.int64), db_l.astype(np.int64)
        nq, ndb, b, L, R = q_f.shape[0], db_f.shape[0], db_f.shape[1], db_l.shape[1], int(self.R)
        _native.code_words(b)
        _native.label_words(L)
        if R > ndb:
            raise ValueError(f"operands could not be broadcast together: R={R} exceeds the database size {ndb}")
        if R <= 0:
            raise ValueError("R must be positive")
        lib = _native.lib()
        if nq == 0:
            return np.empty((0,), dtype=np.float64)
        need = lib.hg_hamming_map_workspace_bytes(nq, ndb, b, L, R)
        if need == 0 or 2 * need > self.workspace_limit:
            return None  # needs query chunking: device path
        out = C.c_double(0.0)
        ap = np.empty((nq,), dtype=np.float64)
        with torch.cuda.device(device):
            rc = lib.hg_maps_by_feature_host(db_f.ctypes.data, db_l.ctypes.data, ndb, q_f.ctypes.data, q_l.ctypes.data, nq,
                                             b, L, db_l.dtype.itemsize, R, self.flags, C.byref(out), ap.ctypes.data)
        if rc == _native.HG_ELABEL:
            raise ValueError("labels must be 0/1 integers (lib/metric.py:17-19 is only defined for 0/1 labels)")
        _native.check(rc)
        return ap

    def per_query_ap(self, database, query, *, want_ids: bool = False):
        """Per-query AP@R as a NumPy float64 vector (NaN where the reference would skip the query).
        With ``want_ids`` also returns (ids [Nq, R] int64, dist [Nq, R] int32) in rank order."""
        database, query = _as_record(database), _as_record(query)
        if not self.binarize:
            return self._per_query_ap_real(database, query, want_ids)
        if not want_ids and not self.collect_stats:
            host = self._host_call(database, query)
            if host is not None:
                return host
        device, bad, db_rows, q_rows, b, L = self._pack_all(database, query)
        R = int(self.R)
        self.last_stats = {} if self.collect_stats else None
        ap, ids, dist, _ = hamming_map_device(q_rows, db_rows, b, L, R, flags=self.flags, want_ids=want_ids,
                                              workspace_limit=self.workspace_limit, stats=self.last_stats)
        ap_h = ap.cpu().numpy()
        if int(bad.item()) != 0:
            raise ValueError("labels must be 0/1 integers (lib/metric.py:17-19 is only defined for 0/1 labels)")
        if want_ids:
            return ap_h, ids.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, dist.cpu().numpy().astype(np.int32) & 0xFFFF
        return ap_h

    def _per_query_ap_real(self, database, query, want_ids: bool):
        """binarize=False: (ip descending, row ascending) ranking of the raw features; with want_ids returns
        (ap, ids [Nq, R] int64, ips [Nq, R] float32)."""
        torch = _torch()
        device, bad, db_rows, q_rows, b, L = self._pack_all(database, query)
        with torch.cuda.device(device):
            db_f = _features_to_device(torch, database.output, device)
            q_f = _features_to_device(torch, query.output, device)
            ap, ids, ips, _ = ip_map_device(q_f, q_rows, db_f, db_rows, b, L, int(self.R), want_ids=want_ids,
                                            workspace_limit=self.workspace_limit)
            ap_h = ap.cpu().numpy()
        if int(bad.item()) != 0:
            raise ValueError("labels must be 0/1 integers (lib/metric.py:17-19 is only defined for 0/1 labels)")
        if want_ids:
            return ap_h, ids.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, ips.cpu().numpy()
        return ap_h

    def precision_recall(self, database, query):
        """precision@R, recall@R and mAP@R on ONE ranking (the reference reports mAP only).  Per query, with
        rel = relevant rows inside the top-R (lib/metric.py:20) and total = relevant rows in the whole database under the same
        relevance test (lib/metric.py:17-19):  precision = rel / R,  recall = rel / total (NaN when the database holds no
        relevant row for the query).  Returns a dict: 'precision' / 'recall' / 'mAP' are means (recall over the queries that
        have a relevant row at all, mAP as lib/metric.py:22-24), 'per_query' holds the vectors (ap, rel, total)."""
        torch = _torch()
        database, query = _as_record(database), _as_record(query)
        device, bad, db_rows, q_rows, b, L = self._pack_all(database, query)
        R = int(self.R)
        with torch.cuda.device(device):
            if self.binarize:
                ap, _, _, rel = hamming_map_device(q_rows, db_rows, b, L, R, flags=self.flags, want_rel=True, workspace_limit=self.workspace_limit)
            else:
                db_f = _features_to_device(torch, database.output, device)
                q_f = _features_to_device(torch, query.output, device)
                ap, _, _, rel = ip_map_device(q_f, q_rows, db_f, db_rows, b, L, R, want_rel=True, workspace_limit=self.workspace_limit)
            total = relevant_totals_device(q_rows, db_rows, b, L)
            ap_h, rel_h, total_h = ap.cpu().numpy(), rel.cpu().numpy().astype(np.int64), total.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        if int(bad.item()) != 0:
            raise ValueError("labels must be 0/1 integers (lib/metric.py:17-19 is only defined for 0/1 labels)")
        precision = rel_h / float(R)
        with np.errstate(divide="ignore", invalid="ignore"):
            recall = np.where(total_h > 0, rel_h / np.maximum(total_h, 1), np.nan)
        kept = ap_h[~np.isnan(ap_h)]
        has = ~np.isnan(recall)
        return {"precision": np.float64(np.mean(precision)) if len(precision) else np.float64("nan"),
                "recall": np.float64(np.mean(recall[has])) if has.any() else np.float64("nan"),
                "mAP": np.mean(kept) if len(kept) else np.float64("nan"), "R": R,
                "per_query": {"ap": ap_h, "rel": rel_h, "total": total_h, "precision": precision, "recall": recall}}

    def get_maps_by_feature(self, database, query):
        """mAP@R; same call and return type (numpy.float64) as lib/metric.py:12-24."""
        ap = self.per_query_ap(database, query)
        kept = ap[~np.isnan(ap)]
        return np.mean(kept)  # lib/metric.py:24 (nan + RuntimeWarning when every query was skipped)

    def get_maps_by_feature_host(self, database, query, return_ap: bool = False):
        """Same result through the single C call hg_maps_by_feature_host (host pointers in, scalar out)."""
        lib = _native.lib()
        db_f = np.ascontiguousarray(np.asarray(database.output), dtype=np.float32)
        q_f = np.ascontiguousarray(np.asarray(query.output), dtype=np.float32)
        db_l = np.ascontiguousarray(np.asarray(database.label))
        q_l = np.ascontiguousarray(np.asarray(query.label))
        if db_l.dtype != q_l.dtype or db_l.dtype.kind not in "iub" or db_l.dtype.itemsize not in (1, 4, 8):
            db_l = db_l.astype(np.int64)
            q_l = q_l.astype(np.int64)
        if db_f.ndim != 2 or q_f.ndim != 2 or db_f.shape[1] != q_f.shape[1]:
            raise ValueError("features must be [N, b] with equal b")
        if db_l.shape[1] != q_l.shape[1]:
            raise ValueError("label widths differ")
        R = int(self.R)
        if R > db_f.shape[0]:
            raise ValueError(f"operands could not be broadcast together: R={R} exceeds the database size {db_f.shape[0]}")
        out = C.c_double(0.0)
        ap = np.empty((q_f.shape[0],), dtype=np.float64)
        rc = lib.hg_maps_by_feature_host(db_f.ctypes.data, db_l.ctypes.data, db_f.shape[0], q_f.ctypes.data, q_l.ctypes.data,
                                         q_f.shape[0], db_f.shape[1], db_l.shape[1], db_l.dtype.itemsize, R, self.flags,
                                         C.byref(out), ap.ctypes.data)
        if rc == _native.HG_ELABEL:
            raise ValueError("labels must be 0/1 integers")
        _native.check(rc)
        val = np.float64(out.value)
        return (val, ap) if return_ap else val


# north_star asks for a `MAPs_CQ()` call signature; no such symbol exists in thuml/HashGAN (it belongs to
# thuml/DeepHash).  Exported as a plain alias of the class the reference really has.
MAPs_CQ = MAPs
