"""CPU oracle for the retrieval metric -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (hashgan_b200.metric) never does.

This is a NumPy restatement of the reference's metric, thuml/HashGAN
``lib/metric.py:12-24`` (``MAPs.get_maps_by_feature``).  Parity is PINNED: the
restatement is checked against (a) the unmodified reference imported from
/root/reference in the build container (tests/test_oracle.py::test_restatement_equals_unmodified_reference,
skipped where the reference is not mounted) and (b) golden vectors generated from the
unmodified reference by oracle/gen_golden.py and committed under tests/golden/.

Reference line map
    lib/metric.py:13   ips = np.dot(query.output, database.output.T)        -> _inner_products
    lib/metric.py:14   ids = np.argsort(-ips, 1)                            -> _rank (kind selectable)
    lib/metric.py:17-18 label = query.label[i].copy(); label[label==0] = -1 -> _relevance
    lib/metric.py:19   imatch = sum(db.label[ids[:R]] == label, 1) > 0      -> _relevance
    lib/metric.py:20   rel = sum(imatch)
    lib/metric.py:21   px = cumsum(imatch).astype(float) / arange(1, R+1)
    lib/metric.py:22-23 if rel != 0: apx.append(sum(px*imatch)/rel)
    lib/metric.py:24   mean(array(apx))

Tie order.  The reference's argsort is NumPy's default (unstable) sort; with b-bit
codes only b+1 distinct keys exist, so the reference's own output depends on the
NumPy build.  The build fixes the order (Hamming distance ascending, database row
ascending) == ``kind='stable'``.  ``tie='stable'`` below restates exactly that;
``tie='reference'`` keeps the literal default argsort (used for the timed CPU
baseline, where the cost of the reference's own call is what is measured).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

__all__ = [
    "OracleMAPs",
    "per_query_ap",
    "rank_ids",
    "hamming_from_pm1",
    "pack_sign_bits",
    "pack_label_bits",
    "eps_augment",
]


def _inner_products(q_out: np.ndarray, db_out: np.ndarray) -> np.ndarray:
    # lib/metric.py:13
    return np.dot(q_out, db_out.T)


def _rank(ips: np.ndarray, tie: str) -> np.ndarray:
    # lib/metric.py:14 ; 'stable' == (ip desc, db row asc) == (d_H asc, db row asc)
    if tie == "reference":
        return np.argsort(-ips, 1)
    if tie == "stable":
        return np.argsort(-ips, 1, kind="stable")
    raise ValueError("tie must be 'reference' or 'stable'")


def _relevance(db_label: np.ndarray, q_label_row: np.ndarray, top: np.ndarray) -> np.ndarray:
    # lib/metric.py:17-19
    lab = q_label_row.copy()
    lab[lab == 0] = -1
    return np.sum(db_label[top, :] == lab, 1) > 0


def _ap_from_matches(imatch: np.ndarray, R: int):
    # lib/metric.py:20-23 ; returns (ap or None, rel)
    rel = int(np.sum(imatch))
    px = np.cumsum(imatch).astype(float) / np.arange(1, R + 1, 1)
    if rel != 0:
        return float(np.sum(px * imatch) / rel), rel
    return None, rel


class OracleMAPs:
    """Same constructor/method shape as the reference class (lib/metric.py:4-24)."""

    def __init__(self, r: int, tie: str = "stable"):
        self.R = r
        self.tie = tie

    def get_maps_by_feature(self, database, query):
        ips = _inner_products(np.asarray(query.output), np.asarray(database.output))
        ids = _rank(ips, self.tie)
        db_label = np.asarray(database.label)
        q_label = np.asarray(query.label)
        apx = []
        for i in range(ips.shape[0]):
            imatch = _relevance(db_label, q_label[i, :], ids[i, : self.R])
            ap, _ = _ap_from_matches(imatch, self.R)
            if ap is not None:
                apx.append(ap)
        return np.mean(np.array(apx))


def per_query_ap(db_out, db_label, q_out, q_label, R: int, tie: str = "stable", chunk: int = 64):
    """Per-query AP (NaN where the top-R holds no relevant item), query-chunked so the
    [chunk, Ndb] temporaries stay small.  Same arithmetic as lib/metric.py:13-23."""
    db_out = np.asarray(db_out)
    q_out = np.asarray(q_out)
    db_label = np.asarray(db_label)
    q_label = np.asarray(q_label)
    nq = q_out.shape[0]
    out = np.full(nq, np.nan, dtype=np.float64)
    for s in range(0, nq, chunk):
        ips = _inner_products(q_out[s : s + chunk], db_out)
        ids = _rank(ips, tie)
        for i in range(ips.shape[0]):
            imatch = _relevance(db_label, q_label[s + i, :], ids[i, :R])
            ap, _ = _ap_from_matches(imatch, R)
            if ap is not None:
                out[s + i] = ap
    return out


def rank_ids(db_out, q_out, R: int, tie: str = "stable", chunk: int = 64):
    """Top-R database rows and their Hamming distances per query, (d asc, row asc)."""
    db_out = np.asarray(db_out)
    q_out = np.asarray(q_out)
    b = db_out.shape[1]
    nq = q_out.shape[0]
    ids = np.empty((nq, R), dtype=np.int64)
    dist = np.empty((nq, R), dtype=np.int32)
    for s in range(0, nq, chunk):
        ips = _inner_products(q_out[s : s + chunk], db_out)
        order = _rank(ips, tie)[:, :R]
        ids[s : s + chunk] = order
        top_ip = np.take_along_axis(ips, order, 1)
        dist[s : s + chunk] = np.rint((b - top_ip) / 2).astype(np.int32)
    return ids, dist


def exact_mean_ap(ap: np.ndarray) -> float:
    """lib/metric.py:24 over the non-NaN entries (queries with rel != 0)."""
    kept = ap[~np.isnan(ap)]
    return float(np.mean(kept)) if kept.size else float("nan")


def fsum_ap_from_matches(imatch: np.ndarray) -> float:
    """Correctly rounded AP of one ranked 0/1 relevance vector (math.fsum): used to bound
    the summation-order error of both the reference's pairwise np.sum and the GPU sum."""
    rel = int(imatch.sum())
    if rel == 0:
        return float("nan")
    cum = np.cumsum(imatch)
    pos = np.nonzero(imatch)[0]
    return math.fsum(float(cum[p]) / float(p + 1) for p in pos) / rel


def eps_augment(db_codes: np.ndarray, q_codes: np.ndarray):
    """SURVEY 8(c) tie-deterministic protocol: one extra fp64 column makes every inner
    product unique so that the UNMODIFIED reference ranks by (ip desc, db row asc).
    Valid while Ndb < 2**32 and b <= 2**20 (all partial sums exact in fp64)."""
    nq, nd = len(q_codes), len(db_codes)
    qa = np.concatenate([q_codes.astype(np.float64), np.ones((nq, 1))], 1)
    da = np.concatenate([db_codes.astype(np.float64), -(np.arange(nd, dtype=np.float64)[:, None]) * 2.0 ** -32], 1)
    return da, qa


# --- bit-domain restatements used to check the packers and the XOR/POPC identity ---------

def pack_sign_bits(feat: np.ndarray) -> np.ndarray:
    """bit j of word w = (feat[:, 32w+j] > 0); pad bits zero.  uint32 [N, ceil(b/32)]."""
    feat = np.asarray(feat)
    n, b = feat.shape
    W = (b + 31) // 32
    bits = np.zeros((n, W * 32), dtype=np.uint8)
    bits[:, :b] = feat > 0
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    words = (bits.reshape(n, W, 32).astype(np.uint64) * weights).sum(-1)
    return words.astype(np.uint32)


def pack_label_bits(lab: np.ndarray) -> np.ndarray:
    """bit j of word w = (lab[:, 32w+j] != 0 and == 1 in the reference's sense)."""
    lab = np.asarray(lab)
    return pack_sign_bits((lab == 1).astype(np.float32))


def hamming_from_pm1(q_codes: np.ndarray, db_codes: np.ndarray) -> np.ndarray:
    """d_H = (b - ip)/2 on +-1 inputs, int32 [Nq, Ndb]."""
    b = q_codes.shape[1]
    ips = _inner_products(q_codes.astype(np.float32), db_codes.astype(np.float32))
    return np.rint((b - ips) / 2).astype(np.int32)


def as_record(output, label):
    return SimpleNamespace(output=output, label=label)
