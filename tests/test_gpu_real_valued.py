"""Real-valued ranking mode (MAPs(R, binarize=False) -> hg_ip_map): the reference's literal lib/metric.py:13-23 on
un-binarised features.  Oracle = the NumPy restatement with the stable tie order (ip descending, row ascending); the
`real16` golden value comes from the UNMODIFIED reference (oracle/gen_golden.py).

Exactness: features on a coarse dyadic grid make every fp32 inner product exact in any summation order, so ids, inner
products and ties must match bit for bit; for continuous (tanh) features the fp32 sums of GPU and BLAS may differ in
the last bit, so the ranking is compared through the AP (tolerance 1e-6, the north_star's fp tolerance)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

from oracle import maps_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import torch

    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    import hashgan_b200
    from hashgan_b200 import _native

    _native.lib()
    return hashgan_b200


def _grid(rng, n, b, levels=9):
    return (rng.integers(0, levels, (n, b)).astype(np.float32) - (levels // 2)) / 4.0  # multiples of 1/4 in [-1, 1]


def _one_hot(rng, n, L):
    return np.eye(L, dtype=np.int64)[rng.integers(0, L, n)]


def _oracle(db, q, R):
    ips = q.output.astype(np.float64) @ db.output.astype(np.float64).T
    ids = np.argsort(-ips, 1, kind="stable")[:, :R]
    ap = maps_oracle.per_query_ap(db.output.astype(np.float64), db.label, q.output.astype(np.float64), q.label, R, tie="stable")
    return ap, ids, np.take_along_axis(ips, ids, 1)


@pytest.mark.parametrize("b,ndb,nq,L,R", [(16, 3000, 37, 10, 200), (64, 50000, 130, 10, 5000), (48, 20011, 70, 7, 1),
                                           (100, 30000, 20, 81, 30000), (64, 40000, 9, 10, 20000)])
def test_exact_on_dyadic_grid(hb, b, ndb, nq, L, R):
    """Exactly representable inner products: rows, inner products and AP identical to the oracle, heavy ties included;
    R == ndb (the whole ranking, cifar_evaluation.yaml:9,12) and R > 10240 (top-R sorted in global memory) covered."""
    rng = np.random.default_rng(b + R)
    lab = _one_hot if L <= 10 else (lambda r, n, l: (r.random((n, l)) < 0.03).astype(np.int64))
    db = NS(output=_grid(rng, ndb, b), label=lab(rng, ndb, L))
    q = NS(output=_grid(rng, nq, b), label=lab(rng, nq, L))
    ap, ids, ips = hb.MAPs(R, binarize=False).per_query_ap(db, q, want_ids=True)
    ref_ap, ref_ids, ref_ips = _oracle(db, q, R)
    assert np.array_equal(ids, ref_ids)
    assert np.array_equal(ips.astype(np.float64), ref_ips)
    assert np.array_equal(np.isnan(ap), np.isnan(ref_ap))
    keep = ~np.isnan(ap)
    assert np.max(np.abs(ap[keep] - ref_ap[keep])) <= 1e-12
    got = hb.MAPs(R, binarize=False).get_maps_by_feature(db, q)
    assert abs(got - maps_oracle.exact_mean_ap(ref_ap)) <= 1e-12


def test_query_chunks_are_invisible(hb):
    """A workspace that holds the keys of only a few queries at a time: same rows, inner products and AP."""
    from hashgan_b200 import _native

    rng = np.random.default_rng(77)
    b, ndb, nq, L, R = 32, 9001, 41, 5, 300
    db = NS(output=_grid(rng, ndb, b), label=_one_hot(rng, ndb, L))
    q = NS(output=_grid(rng, nq, b), label=_one_hot(rng, nq, L))
    one = _native.lib().hg_ip_map_workspace_bytes(1, ndb, b, L, R)
    small = hb.MAPs(R, binarize=False, workspace_limit=one * 6 + 512)   # chunks of ~6 queries
    ap_s, ids_s, ips_s = small.per_query_ap(db, q, want_ids=True)
    ap, ids, ips = hb.MAPs(R, binarize=False).per_query_ap(db, q, want_ids=True)
    assert np.array_equal(ids_s, ids) and np.array_equal(ips_s, ips) and np.array_equal(ap_s, ap)
    ref_ap, ref_ids, _ = _oracle(db, q, R)
    assert np.array_equal(ids, ref_ids) and np.max(np.abs(ap - ref_ap)) <= 1e-12


def test_pm1_codes_rank_like_the_hamming_path(hb):
    """On {-1,+1} codes ip = b - 2 d_H: the real-valued mode must return the very ranking of the binarised hot path."""
    from hashgan_b200.synthetic import make_workload

    wl, db, q = make_workload("C2", nq=150, ndb=60000)
    ap_h, ids_h, dist_h = hb.MAPs(1000).per_query_ap(db, q, want_ids=True)
    ap_r, ids_r, ips_r = hb.MAPs(1000, binarize=False).per_query_ap(db, q, want_ids=True)
    assert np.array_equal(ids_h, ids_r)
    assert np.array_equal(wl.b - 2 * dist_h, ips_r.astype(np.int64))
    assert np.array_equal(np.isnan(ap_h), np.isnan(ap_r))
    assert np.nanmax(np.abs(ap_h - ap_r)) <= 1e-12


def test_reference_golden_real_valued(hb, golden):
    """`real16`: mAP of the unmodified reference on tanh features (no ties)."""
    db = NS(output=golden["real16/db"], label=golden["real16/db_lab"].astype(np.int64))
    q = NS(output=golden["real16/q"], label=golden["real16/q_lab"].astype(np.int64))
    got = hb.MAPs(int(golden["real16/R"]), binarize=False).get_maps_by_feature(db, q)
    assert abs(got - float(golden["real16/map"])) <= 1e-6
    binarised = hb.MAPs(int(golden["real16/R"])).get_maps_by_feature(db, q)
    assert binarised != got  # the two modes are different metrics on real-valued features


def test_continuous_features_within_fp32_rounding(hb):
    rng = np.random.default_rng(5)
    ndb, nq, b, L, R = 200000, 64, 64, 10, 5000
    proto = rng.normal(size=(L, b))
    dl, ql = rng.integers(0, L, ndb), rng.integers(0, L, nq)
    db = NS(output=np.tanh(proto[dl] * 0.5 + rng.normal(size=(ndb, b))).astype(np.float32), label=np.eye(L, dtype=np.int64)[dl])
    q = NS(output=np.tanh(proto[ql] * 0.5 + rng.normal(size=(nq, b))).astype(np.float32), label=np.eye(L, dtype=np.int64)[ql])
    ap, ids, ips = hb.MAPs(R, binarize=False).per_query_ap(db, q, want_ids=True)
    ref_ap, ref_ids, ref_ips = _oracle(db, q, R)
    # neighbours whose inner products differ by less than the fp32 rounding of a 64-term sum may swap relative to the
    # float64 oracle (the reference's own float32 np.dot has the same freedom): a per-query AP moves by ~1e-6, the mean less
    assert np.max(np.abs(ap - ref_ap)) <= 1e-5
    assert abs(np.mean(ap) - np.mean(ref_ap)) <= 1e-6
    assert np.max(np.abs(ips - ref_ips)) <= 1e-4           # fp32 sums of 64 products of magnitude <= 1
    assert np.mean(ids == ref_ids) >= 0.99                  # only near-ties may swap neighbours ...
    mine = np.einsum("qb,qrb->qr", q.output.astype(np.float64), db.output.astype(np.float64)[ids])
    assert np.max(np.abs(mine - ref_ips)) <= 1e-5           # ... i.e. every rank holds a row whose exact inner product is the oracle's
    assert (np.diff(ips, axis=1) <= 0).all()                # sorted by inner product, descending
    assert 0.2 < np.mean(ap) < 0.99


def test_edge_cases(hb):
    rng = np.random.default_rng(9)
    db = NS(output=_grid(rng, 500, 32), label=_one_hot(rng, 500, 4))
    q = NS(output=_grid(rng, 5, 32), label=_one_hot(rng, 5, 4))
    with pytest.raises(ValueError, match="could not be broadcast"):
        hb.MAPs(501, binarize=False).get_maps_by_feature(db, q)
    # zero features: every inner product is +0 (or -0): one big tie, rows in ascending order
    z = NS(output=np.zeros((500, 32), np.float32), label=db.label)
    zq = NS(output=-np.zeros((5, 32), np.float32), label=q.label)
    ap, ids, ips = hb.MAPs(100, binarize=False).per_query_ap(z, zq, want_ids=True)
    assert np.array_equal(ids, np.tile(np.arange(100), (5, 1))) and not ips.any()
    # a query without relevant rows is skipped (NaN), all skipped -> nan like the reference
    nolab = NS(output=q.output, label=np.zeros((5, 4), np.int64))
    ap = hb.MAPs(50, binarize=False).per_query_ap(db, nolab)
    assert np.isnan(ap).all()
