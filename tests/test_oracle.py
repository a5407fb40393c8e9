"""The oracle (NumPy restatement + C restatement) pinned against the golden vectors made from the unmodified
reference (oracle/gen_golden.py), and, where /root/reference is mounted, against the reference itself."""
import warnings
from types import SimpleNamespace as NS

import numpy as np
import pytest

from oracle import maps_oracle
from tests import helpers


def _names():
    import os
    g = np.load(os.path.join(helpers.ROOT, "tests", "golden", "metric_golden.npz"))
    return helpers.golden_names(g)


@pytest.mark.parametrize("name", _names())
def test_numpy_oracle_matches_golden(golden, name):
    c = helpers.golden_case(golden, name)
    ap = maps_oracle.per_query_ap(c.db.output, c.db.label, c.q.output, c.q.label, c.R, tie="stable")
    assert np.array_equal(np.isnan(ap), np.isnan(c.ap_eps))
    # same integers, same fp64 divides, same pairwise np.sum -> bit-identical per-query AP
    assert np.array_equal(ap[~np.isnan(ap)], c.ap_eps[~np.isnan(c.ap_eps)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = maps_oracle.OracleMAPs(c.R, tie="stable").get_maps_by_feature(c.db, c.q)
    assert m == c.map_eps or (np.isnan(m) and np.isnan(c.map_eps))


@pytest.mark.parametrize("name", _names())
def test_c_oracle_matches_golden(golden, c_oracle, name):
    c = helpers.golden_case(golden, name)
    ap, rel, ids, dist = c_oracle.hamming_map(c.db, c.q, c.R, want_ids=True)
    assert np.array_equal(np.isnan(ap), np.isnan(c.ap_eps))
    np.testing.assert_allclose(ap[~np.isnan(ap)], c.ap_eps[~np.isnan(ap)], rtol=0, atol=1e-13)
    ref_ids, ref_dist = maps_oracle.rank_ids(c.db.output, c.q.output, c.R, tie="stable")
    assert np.array_equal(ids.astype(np.int64), ref_ids)
    assert np.array_equal(dist.astype(np.int32), ref_dist)
    assert abs(maps_oracle.exact_mean_ap(ap) - c.map_eps) <= 1e-13


def test_real_valued_golden(golden):
    db = NS(output=golden["real16/db"], label=golden["real16/db_lab"].astype(np.int64))
    q = NS(output=golden["real16/q"], label=golden["real16/q_lab"].astype(np.int64))
    R = int(golden["real16/R"])
    for tie in ("stable", "reference"):  # no ties in real-valued features: both orders agree
        assert maps_oracle.OracleMAPs(R, tie=tie).get_maps_by_feature(db, q) == float(golden["real16/map"])


def test_packers_and_xor_popc_identity(c_oracle):
    rng = np.random.default_rng(7)
    for b in (1, 31, 32, 33, 48, 64, 96, 128, 200):
        q = (rng.integers(0, 2, (17, b)) * 2 - 1).astype(np.float32)
        d = (rng.integers(0, 2, (101, b)) * 2 - 1).astype(np.float32)
        qp, dp = maps_oracle.pack_sign_bits(q), maps_oracle.pack_sign_bits(d)
        assert np.array_equal(qp, c_oracle.pack_sign(q))
        x = qp[:, None, :] ^ dp[None, :, :]
        pop = np.unpackbits(x.view(np.uint8), axis=-1).sum(-1)
        assert np.array_equal(pop, maps_oracle.hamming_from_pm1(q, d))  # ip = b - 2 d_H  
    lab = (rng.random((50, 81)) < 0.05).astype(np.int64)
    assert np.array_equal(maps_oracle.pack_label_bits(lab), c_oracle.pack_labels(lab))


def test_edge_cases(c_oracle):
    rng = np.random.default_rng(3)
    db = NS(output=(rng.integers(0, 2, (40, 32)) * 2 - 1).astype(np.float32), label=np.eye(4, dtype=np.int64)[rng.integers(0, 4, 40)])
    q = NS(output=(rng.integers(0, 2, (3, 32)) * 2 - 1).astype(np.float32), label=np.eye(4, dtype=np.int64)[rng.integers(0, 4, 3)])
    with pytest.raises(ValueError):  # lib/metric.py:21 broadcast error when R > Ndb
        maps_oracle.OracleMAPs(41).get_maps_by_feature(db, q)
    with pytest.raises(ValueError):
        c_oracle.hamming_map(db, q, 41)
    q0 = NS(output=q.output, label=np.zeros_like(q.label))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.isnan(maps_oracle.OracleMAPs(10).get_maps_by_feature(db, q0))  # lib/metric.py:24 on an empty list
    ap, rel, _, _ = c_oracle.hamming_map(db, q0, 10)
    assert np.isnan(ap).all() and (rel == 0).all()
    before = q.label.copy()
    maps_oracle.OracleMAPs(10).get_maps_by_feature(db, q)
    assert np.array_equal(before, q.label)  # inputs are not mutated (lib/metric.py:17 copies)


@pytest.mark.skipif(not helpers.have_reference(), reason="/root/reference not mounted (GPU box)")
def test_restatement_equals_unmodified_reference():
    MAPs = helpers.reference_maps_class()
    rng = np.random.default_rng(11)
    for b, nq, ndb, L, R in [(32, 12, 700, 10, 700), (64, 9, 1500, 10, 300), (48, 7, 900, 6, 128), (128, 5, 800, 81, 200)]:
        dbc = (rng.integers(0, 2, (ndb, b)) * 2 - 1).astype(np.float32)
        qc = (rng.integers(0, 2, (nq, b)) * 2 - 1).astype(np.float32)
        if L == 81:
            dl, ql = (rng.random((ndb, L)) < 0.03).astype(np.int64), (rng.random((nq, L)) < 0.03).astype(np.int64)
        else:
            dl, ql = np.eye(L, dtype=np.int64)[rng.integers(0, L, ndb)], np.eye(L, dtype=np.int64)[rng.integers(0, L, nq)]
        da, qa = maps_oracle.eps_augment(dbc, qc)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = MAPs(R).get_maps_by_feature(NS(output=da, label=dl), NS(output=qa, label=ql))  # eps protocol, SURVEY 8(c)
            got = maps_oracle.OracleMAPs(R, tie="stable").get_maps_by_feature(NS(output=dbc, label=dl), NS(output=qc, label=ql))
            # literal restatement, default argsort: identical call -> identical result on the same NumPy
            want_d = MAPs(R).get_maps_by_feature(NS(output=dbc, label=dl), NS(output=qc, label=ql))
            got_d = maps_oracle.OracleMAPs(R, tie="reference").get_maps_by_feature(NS(output=dbc, label=dl), NS(output=qc, label=ql))
        assert got == want
        assert got_d == want_d
