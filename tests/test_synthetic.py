"""Synthetic workloads: shapes of BASELINE.json's configs and the REAL label rows of the
reference's list files (committed bit-packed: hashgan_b200/data/label_rows.npz, made by oracle/gen_label_rows.py)."""
import os

import numpy as np
import pytest

from hashgan_b200 import synthetic
from tests import helpers


def test_cifar10_rows_are_the_list_files():
    db = synthetic.real_label_rows("cifar10", "database")
    te = synthetic.real_label_rows("cifar10", "test")
    assert db.shape == (54000, 10) and te.shape == (1000, 10) and db.dtype == np.int64
    assert (db.sum(1) == 1).all() and (te.sum(1) == 1).all()
    assert (db.sum(0) == 5400).all() and (te.sum(0) == 100).all()  # exactly balanced


def test_nuswide_rows_statistics():
    db = synthetic.real_label_rows("nuswide_81", "database")
    te = synthetic.real_label_rows("nuswide_81", "test")
    assert db.shape == (168692, 81) and te.shape == (5000, 81)
    assert abs(db.sum(1).mean() - 2.432) < 1e-3 and db.sum(1).max() == 13 and db.sum(1).min() >= 1
    card = np.bincount(db.sum(1))
    assert list(card[1:6]) == [62349, 42763, 27036, 16903, 10042]  # cardinality histogram quoted by SURVEY 8(d)
    top = np.sort(db.mean(0))[::-1][:5]
    assert np.allclose(top, [.36, .26, .23, .17, .17], atol=6e-3)


@pytest.mark.skipif(not helpers.have_reference(), reason="reference lists not mounted")
def test_rows_equal_the_mounted_reference_lists():
    for name, split in (("cifar10", "test"), ("nuswide_81", "test")):
        path = os.path.join(helpers.REFERENCE_DIR, "data_list", name, split + ".txt")
        want = np.array([[int(x) for x in ln.split()[1:]] for ln in open(path) if ln.split()], dtype=np.int64)
        assert np.array_equal(synthetic.real_label_rows(name, split), want)


def test_workload_shapes_and_label_sources():
    wl, db, q = synthetic.make_workload("C1")
    assert (wl.nq, wl.ndb, wl.b, wl.L, wl.R) == (1000, 54000, 32, 10, 54000)
    assert np.array_equal(db.label, synthetic.real_label_rows("cifar10", "database"))
    assert set(np.unique(db.output)) == {-1.0, 1.0} and db.output.dtype == np.float32
    wl, db, q = synthetic.make_workload("C5", nq=64, ndb=200000)  # > 168,692 rows: seeded bootstrap of the real rows
    assert db.label.shape == (200000, 81) and q.label.shape == (64, 81)
    rows = {r.tobytes() for r in np.packbits(synthetic.real_label_rows("nuswide_81", "database").astype(np.uint8), axis=1)}
    assert all(r.tobytes() in rows for r in np.packbits(db.label[:2000].astype(np.uint8), axis=1))
    wl2, db2, _ = synthetic.make_workload("C5", nq=64, ndb=200000)
    assert np.array_equal(db.label, db2.label) and np.array_equal(db.output, db2.output)  # seeded
    wl, db, q = synthetic.make_workload("C4", nq=16, ndb=1000)
    assert (db.label.sum(1) == 1).all() and wl.R == 1000
