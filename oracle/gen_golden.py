#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference mounted); the GPU box and the
test-suite read the committed .npz files, never the reference.

    python oracle/gen_golden.py            # rewrites tests/golden/metric_golden.npz

What is recorded per case: the inputs (sign bits of the +-1 codes and the 0/1 labels,
bit-packed) and the outputs of ``lib.metric.MAPs(R).get_maps_by_feature`` (lib/metric.py:4-24)
called unmodified:
  * ``map_eps``   -- the call on the eps-augmented fp64 features of SURVEY 8(c): every inner
                     product is unique, so the reference ranks by (ip desc, db row asc)
                     independent of the NumPy build;
  * ``ap_eps``    -- per-query AP from the same call with one query row at a time
                     (NaN when the reference's ``apx`` list stays empty -> mean of empty);
  * ``map_default`` -- the plain call on float32 codes with NumPy's default (unstable)
                     argsort: informational only, depends on the NumPy build.
A real-valued case (no ties) pins the NumPy restatement outside the +-1 domain.
"""
from __future__ import annotations

import os
import sys
import warnings
from types import SimpleNamespace as NS

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HASHGAN_REFERENCE", "/root/reference")


def _ref_maps():
    sys.path.insert(0, REF)
    try:
        from lib.metric import MAPs  # lib/metric.py:4 (NumPy only)
    finally:
        sys.path.pop(0)
    return MAPs


def pm1(rng, n, b):
    return (rng.integers(0, 2, (n, b), dtype=np.int8) * 2 - 1).astype(np.int8)


def one_hot(rng, n, L):
    lab = np.zeros((n, L), dtype=np.int64)
    lab[np.arange(n), rng.integers(0, L, n)] = 1
    return lab


def multi_hot(rng, n, L, p):
    lab = (rng.random((n, L)) < p).astype(np.int64)
    return lab


def class_codes(rng, lab, b, flip):
    L = lab.shape[1]
    proto = pm1(rng, L, b)
    cls = lab.argmax(1)
    codes = proto[cls].copy()
    mask = rng.random(codes.shape) < flip
    codes[mask] *= -1
    return codes.astype(np.int8)


def eps_augment(db_codes, q_codes):
    nq, nd = len(q_codes), len(db_codes)
    qa = np.concatenate([q_codes.astype(np.float64), np.ones((nq, 1))], 1)
    da = np.concatenate([db_codes.astype(np.float64), -(np.arange(nd, dtype=np.float64)[:, None]) * 2.0 ** -32], 1)
    return da, qa


def run_case(MAPs, name, db_codes, db_lab, q_codes, q_lab, R):
    da, qa = eps_augment(db_codes, q_codes)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        map_eps = MAPs(R).get_maps_by_feature(NS(output=da, label=db_lab), NS(output=qa, label=q_lab))
        ap_eps = np.array([
            MAPs(R).get_maps_by_feature(NS(output=da, label=db_lab), NS(output=qa[i:i + 1], label=q_lab[i:i + 1]))
            for i in range(len(qa))
        ], dtype=np.float64)
        map_default = MAPs(R).get_maps_by_feature(
            NS(output=db_codes.astype(np.float32), label=db_lab),
            NS(output=q_codes.astype(np.float32), label=q_lab))
    out = {
        "b": np.int64(db_codes.shape[1]),
        "L": np.int64(db_lab.shape[1]),
        "R": np.int64(R),
        "ndb": np.int64(len(db_codes)),
        "nq": np.int64(len(q_codes)),
        "db_bits": np.packbits(db_codes > 0, axis=1),
        "q_bits": np.packbits(q_codes > 0, axis=1),
        "db_lab_bits": np.packbits(db_lab.astype(bool), axis=1),
        "q_lab_bits": np.packbits(q_lab.astype(bool), axis=1),
        "map_eps": np.float64(map_eps),
        "ap_eps": ap_eps,
        "map_default": np.float64(map_default),
    }
    print(f"{name:>14s}: b={out['b']} L={out['L']} nq={out['nq']} ndb={out['ndb']} R={R} "
          f"map_eps={map_eps!r} map_default={map_default!r} nan_queries={int(np.isnan(ap_eps).sum())}")
    return {f"{name}/{k}": v for k, v in out.items()}


def main():
    MAPs = _ref_maps()
    blob = {}
    names = []

    rng = np.random.default_rng(101)
    dl, ql = one_hot(rng, 200, 10), one_hot(rng, 8, 10)
    blob.update(run_case(MAPs, "tiny32_full", pm1(rng, 200, 32), dl, pm1(rng, 8, 32), ql, 200)); names.append("tiny32_full")
    rng = np.random.default_rng(102)
    dl, ql = one_hot(rng, 200, 10), one_hot(rng, 8, 10)
    blob.update(run_case(MAPs, "tiny32_r50", pm1(rng, 200, 32), dl, pm1(rng, 8, 32), ql, 50)); names.append("tiny32_r50")

    rng = np.random.default_rng(103)
    dl, ql = one_hot(rng, 2000, 10), one_hot(rng, 16, 10)
    blob.update(run_case(MAPs, "class64", class_codes(rng, dl, 64, 0.25), dl, pm1(rng, 16, 64), ql, 500)); names.append("class64")

    # class-correlated queries and db from the same prototypes: mAP far from chance, heavy low-distance ties
    rng = np.random.default_rng(104)
    L = 10
    proto = pm1(rng, L, 64)
    dl, ql = one_hot(rng, 3000, L), one_hot(rng, 24, L)

    def noisy(lab):
        c = proto[lab.argmax(1)].copy()
        c[rng.random(c.shape) < 0.25] *= -1
        return c.astype(np.int8)
    blob.update(run_case(MAPs, "proto64", noisy(dl), dl, noisy(ql), ql, 1000)); names.append("proto64")

    rng = np.random.default_rng(105)
    dl, ql = multi_hot(rng, 1500, 81, 0.03), multi_hot(rng, 8, 81, 0.03)
    blob.update(run_case(MAPs, "multi128", pm1(rng, 1500, 128), dl, pm1(rng, 8, 128), ql, 300)); names.append("multi128")

    rng = np.random.default_rng(106)
    dl, ql = one_hot(rng, 1000, 10), one_hot(rng, 8, 10)
    blob.update(run_case(MAPs, "pad48", pm1(rng, 1000, 48), dl, pm1(rng, 8, 48), ql, 100)); names.append("pad48")

    # duplicates: every db code identical -> one bucket holds everything, order is purely by row
    rng = np.random.default_rng(107)
    dl, ql = one_hot(rng, 400, 5), one_hot(rng, 6, 5)
    same = np.repeat(pm1(rng, 1, 32), 400, 0)
    blob.update(run_case(MAPs, "allsame32", same, dl, pm1(rng, 6, 32), ql, 123)); names.append("allsame32")

    # some queries have no relevant item at all (label column never set in the db)
    rng = np.random.default_rng(108)
    dl = one_hot(rng, 500, 6); dl[:, 5] = 0; dl[dl.sum(1) == 0, 0] = 1
    ql = one_hot(rng, 10, 6); ql[:3] = 0; ql[:3, 5] = 1
    blob.update(run_case(MAPs, "norel64", pm1(rng, 500, 64), dl, pm1(rng, 10, 64), ql, 77)); names.append("norel64")

    # real-valued features (tanh-like), no ties: pins the NumPy restatement outside +-1
    rng = np.random.default_rng(109)
    dbf = np.tanh(rng.normal(size=(300, 16))).astype(np.float32)
    qf = np.tanh(rng.normal(size=(8, 16))).astype(np.float32)
    dl, ql = one_hot(rng, 300, 10), one_hot(rng, 8, 10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MAPs(120).get_maps_by_feature(NS(output=dbf, label=dl), NS(output=qf, label=ql))
    blob.update({"real16/db": dbf, "real16/q": qf, "real16/db_lab": dl.astype(np.uint8), "real16/q_lab": ql.astype(np.uint8),
                 "real16/R": np.int64(120), "real16/map": np.float64(m)})
    print(f"        real16: map={m!r}")

    blob["cases"] = np.array(names)
    blob["numpy_version"] = np.array(np.__version__)
    out = os.path.join(ROOT, "tests", "golden", "metric_golden.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
