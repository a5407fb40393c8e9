"""Shared test helpers: golden-vector decoding and the ctypes view of the C oracle (oracle/hamming_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = os.environ.get("HASHGAN_REFERENCE", "/root/reference")


def have_reference() -> bool:
    return os.path.exists(os.path.join(REFERENCE_DIR, "lib", "metric.py"))


def reference_maps_class():
    """The UNMODIFIED reference class (only where /root/reference is mounted, i.e. the build container)."""
    sys.path.insert(0, REFERENCE_DIR)
    try:
        saved = sys.modules.pop("lib", None), sys.modules.pop("lib.metric", None)
        from lib.metric import MAPs  # noqa
    finally:
        sys.path.pop(0)
        for name in ("lib", "lib.metric"):
            sys.modules.pop(name, None)
        if saved[0] is not None:
            sys.modules["lib"] = saved[0]
        if saved[1] is not None:
            sys.modules["lib.metric"] = saved[1]
    return MAPs


def golden_case(golden, name):
    """Decode one case of tests/golden/metric_golden.npz into +-1 float32 codes and int64 labels."""
    g = lambda k: golden[f"{name}/{k}"]
    b, L = int(g("b")), int(g("L"))
    unbits = lambda a, n: np.unpackbits(a, axis=1)[:, :n]
    db_codes = unbits(g("db_bits"), b).astype(np.float32) * 2 - 1
    q_codes = unbits(g("q_bits"), b).astype(np.float32) * 2 - 1
    db_lab = unbits(g("db_lab_bits"), L).astype(np.int64)
    q_lab = unbits(g("q_lab_bits"), L).astype(np.int64)
    return SimpleNamespace(
        name=name, b=b, L=L, R=int(g("R")),
        db=SimpleNamespace(output=db_codes, label=db_lab),
        q=SimpleNamespace(output=q_codes, label=q_lab),
        map_eps=float(g("map_eps")), ap_eps=np.array(g("ap_eps")), map_default=float(g("map_default")))


def golden_names(golden):
    return [str(x) for x in golden["cases"]]


class COracle:
    """ctypes wrapper of oracle/_build/libhamming_oracle.so (built by __graft_entry__.build_oracle)."""

    def __init__(self):
        import __graft_entry__ as entry

        self.lib = C.CDLL(entry.build_oracle())
        i64, vp = C.c_int64, C.c_void_p
        self.lib.hgo_pack_sign_f32.argtypes = [vp, i64, C.c_int, vp]
        self.lib.hgo_pack_labels_i64.argtypes = [vp, i64, C.c_int, vp]
        self.lib.hgo_hamming_map.argtypes = [vp, vp, i64, vp, vp, i64, C.c_int, C.c_int, i64, vp, vp, vp, vp, C.c_int]

    def pack_sign(self, feat):
        feat = np.ascontiguousarray(feat, dtype=np.float32)
        n, b = feat.shape
        out = np.zeros((n, (b + 31) // 32), dtype=np.uint32)
        assert self.lib.hgo_pack_sign_f32(feat.ctypes.data, n, b, out.ctypes.data) == 0
        return out

    def pack_labels(self, lab):
        lab = np.ascontiguousarray(lab, dtype=np.int64)
        n, L = lab.shape
        out = np.zeros((n, (L + 31) // 32), dtype=np.uint32)
        assert self.lib.hgo_pack_labels_i64(lab.ctypes.data, n, L, out.ctypes.data) == 0
        return out

    def hamming_map(self, db, q, R, want_ids=False, threads=0):
        """db / q: records with +-1 .output and 0/1 .label.  Returns (ap, rel, ids, dist)."""
        b, L = db.output.shape[1], db.label.shape[1]
        dbc, qc = self.pack_sign(db.output), self.pack_sign(q.output)
        dbl, ql = self.pack_labels(db.label), self.pack_labels(q.label)
        nq, ndb = len(qc), len(dbc)
        ap = np.empty(nq, dtype=np.float64)
        rel = np.empty(nq, dtype=np.int64)
        ids = np.empty((nq, R), dtype=np.uint32) if want_ids else None
        dist = np.empty((nq, R), dtype=np.uint16) if want_ids else None
        rc = self.lib.hgo_hamming_map(qc.ctypes.data, ql.ctypes.data, nq, dbc.ctypes.data, dbl.ctypes.data, ndb, b, L, R,
                                      ap.ctypes.data, rel.ctypes.data,
                                      ids.ctypes.data if want_ids else None, dist.ctypes.data if want_ids else None, threads)
        if rc == 2:
            raise ValueError("R exceeds the database size")
        assert rc == 0, rc
        return ap, rel, ids, dist
