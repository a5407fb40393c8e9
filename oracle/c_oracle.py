"""ctypes view of oracle/hamming_oracle.c -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (its cpu_baseline leg and its parity self-check, where the oracle is the
checker, never the thing measured) may import this module; the product (hashgan_b200/) never does.  The C file restates
thuml/HashGAN lib/metric.py:12-24 in the packed-bit domain (line map in its header) and is pinned against the golden vectors
of the unmodified reference by tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build() -> str:
    """gcc build of the C restatement into oracle/_build/ (building the checker is not using it)."""
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libhamming_oracle.so")
    src = os.path.join(HERE, "hamming_oracle.c")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        cmd = ["gcc", "-O3", "-march=x86-64-v2", "-mpopcnt", "-fopenmp", "-shared", "-fPIC", "-o", out, src]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"command failed ({proc.returncode}): {' '.join(cmd)}\n{proc.stdout}")
    return out


class COracle:
    """oracle/_build/libhamming_oracle.so: pack + (distance, row)-ordered top-R + AP on the host CPU."""

    def __init__(self):
        self.lib = C.CDLL(build())
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
        return self.hamming_map_packed(dbc, dbl, qc, ql, b, L, R, want_ids, threads)

    def hamming_map_packed(self, dbc, dbl, qc, ql, b, L, R, want_ids=False, threads=0):
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
