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


from oracle.c_oracle import COracle  # noqa: E402,F401  (ctypes view of oracle/hamming_oracle.c)
