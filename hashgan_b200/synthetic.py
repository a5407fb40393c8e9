"""Seeded synthetic inputs of the shapes BASELINE.json names.

Codes are uniform random {-1,+1} float32 (or class-correlated).  Labels:
  C1      the REAL label rows of the reference's data_list/cifar10/{database,test}.txt (10-way one-hot, exactly balanced);
  C5      bootstrap (seeded, with replacement) of the REAL rows of data_list/nuswide_81/database.txt (168,692 rows, 81-way
          multi-label, mean 2.43 labels per image) and the rows of test.txt for the queries -- label co-occurrence is kept;
  C2, C4  one-hot draws from the generator, as the survey specifies.
Nothing here reads /root/reference: the label matrices ship bit-packed in hashgan_b200/data/label_rows.npz (made by
oracle/gen_label_rows.py in the build container).  `multi_hot_labels` (independent per-class rates measured from the same
list) remains for label widths other than 81 and for the synthetic image loader.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from functools import lru_cache
from types import SimpleNamespace

import numpy as np

__all__ = ["Workload", "WORKLOADS", "make_workload", "pm1_codes", "one_hot_labels", "multi_hot_labels", "proto_codes",
           "real_label_rows", "list_labels"]


@dataclass(frozen=True)
class Workload:
    name: str
    nq: int
    ndb: int
    b: int
    L: int
    R: int
    labels: str  # "onehot" (generator) | "cifar10" (real rows) | "nuswide" (bootstrap of the real rows)
    seed: int
    note: str = ""


# C1..C5 of SURVEY.md section 8 (C3 is the encoder workload and lives in hashgan_b200.encoder)
WORKLOADS = {
    "C1": Workload("C1", 1000, 54000, 32, 10, 54000, "cifar10", 1, "cifar_evaluation.yaml shape, MODEL.HASH_DIM=32, MAP_R=DB_SIZE"),
    "C1_64": Workload("C1_64", 1000, 54000, 64, 10, 54000, "cifar10", 11, "cifar_evaluation.yaml as shipped (HASH_DIM default 64)"),
    "C2": Workload("C2", 10000, 100000, 48, 10, 5000, "onehot", 2, "48-bit codes: 2 words with 16 zero pad bits"),
    "C4": Workload("C4", 10000, 1000000, 64, 10, 5000, "onehot", 4, "headline: 10k queries x 1M database, 64-bit, mAP@5000"),
    "C5": Workload("C5", 5000, 2000000, 128, 81, 5000, "nuswide", 5, "NUS-WIDE_81-shaped multi-label, 128-bit"),
}

# per-class positive rates of data_list/nuswide_81/database.txt (168,692 rows), sorted, rounded to 3 digits;
# they sum to 2.43 labels per image.  Summary statistics only -- no row of the list file is reproduced.
_NUSWIDE_RATES = [
    0.36, 0.264, 0.229, 0.172, 0.168, 0.109, 0.087, 0.072, 0.071, 0.066, 0.055, 0.045, 0.042, 0.042, 0.039, 0.031, 0.029,
    0.026, 0.026, 0.025, 0.025, 0.02, 0.02, 0.019, 0.019, 0.019, 0.018, 0.016, 0.015, 0.014, 0.013, 0.013, 0.013, 0.012,
    0.012, 0.012, 0.012, 0.011, 0.011, 0.01, 0.009, 0.009, 0.008, 0.008, 0.008, 0.008, 0.007, 0.007, 0.007, 0.006, 0.006,
    0.006, 0.005, 0.005, 0.005, 0.005, 0.005, 0.004, 0.004, 0.004, 0.004, 0.004, 0.003, 0.003, 0.003, 0.003, 0.003, 0.003,
    0.002, 0.002, 0.002, 0.002, 0.002, 0.002, 0.002, 0.002, 0.002, 0.001, 0.001, 0.0005, 0.0005,
]


def _nuswide_rates(L: int = 81) -> np.ndarray:
    base = np.array(_NUSWIDE_RATES, dtype=np.float64)
    if L <= len(base):
        return base[:L]
    return np.concatenate([base, np.full(L - len(base), 0.0005)])


def pm1_codes(rng: np.random.Generator, n: int, b: int) -> np.ndarray:
    """codes = (rng.integers(0,2,(N,b))*2-1).astype(float32)."""
    return (rng.integers(0, 2, (n, b), dtype=np.int8) * 2 - 1).astype(np.float32)


def one_hot_labels(rng: np.random.Generator, n: int, L: int, balanced: bool = True) -> np.ndarray:
    if balanced:
        cls = np.arange(n) % L
        rng.shuffle(cls)
    else:
        cls = rng.integers(0, L, n)
    lab = np.zeros((n, L), dtype=np.int64)
    lab[np.arange(n), cls] = 1
    return lab


def multi_hot_labels(rng: np.random.Generator, n: int, L: int = 81) -> np.ndarray:
    rates = _nuswide_rates(L)
    lab = (rng.random((n, L), dtype=np.float32) < rates.astype(np.float32)[None, :]).astype(np.int64)
    empty = lab.sum(1) == 0  # every NUS-WIDE row kept by the reference lists has >= 1 label
    lab[empty, 0] = 1
    return lab


@lru_cache(maxsize=None)
def real_label_rows(dataset: str, split: str) -> np.ndarray:
    """Label matrix [n, L] (int64 0/1) of the reference's list file data_list/<dataset>/<split>.txt, from the committed
    bit-packed copy (oracle/gen_label_rows.py)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "label_rows.npz")
    with np.load(path) as z:
        bits, L = z[f"{dataset}/{split}_bits"], int(z[f"{dataset}/{split}_L"])
    return np.unpackbits(bits, axis=1)[:, :L].astype(np.int64)


def list_labels(rng: np.random.Generator, dataset: str, split: str, n: int) -> np.ndarray:
    """n label rows of a reference list: the rows themselves (in list order) while n fits, else a seeded bootstrap with
    replacement (C5's 2M database rows are resampled from the 168,692 real ones)."""
    rows = real_label_rows(dataset, split)
    if n <= rows.shape[0]:
        return rows[:n].copy()
    return rows[rng.integers(0, rows.shape[0], n)]


def proto_codes(rng: np.random.Generator, lab: np.ndarray, b: int, flip: float, proto: np.ndarray | None = None):
    """Class-correlated codes: prototype of the (first) class with i.i.d. bit flips; returns (codes, proto)."""
    L = lab.shape[1]
    if proto is None:
        proto = pm1_codes(rng, L, b)
    cls = lab.argmax(1)
    codes = proto[cls].copy()
    codes[rng.random(codes.shape, dtype=np.float32) < flip] *= -1
    return codes.astype(np.float32), proto


def make_workload(name_or_wl, *, nq: int | None = None, ndb: int | None = None, correlated: float | None = None):
    """Returns (workload, database, query) with .output float32 [N,b] and .label int64 [N,L].
    ``nq`` / ``ndb`` override the sizes (prefix-stable: the first rows are the same for any size because each
    array is drawn from its own generator).  ``correlated`` = bit-flip probability for class-correlated codes."""
    wl = WORKLOADS[name_or_wl] if isinstance(name_or_wl, str) else name_or_wl
    nq = wl.nq if nq is None else nq
    ndb = wl.ndb if ndb is None else ndb
    r_db_c, r_q_c = np.random.default_rng(1000 + wl.seed), np.random.default_rng(2000 + wl.seed)
    r_db_l, r_q_l = np.random.default_rng(3000 + wl.seed), np.random.default_rng(4000 + wl.seed)
    if wl.labels == "onehot":
        db_lab = one_hot_labels(r_db_l, ndb, wl.L)
        q_lab = one_hot_labels(r_q_l, nq, wl.L)
    elif wl.labels == "cifar10" and wl.L == 10:
        db_lab = list_labels(r_db_l, "cifar10", "database", ndb)
        q_lab = list_labels(r_q_l, "cifar10", "test", nq)
    elif wl.labels == "nuswide" and wl.L == 81:
        db_lab = list_labels(r_db_l, "nuswide_81", "database", ndb)
        q_lab = list_labels(r_q_l, "nuswide_81", "test", nq)
    else:
        db_lab = multi_hot_labels(r_db_l, ndb, wl.L)
        q_lab = multi_hot_labels(r_q_l, nq, wl.L)
    if correlated is None:
        db_codes = pm1_codes(r_db_c, ndb, wl.b)
        q_codes = pm1_codes(r_q_c, nq, wl.b)
    else:
        db_codes, proto = proto_codes(r_db_c, db_lab, wl.b, correlated)
        q_codes, _ = proto_codes(r_q_c, q_lab, wl.b, correlated, proto)
    eff = Workload(wl.name, nq, ndb, wl.b, wl.L, min(wl.R, ndb), wl.labels, wl.seed, wl.note)
    return eff, SimpleNamespace(output=db_codes, label=db_lab), SimpleNamespace(output=q_codes, label=q_lab)
