#!/usr/bin/env python
"""Generates hashgan_b200/data/label_rows.npz from the reference's list files (run in the build container only).

SURVEY.md section 8(d) asks the C1 / C5 benchmarks to use the REAL label rows of the reference's lists
(/root/reference/data_list/cifar10/{database,test}.txt -- 10-way one-hot, exactly balanced -- and
data_list/nuswide_81/{database,test}.txt -- 81-way multi-label, 168,692 + 5,000 rows) instead of independent per-class
Bernoulli draws, which lose the label co-occurrence.  /root/reference does not exist on the GPU box, so the label matrices
(only the 0/1 columns after the image path; lib/dataloader.py:45 parses exactly these) are committed bit-packed:

    <set>/<split>_bits : uint8 [n, ceil(L/8)]  np.packbits(labels, axis=1)
    <set>/<split>_L    : label width

hashgan_b200/synthetic.py reads the file; tests/test_synthetic.py checks the row statistics SURVEY.md quotes.
"""
import os
import sys

import numpy as np

REF = os.environ.get("HASHGAN_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_labels(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            parts = line.split()
            if parts:
                rows.append([int(x) for x in parts[1:]])  # lib/dataloader.py:45: everything after the path
    return np.array(rows, dtype=np.uint8)


def main():
    out = {}
    for name in ("cifar10", "nuswide_81"):
        for split in ("database", "test"):
            lab = read_labels(os.path.join(REF, "data_list", name, split + ".txt"))
            assert set(np.unique(lab)) <= {0, 1}
            out[f"{name}/{split}_bits"] = np.packbits(lab, axis=1)
            out[f"{name}/{split}_L"] = np.int32(lab.shape[1])
            print(name, split, lab.shape, "labels/row %.3f" % lab.sum(1).mean())
    dst = os.path.join(ROOT, "hashgan_b200", "data", "label_rows.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
