#!/usr/bin/env python
"""TEST INFRASTRUCTURE: golden vectors for the host input pipeline, made by running the UNMODIFIED reference
loader (/root/reference/lib/dataloader.py: list-file parser, cv2 decode + INTER_AREA resize, shuffled epoch with wrap-around last
batch, NHWC->NCHW, BGR->RGB, flatten) on a tiny synthetic image folder.

    python oracle/gen_golden_dataloader.py        # writes tests/golden/dataloader_golden.npz

The fixture stores the source images (the test writes them to lossless PNG files itself), the list files and every batch the
reference yields for two epochs under np.random.seed(7).  Needs /root/reference and cv2; runs in the build container only."""
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_folder(root, images, labels, split):
    os.makedirs(os.path.join(root, "img", "sub"), exist_ok=True)
    os.makedirs(os.path.join(root, "lists"), exist_ok=True)
    lines = []
    for i, (img, lab) in enumerate(zip(images, labels)):
        rel = f"sub/im_{split}_{i}.png" if i % 2 else f"im_{split}_{i}.png"
        cv2.imwrite(os.path.join(root, "img", rel), img)
        lines.append(rel + " " + " ".join(str(int(v)) for v in lab))
    with open(os.path.join(root, "lists", split + ".txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return lines


def main():
    sys.path.insert(0, "/root/reference")
    from lib.dataloader import Dataloader  # the reference itself

    rng = np.random.default_rng(42)
    blob = {}
    with tempfile.TemporaryDirectory() as root:
        for split, n in (("database", 11), ("test", 5)):
            sizes = [(int(rng.integers(9, 40)), int(rng.integers(9, 40))) for _ in range(n)]
            images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
            labels = (rng.random((n, 6)) < 0.4).astype(np.int64)
            lines = build_folder(root, images, labels, split)
            blob[f"{split}/n"] = np.int64(n)
            for i, img in enumerate(images):
                blob[f"{split}/img{i}"] = img
            blob[f"{split}/lines"] = np.array(lines)
        for wh, batch in ((8, 4), (16, 3)):
            dl = Dataloader(batch, wh, os.path.join(root, "lists"), os.path.join(root, "img"))
            np.random.seed(7)
            for split, gen in (("database", dl.db_gen), ("test", dl.test_gen)):
                for epoch in range(2):  # second epoch: the reference serves cached arrays (its _status switch)
                    for k, (data, label) in enumerate(gen()):
                        blob[f"out/{wh}_{batch}/{split}/e{epoch}/b{k}/data"] = np.asarray(data)
                        blob[f"out/{wh}_{batch}/{split}/e{epoch}/b{k}/label"] = np.asarray(label)
    out = os.path.join(ROOT, "tests", "golden", "dataloader_golden.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes,", len(blob), "arrays")


if __name__ == "__main__":
    main()
