"""Input pipeline of the evaluation path -- host mirror of lib/dataloader.py (the encoder's input-layout contract).

Reference behaviour kept (lib/dataloader.py:22-140):
  * list files `<list_root>/<split>.txt`, one line per image: `relative/path.jpg l0 l1 ... lL-1`   (:24, :38-45)
  * cv2.imread (BGR) -> cv2.resize(..., INTER_AREA) to wh x wh                                       (:38-42)
  * one epoch = ceil(n / batch) batches over a fresh np.random.shuffle permutation; the last batch wraps around to
    the start of the permutation so every batch is full                                             (:90-104)
  * NHWC -> NCHW, BGR -> RGB, flatten to [batch, 3*wh*wh]                                           (:109-113)
  * generators: train_gen / test_gen / db_gen / unlabeled_db_gen, inf_gen                          (:118-140)
Differences: images are cached as uint8 (the reference caches the same arrays), an unreadable image raises instead of
being silently dropped (the reference's bare `except` leaves a short batch that breaks the later reshape, :55-69).
`SyntheticDataloader` serves seeded random images with the label statistics of the list files when the image
directory is absent (config key EVAL.SYNTHETIC).
"""
from __future__ import annotations

import math
import os

import numpy as np

__all__ = ["Dataset", "Dataloader", "SyntheticDataloader"]


class Dataset:
    def __init__(self, list_path, image_root, height_width=256):
        with open(list_path, "r") as fh:
            self.lines = [ln for ln in fh.read().splitlines() if ln.strip()]
        self.image_root = image_root
        self.n_samples = len(self.lines)
        self.height_width = height_width
        self._img = [None] * self.n_samples
        self._label = np.array([[int(j) for j in ln.split()[1:]] for ln in self.lines], dtype=np.int64)  # :44-45

    def read_image_at(self, index):
        import cv2

        path = os.path.join(self.image_root, self.lines[index].split()[0])
        img = cv2.imread(path)
        if img is None:
            raise FileNotFoundError(f"cannot open {path}")
        return cv2.resize(img, (self.height_width, self.height_width), interpolation=cv2.INTER_AREA)

    def data(self, index):
        imgs = []
        for i in index:
            if self._img[i] is None:
                self._img[i] = self.read_image_at(i)
            imgs.append(self._img[i])
        return np.asarray(imgs), self._label[np.asarray(index)]


def _epoch(n_samples, batch_size, fetch, batch_range=None):
    """lib/dataloader.py:90-114.  `batch_range` = (lo, hi) (additive, multi-GPU evaluation): only batches lo <= i < hi of the
    epoch are fetched and yielded -- the permutation is drawn in full, so every rank sees the same epoch, but a rank never
    decodes the images of another rank's block."""
    perm = np.arange(n_samples)
    np.random.shuffle(perm)
    pos = 0
    lo, hi = batch_range if batch_range is not None else (0, None)
    for i in range(int(math.ceil(n_samples / batch_size))):
        start = pos
        pos += batch_size
        if i < lo or (hi is not None and i >= hi):
            continue
        if pos > n_samples:  # wrap around: the last batch is completed from the head of the permutation
            idx = np.concatenate([perm[start:], perm[:pos - n_samples]])
        else:
            idx = perm[start:pos]
        data, label = fetch(idx)
        data = np.transpose(data, (0, 3, 1, 2))[:, ::-1, :, :]  # NHWC -> NCHW, BGR -> RGB
        yield np.reshape(data, (batch_size, -1)), label


class Dataloader:
    def __init__(self, batch_size, width_height, list_root, image_root):
        self.batch_size = batch_size
        self.width_height = width_height
        self.data_root = list_root
        self.image_root = image_root

    def data_generator(self, split):
        ds = Dataset(os.path.join(self.data_root, split + ".txt"), self.image_root, self.width_height)
        return lambda batch_range=None: _epoch(ds.n_samples, self.batch_size, ds.data, batch_range)

    @property
    def train_gen(self):
        return self.data_generator("train")

    @property
    def test_gen(self):
        return self.data_generator("test")

    @property
    def db_gen(self):
        return self.data_generator("database")

    @property
    def unlabeled_db_gen(self):
        return self.data_generator("database_nolabel")

    @staticmethod
    def inf_gen(gen):
        def generator():
            while True:
                for images, labels in gen():
                    return images, labels
        return generator


class SyntheticDataloader:
    """Same generator protocol, seeded uint8 images (BGR HWC like cv2) and one-hot / multi-hot labels."""

    def __init__(self, batch_size, width_height, label_dim, sizes, seed=0):
        self.batch_size, self.width_height, self.label_dim, self.seed = batch_size, width_height, label_dim, seed
        self.sizes = dict(sizes)  # split -> number of images
        self._cache = {}

    def _split(self, split):
        if split not in self._cache:
            from .synthetic import multi_hot_labels, one_hot_labels

            n = self.sizes[split]
            rng = np.random.default_rng(self.seed + sum(map(ord, split)))
            imgs = rng.integers(0, 256, (n, self.width_height, self.width_height, 3), dtype=np.uint8)
            lab = one_hot_labels(rng, n, self.label_dim) if self.label_dim <= 20 else multi_hot_labels(rng, n, self.label_dim)
            self._cache[split] = (imgs, lab)
        return self._cache[split]

    def data_generator(self, split):
        imgs, lab = self._split(split)
        return lambda batch_range=None: _epoch(len(imgs), self.batch_size, lambda idx: (imgs[idx], lab[idx]), batch_range)

    test_gen = property(lambda self: self.data_generator("test"))
    db_gen = property(lambda self: self.data_generator("database"))
    train_gen = property(lambda self: self.data_generator("train"))
