"""Evaluation loop -- host mirror of main.py:151-164 (`forward_all`, `evaluate`).

`session, model` of the reference are replaced by one encoder handle (hashgan_b200.encoder.AlexNetHashEncoder);
generator protocol, `size` truncation (the last batch wraps around, lib/dataloader.py:99-104) and the returned
{output, label} record are kept.  Outputs stay on the GPU between the encoder and the metric.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from .metric import MAPs

__all__ = ["forward_all", "evaluate"]


def forward_all(encoder, data_generator, size, cfg):
    """main.py:151-158: one epoch of `data_generator`, encode every batch, stack, truncate to `size`."""
    import torch

    outputs, labels = [], []
    for image, label in data_generator():
        outputs.append(encoder(image))          # main.py:154-155: session.run(model.disc_real_acgan, feed_dict)
        labels.append(np.asarray(label))
    output = torch.cat(outputs, 0).reshape(-1, cfg.MODEL.HASH_DIM)[:size]
    label = np.concatenate(labels, 0).reshape(-1, cfg.DATA.LABEL_DIM)[:size]
    return SimpleNamespace(output=output, label=label)


def evaluate(encoder, dataloader, cfg, metric=None):
    """main.py:161-164."""
    db = forward_all(encoder, dataloader.db_gen, cfg.DATA.DB_SIZE, cfg)
    test = forward_all(encoder, dataloader.test_gen, cfg.DATA.TEST_SIZE, cfg)
    if metric is None:
        ev = getattr(cfg, "EVAL", None)
        # EVAL.BINARIZE False = the reference's literal ranking of the raw tanh outputs (lib/metric.py:13-14)
        metric = MAPs(cfg.DATA.MAP_R, binarize=bool(getattr(ev, "BINARIZE", True)))
    return metric.get_maps_by_feature(db, test)
