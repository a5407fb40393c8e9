"""Evaluation loop -- host mirror of main.py:151-164 (`forward_all`, `evaluate`).

`session, model` of the reference are replaced by one encoder handle (hashgan_b200.encoder.AlexNetHashEncoder);
generator protocol, `size` truncation (the last batch wraps around, lib/dataloader.py:99-104) and the returned
{output, label} record are kept.  Outputs stay on the GPU between the encoder and the metric.

Multi-GPU (new: the reference is single-GPU, main.py:263): under torchrun every rank encodes a contiguous block of the
batches of each split -- the global row order, and with it the tie order of the ranking, is the single-process one --
and the metric is hashgan_b200.sharding.ShardedMAPs.  The loader's per-epoch shuffle
(lib/dataloader.py:93-94) is seeded identically on all ranks so that they agree on the permutation.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from .metric import MAPs

__all__ = ["forward_all", "evaluate", "periodic_evaluate", "scalar_summary", "ScalarLog"]


def _dist_info(group=None):
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def _block(size: int, batch: int, shard):
    """Batches [lo, hi) and rows [row_lo, row_hi) of this rank for a split of `size` rows."""
    from .sharding import shard_bounds

    nb = int(math.ceil(size / batch))
    if shard is None:
        return 0, nb, 0, size
    rank, world = shard
    lo, hi = shard_bounds(nb, world)[rank]
    return lo, hi, min(lo * batch, size), min(hi * batch, size)


def forward_all(encoder, data_generator, size, cfg, shard=None):
    """main.py:151-158: one epoch of `data_generator`, encode every batch, stack, truncate to `size`.
    shard = (rank, world): only this rank's contiguous block of batches is encoded (rows [row_lo, row_hi) of the split)."""
    import torch

    lo, hi, row_lo, row_hi = _block(size, cfg.TRAIN.BATCH_SIZE, shard)
    outputs, labels = [], []
    first, batches = 0, None
    if shard is not None:
        try:  # this package's loaders skip the fetch (image decode) of other ranks' batches; any other generator is filtered below
            batches, first = data_generator(batch_range=(lo, hi)), lo
        except TypeError:
            batches = None
    if batches is None:
        batches = data_generator()
    for i, (image, label) in enumerate(batches, start=first):
        if i >= hi:
            break
        if i < lo:
            continue
        outputs.append(encoder(image))          # main.py:154-155: session.run(model.disc_real_acgan, feed_dict)
        labels.append(np.asarray(label))
    n = row_hi - row_lo
    if outputs:
        output = torch.cat(outputs, 0).reshape(-1, cfg.MODEL.HASH_DIM)[:n]
        label = np.concatenate(labels, 0).reshape(-1, cfg.DATA.LABEL_DIM)[:n]
    else:  # more ranks than batches
        output = torch.empty((0, cfg.MODEL.HASH_DIM), dtype=torch.float32, device=getattr(encoder, "device", "cpu"))
        label = np.empty((0, cfg.DATA.LABEL_DIM), dtype=np.int64)
    return SimpleNamespace(output=output, label=label)


def evaluate(encoder, dataloader, cfg, metric=None, group=None, precision_recall: bool = False):
    """main.py:161-164.  With an initialised process group of more than one rank the splits are encoded in contiguous blocks
    per rank and ranked by ShardedMAPs; every rank returns the same value.  precision_recall=True (single process) returns the
    dict of MAPs.precision_recall (precision@R, recall@R, mAP on one ranking) instead of the scalar."""
    rank, world = _dist_info(group)
    ev = getattr(cfg, "EVAL", None)
    shard = (rank, world) if world > 1 else None
    if shard is not None or bool(getattr(ev, "DETERMINISTIC", False)):
        # the loader shuffles every epoch (lib/dataloader.py:93-94) and the row order decides ties: the same shuffle on every
        # rank, and from run to run in the deterministic mode (1 GPU and N GPUs then print the same map_val)
        np.random.seed(int(getattr(ev, "SEED", 0)))
    db = forward_all(encoder, dataloader.db_gen, cfg.DATA.DB_SIZE, cfg, shard)
    test = forward_all(encoder, dataloader.test_gen, cfg.DATA.TEST_SIZE, cfg, shard)
    if world > 1:
        from .sharding import ShardedMAPs

        B = cfg.TRAIN.BATCH_SIZE
        db_counts = [(lambda t: t[3] - t[2])(_block(cfg.DATA.DB_SIZE, B, (r, world))) for r in range(world)]
        q_counts = [(lambda t: t[3] - t[2])(_block(cfg.DATA.TEST_SIZE, B, (r, world))) for r in range(world)]
        if metric is None:
            on_gpu = getattr(db.output, "is_cuda", False)
            # EVAL.BINARIZE False = the reference's literal ranking of the raw tanh outputs (lib/metric.py:13-14), sharded as well
            binarize = bool(getattr(ev, "BINARIZE", True))
            metric = ShardedMAPs(cfg.DATA.MAP_R, group, db_counts=db_counts, query_counts=q_counts, symmetric=bool(on_gpu) and binarize,
                                 binarize=binarize)
        else:
            if getattr(metric, "db_counts", None) is None:
                metric.db_counts = db_counts
            if getattr(metric, "query_counts", None) is None:
                metric.query_counts = q_counts
    elif metric is None:
        # EVAL.BINARIZE False = the reference's literal ranking of the raw tanh outputs (lib/metric.py:13-14)
        metric = MAPs(cfg.DATA.MAP_R, binarize=bool(getattr(ev, "BINARIZE", True)))
    if precision_recall:
        if not hasattr(metric, "precision_recall"):
            raise NotImplementedError("precision@R / recall@R are computed by MAPs (single process); run without torchrun")
        return metric.precision_recall(db, test)
    return metric.get_maps_by_feature(db, test)


def scalar_summary(tag, value):
    """lib/util.py:60-61 builds a one-value tf.Summary; without TensorFlow the record is the (tag, simple_value) pair itself."""
    return SimpleNamespace(tag=tag, simple_value=float(value))


class ScalarLog:
    """Stand-in for the tf.summary.FileWriter of main.py:176: `add_summary(summary, step)` appends one JSON line per scalar to
    <LOG_DIR>/scalars.jsonl (TensorBoard event files are out of scope; the call signature is the reference's)."""

    def __init__(self, log_dir):
        import os

        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, "scalars.jsonl")

    def add_summary(self, summary, global_step=None):
        import json

        with open(self.path, "a") as fh:
            fh.write(json.dumps({"tag": summary.tag, "value": summary.simple_value, "step": None if global_step is None else int(global_step)}) + "\n")


def periodic_evaluate(iteration, encoder, dataloader, cfg, summary_writer=None, metric=None, group=None):
    """The in-training evaluation hook, main.py:236-240: a training loop calls this once per iteration; every
    TRAIN.EVAL_FREQUENCY iterations and on the last one (TRAIN.ITERS) it evaluates, prints `map_val: ...` and logs the scalar
    `mAP_feature` at step `iteration` -- exactly the reference's condition, line and tag.  Returns map_val, or None when this
    iteration does not evaluate.  `encoder` is the handle that stands for (session, model): a trainer refreshes its weights
    before the call (training itself is out of scope)."""
    due = (iteration + 1) % cfg.TRAIN.EVAL_FREQUENCY == 0 or iteration + 1 == cfg.TRAIN.ITERS   # main.py:237
    if not due:
        return None
    map_val = evaluate(encoder, dataloader, cfg, metric=metric, group=group)                      # main.py:238
    rank, _ = _dist_info(group)
    if rank == 0:
        print('map_val: {}'.format(map_val))                                                      # main.py:239
        if summary_writer is not None:
            summary_writer.add_summary(scalar_summary("mAP_feature", map_val), iteration)         # main.py:240
    return map_val
