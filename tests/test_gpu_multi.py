"""Multi-GPU parity ON HARDWARE: two ranks, NCCL, the fused pack+push exchange over symmetric memory
(hg_pack_rows_push) and the all-gather fallback, checked against oracle/hamming_oracle.c.  Skipped below 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        from types import SimpleNamespace as NS
        from hashgan_b200.sharding import ShardedMAPs, shard_bounds
        from hashgan_b200.synthetic import make_workload

        name, nq, ndb, R, symmetric, binarize, poison = case
        wl, db, q = make_workload(name, nq=nq, ndb=ndb, correlated=0.3 if name == "C2" else None)
        if not binarize:
            rng = np.random.default_rng(3)
            db.output = np.tanh(rng.normal(size=db.output.shape)).astype(np.float32)
            q.output = np.tanh(rng.normal(size=q.output.shape)).astype(np.float32)
        db_b, q_b = shard_bounds(ndb, world), shard_bounds(nq, world)
        (lo, hi), (qlo, qhi) = db_b[rank], q_b[rank]
        lab = db.label[lo:hi].copy()
        if poison and rank == 1:
            lab[3, 0] = 2                       # a label that is not 0/1 on ONE rank must fail the call on EVERY rank
        m = ShardedMAPs(R, device=device, db_counts=[b - a for a, b in db_b], query_counts=[b - a for a, b in q_b],
                        symmetric=symmetric, binarize=binarize)
        err = ""
        try:
            ap = m.per_query_ap_device(NS(output=db.output[lo:hi], label=lab), NS(output=q.output[qlo:qhi], label=q.label[qlo:qhi])).cpu().numpy()
            val = m.get_maps_by_feature(NS(output=db.output[lo:hi], label=lab), NS(output=q.output[qlo:qhi], label=q.label[qlo:qhi]))
        except ValueError as exc:
            ap, val, err = np.zeros(0), np.nan, str(exc)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ap=ap, val=val, err=err, exchange=m.exchange)
    finally:
        dist.destroy_process_group()


CASES = [
    ("C4", 300, 60000, 2000, True, True, False),     # 64-bit: fused push, tensor-core select
    ("C5", 120, 40000, 1500, True, True, False),     # 128-bit, 81-way multi-label
    ("C2", 200, 30000, 1000, True, True, False),     # 48-bit: ragged hash length -> all-gather fallback of the exchange
    ("C4", 100, 20001, 500, False, True, False),     # NCCL all-gather, ragged shard sizes
    ("C1", 64, 5000, 300, False, False, False),      # EVAL.BINARIZE False under a process group
    ("C4", 64, 8000, 300, True, True, True),         # bad label on one rank
]


@pytest.mark.skipif(_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-{'push' if c[4] else 'gather'}-{'ham' if c[5] else 'real'}{'-badlabel' if c[6] else ''}")
def test_two_gpu_sharded_map_equals_oracle(tmp_path, c_oracle, case):
    import torch.multiprocessing as mp
    from types import SimpleNamespace as NS
    from hashgan_b200.synthetic import make_workload
    from oracle import maps_oracle

    name, nq, ndb, R, symmetric, binarize, poison = case
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    if poison:
        assert all("0/1" in str(o["err"]) for o in outs), [str(o["err"]) for o in outs]
        return
    wl, db, q = make_workload(name, nq=nq, ndb=ndb, correlated=0.3 if name == "C2" else None)
    if binarize:
        want, _, _, _ = c_oracle.hamming_map(db, q, R)
    else:
        rng = np.random.default_rng(3)
        dbf = np.tanh(rng.normal(size=db.output.shape)).astype(np.float32)
        qf = np.tanh(rng.normal(size=q.output.shape)).astype(np.float32)
        want = maps_oracle.per_query_ap(dbf, db.label, qf, q.label, R, tie="stable")
    for o in outs:
        assert str(o["err"]) == ""
        assert np.array_equal(np.isnan(o["ap"]), np.isnan(want))
        assert np.nanmax(np.abs(o["ap"] - want)) <= (1e-12 if binarize else 1e-6)
        assert abs(float(o["val"]) - float(np.mean(want[~np.isnan(want)]))) <= (1e-12 if binarize else 1e-6)
    assert float(outs[0]["val"]) == float(outs[1]["val"])                   # every rank returns the same mean
    if symmetric and wl.b % 32 == 0:
        assert all(str(o["exchange"]) == "push" for o in outs)
