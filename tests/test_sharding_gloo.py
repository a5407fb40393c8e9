"""Host logic of the multi-GPU path (hashgan_b200/sharding.py) on CPU: world_size 2, gloo backend.
The packers and the ranker are injected with the oracle (tests may use it as the checker); what is
under test is the sharding / all-gather / AP-gather bookkeeping that the NCCL path shares."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(ndb, nq, b, L, seed):
    rng = np.random.default_rng(seed)
    dbc = (rng.integers(0, 2, (ndb, b)) * 2 - 1).astype(np.float32)
    qc = (rng.integers(0, 2, (nq, b)) * 2 - 1).astype(np.float32)
    dl = np.eye(L, dtype=np.int64)[rng.integers(0, L, ndb)]
    ql = np.eye(L, dtype=np.int64)[rng.integers(0, L, nq)]
    return dbc, dl, qc, ql


def _worker(rank, world, port, ndb, nq, b, L, R, seed, out_dir):
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace as NS
        from hashgan_b200.sharding import ShardedMAPs, row_shard
        from oracle import maps_oracle

        co = helpers.COracle()
        dbc, dl, qc, ql = _make(ndb, nq, b, L, seed)
        lo, hi = row_shard(ndb, rank, world)
        qlo, qhi = row_shard(nq, rank, world)

        W, LW = (b + 31) // 32, (L + 31) // 32

        def pack_rows(out, lab):  # [code words | label words], the layout the device path all-gathers
            rows = np.concatenate([maps_oracle.pack_sign_bits(np.asarray(out)), maps_oracle.pack_label_bits(np.asarray(lab))], 1)
            return torch.from_numpy(np.ascontiguousarray(rows).view(np.int32))

        def rank_fn(q_rows, db_rows, b_, L_, R_):
            qn, dn = q_rows.numpy().view(np.uint32), db_rows.numpy().view(np.uint32)
            qc, ql = np.ascontiguousarray(qn[:, :W]), np.ascontiguousarray(qn[:, W:W + LW])
            dc, dl_ = np.ascontiguousarray(dn[:, :W]), np.ascontiguousarray(dn[:, W:W + LW])
            ap = np.empty(len(qn), dtype=np.float64)
            rc = co.lib.hgo_hamming_map(qc.ctypes.data, ql.ctypes.data, len(qn), dc.ctypes.data, dl_.ctypes.data, len(dn),
                                        b_, L_, R_, ap.ctypes.data, None, None, None, 1)
            assert rc == 0
            return torch.from_numpy(ap)

        m = ShardedMAPs(R, pack_rows=pack_rows, rank_fn=rank_fn)
        ap = m.per_query_ap_device(NS(output=dbc[lo:hi], label=dl[lo:hi]), NS(output=qc[qlo:qhi], label=ql[qlo:qhi])).numpy()
        val = m.get_maps_by_feature(NS(output=dbc[lo:hi], label=dl[lo:hi]), NS(output=qc[qlo:qhi], label=ql[qlo:qhi]))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ap=ap, val=val, counts=np.array(m.last_counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ndb,nq", [(1000, 40), (1001, 37)])
def test_two_rank_sharded_map_equals_single_process(tmp_path, c_oracle, ndb, nq):
    from types import SimpleNamespace as NS

    b, L, R, seed, world = 48, 6, 200, 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ndb, nq, b, L, R, seed, str(tmp_path)), nprocs=world, join=True)
    dbc, dl, qc, ql = _make(ndb, nq, b, L, seed)
    ref_ap, _, _, _ = c_oracle.hamming_map(NS(output=dbc, label=dl), NS(output=qc, label=ql), R)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for o in outs:
        assert np.array_equal(o["ap"], ref_ap, equal_nan=True)     # rank order == global query order
        assert float(o["val"]) == float(np.mean(ref_ap[~np.isnan(ref_ap)]))
        assert int(o["counts"].sum()) == ndb
    assert float(outs[0]["val"]) == float(outs[1]["val"])


def _real_worker(rank, world, port, ndb, nq, b, L, R, seed, out_dir):
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace as NS
        from hashgan_b200.sharding import ShardedMAPs, row_shard
        from oracle import maps_oracle

        rng = np.random.default_rng(seed)
        dbf = np.tanh(rng.normal(size=(ndb, b))).astype(np.float32)
        qf = np.tanh(rng.normal(size=(nq, b))).astype(np.float32)
        dl = np.eye(L, dtype=np.int64)[rng.integers(0, L, ndb)]
        ql = np.eye(L, dtype=np.int64)[rng.integers(0, L, nq)]
        lo, hi = row_shard(ndb, rank, world)
        qlo, qhi = row_shard(nq, rank, world)
        LW = (L + 31) // 32

        def pack_rows(out, lab):  # only the label words matter on the real-valued path
            return torch.from_numpy(np.ascontiguousarray(maps_oracle.pack_label_bits(np.asarray(lab))).view(np.int32))

        def unpack(rows):
            bits = np.unpackbits(rows.numpy().view(np.uint8), axis=1, bitorder="little")[:, :L]
            return bits.astype(np.int64)

        def rank_real_fn(q_feat, q_rows, db_feat, db_rows, b_, L_, R_):
            return torch.from_numpy(maps_oracle.per_query_ap(db_feat.numpy(), unpack(db_rows), q_feat.numpy(), unpack(q_rows), R_, tie="stable"))

        m = ShardedMAPs(R, pack_rows=pack_rows, rank_real_fn=rank_real_fn, binarize=False)
        ap = m.per_query_ap_device(NS(output=dbf[lo:hi], label=dl[lo:hi]), NS(output=qf[qlo:qhi], label=ql[qlo:qhi])).numpy()
        np.savez(os.path.join(out_dir, f"real{rank}.npz"), ap=ap, dbf=dbf, qf=qf, dl=dl, ql=ql)
    finally:
        dist.destroy_process_group()


def test_two_rank_real_valued_ranking_equals_single_process(tmp_path):
    """EVAL.BINARIZE False under a process group: the raw features are all-gathered next to the packed label rows
    and ranked by inner product -- the single-process per-query APs, bit for bit, on every rank."""
    from oracle import maps_oracle

    ndb, nq, b, L, R, seed, world = 301, 23, 16, 5, 50, 9, 2
    port = _free_port()
    mp.spawn(_real_worker, args=(world, port, ndb, nq, b, L, R, seed, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"real{r}.npz") for r in range(world)]
    o = outs[0]
    want = maps_oracle.per_query_ap(o["dbf"], o["dl"], o["qf"], o["ql"], R, tie="stable")
    for r in range(world):
        assert np.array_equal(outs[r]["ap"], want, equal_nan=True)


def test_batch_range_skips_the_fetch_of_other_ranks_batches():
    """forward_all(shard=...) asks the loader for its block only: images outside [lo, hi) are never fetched, and the
    batches it gets are the ones the full epoch yields at those positions."""
    from hashgan_b200 import dataloader as dl

    fetched = []

    def fetch(idx):
        fetched.append(np.array(idx))
        return np.zeros((len(idx), 2, 2, 3), np.uint8) + np.asarray(idx, np.uint8)[:, None, None, None], np.asarray(idx)[:, None]

    np.random.seed(4)
    full = list(dl._epoch(50, 8, fetch))
    n_full = len(fetched)
    fetched.clear()
    np.random.seed(4)
    part = list(dl._epoch(50, 8, fetch, batch_range=(2, 5)))
    assert n_full == 7 and len(fetched) == 3 and len(part) == 3
    for (a, la), (b_, lb) in zip(part, full[2:5]):
        assert np.array_equal(a, b_) and np.array_equal(la, lb)


def _eval_cfg(db_size, test_size, batch, b, L, R):
    from types import SimpleNamespace as NS

    return NS(MODEL=NS(HASH_DIM=b), DATA=NS(DB_SIZE=db_size, TEST_SIZE=test_size, LABEL_DIM=L, MAP_R=R), TRAIN=NS(BATCH_SIZE=batch),
              EVAL=NS(SEED=3, BINARIZE=True))


def _fake_encoder(wh, b):
    proj = np.random.default_rng(11).normal(size=(3 * wh * wh, b)).astype(np.float32)
    return lambda image: torch.from_numpy(np.tanh((np.asarray(image, dtype=np.float32) / 255.0 - 0.5) @ proj))


def _eval_worker(rank, world, port, sizes, out_dir):
    sys.path.insert(0, helpers.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hashgan_b200.dataloader import SyntheticDataloader
        from hashgan_b200.evaluate import evaluate
        from hashgan_b200.sharding import ShardedMAPs
        from oracle import maps_oracle

        db_size, test_size, batch, wh, b, L, R = sizes
        co = helpers.COracle()
        W, LW = (b + 31) // 32, (L + 31) // 32

        def pack_rows(out, lab):
            rows = np.concatenate([maps_oracle.pack_sign_bits(np.asarray(out)), maps_oracle.pack_label_bits(np.asarray(lab))], 1)
            return torch.from_numpy(np.ascontiguousarray(rows).view(np.int32))

        def rank_fn(q_rows, db_rows, b_, L_, R_):
            qn, dn = q_rows.numpy().view(np.uint32), db_rows.numpy().view(np.uint32)
            ap = np.empty(len(qn), dtype=np.float64)
            qc, ql = np.ascontiguousarray(qn[:, :W]), np.ascontiguousarray(qn[:, W:W + LW])
            dc, dl_ = np.ascontiguousarray(dn[:, :W]), np.ascontiguousarray(dn[:, W:W + LW])
            assert co.lib.hgo_hamming_map(qc.ctypes.data, ql.ctypes.data, len(qn), dc.ctypes.data, dl_.ctypes.data, len(dn), b_, L_, R_,
                                          ap.ctypes.data, None, None, None, 1) == 0
            return torch.from_numpy(ap)

        loader = SyntheticDataloader(batch, wh, L, {"database": db_size, "test": test_size}, seed=5)
        val = evaluate(_fake_encoder(wh, b), loader, _eval_cfg(db_size, test_size, batch, b, L, R),
                       metric=ShardedMAPs(R, pack_rows=pack_rows, rank_fn=rank_fn))
        np.savez(os.path.join(out_dir, f"eval{rank}.npz"), val=val)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("db_size,test_size,batch", [(203, 37, 16), (64, 16, 16), (50, 40, 32)])
def test_sharded_evaluate_equals_single_process(tmp_path, c_oracle, db_size, test_size, batch):
    """evaluate() under a process group: contiguous blocks of batches per rank, `size` truncation on the last block,
    identical shuffle on every rank -> the single-process value, bit for bit (also with more ranks than batches)."""
    from types import SimpleNamespace as NS
    from hashgan_b200.dataloader import SyntheticDataloader
    from hashgan_b200.evaluate import forward_all

    wh, b, L, R, world = 4, 40, 5, 30, 2
    port = _free_port()
    mp.spawn(_eval_worker, args=(world, port, (db_size, test_size, batch, wh, b, L, R), str(tmp_path)), nprocs=world, join=True)
    cfg = _eval_cfg(db_size, test_size, batch, b, L, R)
    loader = SyntheticDataloader(batch, wh, L, {"database": db_size, "test": test_size}, seed=5)
    enc = _fake_encoder(wh, b)
    np.random.seed(cfg.EVAL.SEED)  # the shuffle the sharded run used
    db = forward_all(enc, loader.db_gen, db_size, cfg)
    te = forward_all(enc, loader.test_gen, test_size, cfg)
    assert db.output.shape == (db_size, b) and te.output.shape == (test_size, b)
    ref_ap, _, _, _ = c_oracle.hamming_map(NS(output=db.output.numpy(), label=db.label), NS(output=te.output.numpy(), label=te.label), R)
    want = float(np.mean(ref_ap[~np.isnan(ref_ap)]))
    vals = [float(np.load(tmp_path / f"eval{r}.npz")["val"]) for r in range(world)]
    assert vals[0] == vals[1] == want


def test_shard_bounds_cover_everything():
    from hashgan_b200.sharding import shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            bounds = shard_bounds(n, w)
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
