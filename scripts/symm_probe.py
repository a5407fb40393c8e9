"""N-GPU check of the fused pack -> peer-memory push (hg_pack_rows_push + symmetric-memory barrier) against
pack + NCCL all-gather: bit-identical rows, time of both.  torchrun --nproc-per-node N scripts/symm_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hashgan_b200.metric import pack_rows
from hashgan_b200.sharding import SymmetricRows, gather_rows, row_shard, shard_bounds
from hashgan_b200.synthetic import make_workload

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for name in ("C4", "C5"):
    wl, db, q = make_workload(name, nq=64)
    lo, hi = row_shard(wl.ndb, rank, world)
    counts = [b - a for a, b in shard_bounds(wl.ndb, world)]
    f = torch.from_numpy(db.output[lo:hi]).to(dev); l = torch.from_numpy(db.label[lo:hi]).to(dev)
    ref, _ = gather_rows(pack_rows(f, l, dev), counts=counts)
    sr = SymmetricRows(wl.ndb, wl.b, wl.L, dev)
    for it in range(3):
        got = sr.pack(f, l, lo)
        torch.cuda.synchronize()
        assert torch.equal(got, ref), f"{name}: pushed rows differ from pack + all-gather (iteration {it})"
    def timeit(fn, n=20):
        for _ in range(3): fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    t_nccl = timeit(lambda: gather_rows(pack_rows(f, l, dev), counts=counts))
    t_push = timeit(lambda: sr.pack(f, l, lo))
    if rank == 0:
        print(f"{name} x{world}: rows identical; pack + NCCL all-gather {t_nccl*1e3:.1f} us, fused pack+push+barrier {t_push*1e3:.1f} us", flush=True)
dist.destroy_process_group()
