import time, numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from hashgan_b200.metric import MAPs, pack_rows, ip_map_device
from hashgan_b200.synthetic import make_workload
from hashgan_b200 import _native
wl, db, q = make_workload("C4")
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
dbf = torch.from_numpy(np.tanh(db.output * 0.7 + rng.normal(size=db.output.shape).astype(np.float32))).to(dev)
qf = torch.from_numpy(np.tanh(q.output * 0.7 + rng.normal(size=q.output.shape).astype(np.float32))).to(dev)
dbr = pack_rows(dbf, db.label); qr = pack_rows(qf, q.label)
for nq in (1024, 10000):
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ap, _, _, _ = ip_map_device(qf[:nq].contiguous(), qr[:nq].contiguous(), dbf, dbr, wl.b, wl.L, wl.R)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"real-valued C4 shape nq={nq}: {dt*1e3:.1f} ms  {nq/dt:.0f} q/s  mAP={np.nanmean(ap.cpu().numpy()):.6f}", flush=True)
