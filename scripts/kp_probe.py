#!/usr/bin/env python
"""GPU probe: tensor-core select vs POPC select for short (b <= 32) and long (b > 128) codes -- AP equality and phase times."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hashgan_b200 import _native
from hashgan_b200.metric import hamming_map_device, pack_rows
from hashgan_b200.synthetic import Workload, make_workload
lib = _native.lib()
def run(b, L, nq, ndb, R, backend):
    os.environ["HG_SELECT_BACKEND"] = backend
    wl = Workload("T", nq, ndb, b, L, R, "onehot", 7)
    _, db, q = make_workload(wl)
    dbr, qr = pack_rows(db.output, db.label), pack_rows(q.output, q.label)
    phase = (C.c_float * 6)(); acc = np.zeros(6); stats = {}
    for i in range(4):
        ap, _, _, rel = hamming_map_device(qr, dbr, b, L, R, flags=_native.FLAG_TIMING, want_rel=True, stats=stats if i == 0 else None)
        torch.cuda.synchronize(); _native.check(lib.hg_hamming_map_phase_ms(phase))
        if i: acc += np.array(phase[:])
    acc /= 3
    print(f"b={b} nq={nq} ndb={ndb} R={R} backend={backend} kp={lib.hg_select_backend_for(nq, ndb, b, L, R)}: sample {acc[0]:.3f} expand {acc[2]:.3f} select {acc[3]:.3f} ap {acc[4]:.3f} ms; "
          f"exact {stats['chunks'][0]['exact_queries']} splits {stats['chunks'][0]['splits']}", flush=True)
    return ap.cpu().numpy(), rel.cpu().numpy()
for b, nq, ndb, R in ((32, 10000, 1000000, 5000), (16, 10000, 1000000, 5000), (256, 5000, 1000000, 5000), (160, 5000, 1000000, 5000), (32, 1000, 54000, 54000)):
    a = run(b, 10, nq, ndb, R, "popc")
    u = run(b, 10, nq, ndb, R, "umma")
    print("   AP equal:", np.array_equal(a[0], u[0], equal_nan=True), " rel equal:", np.array_equal(a[1], u[1]), flush=True)
