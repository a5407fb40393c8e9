#!/usr/bin/env python
"""GPU probe: tensor-core select vs POPC select -- bit-exact ids/dist and phase times."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hashgan_b200 import _native
from hashgan_b200.metric import hamming_map_device, pack_rows
from hashgan_b200.synthetic import make_workload
lib = _native.lib()
def run(name, nq, ndb, backend, want_ids):
    os.environ["HG_SELECT_BACKEND"] = backend
    wl, db, q = make_workload(name, nq=nq, ndb=ndb)
    dbr, qr = pack_rows(db.output, db.label), pack_rows(q.output, q.label)
    phase = (C.c_float * 6)(); acc = np.zeros(6); stats = {}
    for i in range(4):
        ap, ids, dist, rel = hamming_map_device(qr, dbr, wl.b, wl.L, wl.R, flags=_native.FLAG_TIMING, want_ids=want_ids, want_rel=True, stats=stats if i == 0 else None)
        torch.cuda.synchronize(); _native.check(lib.hg_hamming_map_phase_ms(phase))
        if i: acc += np.array(phase[:])
    acc /= 3
    print(f"{name} nq={nq} ndb={ndb} backend={backend}: sample {acc[0]:.3f} expand {acc[2]:.3f} select {acc[3]:.3f} ap {acc[4]:.3f} exact {acc[5]:.3f} ms; stats {stats['chunks'][0]}", flush=True)
    return ap.cpu().numpy(), (ids.cpu().numpy() if want_ids else None), (dist.cpu().numpy() if want_ids else None), rel.cpu().numpy()
for name, nq, ndb, ids in (("C2", 700, 30000, True), ("C4", 1000, 200000, True), ("C5", 600, 300000, True), ("C4", None, None, False), ("C5", None, None, False), ("C2", None, None, False)):
    a = run(name, nq, ndb, "popc", ids)
    b = run(name, nq, ndb, "umma", ids)
    same_ap = np.array_equal(np.isnan(a[0]), np.isnan(b[0])) and np.nanmax(np.abs(a[0] - b[0])) <= 1e-12
    print("   AP equal:", same_ap, " rel equal:", np.array_equal(a[3], b[3]), " ids equal:", (np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])) if ids else "n/a", flush=True)
