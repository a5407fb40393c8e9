#!/usr/bin/env python
"""GPU probe: phase times of hg_hamming_map at a full-size workload under the current environment (HG_SELECT_MODE, HG_BM_DEBUG, ...).
usage: python scripts/phase_probe.py [C4|C5|C2] [correlated | -] [nq]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hashgan_b200 import _native
from hashgan_b200.metric import hamming_map_device, pack_rows
from hashgan_b200.synthetic import make_workload
lib = _native.lib()
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
corr = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
nq = int(sys.argv[3]) if len(sys.argv) > 3 else None
wl, db, q = make_workload(name, correlated=corr, nq=nq)
dbr, qr = pack_rows(db.output, db.label), pack_rows(q.output, q.label)
phase = (C.c_float * 6)(); acc = np.zeros(6); stats = {}
for i in range(5):
    ap, ids, dist, rel = hamming_map_device(qr, dbr, wl.b, wl.L, wl.R, flags=_native.FLAG_TIMING, stats=stats if i == 0 else None)
    torch.cuda.synchronize(); _native.check(lib.hg_hamming_map_phase_ms(phase))
    if i: acc += np.array(phase[:])
acc /= 4
apn = ap.cpu().numpy()
print(f"{name} corr={corr} mode={os.environ.get('HG_SELECT_MODE','')} dbg={os.environ.get('HG_BM_DEBUG','')} K={os.environ.get('HG_DRAIN_LANES','')}: sample {acc[0]:.3f} thr {acc[1]:.3f} expand {acc[2]:.3f} select {acc[3]:.3f} ap {acc[4]:.3f} exact {acc[5]:.3f} ms  mAP {np.nanmean(apn):.12f}", flush=True)
