#!/usr/bin/env python
"""GPU probe: phase times of hg_hamming_map on a CLASS-SORTED database (C4 shape, class-correlated codes): the worst row order for the
per-(query, split) bins -- a query's candidates sit in the few splits that hold its class."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
from hashgan_b200 import _native
from hashgan_b200.metric import hamming_map_device, pack_rows
from hashgan_b200.synthetic import make_workload
lib = _native.lib()
wl, db, q = make_workload("C4", correlated=0.25)
order = np.argsort(np.argmax(db.label, 1), kind="stable")
db = NS(output=db.output[order], label=db.label[order])
dbr, qr = pack_rows(db.output, db.label), pack_rows(q.output, q.label)
for mode in ("queue", "lists"):
    os.environ["HG_SELECT_MODE"] = mode
    phase = (C.c_float * 6)(); acc = np.zeros(6); stats = {}
    for i in range(4):
        ap, ids, dist, rel = hamming_map_device(qr, dbr, wl.b, wl.L, wl.R, flags=_native.FLAG_TIMING, stats=stats if i == 0 else None)
        torch.cuda.synchronize(); _native.check(lib.hg_hamming_map_phase_ms(phase))
        if i: acc += np.array(phase[:])
    acc /= 3
    print(f"class-sorted C4 corr 0.25 mode={mode}: sample {acc[0]:.3f} select {acc[3]:.3f} ap {acc[4]:.3f} exact {acc[5]:.3f} ms  mAP {np.nanmean(ap.cpu().numpy()):.9f} stats {stats['chunks'][0]}", flush=True)
