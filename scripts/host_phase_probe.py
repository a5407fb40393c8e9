#!/usr/bin/env python
"""GPU probe: phases of the pipelined host entry point (hg_maps_by_feature_host) for several chunk counts."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
from hashgan_b200 import MAPs, _native
from hashgan_b200.synthetic import make_workload
lib = _native.lib()
wl, db, q = make_workload(sys.argv[1] if len(sys.argv) > 1 else "C4")
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_db = NS(output=pin(db.output), label=pin(db.label)); h_q = NS(output=pin(q.output), label=pin(q.label))
phase = (C.c_float * 6)()
for k in (1, 8):
    os.environ["HG_HOST_CHUNKS"] = str(k)
    m = MAPs(wl.R, flags=_native.FLAG_TIMING)
    for _ in range(2): v = m.get_maps_by_feature(h_db, h_q)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        v = m.get_maps_by_feature(h_db, h_q)
        ts.append((time.perf_counter() - t0) * 1e3)
    _native.check(lib.hg_hamming_map_phase_ms(phase))
    print("chunks", k, "ms", [round(t, 2) for t in ts], "phases", [round(x, 3) for x in phase[:]], flush=True)
