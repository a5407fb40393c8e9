#!/usr/bin/env python
"""GPU probe: end-to-end time of MAPs.get_maps_by_feature on pinned host buffers for several HG_HOST_CHUNKS."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace as NS
from hashgan_b200 import MAPs
from hashgan_b200.synthetic import make_workload
wl, db, q = make_workload(sys.argv[1] if len(sys.argv) > 1 else "C4")
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_db = NS(output=pin(db.output), label=pin(db.label)); h_q = NS(output=pin(q.output), label=pin(q.label))
d = torch.empty(db.output.shape, dtype=torch.float32, device="cuda"); l = torch.empty(db.label.shape, dtype=torch.int64, device="cuda")
for _ in range(2): d.copy_(h_db.output, non_blocking=True); l.copy_(h_db.label, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(h_db.output, non_blocking=True); l.copy_(h_db.label, non_blocking=True)
torch.cuda.synchronize(); print("pure H2D ms", (time.perf_counter() - t0) / 5 * 1e3)
for k in (1, 2, 3, 4, 6, 8):
    os.environ["HG_HOST_CHUNKS"] = str(k)
    m = MAPs(wl.R)
    for _ in range(2): v = m.get_maps_by_feature(h_db, h_q)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): v = m.get_maps_by_feature(h_db, h_q)
    torch.cuda.synchronize(); print("chunks", k, "ms", (time.perf_counter() - t0) / 5 * 1e3, "map", v)
