#!/usr/bin/env python
"""GPU tuning sweep: phase times of hg_hamming_map for plan overrides (HG_SELECT_QT, HG_SELECT_CTAS_PER_SM)."""
import ctypes as C
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hashgan_b200 import _native  # noqa: E402
from hashgan_b200.metric import hamming_map_device, pack_rows  # noqa: E402
from hashgan_b200.synthetic import make_workload  # noqa: E402


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C4"]
    qts = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4"])]
    ctas = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["6", "8", "9", "12", "16"])]
    lib = _native.lib()
    out = []
    for name in names:
        wl, db, q = make_workload(name)
        dbr, qr = pack_rows(db.output, db.label), pack_rows(q.output, q.label)
        ref = None
        ilps = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["0"])]  # here: threads per query of the AP kernel (0 = auto)
        for qt, c, ilp in itertools.product(qts, ctas, ilps):
            if ilp > 0:
                os.environ["HG_AP_G"] = str(ilp)
            else:
                os.environ.pop("HG_AP_G", None)
            os.environ["HG_SELECT_QT"] = str(qt)
            os.environ["HG_SELECT_CTAS_PER_SM"] = str(c)
            acc = np.zeros(6)
            phase = (C.c_float * 6)()
            stats = {}
            n = 4
            for i in range(n + 1):
                ap, _, _, _ = hamming_map_device(qr, dbr, wl.b, wl.L, wl.R, flags=_native.FLAG_TIMING, stats=stats if i == 0 else None)
                torch.cuda.synchronize()
                _native.check(lib.hg_hamming_map_phase_ms(phase))
                if i > 0:
                    acc += np.array(phase[:])
            acc /= n
            a = ap.cpu().numpy()
            if ref is None:
                ref = a
            same = bool(np.array_equal(np.isnan(a), np.isnan(ref)) and np.nanmax(np.abs(a - ref)) <= 1e-12)
            st = stats["chunks"][0]
            rec = dict(wl=name, G=ilp, qt=qt, ctas_per_sm=c, P=st["splits"], SL=st["rows_per_split"], cap=st["bin_entries"], exact=st["exact_queries"],
                       sample=round(acc[0], 4), expand=round(acc[2], 4), select=round(acc[3], 4), ap=round(acc[4], 4), exact_ms=round(acc[5], 4),
                       total=round(acc.sum(), 4), same_ap=same)
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tune_select.json", "w"), indent=1)


if __name__ == "__main__":
    main()
