"""Time the threshold fit on the histograms of a library as deep as BASELINE.json configs[3] (tools/em_deep_hist.json: window
counts of an 8 Mb genome at 2 B fragments / 3.1 Gb, weights scaled to chr1; built on the CPU with the oracle's fit beside it)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cellranger-atac_b200"))
from cratac import _ffi

cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "em_deep_hist.json")))
for name in sys.argv[1:] or sorted(cases):
    c = cases[name]
    ctx = _ffi.Context([("chr1", 1000)])
    out = {"case": name, "bins": len(c["values"]), "max": c["values"][-1], "oracle": c["oracle"]}
    for rep in range(2):
        t = time.time()
        try:
            thr, params, _ = ctx.fit_threshold(c["values"], c["weights"], 0.2)
            out["threshold"], out["params"] = thr, list(params) if params is not None else None
        except _ffi.CratacError as exc:
            out["error"] = str(exc)
        out["seconds_%d" % rep] = round(time.time() - t, 4)
    print(json.dumps(out))
    sys.stdout.flush()
    ctx.close()
