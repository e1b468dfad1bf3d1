#!/usr/bin/env python
"""Append a fit case to tests/golden/em.json from a histogram saved as npz (v, w) -- the threshold / parameters
come from the UNMODIFIED reference module (/root/reference/lib/python/analysis/peaks.py), run in the build container.

  python tests/golden/add_em_case_from_npz.py <name> <hist.npz>

Used for `bench_250m`: the window-count histogram of the 250M-fragment benchmark input (4213 bins), produced on the
B200 by tools/hist_probe.py (the CPU oracle reproduces the same histogram on small inputs; this one is too big
to pile up on the CPU in the test budget)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, "/root/reference/lib/python")
import analysis.peaks as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
name, path = sys.argv[1], sys.argv[2]
z = np.load(path)
cd = {int(v): int(w) for v, w in zip(z["v"], z["w"])}
t0 = time.time()
thr, params = ref.estimate_final_threshold(dict(cd), 1 / 5)
print("reference fit: threshold", thr, "params", params, "%.1fs" % (time.time() - t0))
cases = json.load(open(os.path.join(HERE, "em.json")))
cases = [c for c in cases if c["name"] != name]
cases.append(dict(name=name, values=sorted(cd), weights=[cd[k] for k in sorted(cd)], threshold=int(thr),
                  params=[float(x) for x in params]))
json.dump(cases, open(os.path.join(HERE, "em.json"), "w"), separators=(",", ":"))
