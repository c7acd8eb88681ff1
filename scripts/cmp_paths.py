"""GPU-vs-GPU trajectory comparison of kernel families on a bench workload (run on the GPU box):
    python scripts/cmp_paths.py cfg3 gather seg seg32
The first family is the baseline; prints max relative differences of ELBO trace, ID_prob, GT_prob after
T iterations from the same random start, and argmax agreement."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vireo_b200 as vb  # noqa: E402
from vireo_b200 import _lib  # noqa: E402

wl = sys.argv[1]
paths = sys.argv[2:]
T = int(os.environ.get("T", "20"))
AD, DP, w = bench.load_workload(wl)
C_, V, K = w["C"], w["V"], w["K"]
inits = bench.draw_inits(w, 1)
counts = vb.stage(AD, DP)
out = {}
for pth in paths:
    _lib.set_path(pth)
    m = vb.Vireo(n_cell=C_, n_var=V, n_donor=K, ID_prob_init=inits[0][0].copy(), GT_prob_init=inits[0][1].copy())
    m.ID_prob, m.GT_prob = inits[0][0].copy(), inits[0][1].copy()
    t0 = time.perf_counter()
    m.fit(counts, None, max_iter=T, min_iter=T, delay_fit_theta=3, verbose=False)
    dt = time.perf_counter() - t0
    out[pth] = (m.ELBO_.copy(), m.ID_prob.copy(), m.GT_prob.copy(), m.beta_mu.copy(), m.beta_sum.copy())
    print("%-8s fit %.2fs  ELBO[-1] %.6f" % (pth, dt, m.ELBO_[-1]), flush=True)
    if pth in ("seg", "seg32"):
        lib = _lib.load()
        b = 20 if pth == "seg" else 30
        print("   seg info:", {k: int(lib.vb_counts_info(counts.handle, b + i)) for i, k in enumerate(
            ["built", "stepsA", "stepsB", "max_readsA", "max_readsB", "gridA", "gridB", "bytes", "heavyA", "lightA"])})
base = out[paths[0]]
for pth in paths[1:]:
    e, r, g, mu, sm = out[pth]
    def rel(a, b):
        nz = np.abs(b) > 1e-300
        return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])))
    print("%-8s vs %s: ELBO rel %.3e | ID_prob rel %.3e | GT_prob rel %.3e | mu %.3e sum %.3e | argmax mismatches %d"
          % (pth, paths[0], rel(e, base[0]), rel(r, base[1]), rel(g, base[2]), rel(mu, base[3]), rel(sm, base[4]),
             int(np.sum(r.argmax(1) != base[1].argmax(1)))))
