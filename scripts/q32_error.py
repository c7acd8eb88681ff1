"""Error of the fixed-point family ("seg32") against an FP64 family on a bench workload (run on the GPU box):
    python scripts/q32_error.py cfg2 [T]
free-running: both families iterate from the same start, difference after every iteration;
teacher-forced: one seg32 iteration from the FP64 family's state at that iteration."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import vireo_b200 as vb
from vireo_b200 import _lib

wl = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
AD, DP, w = bench.load_workload(wl)
C_, V, K = w["C"], w["V"], w["K"]
inits = bench.draw_inits(w, 1)
counts = vb.stage(AD, DP)

def fresh():
    m = vb.Vireo(n_cell=C_, n_var=V, n_donor=K, ID_prob_init=inits[0][0].copy(), GT_prob_init=inits[0][1].copy())
    m.ID_prob, m.GT_prob = inits[0][0].copy(), inits[0][1].copy()
    return m

def rel(a, b):
    nz = np.abs(b) > 1e-300
    return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])))

def clone(m):
    c = vb.Vireo(n_cell=C_, n_var=V, n_donor=K, ID_prob_init=m.ID_prob.copy(), GT_prob_init=m.GT_prob.copy(),
                 beta_mu_init=m.beta_mu.copy(), beta_sum_init=m.beta_sum.copy())
    c.ID_prob, c.GT_prob = m.ID_prob.copy(), m.GT_prob.copy()
    return c

a, b = fresh(), fresh()
for it in range(T):
    delay = 0 if it >= 3 else 100
    tf = clone(a)
    _lib.set_path("rows"); a.fit(counts, None, max_iter=1, min_iter=1, delay_fit_theta=delay, verbose=False)
    _lib.set_path("seg32"); b.fit(counts, None, max_iter=1, min_iter=1, delay_fit_theta=delay, verbose=False)
    tf.fit(counts, None, max_iter=1, min_iter=1, delay_fit_theta=delay, verbose=False)
    print("it %2d free: ID %.2e GT %.2e | forced: ID %.2e GT %.2e | argmax diff %d" % (
        it, rel(b.ID_prob, a.ID_prob), rel(b.GT_prob, a.GT_prob),
        rel(tf.ID_prob, a.ID_prob), rel(tf.GT_prob, a.GT_prob), int(np.sum(a.ID_prob.argmax(1) != b.ID_prob.argmax(1)))), flush=True)
