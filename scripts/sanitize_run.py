"""Small runs for compute-sanitizer (round 2): every kernel family on the reference's own fixture shape plus a K = 16
problem, the doublet pass and the sharded-fit loop.  Prints a digest of the results so that two builds of the library
(default: windows filled by cp.async.bulk; VIREO_B200_LIB=..._plainfill.so: same protocol, plain stores) can be shown to
agree bit for bit.

    python scripts/sanitize_run.py [rows|seg|seg32|all] [small]
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vireo_b200 as vb                                     # noqa: E402
from vireo_b200 import _lib                                 # noqa: E402
from scipy.sparse import csc_matrix                         # noqa: E402


def synth(C, V, K, density, seed):
    rng = np.random.default_rng(seed)
    mask = rng.random((V, C)) < density
    dp = (rng.geometric(0.7, size=(V, C)) * mask).astype(np.int64)
    donor = rng.integers(0, K, C)
    gt = rng.integers(0, 3, size=(V, K))
    p = np.array([0.01, 0.5, 0.99])[gt[:, donor]]
    ad = rng.binomial(dp, p)
    return csc_matrix(ad), csc_matrix(dp)


def digest(*arrs):
    h = hashlib.sha1()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


which = sys.argv[1] if len(sys.argv) > 1 else "all"
small = len(sys.argv) > 2
paths = ["rows", "seg", "seg32"] if which == "all" else [which]
C, V = (600, 500) if small else (1500, 1200)
AD, DP = synth(C, V, 16, 0.06, 2)
AD4, DP4 = synth(C, V // 2, 4, 0.06, 3)
ADs, DPs = synth(200, 12500, 16, 0.05, 5)        # few cells, long table: the row-split cell pass of the segment family


def synth_heavy(C, V, K, seed):
    """power-law SNP coverage: rows far heavier than the mean, which the segment family cuts into parts"""
    rng = np.random.default_rng(seed)
    pv = np.minimum(1.0, 12.0 / (1 + np.arange(V)) ** 0.9)
    mask = rng.random((V, C)) < pv[:, None]
    dp = np.where(mask, rng.integers(1, 4, size=(V, C)), 0)
    donor = rng.integers(0, K, C)
    gt = rng.integers(0, 3, size=(V, K))
    return csc_matrix(rng.binomial(dp, np.array([0.01, 0.5, 0.99])[gt[:, donor]])), csc_matrix(dp)


ADh, DPh = synth_heavy(700, 900, 16, 7)
for path in paths:
    _lib.set_path(path)
    np.random.seed(1)
    m = vb.Vireo(n_cell=C, n_var=V, n_donor=16)
    m.fit(AD, DP, max_iter=3, min_iter=3, delay_fit_theta=1, verbose=False)
    dbl, sgl, llr = vb.predict_doublet(m, AD, DP)
    np.random.seed(1)
    s = vb.Vireo(n_cell=C, n_var=V, n_donor=16)
    vb.fit_cell_sharded(s, AD, DP, max_iter=3, min_iter=3, delay_fit_theta=1, verbose=False)
    np.random.seed(2)
    m4 = vb.Vireo(n_cell=C, n_var=V // 2, n_donor=4)
    m4.fit(AD4, DP4, max_iter=3, min_iter=3, verbose=False)
    b = vb.BinomMixtureVB(n_cell=C, n_var=V // 2, n_donor=4)
    b.fit(AD4, DP4, n_init=3, max_iter=4, max_iter_pre=3, min_iter=2, random_seed=1, verbose=False)
    print("%-5s ELBO %.10f  digest fit %s doublet %s sharded %s k4 %s bmm %s" % (
        path, m.ELBO_[-1], digest(m.ELBO_, m.ID_prob, m.GT_prob), digest(dbl, sgl, llr),
        digest(s.ELBO_, s.ID_prob), digest(m4.ELBO_, m4.ID_prob), digest(b.ELBO_iters, b.ID_prob)))
    np.random.seed(4)
    mh = vb.Vireo(n_cell=700, n_var=900, n_donor=16)
    mh.fit(ADh, DPh, max_iter=3, min_iter=3, delay_fit_theta=1, verbose=False)
    print("%-5s heavy rows: parts %d digest %s" % (path, int(_lib.load().vb_counts_info(vb.stage(ADh, DPh).handle, 63)),
                                                    digest(mh.ELBO_, mh.ID_prob, mh.GT_prob)))
    if os.environ.get("VIREO_B200_SAN_NOSPLIT") != "1":      # the long-table case (window buffers are refilled)
        np.random.seed(3)
        ms = vb.Vireo(n_cell=200, n_var=12500, n_donor=16)
        ms.fit(ADs, DPs, max_iter=3, min_iter=3, delay_fit_theta=1, verbose=False)
        print("%-5s split launches %d digest %s" % (path, int(_lib.load().vb_counts_info(vb.stage(ADs, DPs).handle, 61)),
                                                     digest(ms.ELBO_, ms.ID_prob, ms.GT_prob)))
    vb.clear_cache()
_lib.set_path("auto")
