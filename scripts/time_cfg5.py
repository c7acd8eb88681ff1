"""BASELINE cfg5 (clone mode): BinomMixtureVB, 2k cells x 300 mito SNPs x 6 clones, n_init = 50, on one GPU."""
import os, sys, time, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vireo_b200 as vb
from oracle import vireo_oracle as O
AD, DP, _ = O.synth_clones(2000, 300, 6, seed=0)
vb.stage(AD, DP)
for rep in range(3):
    m = vb.BinomMixtureVB(n_var=300, n_cell=2000, n_donor=6)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        m.fit(AD, DP, min_iter=30, n_init=50, random_seed=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("BinomMixtureVB.fit n_init=50: %.3f s, final ELBO %.4f, %d trace entries" % (dt, m.ELBO_iters[-1], len(m.ELBO_iters)))
if "--oracle" in sys.argv:
    o = O.bmm_new(2000, 300, 6)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        O.bmm_fit(o, AD, DP, min_iter=30, n_init=50, random_seed=1)
    print("oracle (numpy/scipy, 1 core): %.1f s, final ELBO %.4f" % (time.perf_counter() - t0, o.ELBO_iters[-1]))
