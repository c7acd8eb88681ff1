"""Where vireo_wrap spends its time at a bench workload (run on the GPU box): python scripts/time_wrap.py cfg3 [n_init]"""
import os, sys, time, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import vireo_b200 as vb
from vireo_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n_init = int(sys.argv[2]) if len(sys.argv) > 2 else 2
AD, DP, w = bench.load_workload(wl)
vb.stage(AD, DP)
for rep in range(2):
    _lib.load().vb_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = vb.vireo_wrap(AD, DP, n_donor=w["K"], n_init=n_init, random_seed=1, nproc=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    prof = _lib.profile_read(); _lib.load().vb_profile_enable(0)
    print("vireo_wrap n_init=%d: %.3f s; LB_list %s" % (n_init, dt, np.round(res["LB_list"], 1)))
    for k, (ms, n) in prof.items():
        if n: print("   %-12s %8.1f ms in %4d launches" % (k, ms, n))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    res = vb.vireo_wrap(AD, DP, n_donor=w["K"], n_init=n_init, random_seed=1, nproc=1)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
