"""cProfile of the public-API path (Vireo.fit with host buffers) at a bench workload."""
import cProfile, pstats, sys, time, io
sys.path.insert(0, ".")
import numpy as np, torch
import bench, vireo_b200 as vb
W = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
AD, DP, w = bench.load_workload(W)
inits = bench.draw_inits(w, 1)
inits = [(torch.from_numpy(a).pin_memory().numpy(), torch.from_numpy(b).pin_memory().numpy()) for a, b in inits]
t0 = time.perf_counter(); counts = vb.stage(AD, DP); torch.cuda.synchronize(); print("stage %.1f ms" % (1e3 * (time.perf_counter() - t0)))
m = vb.Vireo(n_cell=w["C"], n_var=w["V"], n_donor=w["K"], ID_prob_init=inits[0][0], GT_prob_init=inits[0][1])
def one():
    m.ID_prob, m.GT_prob = inits[0][0], inits[0][1]
    m.beta_mu = np.ones((1, 3)) * np.linspace(0.01, 0.99, 3).reshape(1, -1); m.beta_sum = np.ones((1, 3)) * 50
    m.ELBO_ = np.zeros(0)
    m.fit(counts, None, max_iter=20, min_iter=20, delay_fit_theta=3, verbose=False)
for i in range(3):
    t0 = time.perf_counter(); one(); torch.cuda.synchronize(); print("fit %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable(); one(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:6000])
