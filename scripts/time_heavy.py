"""Per-kernel durations on a matrix with heavy-tailed SNP coverage (power law: the top SNPs are seen in every cell, most
SNPs in a few per mille of the cells), with and without cutting heavy rows into parts (VIREO_B200_SEG_OWNER_SPLIT).
usage: python scripts/time_heavy.py [K]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy.sparse import coo_matrix

import vireo_b200 as vb
from vireo_b200 import _engine, _lib

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C, V = 50000, 20000
rng = np.random.default_rng(3)
pv = np.minimum(1.0, 20.0 / (1.0 + np.arange(V)) ** 0.8)
n_i = rng.binomial(C, pv)
rows = np.repeat(np.arange(V), n_i)
cols = np.concatenate([rng.choice(C, n, replace=False) for n in n_i])
dp = rng.integers(1, 4, size=rows.size)
donor = rng.integers(0, K, C)
gt = rng.integers(0, 3, size=(V, K))
ad = rng.binomial(dp, np.array([0.01, 0.5, 0.99])[gt[rows, donor[cols]]])
DP = coo_matrix((dp, (rows, cols)), shape=(V, C)).tocsc()
AD = coo_matrix((ad, (rows, cols)), shape=(V, C)).tocsc()
AD.eliminate_zeros()
torch.cuda.set_device(0)
_lib.set_path("seg")
counts = vb.stage(AD, DP)
np.random.seed(1)
m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
batch = _engine.VireoBatch(counts, [m])
init_dev = batch.state.clone()
iters = 10


def step():
    batch.state.copy_(init_dev)
    batch.run_fit(iters, iters, 1e-2, 3, poll_every=iters + 1)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms_iter = e0.elapsed_time(e1) / 3 / iters
_lib.load().vb_profile_enable(1)
step()
torch.cuda.synchronize()
prof = _lib.profile_read()
lib = _lib.load()
print(json.dumps({"nnz": int(DP.nnz), "K": K, "owner_split": os.environ.get("VIREO_B200_SEG_OWNER_SPLIT", "1"),
                  "ms_per_iteration": round(ms_iter, 4), "kernels_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items() if v[1]},
                  "row_imbalance": int(lib.vb_counts_info(counts.handle, 62)) / 1000.0, "parts": int(lib.vb_counts_info(counts.handle, 63)),
                  "longest_row": int(n_i.max()), "mean_row": float(n_i.mean())}))
