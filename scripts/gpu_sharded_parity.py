"""torchrun --nproc-per-node N scripts/gpu_sharded_parity.py -- the NCCL leg of the cell-sharded fit and of the sharded
doublet pass against the single-GPU path on the same inputs (reference fixture shape and a K = 16 problem), and the whole
vireo_wrap call sharded vs unsharded.  Rank 0 prints one JSON line."""
import contextlib
import io
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                              # noqa: E402
import torch.distributed as dist                          # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
import vireo_b200 as vb                                   # noqa: E402
from oracle import vireo_oracle as O                      # noqa: E402  (generator only)
from scipy.sparse import csc_matrix                       # noqa: E402

z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixture_cellsnp.npz"))
AD1 = csc_matrix((z["AD_data"], z["AD_indices"], z["AD_indptr"]), shape=tuple(z["AD_shape"]))
DP1 = csc_matrix((z["DP_data"], z["DP_indices"], z["DP_indptr"]), shape=tuple(z["DP_shape"]))
AD2, DP2, _, _ = O.synth_counts(20000, 8000, 16, density=0.03, seed=5)


def rel(a, b):
    nz = np.abs(b) > 1e-300
    return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))) if nz.any() else 0.0


out = {"world": world}
for name, (AD, DP, K, kw) in {"cfg1": (AD1, DP1, 4, dict(max_iter=60, min_iter=5, delay_fit_theta=3)),
                              "k16": (AD2, DP2, 16, dict(max_iter=25, min_iter=25, delay_fit_theta=3))}.items():
    V, C = AD.shape
    vb.dist.disable()
    np.random.seed(7)
    a = vb.Vireo(n_var=V, n_cell=C, n_donor=K)
    a.fit(AD, DP, verbose=False, **kw)
    a_state = (a.ID_prob.copy(), a.GT_prob.copy(), a.beta_mu.copy(), a.beta_sum.copy())
    da = vb.predict_doublet(a, AD, DP)
    if world > 1:
        vb.dist.enable()
    np.random.seed(7)
    b = vb.Vireo(n_var=V, n_cell=C, n_donor=K)
    vb.fit_cell_sharded(b, AD, DP, verbose=False, **kw)
    res = {"n_elbo": [len(a.ELBO_), len(b.ELBO_)], "elbo_rel": rel(b.ELBO_, a.ELBO_) if len(a.ELBO_) == len(b.ELBO_) else None,
           "id_prob_rel": rel(b.ID_prob, a_state[0]), "gt_prob_rel": rel(b.GT_prob, a_state[1]),
           "beta_sum_rel": rel(b.beta_sum, a_state[3]),
           "same_argmax": bool(np.array_equal(b.ID_prob.argmax(1), a_state[0].argmax(1)))}
    db = vb.predict_doublet_sharded(b, AD, DP)
    res.update(doublet_prob_rel=rel(db[0], da[0]), singlet_prob_rel=rel(db[1], da[1]),
               llr_abs=float(np.max(np.abs(db[2] - da[2]))), gt_after_doublet_rel=rel(b.GT_prob, a.GT_prob))
    out[name] = res

# whole wrapper: sharded (restarts by rank, final fit + doublet by cell) vs one GPU
vb.dist.disable()
with contextlib.redirect_stdout(io.StringIO()):
    r1 = vb.vireo_wrap(AD1, DP1, n_donor=4, n_init=6, random_seed=2)
if world > 1:
    vb.dist.enable()
with contextlib.redirect_stdout(io.StringIO()):
    rN = vb.vireo_wrap(AD1, DP1, n_donor=4, n_init=6, random_seed=2)
out["wrap_cfg1"] = {"LB_list_rel": rel(rN["LB_list"], r1["LB_list"]), "LB_doublet_rel": abs(rN["LB_doublet"] - r1["LB_doublet"]) / abs(r1["LB_doublet"]),
                    "id_prob_rel": rel(rN["ID_prob"], r1["ID_prob"]), "doublet_prob_rel": rel(rN["doublet_prob"], r1["doublet_prob"]),
                    "gt_prob_rel": rel(rN["GT_prob"], r1["GT_prob"]),
                    "same_argmax": bool(np.array_equal(rN["ID_prob"].argmax(1), r1["ID_prob"].argmax(1)))}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
