"""vireo_wrap on 2+ GPUs with the final fit sharded over cells (VIREO_B200_SHARD_CELLS=1) against the same call with
the final fit replicated; run with torchrun.  Prints the parity of the two result dicts on rank 0."""
import contextlib, io, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vireo_b200 as vb
from oracle import vireo_oracle as O

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
AD, DP, donor, _ = O.synth_counts(30000, 12000, 8, seed=5)
out = {}
for flag in ("0", "1"):
    os.environ["VIREO_B200_SHARD_CELLS"] = flag
    with contextlib.redirect_stdout(io.StringIO()):
        out[flag] = vb.vireo_wrap(AD, DP, n_donor=8, n_init=4, random_seed=3, nproc=1)
a, b = out["0"], out["1"]
def rel(x, y):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    nz = np.abs(y) > 1e-300
    return float(np.max(np.abs(x[nz] - y[nz]) / np.abs(y[nz]))) if nz.any() else 0.0
res = {"rank": rank, "LB_list_rel": rel(b["LB_list"], a["LB_list"]), "LB_doublet_rel": rel([b["LB_doublet"]], [a["LB_doublet"]]),
       "ID_prob_rel": rel(b["ID_prob"], a["ID_prob"]), "GT_prob_rel": rel(b["GT_prob"], a["GT_prob"]),
       "same_argmax": bool(np.array_equal(a["ID_prob"].argmax(1), b["ID_prob"].argmax(1))),
       "recovered": float((np.bincount(donor * 8 + b["ID_prob"].argmax(1), minlength=64).reshape(8, 8).max(1).sum()) / len(donor))}
print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
