"""Per-kernel durations of a short fit on one workload, one JSON line -- for A/B runs of kernel variants (build.py
variants through VIREO_B200_LIB, ring geometry through VIREO_B200_SEG_*).  No parity check: diagnostic builds give
garbage results by design.   usage: python scripts/time_passes.py [workload] [iterations] [label]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import vireo_b200 as vb
from vireo_b200 import _engine, _lib

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
label = sys.argv[3] if len(sys.argv) > 3 else ""
torch.cuda.set_device(0)
AD, DP, w = bench.load_workload(name)
counts = vb.stage(AD, DP)
inits = bench.draw_inits(w, 1)
models = bench._new_models(vb, w, inits, [0])
batch = _engine.VireoBatch(counts, models)
init_dev = batch.state.clone()


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
_lib.load().vb_profile_enable(0)
info = {k: int(_lib.load().vb_counts_info(counts.handle, 20 + i)) for i, k in enumerate(
    ["built", "steps_cell", "steps_snp", "reads_cell", "reads_snp", "grid_cell", "grid_snp", "bytes", "residual", "stream_pairs"])}
print(json.dumps({"label": label, "workload": name, "ms_per_iteration": round(ms_iter, 4), "it_per_s": round(1e3 / ms_iter, 1),
                  "kernels_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items() if v[1]},
                  "fill": {"cell": round(info["stream_pairs"] / max(1, 32 * info["steps_cell"]), 4),
                           "snp": round(info["stream_pairs"] / max(1, 32 * info["steps_snp"]), 4)},
                  "format": info}))
