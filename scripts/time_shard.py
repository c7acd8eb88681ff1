"""Kernel durations of one rank's share of a cell-sharded fit, emulated on one GPU: the staged cfg3 matrix is cut to
1/n of its cells (vb_counts_slice) and a short fit on the slice is profiled.  usage: time_shard.py n [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import vireo_b200 as vb
from vireo_b200 import _engine, _lib, sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.cuda.set_device(0)
AD, DP, w = bench.load_workload("cfg3")
counts = vb.stage(AD, DP)
bounds = sharded.cell_shards(counts.indptr, n)
c0, c1 = int(bounds[0]), int(bounds[1])
local = counts.slice_cells(c0, c1)
np.random.seed(1)
m = vb.Vireo(n_cell=c1 - c0, n_var=w["V"], n_donor=w["K"])
batch = _engine.VireoBatch(local, [m])
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
info = {k: int(_lib.load().vb_counts_info(local.handle, 20 + i)) for i, k in enumerate(
    ["built", "steps_cell", "steps_snp", "reads_cell", "reads_snp", "grid_cell", "grid_snp", "bytes", "residual", "stream_pairs"])}
print(json.dumps({"shards": n, "cells": c1 - c0, "ms_per_iteration": round(ms_iter, 4), "split": int(_lib.load().vb_counts_info(local.handle, 61)), "kernels_ms": {k: round(v[0] / v[1], 4) for k, v in prof.items() if v[1]},
                  "fill": {"cell": round(info["stream_pairs"] / max(1, 32 * info["steps_cell"]), 4),
                           "snp": round(info["stream_pairs"] / max(1, 32 * info["steps_snp"]), 4)}, "format": info}))
