"""cProfile of vireo_b200.fit_cell_sharded on rank 0 of a torchrun launch (where do the fixed costs of the call go):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/prof_sharded.py"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vireo_b200 as vb  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
vb.dist.enable()
AD, DP, w = bench.load_workload("cfg3", rank, dist.barrier)
counts = vb.stage(AD, DP)
inits = bench.draw_inits(w, 1)


def fresh():
    m = vb.Vireo(n_cell=w["C"], n_var=w["V"], n_donor=w["K"], ID_prob_init=inits[0][0], GT_prob_init=inits[0][1])
    m.ID_prob, m.GT_prob = inits[0][0], inits[0][1]
    return m


kw = dict(max_iter=20, min_iter=20, delay_fit_theta=3, verbose=False)
for _ in range(2):
    vb.fit_cell_sharded(fresh(), counts, None, **kw)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter(); m = fresh(); t1 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
vb.fit_cell_sharded(m, counts, None, **kw)
torch.cuda.synchronize()
pr.disable()
t2 = time.perf_counter()
if rank == 0:
    print("fresh() %.1f ms, fit_cell_sharded %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
dist.barrier()
