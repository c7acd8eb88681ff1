#!/usr/bin/env python
"""bench.py -- EM iterations/sec of the vireoSNP variational-EM inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3]

Metric (BASELINE.json): EM iterations per second -- one iteration is one pass of the `_fit_VB` body
(theta, GT, ID, ELBO; reference vireoSNP/utils/vireo_model.py:257-264) for one restart -- aggregated
over all restarts of all ranks.  A "step" is one fit of T = 20 fixed iterations
(min_iter = max_iter = 20, delay_fit_theta = 3) of every restart a rank owns.

Workload at every N: BASELINE cfg3 (synthetic 100k cells x 50k SNPs x 16 donors, ~1e8 nnz, learn GT),
one restart per GPU (n_init = N sharded round-robin; weak scaling), the full matrices staged on every
GPU.  Under torchrun the restarts never talk; ONE all-gather of the final ELBOs ends the run.

Beside the headline the line carries (extra keys, same run): `wrap` -- one whole `vireo_wrap(n_init = 8)` call on
the N GPUs (warm-ups sharded by restart, final fit and doublet pass sharded by cell) split into its phases;
`sharded_fit` (N > 1) -- one 100-iteration fit cell-sharded over the N GPUs against the same fit on one GPU, with
its parity figures; `doublet_ms`; `cold_e2e`.  `--workload cfg4` is BASELINE cfg4 as specified (GT-given mode:
learn_GT = False, GT_prior 0.98 / 0.01), `--workload cfg5` the clone mode (BinomMixtureVB, n_init = 50).

`--impl reference` times the CPU restatement of the reference (oracle/, the reference itself is pure
Python and cannot travel to the GPU box) on the host cores, one process per core, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg2": dict(C=10000, V=5000, K=4, seed=0, mode="learn_gt"),
    "cfg3": dict(C=100000, V=50000, K=16, seed=0, mode="learn_gt"),
    "cfg4": dict(C=50000, V=20000, K=8, seed=0, mode="gt_given"),      # learn_GT = False, GT_prior from the planted GT
    "cfg5": dict(C=2000, V=300, K=6, seed=0, mode="bmm"),              # BinomMixtureVB clone mode
    "tiny": dict(C=2000, V=1500, K=4, seed=0, mode="learn_gt"),
}
T_ITERS = 20
DELAY = 3
BMM_INIT = 50
CACHE_DIR = os.environ.get("VIREO_B200_BENCH_CACHE", "/tmp/vireo_b200_bench")


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------

def load_workload(name, rank=0, barrier=None):
    """Synthetic AD/DP of SURVEY 8d (matrix-level donor-pool generator), cached on local disk so the
    ranks of one box -- and the two arms of one round -- generate it once."""
    from scipy.sparse import csc_matrix
    from oracle.vireo_oracle import synth_clones, synth_counts      # generators only; shared with the tests
    w = dict(WORKLOADS[name])
    path = os.path.join(CACHE_DIR, "%s_seed%d_v2.npz" % (name, w["seed"]))
    if rank == 0 and not os.path.exists(path):
        os.makedirs(CACHE_DIR, exist_ok=True)
        if w["mode"] == "bmm":
            AD, DP, _ = synth_clones(w["C"], w["V"], w["K"], seed=w["seed"])
            GT = np.zeros((0, 0), dtype=np.int8)
        else:
            AD, DP, _, GT = synth_counts(w["C"], w["V"], w["K"], seed=w["seed"])
        tmp = path + ".%d.tmp.npz" % os.getpid()
        np.savez(tmp, dp_data=DP.data.astype(np.uint32), dp_idx=DP.indices.astype(np.int32),
                 dp_ptr=DP.indptr.astype(np.int64), ad_data=AD.data.astype(np.uint32),
                 ad_idx=AD.indices.astype(np.int32), ad_ptr=AD.indptr.astype(np.int64), GT=GT.astype(np.int8))
        os.replace(tmp, path)
    if barrier is not None:
        barrier()
    z = np.load(path)
    shape = (w["V"], w["C"])
    DP = csc_matrix((z["dp_data"].astype(np.int64), z["dp_idx"], z["dp_ptr"]), shape=shape)
    AD = csc_matrix((z["ad_data"].astype(np.int64), z["ad_idx"], z["ad_ptr"]), shape=shape)
    w["GT"] = z["GT"].astype(np.int64)
    return AD, DP, w


def gt_prior_of(w):
    """BASELINE cfg4: known donor genotypes, 0.98 on the true genotype and 0.01 elsewhere."""
    prior = np.full((w["V"], w["K"], 3), 0.01)
    np.put_along_axis(prior, w["GT"][:, :, None], 0.98, axis=2)
    return prior


def draw_inits(w, n_init, seed=1):
    """Initial states of n_init restarts in the reference's RNG order (vireo_wrap.py:65-71,
    vireo_model.py:95-104): per model rand(C, K) then rand(V, K, 3)."""
    np.random.seed(seed)
    out = []
    for _ in range(n_init):
        idp = np.random.rand(w["C"], w["K"])
        gtp = np.random.rand(w["V"], w["K"], 3)
        out.append((idp / idp.sum(1, keepdims=True), gtp / gtp.sum(2, keepdims=True)))
    return out


def algorithmic_bytes(w, nnz, wide=False):
    """Compulsory HBM traffic per launch, 8 B per nnz record (12 B wide)."""
    C, V, K, G = w["C"], w["V"], w["K"], 3
    e = 12 if wide else 8
    return {
        "k_cell": nnz * e + 16 * V * K + 8 * C * K + 4 * (C + 1),          # B_ID: stream + Wa/Wb read + ID_prob write
        "k_snp": nnz * e + 8 * C * K + 16 * V * K + 4 * (V + 1),           # stream + ID_prob read + S1/S2 write
        "k_gt": 24 * V * K * G + 32 * V * K,                               # prior read, GT write, S1/S2 read, Wa/Wb write
        "iter": 2 * nnz * e + 16 * C * K + 24 * V * K * G + 32 * V * K + 4 * (C + V + 2),
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.hw_slowdown,"
              "clocks_throttle_reasons.hw_thermal_slowdown,clocks_throttle_reasons.sw_thermal_slowdown,"
              "clocks_throttle_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_id):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_id), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.2):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
                    "error": " | ".join(l for _, l in self.rows[:2])[:300]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm (CPU)
# ------------------------------------------------------------------------------------------------

_REF = {}


def _ref_worker(seed):
    """One EM iteration (theta, GT, ID, ELBO) of one restart on the shared sample -- oracle/ restates
    Vireo._fit_VB / BinomMixtureVB._fit_BV with the reference's own scipy/numpy operations."""
    from oracle import vireo_oracle as O
    AD, DP, K, mode = _REF["AD"], _REF["DP"], _REF["K"], _REF["mode"]
    np.random.seed(seed)
    if mode == "bmm":
        st = O.bmm_new(AD.shape[1], AD.shape[0], K)
        t0 = time.perf_counter()
        O.bmm_fit_vb(st, AD, DP, max_iter=_REF["iters"], min_iter=_REF["iters"], verbose=False)
        return time.perf_counter() - t0
    if mode == "gt_given":
        prior = _REF["prior"]
        st = O.vireo_new(AD.shape[1], AD.shape[0], K, learn_GT=False, GT_prob_init=prior.copy())
        O.vireo_set_prior(st, GT_prior=prior.copy())
    else:
        st = O.vireo_new(AD.shape[1], AD.shape[0], K)
    t0 = time.perf_counter()
    O.vireo_fit_vb(st, AD, DP, max_iter=_REF["iters"], min_iter=_REF["iters"], delay_fit_theta=0, verbose=False)
    return time.perf_counter() - t0


def cpu_sample(AD, DP, w, budget_s, n_steps, workers):
    """Column (cell) subset sized so that n_steps single-iteration steps fit the time budget."""
    full_nnz = DP.nnz
    est_iter_s = 23.4 * (full_nnz / 1.0e8) * (w["K"] / 16.0) * (1.0 + 0.04 * max(0, workers - 1))
    frac = min(1.0, budget_s / max(1e-9, n_steps * est_iter_s))
    n_cells = max(min(w["C"], 500), int(w["C"] * frac))
    if n_cells >= w["C"]:
        return AD, DP, 1.0, w["C"]
    ADs, DPs = AD[:, :n_cells], DP[:, :n_cells]
    return ADs, DPs, DPs.nnz / float(full_nnz), n_cells


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    AD, DP, w = load_workload(args.workload)
    cores = os.cpu_count() or 1
    try:
        import psutil
        ram_gb = psutil.virtual_memory().available / 1e9
    except Exception:
        ram_gb = 32.0
    n_steps = args.steps + args.warmup
    workers = max(1, min(cores, 64))
    ADs, DPs, frac, n_cells = cpu_sample(AD, DP, w, args.ref_budget_s, n_steps, workers)
    per_worker_gb = 6.1 * frac * (DP.nnz / 1.0e8) + 0.2       # measured RSS of the reference at cfg3: 6.1 GB
    workers = max(1, min(workers, int(0.6 * ram_gb / per_worker_gb)))
    iters = 10 if w["mode"] == "bmm" else 1                   # clone mode: 24 ms per iteration on one core
    _REF.update(AD=ADs, DP=DPs, K=w["K"], iters=iters, mode=w["mode"],
                prior=gt_prior_of(w) if w["mode"] == "gt_given" else None)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(workers) as pool:
        for s in range(n_steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [1000 * s + i for i in range(workers)])
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = float(sum(times))
    # iterations of the FULL-shape workload per second: each worker ran `iters` iterations on `frac` of the nnz
    value = workers * iters * len(times) * frac / total
    sample = ("%d of %d cells (%.1f%% of the nnz) x %d SNPs x %d donors; each step = %d EM iteration(s) "
              "(theta+GT+ID+ELBO) in each of %d forked processes (the reference's restart parallelism, "
              "vireo_wrap.py:74-83); value scaled to full-shape iterations by the nnz fraction -- the V-proportional "
              "part of an iteration does not shrink with the cell sample, so the scaled figure is ~10%% pessimistic; "
              "unscaled anchor: one core, full data = 10.2 s per iteration at cfg3 (0.098 it/s, BENCH_r01)"
              % (n_cells, w["C"], 100 * frac, w["V"], w["K"], iters, workers))
    line = {
        "impl": "reference", "metric": "EM iterations/sec", "value": value, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
        "higher_is_better": True, "scaling": "strong" if w["mode"] == "bmm" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w, DP.nnz),
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": workers, "kind": "port", "sample": sample,
                         "host_cores": cores},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cells_snps_donors_per_s": value * w["C"] * w["V"] * w["K"],
    }
    print(json.dumps(line))
    return 0


def workload_config(args, w, nnz):
    world = args.gpus
    if w["mode"] == "bmm":
        return {
            "workload": "%s: BinomMixtureVB clone mode, synthetic %d cells x %d mito SNPs x %d clones, nnz(DP)=%d; "
                        "step = %d EM iterations (min_iter=max_iter=%d) of each of the n_init = %d restarts, sharded "
                        "round-robin over the GPUs" % (args.workload, w["C"], w["V"], w["K"], nnz, T_ITERS, T_ITERS, BMM_INIT),
            "n_init": BMM_INIT, "parallelism": "restart-sharded x%d, full matrices on every GPU" % world,
            "l2": "the matrices fit L2 (%.1f MB): the path is latency bound; every step starts from fresh state" % (nnz * 12 / 1e6),
        }
    mode = ("learn_GT, no donor GT" if w["mode"] == "learn_gt" else
            "GT-given mode: learn_GT=False, GT_prior = GT_prob_init = 0.98 on the planted genotype, 0.01 elsewhere")
    return {
        "workload": "%s: synthetic %d cells x %d SNPs x %d donors, nnz(DP)=%d, %s; "
                    "step = %d EM iterations (min_iter=max_iter=%d, delay_fit_theta=%d) of %d restart(s) per GPU"
                    % (args.workload, w["C"], w["V"], w["K"], nnz, mode, T_ITERS, T_ITERS, DELAY, args.restarts),
        "n_init": args.gpus * args.restarts,
        "restarts_per_gpu": args.restarts,
        "parallelism": "restart-sharded x%d, full matrices on every GPU" % args.gpus,
        "l2": "inputs larger than L2: every pass streams its whole record stream from HBM (measured 0.52 GB per pass "
              "at cfg3, 126 MB L2; %.2f GB in the 8 B/nnz row format) -- no flush needed" % (nnz * 8 / 1e9),
    }


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

def _setup(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import vireo_b200 as vb
    if world > 1:
        vb.dist.enable()                 # restart / cell sharding over the ranks is opt-in

    def barrier():
        if world > 1:
            dist.barrier()

    return torch, dist, vb, rank, local_rank, world, barrier


def _max_over_ranks(torch, dist, world, x):
    if world > 1:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def _peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    return peak, src, float(peaks.get("sm_max_mhz", 1965.0))


def _new_models(vb, w, inits, mine):
    C_, V, K = w["C"], w["V"], w["K"]
    models = []
    prior = gt_prior_of(w) if w["mode"] == "gt_given" else None
    for i in mine:
        if prior is None:
            m = vb.Vireo(n_cell=C_, n_var=V, n_donor=K, ID_prob_init=inits[i][0], GT_prob_init=inits[i][1])
            m.ID_prob, m.GT_prob = inits[i][0], inits[i][1]
        else:       # BASELINE cfg4: genotypes known and fixed (vireo_wrap.py:48-50 forces one restart per problem)
            m = vb.Vireo(n_cell=C_, n_var=V, n_donor=K, learn_GT=False, ID_prob_init=inits[i][0],
                         GT_prob_init=prior.copy())
            m.set_prior(GT_prior=prior.copy())
            m.ID_prob = inits[i][0]
        models.append(m)
    return models


def run_b200(args):
    torch, dist, vb, rank, local_rank, world, barrier = _setup(args)
    from vireo_b200 import _engine, _lib
    dev = local_rank
    AD, DP, w = load_workload(args.workload, rank, barrier if world > 1 else None)
    if w["mode"] == "bmm":
        return run_b200_bmm(args, torch, dist, vb, rank, dev, world, barrier, AD, DP, w)
    C_, V, K = w["C"], w["V"], w["K"]
    R = args.restarts
    n_init = world * R
    learn_gt = w["mode"] == "learn_gt"

    # ---- staging (one-off): host CSC -> HBM, both orientations; then the first fit through the public API: `cold_e2e`
    torch.cuda.synchronize()
    t_cold = time.perf_counter()
    counts = vb.stage(AD, DP)
    torch.cuda.synchronize()
    staging_ms = 1e3 * (time.perf_counter() - t_cold)
    t0 = time.perf_counter()
    binom = float(counts.binom_const())
    binom_ms = 1e3 * (time.perf_counter() - t0)

    # ---- restarts: every rank draws all inits in the reference's RNG order and keeps its share
    inits = draw_inits(w, n_init)
    mine = [i for i in range(n_init) if i % world == rank]
    # this rank's input states live in pinned host memory (the contract's "host->device copy of that step's inputs from
    # pinned host memory"): numpy views of page-locked blocks, handed to the public API like any other array
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    inits = [(pinned(a), pinned(b)) if i in mine else (a, b) for i, (a, b) in enumerate(inits)]
    models = _new_models(vb, w, inits, mine)

    def reset_models():
        for m, i in zip(models, mine):
            m.ID_prob = inits[i][0]
            if learn_gt:
                m.GT_prob = inits[i][1]
            m.beta_mu = np.ones((1, 3)) * np.linspace(0.01, 0.99, 3).reshape(1, -1)
            m.beta_sum = np.ones((1, 3)) * 50
            m.ELBO_ = np.zeros(0)

    # ---- end to end through the public API: host numpy state in, host results out.  The count matrices are staged
    #      once and the handle is passed (`Vireo.fit(counts, None)`, the documented fast path; `fit(AD, DP)` with the
    #      scipy matrices re-uses the same copy after checksumming both matrices, timed separately as `raw_matrices`).
    def e2e_step(handle=True):
        reset_models()
        for m in models:
            if handle:
                m.fit(counts, None, max_iter=T_ITERS, min_iter=T_ITERS, delay_fit_theta=DELAY, verbose=False)
            else:
                m.fit(AD, DP, max_iter=T_ITERS, min_iter=T_ITERS, delay_fit_theta=DELAY, verbose=False)

    e2e_step()                                   # first fit: format builds, first-use allocations
    torch.cuda.synchronize()
    cold_ms = 1e3 * (time.perf_counter() - t_cold)

    reset_models()
    batch = _engine.VireoBatch(counts, models)
    init_dev = batch.state.clone()

    def step():
        batch.state.copy_(init_dev)
        batch.run_fit(T_ITERS, T_ITERS, 1e-2, DELAY, poll_every=T_ITERS + 1)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    gpu_id = local_rank
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", "")
    except Exception:
        pass
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    time.sleep(0.25)

    # ---- timed region: inputs resident in HBM, CUDA events on the launching (current) stream
    barrier()
    torch.cuda.synchronize()
    before = _lib.launch_counts()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    after = _lib.launch_counts()
    ms = _max_over_ranks(torch, dist, world, ev0.elapsed_time(ev1))
    launches = sum(after[k] - before[k] for k in after)
    iters_total = world * R * T_ITERS * args.steps
    value = iters_total / (ms / 1e3)
    traces = batch.traces()
    elbo_last = np.array([tr[0][tr[1] - 1] for tr in traces]) + binom

    # ---- per-kernel durations, live, CUDA events around every launch (separate steps: the event pairs
    #      serialise launches slightly, so they stay out of the timed region above)
    _lib.load().vb_profile_enable(1)
    for _ in range(max(1, min(args.steps, 2))):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.load().vb_profile_enable(0)
    alg = algorithmic_bytes(w, counts.nnz, counts.wide)
    if not learn_gt:
        alg["iter"] -= 16 * V * K * 3             # GT_prob is read once, never re-written
        alg["k_gt"] -= 16 * V * K * 3

    def kernel_table(prof):
        out = {}
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        for name, (kms, n) in prof.items():
            if n:
                per = kms / n / R                    # grid.y = restarts: one launch covers R restarts
                out[name] = {"ms_per_launch_per_restart": per, "launches": n, "share": kms / tot_ms}
                if name in alg:
                    out[name]["achieved_gbs"] = alg[name] / (per * 1e-3) / 1e9
        return out

    kernels = kernel_table(prof)
    family = os.environ.get("VIREO_B200_PATH", "auto").lower()

    # ---- the opt-in fixed-point family (VIREO_B200_PATH=seg32): same workload, same timing rules, and its
    #      distance from the default family's result after the same 20 iterations from the same start.  Reported
    #      beside the headline, never as the headline: its gather tables are 32-bit fixed point (exact integer
    #      accumulation), everything else FP64.
    fixed32 = None
    if family == "auto" and not args.no_fixed32 and learn_gt:
        step()
        torch.cuda.synchronize()
        ref_state = [t.clone() for t in (batch.id_prob, batch.gt_prob)]
        ref_elbo = np.array([tr[0][:tr[1]] for tr in batch.traces()])
        _lib.set_path("seg32")
        try:
            batch32 = _engine.VireoBatch(counts, models)

            def step32():
                batch32.state.copy_(init_dev)
                batch32.run_fit(T_ITERS, T_ITERS, 1e-2, DELAY, poll_every=T_ITERS + 1)

            for _ in range(args.warmup):
                step32()
            barrier()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.steps):
                step32()
            f1.record()
            torch.cuda.synchronize()
            ms32 = _max_over_ranks(torch, dist, world, f0.elapsed_time(f1))

            def max_rel(a, b_):
                nz = b_.abs() > 1e-300
                return float(((a[nz] - b_[nz]).abs() / b_[nz].abs()).max().item())

            elbo32 = np.array([tr[0][:tr[1]] for tr in batch32.traces()])
            same_argmax = bool((batch32.id_prob.view(R, C_, K).argmax(2) == ref_state[0].view(R, C_, K).argmax(2)).all().item())
            _lib.load().vb_profile_enable(1)
            step32()
            torch.cuda.synchronize()
            prof32 = _lib.profile_read()
            _lib.load().vb_profile_enable(0)
            fixed32 = {
                "value": iters_total / (ms32 / 1e3), "unit": "it/s", "ms_per_iteration_per_restart": ms32 / args.steps / T_ITERS,
                "dtype": "u32 fixed-point gather tables, exact i64 accumulation; f64 state, softmax, reductions",
                "kernels": kernel_table(prof32),
                "vs_default_family_after_%d_iterations" % T_ITERS: {
                    "id_prob_max_rel_diff": max_rel(batch32.id_prob, ref_state[0]),
                    "gt_prob_max_rel_diff": max_rel(batch32.gt_prob, ref_state[1]),
                    "elbo_max_rel_diff": float(np.max(np.abs(elbo32 - ref_elbo) / np.abs(ref_elbo))),
                    "identical_argmax_donor": same_argmax},
            }
            batch32._bufs = None
            del batch32
        except Exception as exc:                      # the opt-in leg must never take the headline down with it
            fixed32 = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
        finally:
            _lib.set_path("auto")

    peak, peak_src, sm_mhz = _peaks()
    dom = max((k for k in kernels if k in ("k_cell", "k_snp")), key=lambda k: kernels[k]["share"], default=None)
    roofline = None
    traffic = wavefronts = None
    try:   # per launch, from the committed ncu --set full capture: DRAM bytes and shared-memory wavefronts
        tr = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        traffic = tr.get(args.workload, {}).get(dom)
        wavefronts = tr.get(args.workload + "_smem_wavefronts", {}).get(dom)
    except Exception:
        pass
    if dom:
        per_ms = kernels[dom]["ms_per_launch_per_restart"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom],
                    "ms_per_launch": per_ms,
                    "iteration": {"algorithmic_bytes": alg["iter"],
                                  "achieved_gbs": alg["iter"] * (iters_total / world) / (ms / 1e3) / 1e9,
                                  "frac": alg["iter"] * (iters_total / world) / (ms / 1e3) / 1e9 / peak}}
        if wavefronts:
            # the unit that actually bounds these kernels: every (owner, row) pair moves one table row
            # through the shared-memory crossbar, 128 B per clock per SM
            sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
            xbar_peak = sm_count * 128 * sm_mhz * 1e6 / 1e9
            xbar = wavefronts * 128 / (per_ms * 1e-3) / 1e9
            roofline["crossbar"] = {"wavefronts_per_launch": wavefronts, "achieved_gbs": xbar, "peak_gbs": xbar_peak,
                                    "frac": xbar / xbar_peak,
                                    "how": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of the committed ncu capture x "
                                           "128 B / live launch duration; peak = SMs x 128 B/clk x max SM clock"}

    # ---- warm end-to-end (two untimed calls first: the legs above have churned the device and pinned-host caches)
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(2):
        e2e_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(torch, dist, world, time.perf_counter() - t0)
    e2e_value = world * R * T_ITERS * e2e_steps / e2e_s
    e2e_elbo = np.array([m.ELBO_[-1] for m in models])
    t0 = time.perf_counter()
    e2e_step(handle=False)
    torch.cuda.synchronize()
    raw_s = _max_over_ranks(torch, dist, world, time.perf_counter() - t0)
    G = 3
    # state of every restart + priors: the default genotype prior is one row (replicated on the device), the donor
    # prior one row, the theta prior 2 x G values; GT-given mode uploads the full genotype prior once (cached)
    h2d = R * 8 * (C_ * K + V * K * G + 2 * G) + 8 * (G + K + 2 * G)
    d2h = R * 8 * (C_ * K + (V * K * G if learn_gt else 0) + 2 * G + T_ITERS) + R * 16

    # ---- the single collective of the restart-sharded path: all-gather of final ELBOs -> model selection
    from vireo_b200.dist import allgather_elbo
    final = np.full(n_init, -np.inf)
    final[mine] = elbo_last
    final = allgather_elbo(final, dev)

    # ---- the doublet pass (on by default in vireo_wrap, reference vireo_wrap.py:151-152) on the fitted state
    doublet_ms = None
    try:
        m0 = models[0]
        torch.cuda.synchronize()
        vb.predict_doublet(m0, counts, None, update_GT=False, update_ID=False)      # first use: workspaces
        torch.cuda.synchronize()
        before_d = _lib.launch_counts()
        t0 = time.perf_counter()
        vb.predict_doublet(m0, counts, None, update_GT=False, update_ID=False)
        torch.cuda.synchronize()
        whole = 1e3 * (time.perf_counter() - t0)
        _lib.load().vb_profile_enable(1)
        vb.predict_doublet(m0, counts, None, update_GT=False, update_ID=False)
        torch.cuda.synchronize()
        pd = _lib.profile_read()
        _lib.load().vb_profile_enable(0)
        after_d = _lib.launch_counts()
        doublet_ms = {"kernels_ms": sum(v[0] for v in pd.values()), "cell_pass_ms": pd["k_cell"][0],
                      "cell_pass_launches": pd["k_cell"][1], "columns": K + K * (K - 1) // 2,
                      "call_ms_incl_upload_and_download": whole,
                      "one_id_update_ms": kernels.get("k_cell", {}).get("ms_per_launch_per_restart"),
                      "launches": sum(after_d[k] - before_d[k] for k in after_d) // 2}
    except Exception as exc:
        doublet_ms = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    # ---- one fit over all N GPUs: 100 fixed iterations, cells sharded over the ranks, against the same fit on one GPU
    sharded = None
    if world > 1 and learn_gt:
        try:
            sharded = measure_sharded_fit(torch, dist, vb, rank, world, barrier, counts, w, inits[0])
        except Exception as exc:
            sharded = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    # ---- the whole vireo_wrap call (n_init = 8): warm-ups sharded by restart, final fit + doublet pass by cell
    wrap = None
    if not args.no_wrap:
        try:
            wrap = measure_wrap(torch, dist, vb, rank, world, barrier, counts, w)
        except Exception as exc:
            wrap = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    clocks = sampler.stop(t_load0, None) if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle restatement, one core, bounded sample.
    #      The same single iteration (theta+GT+ID+ELBO from restart 0's initial state) is run on the GPU,
    #      so the line also carries a same-inputs parity figure.
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        from oracle import vireo_oracle as O
        ADs, DPs, frac, n_cells = cpu_sample(AD, DP, w, args.cpu_budget_s, 1, 1)
        if learn_gt:
            st = O.vireo_new(n_cells, V, K, ID_prob_init=inits[0][0][:n_cells], GT_prob_init=inits[0][1])
        else:
            prior = gt_prior_of(w)
            st = O.vireo_new(n_cells, V, K, learn_GT=False, ID_prob_init=inits[0][0][:n_cells], GT_prob_init=prior.copy())
            O.vireo_set_prior(st, GT_prior=prior.copy())
        chk = []
        t0 = time.perf_counter()
        O.vireo_fit_vb(st, ADs, DPs, max_iter=1, min_iter=1, delay_fit_theta=0, verbose=False, trace=chk)
        dt = time.perf_counter() - t0
        cpu = {"value": frac / dt, "unit": "it/s", "cores": 1, "kind": "port",
               "sample": "1 EM iteration (theta+GT+ID+ELBO) on %d of %d cells (%.1f%% of the nnz), scipy/numpy "
                         "single thread, scaled to full-shape iterations by the nnz fraction; host has %d cores"
                         % (n_cells, C_, 100 * frac, os.cpu_count() or 1),
               "seconds": dt}
        if frac == 1.0:
            batch.state.copy_(init_dev)
            batch.run_fit(1, 1, 1e-2, 0)
            g_elbo = float(batch.traces()[0][0][0])
            g_id = batch.id_prob.cpu().numpy().reshape(-1, C_, K)[0]
            ref_id = chk[0]["ID_prob"]
            nz = ref_id > 1e-300
            parity = {"elbo_iter0_rel_diff": abs(chk[0]["ELBO"] - g_elbo) / abs(chk[0]["ELBO"]),
                      "id_prob_max_rel_diff": float(np.max(np.abs(g_id[nz] - ref_id[nz]) / ref_id[nz])),
                      # after ONE iteration from a random start most posteriors are near-uniform: report
                      # argmax mismatches only where the top-2 gap is meaningful
                      "argmax_mismatch_gap_gt_1e-9": int(np.sum((g_id.argmax(1) != ref_id.argmax(1)) &
                                                                (np.diff(np.sort(ref_id, 1)[:, -2:], axis=1)[:, 0] > 1e-9)))}

    line = {
        "metric": "EM iterations/sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w, counts.nnz),
        "cells_snps_donors_per_s": value * C_ * V * K,
        "ms_per_iteration_per_restart": ms / args.steps / T_ITERS,
        "roofline": roofline, "kernels": kernels, "kernel_family": family, "fixed32": fixed32, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "vireo_b200.Vireo.fit(counts, None, ...): numpy state in (arrays in pinned host memory), numpy results out, "
                "counts = vireo_b200.stage(AD, DP) staged to HBM once (staging_ms)",
                "raw_matrices": {"value": world * R * T_ITERS / raw_s, "unit": "it/s",
                                 "api": "Vireo.fit(AD, DP) with the scipy matrices: the staged copy is re-used after a "
                                        "checksum over the full contents of both matrices"}},
        "cold_e2e": {"ms": cold_ms, "what": "stage(AD, DP) + binomial constant + format builds + the first "
                                            "Vireo.fit of %d iterations, from host matrices to host results" % T_ITERS,
                     "it_per_s": R * T_ITERS / (cold_ms / 1e3) * world},
        "staging_ms": staging_ms, "binom_const_ms": binom_ms,
        "doublet_ms": doublet_ms, "sharded_fit": sharded, "wrap": wrap,
        "gpu_launches": launches, "clocks": clocks,
        "elbo_final": [float(x) for x in final], "winner": int(np.argmax(final)),
        "e2e_matches_resident": bool(np.allclose(e2e_elbo, elbo_last, rtol=1e-12)),
        "parity_check": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_sharded_fit(torch, dist, vb, rank, world, barrier, counts, w, init, iters=100):
    """One fit of `iters` fixed iterations through the public API (host state in, host results out): cells sharded over
    all ranks (vb.fit_cell_sharded) vs the same fit on one GPU (rank 0 alone).  Both legs exclude the one-off cut of
    the cell shard (timed separately) and the construction of the model object."""
    C_, V, K = w["C"], w["V"], w["K"]
    from vireo_b200.sharded import shard_of

    def fresh():
        m = vb.Vireo(n_cell=C_, n_var=V, n_donor=K, ID_prob_init=init[0], GT_prob_init=init[1])
        m.ID_prob, m.GT_prob = init[0], init[1]
        return m

    kw = dict(max_iter=iters, min_iter=iters, delay_fit_theta=DELAY, verbose=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    shard_of(counts)
    torch.cuda.synchronize()
    cut_ms = 1e3 * (time.perf_counter() - t0)
    vb.fit_cell_sharded(fresh(), counts, None, max_iter=3, min_iter=3, verbose=False)     # first use: formats, NCCL
    import importlib
    sh = importlib.import_module("vireo_b200.sharded")
    ms_ = fresh()                                # the model object (host-side normalisation of the initial state) is not the fit
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vb.fit_cell_sharded(ms_, counts, None, **kw)
    torch.cuda.synchronize()
    t_sh = _max_over_ranks(torch, dist, world, time.perf_counter() - t0)
    barrier()
    sh.PHASES["on"] = True                       # a second, instrumented run: where the time goes (syncs at every mark)
    try:
        vb.fit_cell_sharded(fresh(), counts, None, **kw)
    finally:
        sh.PHASES["on"] = False
    phases = {k: _max_over_ranks(torch, dist, world, v) for k, v in sorted(sh.PHASES["t"].items())}
    barrier()
    out = None
    if rank == 0:
        fresh().fit(counts, None, max_iter=3, min_iter=3, verbose=False)
        m1 = fresh()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m1.fit(counts, None, **kw)
        torch.cuda.synchronize()
        t_one = time.perf_counter() - t0
        nz = m1.ID_prob > 1e-300
        out = {"iterations": iters, "n_gpus": world, "sharded_s": t_sh, "one_gpu_s": t_one, "speedup": t_one / t_sh,
               "it_per_s_sharded": iters / t_sh, "it_per_s_one_gpu": iters / t_one, "cut_shard_ms_one_off": cut_ms,
               "phases_s": phases, "ms_per_iteration_in_loop": 1e3 * phases.get("loop", 0.0) / iters,
               "parity_vs_one_gpu": {
                   "elbo_max_rel_diff": float(np.max(np.abs(ms_.ELBO_ - m1.ELBO_) / np.abs(m1.ELBO_))),
                   "id_prob_max_rel_diff": float(np.max(np.abs(ms_.ID_prob[nz] - m1.ID_prob[nz]) / m1.ID_prob[nz])),
                   "gt_prob_max_abs_diff": float(np.max(np.abs(ms_.GT_prob - m1.GT_prob))),
                   "identical_argmax_donor": bool(np.array_equal(ms_.ID_prob.argmax(1), m1.ID_prob.argmax(1)))}}
    barrier()
    return out


def measure_wrap(torch, dist, vb, rank, world, barrier, counts, w, n_init=8):
    """Wall time of one whole vireo_wrap call (reference vireo_wrap.py:52-152: n_init warm-ups of 20 iterations ->
    argmax -> final fit of up to 200 iterations -> doublet pass) on the N GPUs, split into its phases."""
    import contextlib
    import io
    import importlib
    vw = importlib.import_module("vireo_b200.vireo_wrap")     # the module (the package re-exports the function by that name)
    K = w["K"]
    kw = dict(n_donor=K, n_init=n_init, random_seed=1)
    if w["mode"] == "gt_given":
        kw.update(GT_prior=gt_prior_of(w), learn_GT=False)
    vw.PHASES["on"] = True
    res = []
    try:
        for rep in range(2):                 # first call: one-off shard cuts / format builds / allocations
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                rv = vb.vireo_wrap(counts, None, **kw)
            torch.cuda.synchronize()
            total = _max_over_ranks(torch, dist, world, time.perf_counter() - t0)
            phases = {k: _max_over_ranks(torch, dist, world, v) for k, v in sorted(vw.PHASES["t"].items())}
            res.append((total, phases, rv))
    finally:
        vw.PHASES["on"] = False
    total, phases, rv = res[1]
    return {"n_init": n_init, "n_gpus": world, "total_s": total, "first_call_s": res[0][0],
            "phases_s": phases,
            "LB_doublet": float(rv["LB_doublet"]), "winner": int(np.argmax(rv["LB_list"])),
            "what": "vireo_wrap(counts, None, n_donor=%d, n_init=%d, random_seed=1): warm-ups sharded by restart, "
                    "final fit and doublet pass sharded by cell when N > 1; phases are max over ranks" % (K, n_init)}


def run_b200_bmm(args, torch, dist, vb, rank, dev, world, barrier, AD, DP, w):
    """BASELINE cfg5: BinomMixtureVB clone mode, n_init = 50 restarts sharded round-robin over the GPUs (strong scaling)."""
    from vireo_b200 import _engine, _lib
    C_, V, K = w["C"], w["V"], w["K"]
    t0 = time.perf_counter()
    counts = vb.stage(AD, DP)
    torch.cuda.synchronize()
    staging_ms = 1e3 * (time.perf_counter() - t0)
    binom = float(counts.binom_const())
    np.random.seed(1)
    model = vb.BinomMixtureVB(n_cell=C_, n_var=V, n_donor=K)
    starts = [model._draw_state(None, None, None) for _ in range(BMM_INIT)]
    mine = [i for i in range(BMM_INIT) if i % world == rank]
    batch = _engine.BmmBatch(counts, model, [starts[i] for i in mine])
    init_dev = [t.clone() for t in (batch.id_prob, batch.beta_mu, batch.beta_sum)]

    def step():
        for dst, src in zip((batch.id_prob, batch.beta_mu, batch.beta_sum), init_dev):
            dst.copy_(src)
        batch.run_fit(T_ITERS, T_ITERS, 1e-2, poll_every=T_ITERS + 1)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(dev) if rank == 0 else None
    barrier()
    torch.cuda.synchronize()
    before = _lib.launch_counts()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    after = _lib.launch_counts()
    ms = _max_over_ranks(torch, dist, world, ev0.elapsed_time(ev1))
    value = BMM_INIT * T_ITERS * args.steps / (ms / 1e3)
    _lib.load().vb_profile_enable(1)
    step()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.load().vb_profile_enable(0)
    e = 12 if counts.wide else 8
    alg = {"k_snp": counts.nnz * e + 8 * C_ * K + 16 * V * K, "k_cell": counts.nnz * e + 16 * V * K + 8 * C_ * K,
           "iter": 2 * counts.nnz * e + 16 * C_ * K + 32 * V * K + 4 * (C_ + V + 2)}
    nb = len(mine)
    kernels = {}
    tot = sum(v[0] for v in prof.values()) or 1.0
    for name, (kms, n) in prof.items():
        if n:
            kernels[name] = {"ms_per_launch": kms / n, "restarts_per_launch": nb, "launches": n, "share": kms / tot}
            if name in alg:
                kernels[name]["achieved_gbs"] = alg[name] * nb / (kms / n * 1e-3) / 1e9
    peak, peak_src, _ = _peaks()
    dom = max((k for k in kernels if k in ("k_cell", "k_snp")), key=lambda k: kernels[k]["share"], default=None)
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg[dom] * nb, "ms_per_launch": kernels[dom]["ms_per_launch"],
                    "note": "the whole problem (%.1f MB) is L2 resident: latency bound, not HBM bound" % (counts.nnz * e / 1e6)}
    # end to end: the public call of the clone mode, numpy in / numpy out
    import contextlib
    import io

    def call():
        m = vb.BinomMixtureVB(n_cell=C_, n_var=V, n_donor=K)
        with contextlib.redirect_stdout(io.StringIO()):
            m.fit(counts, None, n_init=BMM_INIT, min_iter=30, random_seed=1)
        return m

    call()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = call()
    torch.cuda.synchronize()
    call_s = _max_over_ranks(torch, dist, world, time.perf_counter() - t0)

    def fixed():
        mm = vb.BinomMixtureVB(n_cell=C_, n_var=V, n_donor=K)
        with contextlib.redirect_stdout(io.StringIO()):
            mm.fit(counts, None, n_init=BMM_INIT, max_iter=T_ITERS, max_iter_pre=T_ITERS, min_iter=T_ITERS, random_seed=1)

    fixed()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fixed()
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(torch, dist, world, time.perf_counter() - t0)
    clocks = sampler.stop(t_load0, None) if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import vireo_oracle as O
        np.random.seed(1)
        st = O.bmm_new(C_, V, K)
        t0 = time.perf_counter()
        O.bmm_fit_vb(st, AD, DP, max_iter=T_ITERS, min_iter=T_ITERS, verbose=False)
        dt = time.perf_counter() - t0
        cpu = {"value": T_ITERS / dt, "unit": "it/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": "%d EM iterations of one restart on the full matrices, scipy/numpy single thread" % T_ITERS}
    line = {
        "metric": "EM iterations/sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, w, counts.nnz),
        "ms_per_iteration_all_restarts": ms / args.steps / T_ITERS, "roofline": roofline, "kernels": kernels,
        "cpu_baseline": cpu,
        "e2e": {"value": (BMM_INIT + 1) * T_ITERS / e2e_s, "unit": "it/s",
                "h2d_bytes_per_step": 8 * BMM_INIT * (C_ * K + 2 * V * K) // world, "d2h_bytes_per_step": 8 * BMM_INIT * (C_ * K + 2 * V * K) // world,
                "api": "BinomMixtureVB.fit(counts, None, n_init=50, max_iter=max_iter_pre=min_iter=%d): 50 restarts + "
                       "the final refit, numpy state in / out" % T_ITERS},
        "fit_call": {"s": call_s, "elbo_final": float(m.ELBO_iters[-1]), "iterations_final_fit": int(len(m.ELBO_iters)),
                     "api": "BinomMixtureVB.fit(counts, None, n_init=50, min_iter=30, random_seed=1) -- BASELINE cfg5 "
                            "(39.8 s for the reference on one core, BASELINE.md)"},
        "staging_ms": staging_ms, "binom_const": binom,
        "gpu_launches": sum(after[k] - before[k] for k in after), "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--restarts", type=int, default=1, help="restarts per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-fixed32", action="store_true", help="skip the opt-in fixed-point family's leg")
    ap.add_argument("--no-wrap", action="store_true", help="skip the whole-call vireo_wrap measurement")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--ref-budget-s", type=float, default=120.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
