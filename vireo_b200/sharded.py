"""Cell-sharded single fit: ONE ``Vireo`` fit data-parallel over cells, for the fits that restart
sharding cannot spread -- the final fit of ``vireo_wrap`` (reference vireoSNP/utils/vireo_wrap.py:94) and the
GT-given mode, which has a single restart (vireo_wrap.py:48-50) -- plus the doublet pass that follows
(vireo_doublet.py:39-75), whose likelihood block is independent per cell.

Every rank holds the staged count columns of its own cells (cut on the device from the full staged matrices,
``vb_counts_slice``) and keeps ``ID_prob`` for them; ``GT_prob`` and theta are replicated.  The whole loop runs inside
the library (``vb_vireo_fit_sharded``): per EM iteration (reference vireo_model.py:257-264)

    SNP pass on the local cells      S1_r = AD_r @ ID_prob_r,  S2_r = (DP_r - AD_r) @ ID_prob_r
    ONE all-reduce (NCCL, sum)       S1 | S2 | {LB_p, KL_ID} of the previous iteration   <- the path's exchange step
    ELBO + convergence rule of the previous iteration, on the device, identical on every rank
    theta, GT update                 identical on every rank (same inputs, same kernels)
    cell pass on the local cells     ID_prob_r and the local parts of LB_p and KL_ID

The convergence rule, the ``ELBO[:it]`` quirk and the binomial constant follow the reference exactly
(vireo_model.py:266-276,313); the constant is the single float32 sum over the FULL matrices (every rank has them
staged), as in the reference.  On a single rank the same library loop runs without the exchange.
"""
import ctypes as C

import numpy as np

from . import _engine, _lib
from .dist import comm, world


def cell_shards(indptr, n_shard):
    """Contiguous cell ranges with about equal nnz: bounds[r] .. bounds[r+1] for shard r."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n_cell = indptr.size - 1
    total = int(indptr[-1])
    targets = (np.arange(1, n_shard) * total) // n_shard
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, n_cell), [n_cell]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def shard_of(counts):
    """(StagedCounts of this rank's cells, c0, c1, bounds of all ranks) -- cut once per staged matrix and world."""
    rank, ws = world()
    key = (ws, rank)
    hit = counts._shards.get(key)
    if hit is not None and hit[0]._h is not None:
        return hit
    bounds = cell_shards(counts.indptr, ws)
    c0, c1 = int(bounds[rank]), int(bounds[rank + 1])
    local = counts if ws == 1 else counts.slice_cells(c0, c1)
    counts._shards[key] = (local, c0, c1, bounds)
    return counts._shards[key]


def _gather_rows(cm, local, n_cols, bounds, dev):
    """All-gather of per-rank row blocks [c1 - c0, n_cols] (device tensor `local`) -> host array [n_cell, n_cols]:
    one NCCL all-gather of padded blocks, the blocks are packed on the device, ONE download."""
    t = _engine.torch()
    ws = cm.world
    if ws == 1:
        return _engine._to_host(local.reshape(-1)).reshape(-1, n_cols)
    per = max(int(np.max(np.diff(bounds))) * n_cols, 1)
    send = t.zeros(per, dtype=t.float64, device="cuda:%d" % dev)
    send[:local.numel()].copy_(local.reshape(-1))
    recv = t.empty(per * ws, dtype=t.float64, device="cuda:%d" % dev)
    _lib.check(_lib.load().vb_comm_allgather(cm.handle, _engine._ptr(send), _engine._ptr(recv), per, _engine._stream(dev)))
    out = t.empty(int(bounds[-1]) * n_cols, dtype=t.float64, device="cuda:%d" % dev)
    for r in range(ws):
        c0, c1 = int(bounds[r]), int(bounds[r + 1])
        out[c0 * n_cols:c1 * n_cols].copy_(recv[r * per:r * per + (c1 - c0) * n_cols])
    return _engine._to_host(out).reshape(int(bounds[-1]), n_cols)


# phase timing for bench.py (`sharded_fit.phases_s`): off unless PHASES["on"]; every mark synchronises the device
PHASES = {"on": False, "last": 0.0, "t": {}}


def _mark(name, device, reset=False):
    if PHASES["on"]:
        import time
        _engine.torch().cuda.synchronize(device)
        now = time.perf_counter()
        if reset:
            PHASES["t"] = {}
        else:
            PHASES["t"][name] = PHASES["t"].get(name, 0.0) + now - PHASES["last"]
        PHASES["last"] = now


def fit_cell_sharded(model, AD, DP=None, max_iter=200, min_iter=5, epsilon_conv=1e-2, delay_fit_theta=0, verbose=True,
                     poll_every=0):
    """``model.fit(AD, DP, ...)`` with the cells sharded over the ranks enabled by ``vireo_b200.dist.enable()``.
    Every rank passes the same model state and the same matrices (or their ``StagedCounts``); on return every rank
    holds the complete fitted state and the same ``ELBO_``."""
    t = _engine.torch()
    if getattr(model, "ASE_mode", False):
        raise NotImplementedError("cell-sharded fit: ASE mode keeps theta per SNP; use model.fit")
    counts = _engine.stage(AD, DP)
    _mark("", counts.device, reset=True)
    local, c0, c1, bounds = shard_of(counts)
    dev = counts.device
    cm = comm(dev)
    K, V = int(model.n_donor), int(model.n_var)
    _mark("shard_and_comm", dev)
    batch = _engine.VireoBatch(local, [model], rows=(c0, c1))
    xbuf = t.zeros(2 * V * K + 8, dtype=t.float64, device="cuda:%d" % dev)
    a = batch.args(max_iter, min_iter, epsilon_conv, delay_fit_theta, poll_every)
    _mark("upload", dev)
    _lib.check(_lib.load().vb_vireo_fit_sharded(local.handle, C.byref(a), cm.handle, _engine._ptr(xbuf),
                                                _engine._stream(dev)))
    batch.max_iter = max_iter
    _mark("loop", dev)
    elbo, last = batch.traces()[0]
    _engine.replay_convergence(elbo, last, max_iter, min_iter, epsilon_conv, False, verbose)

    # results: GT_prob and theta are replicated, ID_prob is gathered
    model.ID_prob = _gather_rows(cm, batch.id_prob, K, bounds, dev)
    if model.learn_GT:
        model.GT_prob = _engine._to_host(batch.gt_prob).reshape(V, K, int(model.n_GT))
    th = _engine._to_host(batch.state[batch.id_prob.numel() + batch.gt_prob.numel():])
    n_th = batch.beta_mu.numel()
    model.beta_mu, model.beta_sum = th[:n_th].reshape(1, -1).copy(), th[n_th:].reshape(1, -1).copy()
    batch.close()
    trace = elbo[:last].copy() + counts.binom_const()      # ELBO[:it] (vireo_model.py:276) + the constant (:313)
    model.ELBO_ = np.append(model.ELBO_, trace)
    _mark("download", dev)
    return trace


def predict_doublet_sharded(vobj, AD, DP=None, update_GT=True, update_ID=True, doublet_rate_prior=None):
    """``predict_doublet`` (reference vireoSNP/utils/vireo_doublet.py:11-82) with the cells sharded over the enabled
    ranks: every rank runs the doublet likelihood block and softmax on its own cells, ONE all-gather brings the
    posteriors and LLRs together, and the closing ``update_GT_prob`` (:75) is a local SNP pass + one all-reduce of
    S1 | S2.  Returns what ``predict_doublet`` returns, identical on every rank."""
    from .vireo_doublet import doublet_log_prior
    t = _engine.torch()
    counts = _engine.stage(AD, DP)
    local, c0, c1, bounds = shard_of(counts)
    dev = counts.device
    cm = comm(dev)
    K, V, G = int(vobj.n_donor), int(vobj.n_var), int(vobj.n_GT)
    K2 = K + K * (K - 1) // 2
    log_prior_both = doublet_log_prior(vobj, counts.n_cell, doublet_rate_prior)
    if log_prior_both.shape[0] > 1:
        log_prior_both = log_prior_both[c0:c1]
    pr, llr = _engine.doublet_pass(local, np.asarray(vobj.GT_prob, dtype=np.float64), vobj.beta_mu, vobj.beta_sum,
                                   log_prior_both, vobj.ASE_mode, keep_device=True)
    prob_both = _gather_rows(cm, pr, K2, bounds, dev)
    llr_all = _gather_rows(cm, llr, 1, bounds, dev).reshape(-1)
    if update_ID:
        vobj.ID_prob = prob_both[:, :K].copy()
    if update_GT:
        if update_ID:
            batch = _engine.VireoBatch(local, [vobj], rows=(c0, c1))
            xbuf = t.zeros(2 * V * K + 8, dtype=t.float64, device="cuda:%d" % dev)
            a = batch.args()
            _lib.check(_lib.load().vb_vireo_gt_sharded(local.handle, C.byref(a), cm.handle, _engine._ptr(xbuf),
                                                       _engine._stream(dev)))
            vobj.GT_prob = _engine._to_host(batch.gt_prob).reshape(V, K, G)
            batch.close()
        else:
            print("For update_GT, please turn on update_ID.")
    return prob_both[:, K:], prob_both[:, :K], llr_all
