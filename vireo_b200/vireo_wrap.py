"""``vireo_wrap`` -- drop-in for ``vireoSNP.vireo_wrap`` (reference vireoSNP/utils/vireo_wrap.py:22-183):
multi-restart donor deconvolution with model selection, optional extra-donor search, GT-prior refits
and doublet prediction.  ``vireo_flock`` is the pre-0.3 name of the same function
(reference doc/release.rst:138-140).

What changes underneath: the ``n_init`` warm-up fits run as device batches on the staged count
matrices instead of a ``multiprocessing.Pool`` of pickled models (vireo_wrap.py:74-83); ``nproc`` is
accepted and ignored.  After ``vireo_b200.dist.enable()`` (one process per GPU, see ``dist.py``) the
warm-ups are sharded over the ranks by restart, and the fits that follow model selection -- which restart
sharding cannot spread -- and the doublet pass are sharded by cell (``sharded.py``).
"""
import sys

import os
import threading

import numpy as np

from . import _engine
from .dist import check_same_problem, gather_restarts, shard_restarts, world
from .vireo_base import donor_select, optimal_match
from .vireo_doublet import predict_doublet
from .vireo_model import Vireo


# phase timing for bench.py (`wrap` key): off unless PHASES["on"]; every mark synchronises the device
PHASES = {"on": False, "last": 0.0, "t": {}}


def _mark(name, device):
    if PHASES["on"]:
        import time
        _engine.torch().cuda.synchronize(device)
        now = time.perf_counter()
        PHASES["t"][name] = PHASES["t"].get(name, 0.0) + now - PHASES["last"]
        PHASES["last"] = now


def _fit_batched(counts, models, max_iter, min_iter, delay_fit_theta):
    """``model.fit(..., verbose=False)`` for each model, as few device batches as memory allows."""
    if not models:
        return
    t = _engine.torch()
    m0 = models[0]
    per_restart = 8 * (3 * m0.n_cell * m0.n_donor + m0.n_var * m0.n_donor * (m0.n_GT + 4))
    free, _ = t.cuda.mem_get_info(counts.device)
    chunk = int(max(1, min(len(models), (0.4 * free) // max(per_restart, 1))))
    const = counts.binom_const()
    for lo in range(0, len(models), chunk):
        part = models[lo:lo + chunk]
        traces = _engine.vireo_fit_models(counts, part, max_iter, min_iter, 1e-2, delay_fit_theta, False)
        for m, elbo in zip(part, traces):
            m.ELBO_ = np.append(m.ELBO_, elbo + const)


def vireo_wrap(AD, DP, GT_prior=None, n_donor=None, learn_GT=True, n_init=20,
               random_seed=None, check_doublet=True, max_iter_init=20, delay_fit_theta=3,
               n_extra_donor=0, extra_donor_mode="distance",
               check_ambient=False, nproc=4, **kwargs):
    """Run Vireo with multiple initialisations; returns the reference's result dict
    (ID_prob, GT_prob, doublet_LLR, doublet_prob, theta_shapes, theta_mean, theta_sum, ambient_Psi,
    Psi_var, Psi_LLRatio, LB_list, LB_doublet)."""
    if type(DP) is np.ndarray and np.mean(DP > 0) < 0.3:
        print("Warning: input matrices is %.1f%% sparse, " % (100 - np.mean(DP > 0) * 100) +
              "change to scipy.sparse.csc_matrix")
    if learn_GT == False and n_extra_donor > 0:   # noqa: E712  (mirror the reference's truthiness)
        print("Searching from extra donors only works with learn_GT")
        n_extra_donor = 0
    if n_donor is None:
        if GT_prior is None:
            print("[vireo] Error: requiring n_donor or GT_prior.")
            sys.exit()
        n_donor = GT_prior.shape[1]
    if learn_GT is False and n_init > 1:
        print("GT is fixed, so use a single initialization")
        n_init = 1
    if check_ambient:
        raise NotImplementedError("check_ambient (\"under development\" in the reference, vireo.py:79-81) "
                                  "is outside the accelerated path")

    if PHASES["on"]:
        import time
        PHASES["t"], PHASES["last"] = {}, time.perf_counter()
    counts = _engine.stage(AD, DP)
    n_var, n_cell = counts.shape
    rank, ws = world()
    check_same_problem(n_cell, n_var, counts.nnz, n_donor, n_init, random_seed, n_extra_donor,
                       float(counts.binom_const()), device=counts.device)
    # fits of ONE model are sharded by cell over the ranks (ASE mode keeps theta per SNP and stays on one GPU;
    # VIREO_B200_SHARD_CELLS=0 keeps every rank fitting the full matrices redundantly)
    shard_cells = ws > 1 and not kwargs.get("ASE_mode", False) and os.environ.get("VIREO_B200_SHARD_CELLS", "1") != "0"

    def fit_one(model, **kw):
        if shard_cells:
            from .sharded import fit_cell_sharded
            fit_cell_sharded(model, counts, None, verbose=False, **kw)
        else:
            model.fit(counts, None, verbose=False, **kw)

    _mark("stage_and_check", counts.device)
    if random_seed is not None:
        np.random.seed(random_seed)

    GT_prior_use = None
    n_donor_use = int(n_donor + n_extra_donor)
    if GT_prior is not None and n_donor_use <= GT_prior.shape[1]:
        GT_prior_use = GT_prior.copy()
        n_donor_use = GT_prior.shape[1]

    # Every rank builds ALL models so the numpy RNG is consumed in the reference's order (vireo_wrap.py:65-71).  Drawing
    # and normalising the initial state of one restart takes about as long on the host as its warm-up fit takes on the
    # device (36 ms vs 35 ms at 100k x 50k x 16), so a builder thread draws the models in order while this thread fits
    # the ones that are ready -- whatever has accumulated goes into one device batch (results do not depend on the
    # batching).  numpy's generator, the array arithmetic and the library call all release the GIL.
    mine = set(shard_restarts(n_init))
    models, ready, failed = [None] * n_init, threading.Semaphore(0), []

    def build():
        try:
            for i in range(n_init):
                m = Vireo(n_var=n_var, n_cell=n_cell, n_donor=n_donor_use, learn_GT=learn_GT,
                          GT_prob_init=GT_prior_use, **kwargs)
                m.set_prior(GT_prior=GT_prior_use)
                models[i] = m
                ready.release()
        except BaseException as exc:       # re-raised by the fitting thread
            failed.append(exc)
            for _ in range(n_init):
                ready.release()

    builder = threading.Thread(target=build, name="vireo_wrap-draw", daemon=True)
    builder.start()
    done = 0
    while done < n_init:
        ready.acquire()
        n_new = 1
        while ready.acquire(blocking=False):
            n_new += 1
        if failed:
            builder.join()
            raise failed[0]
        batch = [models[i] for i in range(done, done + n_new) if i in mine]
        done += n_new
        _fit_batched(counts, batch, max_iter_init, 5, delay_fit_theta)
    builder.join()
    mine = sorted(mine)
    _mark("draw_and_warmups", counts.device)

    # model selection: one all-gather of the final ELBOs, winner's state broadcast by its owner
    final = np.array([models[i].ELBO_[-1] if i in mine else -np.inf for i in range(n_init)])
    results = {i: dict(ID_prob=models[i].ID_prob, GT_prob=models[i].GT_prob, beta_mu=models[i].beta_mu,
                       beta_sum=models[i].beta_sum, ELBO_=models[i].ELBO_) for i in mine}
    elbo_all, best, state = gather_restarts(final, results, ("ID_prob", "GT_prob", "beta_mu", "beta_sum", "ELBO_"),
                                            counts.device)
    modelCA = models[best]
    if ws > 1:
        modelCA.ID_prob, modelCA.GT_prob = state["ID_prob"], state["GT_prob"]
        modelCA.beta_mu, modelCA.beta_sum = state["beta_mu"], state["beta_sum"]
        modelCA.ELBO_ = np.atleast_1d(state["ELBO_"])
    _mark("select", counts.device)

    if n_extra_donor == 0:
        fit_one(modelCA, min_iter=5)          # <= 200 iterations: the fit that dominates the call (vireo_wrap.py:93-94)
    else:
        _ID_prob = donor_select(modelCA.GT_prob, modelCA.ID_prob, n_donor, mode=extra_donor_mode)
        modelCA = Vireo(n_var=n_var, n_cell=n_cell, n_donor=n_donor, learn_GT=learn_GT,
                        GT_prob_init=GT_prior_use, ID_prob_init=_ID_prob,
                        beta_mu_init=modelCA.beta_mu, beta_sum_init=modelCA.beta_sum, **kwargs)
        modelCA.set_prior(GT_prior=GT_prior_use)
        fit_one(modelCA, min_iter=5, delay_fit_theta=delay_fit_theta)

    _mark("final_fit", counts.device)
    print("[vireo] lower bound ranges [%.1f, %.1f, %.1f]"
          % (np.min(elbo_all), np.median(elbo_all), np.max(elbo_all)))

    # GT prior with more / fewer donors than requested: refit (vireo_wrap.py:111-136)
    if GT_prior is not None and n_donor < GT_prior.shape[1]:
        order = np.argsort(np.sum(modelCA.ID_prob, axis=0))[::-1]
        GT_prior_use = GT_prior[:, order[:n_donor], :]
        modelCA = Vireo(n_var=n_var, n_cell=n_cell, n_donor=n_donor, learn_GT=False,
                        GT_prob_init=GT_prior_use, **kwargs)
        fit_one(modelCA, min_iter=20)
    elif GT_prior is not None and n_donor > GT_prior.shape[1]:
        GT_prior_use = modelCA.GT_prob.copy()
        idx = optimal_match(GT_prior, GT_prior_use)[1]
        GT_prior_use[:, idx, :] = GT_prior
        order = np.append(idx, np.delete(np.arange(n_donor), idx))
        GT_prior_use = GT_prior_use[:, order, :]
        modelCA = Vireo(n_var=n_var, n_cell=n_cell, n_donor=n_donor, learn_GT=learn_GT,
                        ID_prob_init=modelCA.ID_prob[:, order], beta_mu_init=modelCA.beta_mu,
                        beta_sum_init=modelCA.beta_sum, GT_prob_init=GT_prior_use, **kwargs)
        modelCA.set_prior(GT_prior=GT_prior_use)
        fit_one(modelCA, min_iter=20)

    print("[vireo] allelic rate mean and concentrations:")
    print(np.round(modelCA.beta_mu, 3))
    print(np.round(modelCA.beta_sum, 1))

    print("[vireo] donor size before removing doublets:")
    _donor_cnt = np.sum(modelCA.ID_prob, axis=0)
    print("\t".join(["donor%d" % x for x in range(len(_donor_cnt))]))
    print("\t".join(["%.0f" % x for x in _donor_cnt]))

    if check_doublet and shard_cells:
        from .sharded import predict_doublet_sharded
        doublet_prob, ID_prob, doublet_LLR = predict_doublet_sharded(modelCA, counts, None)
    elif check_doublet:
        doublet_prob, ID_prob, doublet_LLR = predict_doublet(modelCA, counts, None)
    else:
        ID_prob = modelCA.ID_prob
        doublet_prob = np.zeros((n_cell, int(n_donor * (n_donor - 1) / 2)))
        doublet_LLR = np.zeros(n_cell)

    _mark("doublet", counts.device)
    theta_shapes = np.append(modelCA.beta_mu * modelCA.beta_sum,
                             (1 - modelCA.beta_mu) * modelCA.beta_sum, axis=0)

    return {
        'ID_prob': ID_prob, 'GT_prob': modelCA.GT_prob,
        'doublet_LLR': doublet_LLR, 'doublet_prob': doublet_prob,
        'theta_shapes': theta_shapes, 'theta_mean': modelCA.beta_mu, 'theta_sum': modelCA.beta_sum,
        'ambient_Psi': None, 'Psi_var': None, 'Psi_LLRatio': None,
        'LB_list': elbo_all, 'LB_doublet': modelCA.ELBO_[-1],
    }


vireo_flock = vireo_wrap
