"""``BinomMixtureVB`` -- drop-in for ``vireoSNP.BinomMixtureVB`` (reference
vireoSNP/utils/bmm_model.py:9-263): clone assignment from deep (mitochondrial) allele counts with a
Beta posterior per (variant, clone).  All restarts of ``fit`` run as ONE device batch (the reference
runs them one after the other in a Python loop, bmm_model.py:242-254).
"""
import numpy as np

from . import _engine, _lib
from .vireo_base import normalize


class BinomMixtureVB():
    """Binomial mixture model with variational inference; same API as the reference class.

    Key properties: ``beta_mu``, ``beta_sum`` (n_var, n_donor); ``ID_prob`` (n_cell, n_donor);
    ``ELBO_iters``, ``ELBO_inits`` after ``fit``.
    """

    def __init__(self, n_cell, n_var, n_donor, fix_beta_sum=False,
                 beta_mu_init=None, beta_sum_init=None, ID_prob_init=None):
        self.n_var, self.n_cell, self.n_donor = n_var, n_cell, n_donor
        self.fix_beta_sum = fix_beta_sum
        self.ID_prob_init, self.beta_mu_init, self.beta_sum_init = ID_prob_init, beta_mu_init, beta_sum_init
        self.set_prior()                                   # priors first, then the random start (bmm_model.py:57-63)
        self.set_initial(beta_mu_init, beta_sum_init, ID_prob_init)

    def _draw_state(self, beta_mu_init=None, beta_sum_init=None, ID_prob_init=None):
        mu = np.full((self.n_var, self.n_donor), 0.5) if beta_mu_init is None else beta_mu_init
        sm = np.full(np.shape(mu), 30.0) if beta_sum_init is None else beta_sum_init
        if ID_prob_init is None:
            ID_prob_init = np.random.rand(self.n_cell, self.n_donor)
        return dict(beta_mu=mu, beta_sum=sm, ID_prob=normalize(ID_prob_init, axis=1))

    def set_initial(self, beta_mu_init=None, beta_sum_init=None, ID_prob_init=None):
        """Start state: mu = 0.5, sum = 30, ID_prob ~ normalised U(0,1) unless given; clears
        ``ELBO_iters`` (reference bmm_model.py:67-85)."""
        st = self._draw_state(beta_mu_init, beta_sum_init, ID_prob_init)
        self.beta_mu, self.beta_sum, self.ID_prob = st["beta_mu"], st["beta_sum"], st["ID_prob"]
        self.ELBO_iters = np.array([])

    def set_prior(self, ID_prior=None, beta_mu_prior=None, beta_sum_prior=None):
        """Beta(1, 1) on every theta and a uniform assignment prior unless given
        (reference bmm_model.py:87-106)."""
        if beta_mu_prior is None:
            beta_mu_prior = np.full((self.n_var, self.n_donor), 0.5)
        if beta_sum_prior is None:
            beta_sum_prior = np.full(np.shape(beta_mu_prior), 2.0)
        self.theta_s1_prior = beta_mu_prior * beta_sum_prior
        self.theta_s2_prior = (1 - beta_mu_prior) * beta_sum_prior
        if ID_prior is None:
            self.ID_prior = normalize(np.ones((self.n_cell, self.n_donor)))
        else:
            self.ID_prior = ID_prior[np.newaxis, :] if ID_prior.ndim == 1 else ID_prior

    @property
    def theta_s1(self):
        return self.beta_mu * self.beta_sum

    @property
    def theta_s2(self):
        return (1 - self.beta_mu) * self.beta_sum

    # -- single updates --------------------------------------------------------------------------
    def _state(self):
        return dict(ID_prob=self.ID_prob, beta_mu=self.beta_mu, beta_sum=self.beta_sum)

    def _batch(self, AD, DP):
        counts = _engine.stage(AD, DP)
        self._last_counts = counts
        return _engine.BmmBatch(counts, self, [self._state()])

    def get_E_logLik(self, AD, DP):
        """E_theta[log P(AD | DP, theta, Z)], shape (n_cell, n_donor) (reference bmm_model.py:118-130)."""
        b = self._batch(AD, DP)
        b.run_step(_lib.PH_LOGLIK)
        return b.loglik_host()[0]

    def update_theta_size(self, AD, DP):
        """Beta posterior update (reference bmm_model.py:133-144)."""
        b = self._batch(AD, DP)
        b.run_step(_lib.PH_SNP | _lib.PH_THETA)
        _, mu, sm = b.download()
        self.beta_mu, self.beta_sum = mu[0], sm[0]

    def update_ID_prob(self, AD=None, DP=None, logLik_ID=None):
        """Assignment update (reference bmm_model.py:147-154)."""
        if logLik_ID is None:
            b = self._batch(AD, DP)
            b.run_step(_lib.PH_ID)
            self.ID_prob = b.download()[0][0]
        else:   # a caller-supplied logLik: plain host softmax, nothing sparse is involved
            lg = logLik_ID + np.log(self.ID_prior)
            lg = lg - lg.max(axis=1, keepdims=True)
            self.ID_prob = normalize(np.exp(lg))

    def get_ELBO(self, AD=None, DP=None, logLik_ID=None):
        """ELBO of the current state (reference bmm_model.py:157-175).  The reference's
        ``logLik_ID=None`` branch discards its own result and then fails; here that branch recomputes
        logLik_ID on the device instead."""
        if AD is not None:
            b = self._batch(AD, DP)
        else:
            counts = getattr(self, "_last_counts", None)
            if counts is None or counts._h is None:
                raise ValueError("get_ELBO needs AD, DP on first use")
            b = _engine.BmmBatch(counts, self, [self._state()])
        if logLik_ID is None:
            b.run_step(_lib.PH_LOGLIK | _lib.PH_ELBO)
        else:
            b.loglik.copy_(_engine._dev(np.asarray(logLik_ID, dtype=np.float64).reshape(-1), b.dev))
            b.run_step(_lib.PH_ELBO)
        return float(b.scalars()[0, 0])

    # -- fit -------------------------------------------------------------------------------------
    def _fit_BV(self, AD, DP, max_iter=200, min_iter=20, epsilon_conv=1e-2, verbose=True):
        """One restart's coordinate ascent on the device (reference bmm_model.py:178-201); appends
        ELBO[:it] to ``ELBO_iters``."""
        counts = _engine.stage(AD, DP)
        self._last_counts = counts
        b = _engine.BmmBatch(counts, self, [self._state()])
        b.run_fit(max_iter, min_iter, epsilon_conv)
        idp, mu, sm = b.download()
        self.ID_prob, self.beta_mu, self.beta_sum = idp[0], mu[0], sm[0]
        elbo, last = b.traces()[0]
        _engine.replay_convergence(elbo, last, max_iter, min_iter, epsilon_conv, True, verbose)
        self.ELBO_iters = np.append(self.ELBO_iters, elbo[:last])

    def fit(self, AD, DP, n_init=10, max_iter=200, max_iter_pre=100, random_seed=None, **kwargs):
        """Multi-restart fit (reference bmm_model.py:204-263): ``n_init`` restarts of at most
        ``max_iter_pre`` iterations, keep the one with the best final ELBO, refit it for up to
        ``max_iter`` more, add the binomial constant.

        The restarts consume the numpy RNG in the reference's order (one ``rand(n_cell, n_donor)`` per
        restart, bmm_model.py:80-83) and then run together as one device batch; after
        ``vireo_b200.dist.enable()`` they are sharded round-robin over the ranks first.
        ``kwargs``: min_iter, epsilon_conv, verbose for the inner loop."""
        if random_seed is not None:
            np.random.seed(random_seed)
        if type(DP) is np.ndarray and np.mean(DP > 0) < 0.3:
            print("Warning: input matrices is %.1f%% sparse, " % (100 - np.mean(DP > 0) * 100) +
                  "change to scipy.sparse.csc_matrix")
        min_iter = kwargs.get("min_iter", 20)
        eps = kwargs.get("epsilon_conv", 1e-2)
        verbose = kwargs.get("verbose", True)
        counts = _engine.stage(AD, DP)
        self._last_counts = counts
        const = counts.binom_const()

        starts = [self._draw_state(self.beta_mu_init, self.beta_sum_init, self.ID_prob_init) for _ in range(n_init)]
        from .dist import check_same_problem, gather_restarts, shard_restarts
        check_same_problem(counts.n_cell, counts.n_var, counts.nnz, self.n_donor, n_init, random_seed, float(const),
                           device=counts.device)
        mine = shard_restarts(n_init)
        results = {}
        if mine:
            b = _engine.BmmBatch(counts, self, [starts[i] for i in mine])
            b.run_fit(max_iter_pre, min_iter, eps)
            idp, mu, sm = b.download()
            for slot, (i, (elbo, last)) in enumerate(zip(mine, b.traces())):
                results[i] = dict(elbo=elbo, last=last, ID_prob=idp[slot], beta_mu=mu[slot], beta_sum=sm[slot])
        final = np.array([results[i]["elbo"][results[i]["last"] - 1] if i in results else -np.inf
                          for i in range(n_init)])
        final, best, winner = gather_restarts(final, results, ("elbo", "last", "ID_prob", "beta_mu", "beta_sum"),
                                              counts.device)
        for i in range(n_init):   # the reference prints each restart's warnings as it goes
            if i in results:
                _engine.replay_convergence(results[i]["elbo"], results[i]["last"], max_iter_pre, min_iter, eps, True,
                                           verbose)
        self.ELBO_inits = list(final)
        self.set_initial(winner["beta_mu"], winner["beta_sum"], winner["ID_prob"])
        self.ELBO_iters = winner["elbo"][:winner["last"]].copy()
        self._fit_BV(counts, None, max_iter=max_iter, min_iter=min_iter, epsilon_conv=eps, verbose=verbose)
        self.ELBO_iters = self.ELBO_iters + const
        self.ELBO_inits = np.array(self.ELBO_inits) + const

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_last_counts", None)
        return d
