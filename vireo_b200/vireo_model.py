"""``Vireo`` -- drop-in for ``vireoSNP.Vireo`` (reference vireoSNP/utils/vireo_model.py:11-315) whose
coordinate-ascent updates, ELBO and fit loop run as sm_100a CUDA kernels (csrc/vb_em.cu).

The object keeps the reference's host-side contract: every attribute (``ID_prob``, ``GT_prob``,
``beta_mu``, ``beta_sum``, the priors, ``ELBO_``) is a float64 numpy array that callers may read or
overwrite between calls; each method uploads the current attributes, runs on the device and writes the
results back.  The count matrices are staged to HBM once and cached (see ``_engine.stage``).
"""
import numpy as np

from . import _engine, _lib
from .vireo_base import normalize


class Vireo():
    """Variational inference for donor deconvolution -- same constructor, attributes and methods as
    ``vireoSNP.Vireo`` (reference vireo_model.py:27-30 for the signature).

    Key properties
    --------------
    beta_mu, beta_sum : (1, n_GT) or (n_var, n_GT) in ASE mode -- Beta posterior of theta
    ID_prob : (n_cell, n_donor) posterior donor assignment
    GT_prob : (n_var, n_donor, n_GT) posterior donor genotype
    """

    def __init__(self, n_cell, n_var, n_donor, n_GT=3, learn_GT=True,
                 learn_theta=True, ASE_mode=False, fix_beta_sum=False,
                 beta_mu_init=None, beta_sum_init=None, ID_prob_init=None,
                 GT_prob_init=None):
        self.n_GT, self.n_var, self.n_cell, self.n_donor = n_GT, n_var, n_cell, n_donor
        self.learn_GT, self.ASE_mode = learn_GT, ASE_mode
        self.learn_theta, self.fix_beta_sum = learn_theta, fix_beta_sum
        self.ELBO_ = np.zeros((0))
        self.set_initial(beta_mu_init, beta_sum_init, ID_prob_init, GT_prob_init)
        self.set_prior()

    # -- state ---------------------------------------------------------------------------------
    def set_initial(self, beta_mu_init=None, beta_sum_init=None, ID_prob_init=None, GT_prob_init=None):
        """Initial values; the legacy numpy RNG is consumed in the reference's order -- ID_prob first,
        then GT_prob, each only when not supplied (reference vireo_model.py:78-104)."""
        n_theta = self.n_var if self.ASE_mode else 1
        grid = np.linspace(0.01, 0.99, self.n_GT).reshape(1, -1)
        self.beta_mu = np.ones((n_theta, self.n_GT)) * grid if beta_mu_init is None else beta_mu_init
        self.beta_sum = np.ones((n_theta, self.n_GT)) * 50 if beta_sum_init is None else beta_sum_init
        if ID_prob_init is None:
            ID_prob_init = np.random.rand(self.n_cell, self.n_donor)
        self.ID_prob = normalize(ID_prob_init, axis=1)
        if GT_prob_init is None:
            GT_prob_init = np.random.rand(self.n_var, self.n_donor, self.n_GT)
        self.GT_prob = normalize(GT_prob_init)

    def set_prior(self, GT_prior=None, ID_prior=None, beta_mu_prior=None, beta_sum_prior=None, min_GP=0.00001):
        """Priors, shaped like their variables (reference vireo_model.py:107-137).  As in the
        reference, a supplied ``GT_prior`` is clipped to [min_GP, 1 - min_GP] IN PLACE before it is
        normalised -- ``vireo_wrap`` relies on that side effect for its later models."""
        if beta_mu_prior is None:
            beta_mu_prior = np.linspace(0.01, 0.99, self.beta_mu.shape[1])[np.newaxis, :]
        if beta_sum_prior is None:
            beta_sum_prior = np.full(beta_mu_prior.shape, 50.0)
        self.theta_s1_prior = beta_mu_prior * beta_sum_prior
        self.theta_s2_prior = (1 - beta_mu_prior) * beta_sum_prior

        if ID_prior is None:      # normalize(np.ones(shape)) of the reference, without the two passes over the array
            self.ID_prior = np.full(self.ID_prob.shape, 1.0 / self.ID_prob.shape[-1])
        else:
            self.ID_prior = ID_prior[np.newaxis, :] if ID_prior.ndim == 1 else ID_prior

        if GT_prior is None:
            self.GT_prior = np.full(self.GT_prob.shape, 1.0 / self.GT_prob.shape[-1])
        else:
            if GT_prior.ndim == 2:
                GT_prior = GT_prior[np.newaxis, :, :]
            np.clip(GT_prior, min_GP, 1 - min_GP, out=GT_prior)
            self.GT_prior = normalize(GT_prior)

    @property
    def theta_s1(self):
        """First Beta shape of theta's posterior."""
        return self.beta_mu * self.beta_sum

    @property
    def theta_s2(self):
        """Second Beta shape of theta's posterior."""
        return (1 - self.beta_mu) * self.beta_sum

    def _psi(self, x):
        # host mirror of the device digamma, for callers that read the reference's digamma properties
        # (reference vireo_model.py:149-162); the kernels compute their own.
        from scipy.special import digamma
        return np.expand_dims(digamma(x), 1)

    @property
    def digamma1_(self):
        return self._psi(self.theta_s1)

    @property
    def digamma2_(self):
        return self._psi(self.theta_s2)

    @property
    def digammas_(self):
        return self._psi(self.theta_s1 + self.theta_s2)

    # -- single updates (each is one or two kernel launches on the staged matrices) -----------------
    def _batch(self, AD, DP):
        self._last_counts = _engine.stage(AD, DP)
        return _engine.VireoBatch(self._last_counts, [self])

    def update_theta_size(self, AD, DP):
        """theta posterior update (reference vireo_model.py:165-185): SNP-major pass + k_theta."""
        b = self._batch(AD, DP)
        b.run_step(_lib.PH_SNP | _lib.PH_THETA)
        b.download(("theta",))

    def update_ID_prob(self, AD, DP):
        """Donor assignment update (reference vireo_model.py:187-201).  Returns logLik_ID (n_cell, n_donor)."""
        b = self._batch(AD, DP)
        b.run_step(_lib.PH_ID)
        b.download(("ID_prob",))
        return b.loglik_host()[0]

    def update_GT_prob(self, AD, DP):
        """Genotype update (reference vireo_model.py:204-219)."""
        b = self._batch(AD, DP)
        b.run_step(_lib.PH_SNP | _lib.PH_GT)
        b.download(("GT_prob",))

    def get_ELBO(self, logLik_ID, AD=None, DP=None):
        """Evidence lower bound of the current state (reference vireo_model.py:222-248).

        ``logLik_ID`` is the array returned by ``update_ID_prob``; when None it is recomputed from
        AD, DP without touching ``ID_prob``."""
        if logLik_ID is None:
            b = self._batch(AD, DP)
            b.run_step(_lib.PH_LOGLIK | _lib.PH_ELBO)
        else:
            counts = self._counts_for_elbo(AD, DP)
            b = _engine.VireoBatch(counts, [self])
            b.loglik.copy_(_engine._dev(np.asarray(logLik_ID, dtype=np.float64).reshape(-1), b.dev))
            b.run_step(_lib.PH_ELBO)
        return float(b.scalars()[0, 0])

    def _counts_for_elbo(self, AD, DP):
        if AD is not None:
            counts = _engine.stage(AD, DP)
            self._last_counts = counts
            return counts
        counts = getattr(self, "_last_counts", None)
        if counts is None or counts._h is None or counts.shape != (self.n_var, self.n_cell):
            raise ValueError("get_ELBO(logLik_ID) needs AD, DP on first use (the device path keeps "
                             "everything next to the staged count matrices)")
        return counts

    # -- fit loop --------------------------------------------------------------------------------
    def _fit_VB(self, AD, DP, max_iter=200, min_iter=5, epsilon_conv=1e-2, delay_fit_theta=0, verbose=True):
        """The whole coordinate-ascent loop on the device (reference vireo_model.py:251-276): theta ->
        GT -> ID -> ELBO per iteration, the convergence rule evaluated by a kernel.  Returns ELBO[:it]
        exactly as the reference does (the last computed value is dropped)."""
        counts = _engine.stage(AD, DP)
        self._last_counts = counts
        return _engine.vireo_fit_models(counts, [self], max_iter, min_iter, epsilon_conv, delay_fit_theta,
                                        verbose)[0]

    def fit(self, AD, DP, max_iter=200, min_iter=5, epsilon_conv=1e-2, delay_fit_theta=0, verbose=True,
            n_inits=50, nproc=1):
        """Fit the model (reference vireo_model.py:278-315).

        AD, DP : scipy.sparse.csc_matrix (n_var, n_cell) -- or a ``StagedCounts`` in place of AD.
        ``n_inits`` and ``nproc`` are accepted and unused, as in the reference."""
        if type(DP) is np.ndarray and np.mean(DP > 0) < 0.3:
            print("Warning: input matrices is %.1f%% sparse, " % (100 - np.mean(DP > 0) * 100) +
                  "change to scipy.sparse.csc_matrix")
        counts = _engine.stage(AD, DP)
        ELBO = self._fit_VB(counts, None, max_iter, min_iter, epsilon_conv, delay_fit_theta, verbose)
        ELBO += counts.binom_const()
        self.ELBO_ = np.append(self.ELBO_, ELBO)

    def __getstate__(self):
        # stay picklable like the reference object: the staged-matrix handle is process-local
        d = dict(self.__dict__)
        d.pop("_last_counts", None)
        return d
