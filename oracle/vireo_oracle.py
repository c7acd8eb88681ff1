"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.

A numpy/scipy restatement of vireoSNP's variational-EM hot path, written as
free functions over a plain state record.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; ``vireo_b200`` never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference (vireoSNP 0.5.9 from /root/reference) in the build container and
commits its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against every one of those vectors, including the
notebook known answer ``-190779.74335041404``
(reference ``examples/vireoSNP_clones.ipynb:103``).

The arithmetic deliberately keeps the reference's operation order (one sparse
product per genotype and per count matrix, a sparse ``DP - AD`` per call), so
that (i) results agree with the reference to rounding and (ii) timing this
module on host cores is a fair stand-in for timing the reference itself, which
cannot travel to the GPU box.

Each function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from scipy.sparse import csc_matrix, issparse
from scipy.special import betaln, binom, digamma
from scipy.stats import entropy

# --------------------------------------------------------------------------
# L1 helpers  (vireoSNP/utils/vireo_base.py)
# --------------------------------------------------------------------------


def unit_sum(x, axis=-1):
    """Divide by the sum along ``axis``.  vireo_base.py:44-56 (normalize)."""
    return x / np.sum(x, axis=axis, keepdims=True)


def shift_max(x, axis=-1):
    """Subtract the max along ``axis``.  vireo_base.py:62-74 (loglik_amplify)."""
    return x - np.max(x, axis=axis, keepdims=True)


def softmax_from_log(logits):
    """normalize(exp(loglik_amplify(.))) as used at vireo_model.py:198-199,218-219."""
    return unit_sum(np.exp(shift_max(logits)))


def binom_log_terms(AD, DP, cap=700):
    """Per-nnz log C(dp, ad), capped, rounded to float32.  vireo_base.py:7-22.

    Only entries with DP > 0 contribute (``:14``); the float64 log is capped at
    ``cap`` (``:19``) and the result is cast to float32 (``:20``) -- the caller
    then sums that float32 array (vireo_model.py:313, bmm_model.py:239).
    """
    mask = DP > 0
    a = AD[mask].astype(np.int64)
    d = DP[mask].astype(np.int64)
    out = np.log(binom(d, a))
    out[out > cap] = cap
    return out.astype(np.float32)


def binom_const(AD, DP):
    """float32 sum of :func:`binom_log_terms` (vireo_model.py:313)."""
    return np.sum(binom_log_terms(AD, DP))


def beta_kl(post, prior):
    """Sum of KL(Beta(post) || Beta(prior)).  vireo_base.py:77-127.

    ``post`` and ``prior`` are (T, 2, G): axis 1 holds (s1, s2).  The reference
    forms cross-entropies H(p, q) = betaln(q) - (q1-1)psi(p1) - (q2-1)psi(p2)
    + (q1+q2-2)psi(p1+p2)  (``:96-105``) and returns sum(H(p,q) - H(p,p)).
    """

    def cross(p, q):
        return (betaln(q[:, 0], q[:, 1])
                - (q[:, 0] - 1) * digamma(p[:, 0])
                - (q[:, 1] - 1) * digamma(p[:, 1])
                + (q.sum(axis=1) - 2) * digamma(p.sum(axis=1)))

    return np.sum(cross(post, prior) - cross(post, post))


# --------------------------------------------------------------------------
# Vireo model state  (vireoSNP/utils/vireo_model.py:27-137)
# --------------------------------------------------------------------------


@dataclass
class VireoState:
    n_cell: int
    n_var: int
    n_donor: int
    n_GT: int = 3
    learn_GT: bool = True
    learn_theta: bool = True
    ASE_mode: bool = False
    fix_beta_sum: bool = False
    beta_mu: np.ndarray = None
    beta_sum: np.ndarray = None
    ID_prob: np.ndarray = None
    GT_prob: np.ndarray = None
    ID_prior: np.ndarray = None
    GT_prior: np.ndarray = None
    theta_s1_prior: np.ndarray = None
    theta_s2_prior: np.ndarray = None
    ELBO_: np.ndarray = field(default_factory=lambda: np.zeros(0))

    # vireo_model.py:139-162
    @property
    def theta_s1(self):
        return self.beta_mu * self.beta_sum

    @property
    def theta_s2(self):
        return (1 - self.beta_mu) * self.beta_sum

    def psi(self):
        """(psi(s1), psi(s2), psi(s1+s2)) each shaped (T, 1, G).  :149-162."""
        s1, s2 = self.theta_s1, self.theta_s2
        return (digamma(s1)[:, None, :], digamma(s2)[:, None, :],
                digamma(s1 + s2)[:, None, :])


def vireo_new(n_cell, n_var, n_donor, n_GT=3, learn_GT=True, learn_theta=True,
              ASE_mode=False, fix_beta_sum=False, beta_mu_init=None,
              beta_sum_init=None, ID_prob_init=None, GT_prob_init=None):
    """Constructor: set_initial then set_prior.  vireo_model.py:27-76.

    The legacy global numpy RNG is consumed in the reference's order: ID_prob
    first (``:95``), then GT_prob (``:103``), each only when no init is given.
    """
    st = VireoState(n_cell, n_var, n_donor, n_GT, learn_GT, learn_theta,
                    ASE_mode, fix_beta_sum)
    vireo_set_initial(st, beta_mu_init, beta_sum_init, ID_prob_init, GT_prob_init)
    vireo_set_prior(st)
    return st


def vireo_set_initial(st, beta_mu_init=None, beta_sum_init=None,
                      ID_prob_init=None, GT_prob_init=None):
    """vireo_model.py:78-104."""
    rows = st.n_var if st.ASE_mode else 1
    if beta_mu_init is None:
        st.beta_mu = np.ones((rows, st.n_GT)) * np.linspace(0.01, 0.99, st.n_GT).reshape(1, -1)
    else:
        st.beta_mu = beta_mu_init
    st.beta_sum = np.ones((rows, st.n_GT)) * 50 if beta_sum_init is None else beta_sum_init
    if ID_prob_init is None:
        st.ID_prob = unit_sum(np.random.rand(st.n_cell, st.n_donor))
    else:
        st.ID_prob = unit_sum(ID_prob_init, axis=1)
    if GT_prob_init is None:
        st.GT_prob = unit_sum(np.random.rand(st.n_var, st.n_donor, st.n_GT))
    else:
        st.GT_prob = unit_sum(GT_prob_init)


def vireo_set_prior(st, GT_prior=None, ID_prior=None, beta_mu_prior=None,
                    beta_sum_prior=None, min_GP=0.00001):
    """vireo_model.py:107-137.  NOTE the in-place clipping of the caller's
    ``GT_prior`` array (``:132-133``) -- quirk Q5 depends on it."""
    if beta_mu_prior is None:
        beta_mu_prior = np.linspace(0.01, 0.99, st.beta_mu.shape[1])[None, :]
    if beta_sum_prior is None:
        beta_sum_prior = np.ones(beta_mu_prior.shape) * 50.0
    st.theta_s1_prior = beta_mu_prior * beta_sum_prior
    st.theta_s2_prior = (1 - beta_mu_prior) * beta_sum_prior

    if ID_prior is None:
        st.ID_prior = unit_sum(np.ones(st.ID_prob.shape))
    else:
        st.ID_prior = ID_prior[None, :] if ID_prior.ndim == 1 else ID_prior

    if GT_prior is None:
        st.GT_prior = unit_sum(np.ones(st.GT_prob.shape))
    else:
        if GT_prior.ndim == 2:
            GT_prior = GT_prior[None, :, :]
        GT_prior[GT_prior < min_GP] = min_GP           # in place, as the reference
        GT_prior[GT_prior > 1 - min_GP] = 1 - min_GP
        st.GT_prior = unit_sum(GT_prior)


# --------------------------------------------------------------------------
# Vireo coordinate-ascent updates
# --------------------------------------------------------------------------


def vireo_update_theta(st, AD, DP):
    """theta posterior (Beta shape) update.  vireo_model.py:165-185.

    Uses the CURRENT (old) GT_prob -- quirk Q6."""
    BD = DP - AD
    S1 = AD @ st.ID_prob
    S2 = BD @ st.ID_prob
    new_s1 = np.zeros(st.beta_mu.shape) + st.theta_s1_prior
    new_s2 = np.zeros(st.beta_mu.shape) + st.theta_s2_prior
    ax = 1 if st.ASE_mode else None
    for g in range(st.n_GT):
        new_s1[:, g:g + 1] += np.sum(S1 * st.GT_prob[:, :, g], axis=ax, keepdims=True)
        new_s2[:, g:g + 1] += np.sum(S2 * st.GT_prob[:, :, g], axis=ax, keepdims=True)
    st.beta_mu = new_s1 / (new_s1 + new_s2)
    if not st.fix_beta_sum:
        st.beta_sum = new_s1 + new_s2


def vireo_loglik_id(AD, DP, GT_prob, psi1, psi2, psis):
    """The "binom logLik" block: 3 sparse products per genotype.
    vireo_model.py:190-196 (same block at :228-234 and vireo_doublet.py:53-62)."""
    BD = DP - AD
    out = np.zeros((AD.shape[1], GT_prob.shape[1]))
    for g in range(GT_prob.shape[2]):
        out += (AD.T @ (GT_prob[:, :, g] * psi1[:, :, g])
                + BD.T @ (GT_prob[:, :, g] * psi2[:, :, g])
                - DP.T @ (GT_prob[:, :, g] * psis[:, :, g]))
    return out


def vireo_update_id(st, AD, DP):
    """vireo_model.py:187-201.  Returns logLik_ID (n_cell, n_donor)."""
    ll = vireo_loglik_id(AD, DP, st.GT_prob, *st.psi())
    st.ID_prob = softmax_from_log(ll + np.log(st.ID_prior))
    return ll


def vireo_update_gt(st, AD, DP):
    """vireo_model.py:204-219."""
    S1 = AD @ st.ID_prob
    SS = DP @ st.ID_prob
    S2 = SS - S1
    psi1, psi2, psis = st.psi()
    ll = np.zeros(st.GT_prior.shape)
    for g in range(st.n_GT):
        ll[:, :, g] = S1 * psi1[:, :, g] + S2 * psi2[:, :, g] - SS * psis[:, :, g]
    st.GT_prob = softmax_from_log(ll + np.log(st.GT_prior))


def vireo_elbo_terms(st, logLik_ID):
    """(LB_p, KL_ID, KL_GT, KL_theta).  vireo_model.py:236-245."""
    LB_p = np.sum(logLik_ID * st.ID_prob)
    KL_ID = np.sum(entropy(st.ID_prob, st.ID_prior, axis=-1))
    KL_GT = np.sum(entropy(st.GT_prob, st.GT_prior, axis=-1))
    post = np.stack([st.theta_s1, st.theta_s2], axis=1)
    prior = np.stack([st.theta_s1_prior, st.theta_s2_prior], axis=1)
    return LB_p, KL_ID, KL_GT, beta_kl(post, prior)


def vireo_elbo(st, logLik_ID, AD=None, DP=None):
    """vireo_model.py:222-248."""
    if logLik_ID is None:
        logLik_ID = vireo_loglik_id(AD, DP, st.GT_prob, *st.psi())
    LB_p, KL_ID, KL_GT, KL_th = vireo_elbo_terms(st, logLik_ID)
    return LB_p - KL_ID - KL_GT - KL_th


def vireo_fit_vb(st, AD, DP, max_iter=200, min_iter=5, epsilon_conv=1e-2,
                 delay_fit_theta=0, verbose=True, trace=None):
    """The EM loop.  vireo_model.py:251-276.

    Quirks kept: strict ``it > min_iter`` (Q2); a decrease only warns; the
    returned trace is ``ELBO[:it]`` -- the last computed value is dropped (Q1).
    ``trace`` (a list) optionally receives a copy of the state after every
    iteration, for teacher-forced tests.
    """
    ELBO = np.zeros(max_iter)
    it = 0
    for it in range(max_iter):
        if st.learn_theta and it >= delay_fit_theta:
            vireo_update_theta(st, AD, DP)
        if st.learn_GT:
            vireo_update_gt(st, AD, DP)
        ll = vireo_update_id(st, AD, DP)
        ELBO[it] = vireo_elbo(st, ll)
        if trace is not None:
            trace.append(dict(ID_prob=st.ID_prob.copy(), GT_prob=st.GT_prob.copy(),
                              beta_mu=st.beta_mu.copy(), beta_sum=st.beta_sum.copy(),
                              logLik_ID=ll.copy(), ELBO=ELBO[it]))
        if it > min_iter:
            if ELBO[it] < ELBO[it - 1] - 1e-6:
                if verbose:
                    print("Warning: Lower bound decreases!\n")
            elif it == max_iter - 1:
                if verbose:
                    print("Warning: VB did not converge!\n")
            elif ELBO[it] - ELBO[it - 1] < epsilon_conv:
                break
    return ELBO[:it]


def _maybe_sparsify(AD, DP):
    """vireo_model.py:300-305 / vireo_wrap.py:29-34 / bmm_model.py:232-237 (Q13)."""
    if type(DP) is np.ndarray and np.mean(DP > 0) < 0.3:
        print("Warning: input matrices is %.1f%% sparse, " % (100 - np.mean(DP > 0) * 100)
              + "change to scipy.sparse.csc_matrix")
        AD, DP = csc_matrix(AD), csc_matrix(DP)
    return AD, DP


def vireo_fit(st, AD, DP, max_iter=200, min_iter=5, epsilon_conv=1e-2,
              delay_fit_theta=0, verbose=True):
    """vireo_model.py:278-315: loop, add the binomial constant, append to ELBO_ (Q8)."""
    AD, DP = _maybe_sparsify(AD, DP)
    ELBO = vireo_fit_vb(st, AD, DP, max_iter, min_iter, epsilon_conv,
                        delay_fit_theta, verbose)
    ELBO += binom_const(AD, DP)
    st.ELBO_ = np.append(st.ELBO_, ELBO)
    return st


# --------------------------------------------------------------------------
# Doublets  (vireoSNP/utils/vireo_doublet.py:11-136)
# --------------------------------------------------------------------------


def doublet_theta(beta_mu, beta_sum):
    """vireo_doublet.py:85-102: pairwise mean of mu, geometric mean of sum."""
    pairs = np.array(list(itertools.combinations(range(beta_mu.shape[1]), 2)))
    mu2 = (beta_mu[:, pairs[:, 0]] + beta_mu[:, pairs[:, 1]]) / 2.0
    sum2 = np.sqrt(beta_sum[:, pairs[:, 0]] * beta_sum[:, pairs[:, 1]])
    return np.append(beta_mu, mu2, axis=-1), np.append(beta_sum, sum2, axis=-1)


def doublet_GT(GT_prob):
    """vireo_doublet.py:105-136: donor pairs x (G + G(G-1)/2) genotype classes."""
    V, K, G = GT_prob.shape
    gp = np.array(list(itertools.combinations(range(G), 2)))
    sp = np.array(list(itertools.combinations(range(K), 2)))
    A = GT_prob[:, sp[:, 0], :]
    B = GT_prob[:, sp[:, 1], :]
    pair = np.zeros((V, sp.shape[0], G + gp.shape[0]))
    pair[:, :, :G] = A * B
    pair[:, :, G:] = A[:, :, gp[:, 0]] * B[:, :, gp[:, 1]] + A[:, :, gp[:, 1]] * B[:, :, gp[:, 0]]
    pair = unit_sum(pair, axis=2)
    single = np.append(GT_prob, np.zeros((V, K, gp.shape[0])), axis=2)
    return np.append(single, pair, axis=1)


def vireo_predict_doublet(st, AD, DP, update_GT=True, update_ID=True,
                          doublet_rate_prior=None):
    """vireo_doublet.py:11-82.  Mutates ``st`` like the reference."""
    GT_both = doublet_GT(st.GT_prob)
    mu_both, sum_both = doublet_theta(st.beta_mu, st.beta_sum)
    n_pair = GT_both.shape[1] - st.GT_prob.shape[1]
    if doublet_rate_prior is None:
        doublet_rate_prior = min(0.5, AD.shape[1] / 100000)
    prior_both = np.append(st.ID_prior * (1 - doublet_rate_prior),
                           np.ones((st.n_cell, n_pair)) / n_pair * doublet_rate_prior,
                           axis=1)
    psi1 = digamma(sum_both * mu_both)[:, None, :]
    psi2 = digamma(sum_both * (1 - mu_both))[:, None, :]
    psis = digamma(sum_both)[:, None, :]
    ll = vireo_loglik_id(AD, DP, GT_both, psi1, psi2, psis)
    llr = ll[:, st.n_donor:].max(1) - ll[:, :st.n_donor].max(1)
    both = softmax_from_log(ll + np.log(prior_both))
    if update_ID:
        st.ID_prob = both[:, :st.n_donor]
    if update_GT:
        if update_ID:
            vireo_update_gt(st, AD, DP)
        else:
            print("For update_GT, please turn on update_ID.")
    return both[:, st.n_donor:], both[:, :st.n_donor], llr


# --------------------------------------------------------------------------
# vireo_wrap  (vireoSNP/utils/vireo_wrap.py:22-183)
# --------------------------------------------------------------------------


def donor_select(GT_prob, ID_prob, n_donor, mode="distance"):
    """vireo_base.py:217-254 (host-side glue for the extra-donor branch)."""
    cnt = np.sum(ID_prob, axis=0)
    K = GT_prob.shape[1]
    if mode == "size":
        order = list(np.argsort(cnt)[::-1])
    else:
        diff = np.zeros((K, K))
        for i in range(K):
            for j in range(K):
                diff[i, j] = np.mean(np.abs(GT_prob[:, i, :] - GT_prob[:, j, :]))
        order = [int(np.argmax(cnt))]
        left = np.delete(np.arange(K), order)
        diff = np.delete(diff, order, axis=1)
        while len(order) < diff.shape[0]:
            pick = np.argmax(np.min(diff[order, :], axis=0))
            order.append(left[pick])
            left = np.delete(left, pick)
            diff = np.delete(diff, pick, axis=1)
    print("[vireo] donor size with searching extra %d donors:" % (K - n_donor))
    print("\t".join(["donor%d" % x for x in order]))
    print("\t".join(["%.0f" % cnt[x] for x in order]))
    out = ID_prob[:, order[:n_donor]]
    out[out < 10 ** -10] = 10 ** -10
    return out


def optimal_match(X, Z, axis=1):
    """Hungarian alignment of Z's slices to X's.  vireo_base.py:187-206."""
    from scipy.optimize import linear_sum_assignment
    cost = np.zeros((X.shape[axis], Z.shape[axis]))
    for i in range(X.shape[axis]):
        for j in range(Z.shape[axis]):
            cost[i, j] = np.mean(np.abs(np.take(X, i, axis=axis) - np.take(Z, j, axis=axis)))
    return linear_sum_assignment(cost)


def vireo_wrap(AD, DP, GT_prior=None, n_donor=None, learn_GT=True, n_init=20,
               random_seed=None, check_doublet=True, max_iter_init=20,
               delay_fit_theta=3, n_extra_donor=0, extra_donor_mode="distance",
               **kwargs):
    """Multi-restart driver, serial.  vireo_wrap.py:22-183 minus the ambient
    branch (``:161-168``, "under development" in the reference).

    Construction order (Q4): all ``n_init`` models are built -- consuming the
    RNG -- before any is fitted (``:65-71``); selection uses ``ELBO_[-1]`` (Q8)
    and the winner is fitted again with ``max_iter=200, delay_fit_theta=0`` (Q10).
    """
    AD, DP = _maybe_sparsify(AD, DP)
    if not learn_GT and n_extra_donor > 0:
        n_extra_donor = 0
    if n_donor is None:
        n_donor = GT_prior.shape[1]
    if learn_GT is False and n_init > 1:
        n_init = 1
    if random_seed is not None:
        np.random.seed(random_seed)

    prior_use = None
    K_use = int(n_donor + n_extra_donor)
    if GT_prior is not None and K_use == GT_prior.shape[1]:
        prior_use = GT_prior.copy()
    elif GT_prior is not None and K_use < GT_prior.shape[1]:
        prior_use = GT_prior.copy()
        K_use = GT_prior.shape[1]

    V, C = AD.shape
    models = []
    for _ in range(n_init):
        m = vireo_new(C, V, K_use, learn_GT=learn_GT, GT_prob_init=prior_use, **kwargs)
        vireo_set_prior(m, GT_prior=prior_use)
        models.append(m)
    for m in models:
        vireo_fit(m, AD, DP, min_iter=5, max_iter=max_iter_init,
                  delay_fit_theta=delay_fit_theta, verbose=False)

    elbo_all = np.array([m.ELBO_[-1] for m in models])
    best = models[int(np.argmax(elbo_all))]
    if n_extra_donor == 0:
        vireo_fit(best, AD, DP, min_iter=5, verbose=False)
    else:
        idp = donor_select(best.GT_prob, best.ID_prob, n_donor, mode=extra_donor_mode)
        nxt = vireo_new(C, V, n_donor, learn_GT=learn_GT, GT_prob_init=prior_use,
                        ID_prob_init=idp, beta_mu_init=best.beta_mu,
                        beta_sum_init=best.beta_sum, **kwargs)
        vireo_set_prior(nxt, GT_prior=prior_use)
        vireo_fit(nxt, AD, DP, min_iter=5, delay_fit_theta=delay_fit_theta, verbose=False)
        best = nxt

    if GT_prior is not None and n_donor < GT_prior.shape[1]:
        cnt = np.sum(best.ID_prob, axis=0)
        keep = np.argsort(cnt)[::-1]
        prior_use = GT_prior[:, keep[:n_donor], :]
        best = vireo_new(C, V, n_donor, learn_GT=False, GT_prob_init=prior_use, **kwargs)
        vireo_fit(best, AD, DP, min_iter=20, verbose=False)

    elif GT_prior is not None and n_donor > GT_prior.shape[1]:
        prior_use = best.GT_prob.copy()
        idx = optimal_match(GT_prior, prior_use)[1]
        prior_use[:, idx, :] = GT_prior
        order = np.append(idx, np.delete(np.arange(n_donor), idx))
        prior_use = prior_use[:, order, :]
        nxt = vireo_new(C, V, n_donor, learn_GT=learn_GT, ID_prob_init=best.ID_prob[:, order],
                        beta_mu_init=best.beta_mu, beta_sum_init=best.beta_sum,
                        GT_prob_init=prior_use, **kwargs)
        vireo_set_prior(nxt, GT_prior=prior_use)
        vireo_fit(nxt, AD, DP, min_iter=20, verbose=False)
        best = nxt

    if check_doublet:
        dbl_prob, ID_prob, dbl_llr = vireo_predict_doublet(best, AD, DP)
    else:
        ID_prob = best.ID_prob
        dbl_prob = np.zeros((C, int(n_donor * (n_donor - 1) / 2)))
        dbl_llr = np.zeros(C)

    return dict(ID_prob=ID_prob, GT_prob=best.GT_prob, doublet_LLR=dbl_llr,
                doublet_prob=dbl_prob,
                theta_shapes=np.append(best.beta_mu * best.beta_sum,
                                       (1 - best.beta_mu) * best.beta_sum, axis=0),
                theta_mean=best.beta_mu, theta_sum=best.beta_sum,
                ambient_Psi=None, Psi_var=None, Psi_LLRatio=None,
                LB_list=elbo_all, LB_doublet=best.ELBO_[-1])


# --------------------------------------------------------------------------
# BinomMixtureVB  (vireoSNP/utils/bmm_model.py:9-263)
# --------------------------------------------------------------------------


@dataclass
class BMMState:
    n_cell: int
    n_var: int
    n_donor: int
    fix_beta_sum: bool = False
    beta_mu_init: Optional[np.ndarray] = None
    beta_sum_init: Optional[np.ndarray] = None
    ID_prob_init: Optional[np.ndarray] = None
    beta_mu: np.ndarray = None
    beta_sum: np.ndarray = None
    ID_prob: np.ndarray = None
    ID_prior: np.ndarray = None
    theta_s1_prior: np.ndarray = None
    theta_s2_prior: np.ndarray = None
    ELBO_iters: np.ndarray = field(default_factory=lambda: np.array([]))
    ELBO_inits: np.ndarray = None

    @property
    def theta_s1(self):
        return self.beta_mu * self.beta_sum

    @property
    def theta_s2(self):
        return (1 - self.beta_mu) * self.beta_sum


def bmm_new(n_cell, n_var, n_donor, fix_beta_sum=False, beta_mu_init=None,
            beta_sum_init=None, ID_prob_init=None):
    """bmm_model.py:24-64: priors first, then the (random) initial state."""
    st = BMMState(n_cell, n_var, n_donor, fix_beta_sum, beta_mu_init,
                  beta_sum_init, ID_prob_init)
    bmm_set_prior(st)
    bmm_set_initial(st, beta_mu_init, beta_sum_init, ID_prob_init)
    return st


def bmm_set_initial(st, beta_mu_init=None, beta_sum_init=None, ID_prob_init=None):
    """bmm_model.py:67-85: mu=0.5, sum=30, ID_prob ~ normalised U(0,1); resets ELBO_iters."""
    st.beta_mu = np.ones((st.n_var, st.n_donor)) * 0.5 if beta_mu_init is None else beta_mu_init
    st.beta_sum = np.ones(st.beta_mu.shape) * 30 if beta_sum_init is None else beta_sum_init
    if ID_prob_init is None:
        st.ID_prob = unit_sum(np.random.rand(st.n_cell, st.n_donor))
    else:
        st.ID_prob = unit_sum(ID_prob_init, axis=1)
    st.ELBO_iters = np.array([])


def bmm_set_prior(st, ID_prior=None, beta_mu_prior=None, beta_sum_prior=None):
    """bmm_model.py:87-106: Beta(1,1) prior on every (variant, clone) theta."""
    if beta_mu_prior is None:
        beta_mu_prior = np.ones((st.n_var, st.n_donor)) * 0.5
    if beta_sum_prior is None:
        beta_sum_prior = np.ones(beta_mu_prior.shape) * 2.0
    st.theta_s1_prior = beta_mu_prior * beta_sum_prior
    st.theta_s2_prior = (1 - beta_mu_prior) * beta_sum_prior
    if ID_prior is None:
        st.ID_prior = unit_sum(np.ones((st.n_cell, st.n_donor)))
    else:
        st.ID_prior = ID_prior[None, :] if ID_prior.ndim == 1 else ID_prior


def bmm_loglik(st, AD, DP):
    """bmm_model.py:118-130."""
    BD = DP - AD
    return (AD.T @ digamma(st.theta_s1) + BD.T @ digamma(st.theta_s2)
            - DP.T @ digamma(st.theta_s1 + st.theta_s2))


def bmm_update_theta(st, AD, DP):
    """bmm_model.py:133-144."""
    BD = DP - AD
    s1 = AD @ st.ID_prob
    s2 = BD @ st.ID_prob
    s1 += st.theta_s1_prior
    s2 += st.theta_s2_prior
    st.beta_mu = s1 / (s1 + s2)
    if not st.fix_beta_sum:
        st.beta_sum = s1 + s2


def bmm_update_id(st, logLik_ID):
    """bmm_model.py:147-154."""
    st.ID_prob = softmax_from_log(logLik_ID + np.log(st.ID_prior))


def bmm_elbo(st, logLik_ID):
    """bmm_model.py:157-175 (only the ``logLik_ID=`` path is live: Q12)."""
    LB_p = np.sum(logLik_ID * st.ID_prob)
    KL_ID = np.sum(entropy(st.ID_prob, st.ID_prior, axis=-1))
    post = np.stack([st.theta_s1, st.theta_s2], axis=1)
    prior = np.stack([st.theta_s1_prior, st.theta_s2_prior], axis=1)
    return LB_p - KL_ID - beta_kl(post, prior)


def bmm_fit_vb(st, AD, DP, max_iter=200, min_iter=20, epsilon_conv=1e-2, verbose=True):
    """bmm_model.py:178-201: same Q1/Q2 quirks as the Vireo loop; appends to ELBO_iters."""
    ELBO = np.zeros(max_iter)
    it = 0
    for it in range(max_iter):
        bmm_update_theta(st, AD, DP)
        ll = bmm_loglik(st, AD, DP)
        bmm_update_id(st, ll)
        ELBO[it] = bmm_elbo(st, ll)
        if it > min_iter:
            if ELBO[it] - ELBO[it - 1] < -1e-6:
                if verbose:
                    print("Warning: ELBO decreases %.8f to %.8f!\n" % (ELBO[it - 1], ELBO[it]))
            elif it == max_iter - 1:
                if verbose:
                    print("Warning: VB did not converge!\n")
            elif ELBO[it] - ELBO[it - 1] < epsilon_conv:
                break
    st.ELBO_iters = np.append(st.ELBO_iters, ELBO[:it])


def bmm_fit(st, AD, DP, n_init=10, max_iter=200, max_iter_pre=100,
            random_seed=None, **kwargs):
    """bmm_model.py:204-263: serial restarts, keep the best, refit it, add the constant."""
    if random_seed is not None:
        np.random.seed(random_seed)
    AD, DP = _maybe_sparsify(AD, DP)
    const = binom_const(AD, DP)
    inits = []
    best = None
    for i in range(n_init):
        bmm_set_initial(st, st.beta_mu_init, st.beta_sum_init, st.ID_prob_init)
        bmm_fit_vb(st, AD, DP, max_iter=max_iter_pre, **kwargs)
        inits.append(st.ELBO_iters[-1])
        if i == 0 or st.ELBO_iters[-1] > np.max(inits[:-1]):
            best = (st.ID_prob + 0, st.beta_mu + 0, st.beta_sum + 0, st.ELBO_iters + 0)
    bmm_set_initial(st, best[1], best[2], best[0])
    st.ELBO_iters = best[3]
    bmm_fit_vb(st, AD, DP, max_iter=max_iter, **kwargs)
    st.ELBO_iters = st.ELBO_iters + const
    st.ELBO_inits = np.array(inits) + const
    return st


# --------------------------------------------------------------------------
# Synthetic workloads -- shared by tests and bench.py
# --------------------------------------------------------------------------


def synth_counts(n_cell, n_var, n_donor, density=0.02, seed=0,
                 theta=(0.01, 0.5, 0.99), gt_freq=(0.45, 0.35, 0.20), dp_p=0.85):
    """Matrix-level donor-pool generator.  Returns (AD, DP, donor, GT) with
    AD, DP csc_matrix (n_var, n_cell) int64, pattern(AD) a subset of pattern(DP)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    total = n_cell * n_var
    n_draw = int(total * density * 1.05) + 1024
    pos = np.cumsum(rng.geometric(density, n_draw).astype(np.int64)) - 1
    pos = pos[pos < total]
    cell = pos // n_var
    snp = (pos % n_var).astype(np.int32)
    donor = rng.integers(0, n_donor, n_cell)
    GT = rng.choice(len(gt_freq), size=(n_var, n_donor), p=np.asarray(gt_freq))
    dp = rng.geometric(dp_p, pos.size).astype(np.int64)
    p = np.asarray(theta)[GT[snp, donor[cell]]]
    ad = rng.binomial(dp, p).astype(np.int64)
    indptr = np.zeros(n_cell + 1, dtype=np.int64)
    np.cumsum(np.bincount(cell, minlength=n_cell), out=indptr[1:])
    idt = np.int32 if pos.size < 2 ** 31 else np.int64
    DP = csc_matrix((dp, snp.copy(), indptr.astype(idt)), shape=(n_var, n_cell))
    keep = ad > 0
    ad_ptr = np.zeros(n_cell + 1, dtype=np.int64)
    np.cumsum(np.bincount(cell[keep], minlength=n_cell), out=ad_ptr[1:])
    AD = csc_matrix((ad[keep], snp[keep].copy(), ad_ptr.astype(idt)), shape=(n_var, n_cell))
    return AD, DP, donor, GT


def synth_clones(n_cell=2000, n_var=300, n_clone=6, presence=0.9, seed=0):
    """Mito clone-mode generator (cfg5): deep counts, per-(variant, clone) AF."""
    rng = np.random.Generator(np.random.PCG64(seed))
    clone = rng.integers(0, n_clone, n_cell)
    lam = np.exp(rng.uniform(np.log(20), np.log(5000), n_var))
    af = rng.beta(0.3, 3.0, size=(n_var, n_clone))
    present = rng.random((n_var, n_cell)) < presence
    dp = (1 + rng.poisson(lam[:, None], size=(n_var, n_cell))) * present
    ad = rng.binomial(dp, af[:, clone])
    return csc_matrix(ad.astype(np.int64)), csc_matrix(dp.astype(np.int64)), clone
