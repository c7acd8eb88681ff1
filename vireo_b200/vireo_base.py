"""Numeric helpers with the names ``vireoSNP.utils.vireo_base`` exports (reference
vireoSNP/utils/vireo_base.py).  ``normalize`` / ``loglik_amplify`` are host conveniences (inside the
EM loop they are fused into the kernels); ``get_binom_coeff`` runs on the device; ``optimal_match`` and
``donor_select`` are the small host-side glue ``vireo_wrap`` needs between fits.
"""
import numpy as np
from scipy.optimize import linear_sum_assignment


def normalize(X, axis=-1):
    """Scale X to sum to one along ``axis`` (reference vireo_base.py:44-56)."""
    X = np.asarray(X)
    n = X.shape[axis] if X.ndim else 0
    if X.ndim >= 2 and axis in (-1, X.ndim - 1) and 1 <= n <= 4 and X.dtype == np.float64:
        # a short last axis (the genotype axis): numpy reduces fewer than 8 elements left to right, so summing
        # the slices in that order is bit-identical and avoids the slow strided reduction
        s = X[..., 0].copy()
        for i in range(1, n):
            s += X[..., i]
        return X / s[..., np.newaxis]
    return X / np.sum(X, axis=axis, keepdims=True)


def tensor_normalize(X, axis=1):
    return normalize(X, axis)


def loglik_amplify(X, axis=-1):
    """Shift log-likelihoods so the max along ``axis`` is zero (reference vireo_base.py:62-74)."""
    return X - np.max(X, axis=axis, keepdims=True)


def get_binom_coeff(AD, DP, max_val=700, is_log=True):
    """Sum-ready binomial constant of the ELBO, computed on the device.

    The reference (vireo_base.py:7-22) returns the per-entry float32 array ``log C(DP, AD)`` capped
    at ``max_val`` and every caller immediately sums it (vireo_model.py:313, bmm_model.py:239).  Here
    the sum is what the device produces, returned as a one-element float32 array so that
    ``np.sum(get_binom_coeff(AD, DP))`` keeps working.
    """
    if max_val != 700 or not is_log:
        raise NotImplementedError("only the reference defaults (max_val=700, is_log=True) are on the device path")
    from . import _engine
    return np.array([_engine.stage(AD, DP).binom_const()], dtype=np.float32)


def beta_entropy(X, X_prior=None, axis=None):
    """KL(Beta(X) || Beta(X_prior)) summed over rows, or the entropy when no prior is given
    (reference vireo_base.py:77-127).  X: (N, 2) or (N, 2, G) shape arrays.  Host helper -- the EM loop
    evaluates the same expression inside k_theta / k_bmm_theta."""
    from scipy.special import betaln, digamma

    def cross(p, q):
        return (betaln(q[:, 0], q[:, 1]) - (q[:, 0] - 1) * digamma(p[:, 0]) - (q[:, 1] - 1) * digamma(p[:, 1])
                + (q.sum(axis=1) - 2) * digamma(p.sum(axis=1)))

    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(-1, 2)
    if X_prior is None:
        return np.sum(cross(X, X), axis=axis)
    X_prior = np.asarray(X_prior)
    if X_prior.ndim == 1:
        X_prior = X_prior.reshape(-1, 2)
    return np.sum(cross(X, X_prior) - cross(X, X), axis=axis)


def match(ref_ids, new_ids, uniq_ref_only=True):
    """For every entry of ``ref_ids`` the index of the equal entry in ``new_ids`` (None when absent);
    same contract as reference vireo_base.py:130-184."""
    lookup = {}
    for j, v in enumerate(new_ids):
        lookup.setdefault(v, j)
    used, out = set(), []
    for v in ref_ids:
        j = lookup.get(v)
        if j is not None and uniq_ref_only:
            if j in used:
                j = None
            else:
                used.add(j)
        out.append(j)
    return np.array(out, dtype=object) if None in out else np.array(out)


def optimal_match(X, Z, axis=1, return_delta=False):
    """Hungarian alignment of the slices of Z to those of X along ``axis`` by mean absolute
    difference (reference vireo_base.py:187-206)."""
    cost = np.zeros((X.shape[axis], Z.shape[axis]))
    for i in range(X.shape[axis]):
        xi = np.take(X, i, axis=axis)
        for j in range(Z.shape[axis]):
            cost[i, j] = np.mean(np.abs(xi - np.take(Z, j, axis=axis)))
    idx0, idx1 = linear_sum_assignment(cost)
    return (idx0, idx1, cost) if return_delta else (idx0, idx1)


def donor_select(GT_prob, ID_prob, n_donor, mode="distance"):
    """Pick ``n_donor`` of the fitted (n_donor + extra) donors: the largest ones (mode="size") or a
    greedy max-min genotype-distance set seeded by the largest (reference vireo_base.py:217-254).
    Prints the same summary lines; returns the kept ID_prob columns floored at 1e-10."""
    size = np.sum(ID_prob, axis=0)
    K = GT_prob.shape[1]
    if mode == "size":
        chosen = list(np.argsort(size)[::-1])
    else:
        dist = np.abs(GT_prob[:, :, None, :] - GT_prob[:, None, :, :]).mean(axis=(0, 3))
        chosen = [int(np.argmax(size))]
        rest = [k for k in range(K) if k != chosen[0]]
        while rest:
            gain = dist[np.ix_(chosen, rest)].min(axis=0)
            chosen.append(rest.pop(int(np.argmax(gain))))
    print("[vireo] donor size with searching extra %d donors:" % (K - n_donor))
    print("\t".join(["donor%d" % x for x in chosen]))
    print("\t".join(["%.0f" % size[x] for x in chosen]))
    kept = ID_prob[:, chosen[:n_donor]]
    kept[kept < 10 ** -10] = 10 ** -10
    return kept
