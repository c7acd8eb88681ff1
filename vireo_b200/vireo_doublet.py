"""Doublet prediction with a fitted ``Vireo`` -- drop-in for ``vireoSNP.utils.vireo_doublet.predict_doublet``
(reference vireoSNP/utils/vireo_doublet.py:11-82).  The donor-pair genotype tables
(``add_doublet_GT`` :105-136, ``add_doublet_theta`` :85-102) are built on the device straight into the
per-(SNP, column) tables of the cell-major pass; the (n_var, K + K(K-1)/2, 6) tensor the reference
materialises is never formed.
"""
import numpy as np

from . import _engine


def doublet_log_prior(vobj, n_cell, doublet_rate_prior=None):
    """log of the prior over singlet + donor-pair columns (reference vireo_doublet.py:42-48): singlet prior scaled by
    (1 - rate), the doublet mass spread evenly over the pairs; rows = rows of ``vobj.ID_prior`` (1 when uniform)."""
    K = int(vobj.n_donor)
    n_pair = K * (K - 1) // 2
    if doublet_rate_prior is None:
        doublet_rate_prior = min(0.5, n_cell / 100000)
    id_prior = np.asarray(vobj.ID_prior, dtype=np.float64)
    id_prior = _engine._compress_rows(id_prior)
    if n_pair:
        pair_prior = np.ones((id_prior.shape[0], n_pair)) / n_pair * doublet_rate_prior
    else:
        pair_prior = np.zeros((id_prior.shape[0], 0))
    prior_both = np.append(id_prior * (1 - doublet_rate_prior), pair_prior, axis=1)
    with np.errstate(divide="ignore"):
        return np.log(prior_both)


def predict_doublet(vobj, AD, DP, update_GT=True, update_ID=True, doublet_rate_prior=None):
    """Returns (doublet_prob, singlet_prob, logLik_ratio) and, like the reference, overwrites
    ``vobj.ID_prob`` with the singlet columns and refreshes ``vobj.GT_prob``.

    doublet_prob : (n_cell, n_donor (n_donor - 1) / 2); singlet_prob : (n_cell, n_donor);
    logLik_ratio : (n_cell,) best doublet minus best singlet log-likelihood.
    """
    counts = _engine.stage(AD, DP)
    K = int(vobj.n_donor)
    log_prior_both = doublet_log_prior(vobj, counts.n_cell, doublet_rate_prior)

    loglik, prob_both, llr = _engine.doublet_pass(counts, np.asarray(vobj.GT_prob, dtype=np.float64),
                                                  vobj.beta_mu, vobj.beta_sum, log_prior_both, vobj.ASE_mode)
    if update_ID:
        vobj.ID_prob = prob_both[:, :K].copy()
    if update_GT:
        if update_ID:
            vobj.update_GT_prob(counts, None)
        else:
            print("For update_GT, please turn on update_ID.")
    return prob_both[:, K:], prob_both[:, :K], llr
