"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (all small .npz files, committed):
  fixture_cellsnp.npz   AD/DP of reference data/cellSNP_mat (952 cells x 3784 SNPs) as CSC arrays
  fixture_mito.npz      AD/DP of reference data/mitoDNA (81 cells x 9 variants; counts up to 85197)
  vireo_*.npz           Vireo runs of the reference (inputs + every output the parity tests compare)
  bmm_*.npz             BinomMixtureVB runs of the reference
  wrap_*.npz            vireo_wrap runs of the reference

The fixtures are data files of the reference repository (Apache-2.0), stored
re-encoded; no reference source code is copied.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
from scipy.io import mmread
from scipy.sparse import csc_matrix

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import vireoSNP  # noqa: E402  (the reference)
from vireoSNP.utils.vireo_doublet import predict_doublet  # noqa: E402
from vireoSNP.utils.vireo_base import get_binom_coeff  # noqa: E402

from oracle.vireo_oracle import synth_counts, synth_clones  # noqa: E402  (generators only)


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def csc_parts(prefix, M):
    M = csc_matrix(M)
    M.sort_indices()
    return {prefix + "_data": M.data.astype(np.int64), prefix + "_indices": M.indices.astype(np.int32),
            prefix + "_indptr": M.indptr.astype(np.int64), prefix + "_shape": np.array(M.shape)}


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        rv = fn(*a, **k)
    return rv, buf.getvalue()


def model_out(m):
    return dict(ID_prob=m.ID_prob, GT_prob=m.GT_prob, beta_mu=m.beta_mu, beta_sum=m.beta_sum, ELBO=m.ELBO_)


def vireo_case(name, AD, DP, K, fit_kw, ctor_kw=None, GT_prior=None, seed=1, store_mats=False, two_fits=False):
    """Construct a reference Vireo under np.random.seed(seed), record its inits, fit, record outputs."""
    ctor_kw = dict(ctor_kw or {})
    np.random.seed(seed)
    m = vireoSNP.Vireo(n_var=AD.shape[0], n_cell=AD.shape[1], n_donor=K, **ctor_kw)
    if GT_prior is not None:
        m.set_prior(GT_prior=GT_prior.copy())
    init = dict(ID_prob_init=m.ID_prob.copy(), GT_prob_init=m.GT_prob.copy(),
                beta_mu_init=m.beta_mu.copy(), beta_sum_init=m.beta_sum.copy(),
                GT_prior=m.GT_prior.copy(), ID_prior=m.ID_prior.copy())
    _, out = quiet(m.fit, AD, DP, **fit_kw)
    extra = {}
    if two_fits:   # warm start: ELBO_ accumulates (Q8)
        extra["ELBO_first_len"] = np.array(len(m.ELBO_))
        quiet(m.fit, AD, DP, min_iter=5, verbose=False)
    mats = {}
    if store_mats:
        mats.update(csc_parts("AD", AD))
        mats.update(csc_parts("DP", DP))
    save(name, K=np.array(K), seed=np.array(seed), stdout=np.array(out),
         fit_kw=np.array(repr(fit_kw)), ctor_kw=np.array(repr({k: v for k, v in ctor_kw.items() if np.isscalar(v)})),
         **init, **model_out(m), **mats, **extra)
    return m


def main():
    # ---------------- fixtures
    AD = mmread(REF + "/data/cellSNP_mat/cellSNP.tag.AD.mtx").tocsc()
    DP = mmread(REF + "/data/cellSNP_mat/cellSNP.tag.DP.mtx").tocsc()
    save("fixture_cellsnp", **csc_parts("AD", AD), **csc_parts("DP", DP))
    mAD = mmread(REF + "/data/mitoDNA/cellSNP.tag.AD.mtx").tocsc()
    mDP = mmread(REF + "/data/mitoDNA/cellSNP.tag.DP.mtx").tocsc()
    save("fixture_mito", **csc_parts("AD", mAD), **csc_parts("DP", mDP))

    # ---------------- binomial constant (Q3)
    save("binom_const",
         cellsnp=np.array(np.sum(get_binom_coeff(AD, DP))), mito=np.array(np.sum(get_binom_coeff(mAD, mDP))),
         mito_terms=np.asarray(get_binom_coeff(mAD, mDP)).reshape(-1))

    # ---------------- cfg1: the SURVEY known answer (19 entries, -41723.09107408444)
    kw20 = dict(max_iter=20, min_iter=5, delay_fit_theta=3)
    m = vireo_case("vireo_cfg1_fit20", AD, DP, 4, kw20)
    assert len(m.ELBO_) == 19 and abs(m.ELBO_[-1] - (-41723.09107408444)) < 1e-6, m.ELBO_[-1]

    # fixed-length loop (min_iter = max_iter): no early break
    vireo_case("vireo_cfg1_fixed20", AD, DP, 4, dict(max_iter=20, min_iter=20, delay_fit_theta=3, verbose=False))
    # long run to convergence + warm-start second fit (Q8)
    vireo_case("vireo_cfg1_converge", AD, DP, 4, dict(max_iter=200, min_iter=5, verbose=False), seed=2, two_fits=True)

    # ---------------- teacher-forced single updates from a mid-trajectory state
    np.random.seed(3)
    t = vireoSNP.Vireo(n_var=AD.shape[0], n_cell=AD.shape[1], n_donor=4)
    quiet(t.fit, AD, DP, max_iter=4, min_iter=4, verbose=False)
    st0 = dict(ID_prob0=t.ID_prob.copy(), GT_prob0=t.GT_prob.copy(), beta_mu0=t.beta_mu.copy(),
               beta_sum0=t.beta_sum.copy())
    t.update_theta_size(AD, DP)
    st1 = dict(beta_mu1=t.beta_mu.copy(), beta_sum1=t.beta_sum.copy())
    t.update_GT_prob(AD, DP)
    st2 = dict(GT_prob2=t.GT_prob.copy())
    ll = t.update_ID_prob(AD, DP)
    st3 = dict(logLik_ID3=ll.copy(), ID_prob3=t.ID_prob.copy(), ELBO3=np.array(t.get_ELBO(ll)),
               ELBO3_none=np.array(t.get_ELBO(None, AD, DP)))
    save("vireo_cfg1_single_updates", **st0, **st1, **st2, **st3)

    # ---------------- option coverage on a small synthetic (fast on CPU; stored with matrices)
    sAD, sDP, donor, GT = synth_counts(300, 400, 3, density=0.05, seed=11)
    base = dict(max_iter=25, min_iter=5, delay_fit_theta=2, verbose=False)
    vireo_case("vireo_small_default", sAD, sDP, 3, base, store_mats=True)
    vireo_case("vireo_small_ase", sAD, sDP, 3, base, ctor_kw=dict(ASE_mode=True))
    vireo_case("vireo_small_fixsum", sAD, sDP, 3, base, ctor_kw=dict(fix_beta_sum=True))
    vireo_case("vireo_small_notheta", sAD, sDP, 3, base, ctor_kw=dict(learn_theta=False))
    vireo_case("vireo_small_k7", sAD, sDP, 7, base)
    # GT-given mode (cfg4 style): 0.98 on the truth, 0.01 elsewhere, learn_GT False
    prior = np.full((400, 3, 3), 0.01)
    np.put_along_axis(prior, GT[:, :, None], 0.98, axis=2)
    vireo_case("vireo_small_gtgiven", sAD, sDP, 3, dict(max_iter=25, min_iter=5, verbose=False),
               ctor_kw=dict(learn_GT=False, GT_prob_init=prior.copy()), GT_prior=prior)
    # GT prior with learn_GT True (CLI mode 3 / --forceLearnGT); includes exact 0/1 entries -> clipping
    hard = np.zeros((400, 3, 3))
    np.put_along_axis(hard, GT[:, :, None], 1.0, axis=2)
    vireo_case("vireo_small_gtprior_learn", sAD, sDP, 3, base,
               ctor_kw=dict(GT_prob_init=hard.copy()), GT_prior=hard.copy())
    # n_GT = 2 (generality of the genotype axis)
    vireo_case("vireo_small_g2", sAD, sDP, 3, base, ctor_kw=dict(n_GT=2))

    # edge cases: empty cell, empty SNP, a wide count (> 65535), explicit zero in DP, a dp>0/ad=0 and ad=dp
    eAD, eDP, _, _ = synth_counts(60, 80, 2, density=0.2, seed=5)
    eAD = eAD.tolil(); eDP = eDP.tolil()
    eAD[:, 7] = 0; eDP[:, 7] = 0          # empty cell
    eAD[13, :] = 0; eDP[13, :] = 0        # empty SNP
    eDP[3, 2] = 70000; eAD[3, 2] = 31234  # wide counts
    eDP[4, 2] = 900; eAD[4, 2] = 450      # binom cap (log C(900,450) > 700? no: ~620) keep as large case
    eDP[5, 2] = 1200; eAD[5, 2] = 600     # log C(1200,600) = 828 > 700 -> capped
    eAD = csc_matrix(eAD); eDP = csc_matrix(eDP)
    eAD.eliminate_zeros(); eDP.eliminate_zeros()
    vireo_case("vireo_edge", eAD, eDP, 2, dict(max_iter=15, min_iter=5, verbose=False), store_mats=True)
    save("binom_const_edge", value=np.array(np.sum(get_binom_coeff(eAD, eDP))))

    # ---------------- predict_doublet on the fitted cfg1 model
    np.random.seed(1)
    d = vireoSNP.Vireo(n_var=AD.shape[0], n_cell=AD.shape[1], n_donor=4)
    quiet(d.fit, AD, DP, **kw20)
    pre = dict(ID_prob_in=d.ID_prob.copy(), GT_prob_in=d.GT_prob.copy(), beta_mu_in=d.beta_mu.copy(),
               beta_sum_in=d.beta_sum.copy())
    dp_, sp_, llr = predict_doublet(d, AD, DP)
    save("doublet_cfg1", **pre, doublet_prob=dp_, singlet_prob=sp_, LLR=llr, GT_prob_out=d.GT_prob.copy())

    # ---------------- vireo_wrap (Q4, Q5, Q8, Q10)
    for name, kw in [
        ("wrap_cfg1_n3", dict(n_donor=4, n_init=3, random_seed=1, nproc=1)),
        ("wrap_cfg1_n1", dict(n_donor=4, n_init=1, random_seed=1, nproc=1)),
        ("wrap_cfg1_nodoublet_ase", dict(n_donor=4, n_init=2, random_seed=4, nproc=1, check_doublet=False,
                                         ASE_mode=True)),
        ("wrap_cfg1_extra", dict(n_donor=4, n_init=3, random_seed=5, nproc=1, n_extra_donor=2)),
    ]:
        rv, out = quiet(vireoSNP.vireo_wrap, AD, DP, **kw)
        keep = {k: v for k, v in rv.items() if v is not None}
        keep["LB_doublet"] = np.array(keep["LB_doublet"])
        save(name, stdout=np.array(out), kw=np.array(repr(kw)), **keep)
    assert abs(np.load(os.path.join(HERE, "wrap_cfg1_n1.npz"))["LB_doublet"] - (-41722.95461102484)) < 1e-6

    # wrap with a GT prior on the small synthetic: given-GT (n_init forced to 1: Q11) and prior+learn (Q5)
    for name, kw in [
        ("wrap_small_gtgiven", dict(GT_prior=prior.copy(), learn_GT=False, n_init=5, random_seed=7, nproc=1)),
        ("wrap_small_gtprior_learn", dict(GT_prior=hard.copy(), learn_GT=True, n_init=3, random_seed=7, nproc=1)),
        ("wrap_small_fewer_donors", dict(GT_prior=prior.copy(), n_donor=2, learn_GT=False, random_seed=7, nproc=1)),
        ("wrap_small_more_donors", dict(GT_prior=prior[:, :2, :].copy(), n_donor=3, learn_GT=True, n_init=3,
                                        random_seed=7, nproc=1)),
    ]:
        rv, out = quiet(vireoSNP.vireo_wrap, sAD, sDP, **kw)
        keep = {k: v for k, v in rv.items() if v is not None}
        keep["LB_doublet"] = np.array(keep["LB_doublet"])
        save(name, stdout=np.array(out), **keep)
    save("small_priors", soft=prior, hard=hard, GT_true=GT, donor_true=donor)

    # ---------------- BinomMixtureVB
    # the notebook known answer (examples/vireoSNP_clones.ipynb:103): unseeded there, reproduced for any seed
    b = vireoSNP.BinomMixtureVB(n_var=mAD.shape[0], n_cell=mAD.shape[1], n_donor=3)
    quiet(b.fit, mAD, mDP, min_iter=30, n_init=50, random_seed=0)
    assert repr(float(b.ELBO_iters[-1])) == "-190779.74335041404", repr(b.ELBO_iters[-1])
    save("bmm_mito_n50", ELBO_iters=b.ELBO_iters, ELBO_inits=b.ELBO_inits, ID_prob=b.ID_prob,
         beta_mu=b.beta_mu, beta_sum=b.beta_sum)

    # a single restart with recorded init, teacher-forced friendly
    np.random.seed(9)
    b1 = vireoSNP.BinomMixtureVB(n_var=mAD.shape[0], n_cell=mAD.shape[1], n_donor=3)
    init = b1.ID_prob.copy()
    quiet(b1._fit_BV, mAD, mDP, max_iter=100, min_iter=30)
    save("bmm_mito_single", ID_prob_init=init, ELBO_iters=b1.ELBO_iters, ID_prob=b1.ID_prob,
         beta_mu=b1.beta_mu, beta_sum=b1.beta_sum)

    cAD, cDP, clone = synth_clones(200, 40, 4, seed=3)
    b2 = vireoSNP.BinomMixtureVB(n_var=cAD.shape[0], n_cell=cAD.shape[1], n_donor=4)
    quiet(b2.fit, cAD, cDP, min_iter=20, n_init=6, random_seed=2)
    save("bmm_small_n6", **csc_parts("AD", cAD), **csc_parts("DP", cDP), ELBO_iters=b2.ELBO_iters,
         ELBO_inits=b2.ELBO_inits, ID_prob=b2.ID_prob, beta_mu=b2.beta_mu, beta_sum=b2.beta_sum, clone=clone)
    b3 = vireoSNP.BinomMixtureVB(n_var=cAD.shape[0], n_cell=cAD.shape[1], n_donor=4, fix_beta_sum=True)
    quiet(b3.fit, cAD, cDP, min_iter=20, n_init=2, random_seed=2)
    save("bmm_small_fixsum", ELBO_iters=b3.ELBO_iters, ELBO_inits=b3.ELBO_inits, ID_prob=b3.ID_prob,
         beta_mu=b3.beta_mu, beta_sum=b3.beta_sum)


if __name__ == "__main__":
    main()
