"""Pin the CPU oracle (oracle/vireo_oracle.py) against vectors produced by the real reference.

Every golden file written by tests/golden/make_golden.py is consumed here.  The oracle keeps the
reference's operation order, so agreement is expected to ~1e-12; the gates are 1e-9 (far inside the
1e-5 / 1e-6 product gates).
"""
import ast
import contextlib
import io

import numpy as np
import pytest

from conftest import csc_from, load_golden, rel_close
from oracle import vireo_oracle as O

TIGHT = 1e-9


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        rv = fn(*a, **k)
    return rv, buf.getvalue()


def _run_vireo_case(z, AD, DP):
    K = int(z["K"])
    fit_kw = ast.literal_eval(str(z["fit_kw"]))
    ctor_kw = ast.literal_eval(str(z["ctor_kw"]))
    st = O.vireo_new(AD.shape[1], AD.shape[0], K, ID_prob_init=z["ID_prob_init"].copy(),
                     GT_prob_init=z["GT_prob_init"].copy(), beta_mu_init=z["beta_mu_init"].copy(),
                     beta_sum_init=z["beta_sum_init"].copy(), **ctor_kw)
    # bypass the constructor's re-normalisation so the start state is bit-identical to the reference's
    st.ID_prob = z["ID_prob_init"].copy()
    st.GT_prob = z["GT_prob_init"].copy()
    st.GT_prior = z["GT_prior"].copy()
    st.ID_prior = z["ID_prior"].copy()
    _, out = _quiet(O.vireo_fit, st, AD, DP, **fit_kw)
    return st, out


def _check_state(st, z):
    assert len(st.ELBO_) == len(z["ELBO"])
    rel_close(st.ELBO_, z["ELBO"], 1e-10, "ELBO")
    rel_close(st.ID_prob, z["ID_prob"], TIGHT, "ID_prob")
    rel_close(st.GT_prob, z["GT_prob"], TIGHT, "GT_prob")
    rel_close(st.beta_mu, z["beta_mu"], TIGHT, "beta_mu")
    rel_close(st.beta_sum, z["beta_sum"], TIGHT, "beta_sum")
    assert np.array_equal(st.ID_prob.argmax(1), z["ID_prob"].argmax(1))


def test_rng_order_and_known_answer(cellsnp):
    """np.random.seed(1) + constructor reproduces the reference's inits (Q4) and the SURVEY known answer."""
    AD, DP = cellsnp
    z = load_golden("vireo_cfg1_fit20")
    np.random.seed(1)
    st = O.vireo_new(AD.shape[1], AD.shape[0], 4)
    assert np.array_equal(st.ID_prob, z["ID_prob_init"])
    assert np.array_equal(st.GT_prob, z["GT_prob_init"])
    _, out = _quiet(O.vireo_fit, st, AD, DP, max_iter=20, min_iter=5, delay_fit_theta=3)
    assert out == str(z["stdout"])
    assert len(st.ELBO_) == 19
    assert abs(st.ELBO_[-1] - (-41723.09107408444)) < 1e-7
    assert list(np.bincount(st.ID_prob.argmax(1))) == [251, 233, 235, 233]
    _check_state(st, z)


@pytest.mark.parametrize("name", ["vireo_cfg1_fixed20"])
def test_cfg1_runs(cellsnp, name):
    z = load_golden(name)
    st, _ = _run_vireo_case(z, *cellsnp)
    _check_state(st, z)


def test_cfg1_converge_and_warm_start(cellsnp):
    AD, DP = cellsnp
    z = load_golden("vireo_cfg1_converge")
    st, _ = _run_vireo_case(z, AD, DP)
    assert len(st.ELBO_) == int(z["ELBO_first_len"])
    O.vireo_fit(st, AD, DP, min_iter=5, verbose=False)   # ELBO_ accumulates (Q8)
    _check_state(st, z)


@pytest.mark.parametrize("name", ["vireo_small_default", "vireo_small_ase", "vireo_small_fixsum",
                                  "vireo_small_notheta", "vireo_small_k7", "vireo_small_gtgiven",
                                  "vireo_small_gtprior_learn", "vireo_small_g2"])
def test_small_options(small, name):
    z = load_golden(name)
    st, _ = _run_vireo_case(z, *small)
    _check_state(st, z)


def test_edge_cases(edge):
    z = load_golden("vireo_edge")
    st, _ = _run_vireo_case(z, *edge)
    _check_state(st, z)
    assert float(O.binom_const(*edge)) == float(load_golden("binom_const_edge")["value"])


def test_single_updates(cellsnp):
    AD, DP = cellsnp
    z = load_golden("vireo_cfg1_single_updates")
    st = O.vireo_new(AD.shape[1], AD.shape[0], 4, ID_prob_init=z["ID_prob0"].copy(),
                     GT_prob_init=z["GT_prob0"].copy(), beta_mu_init=z["beta_mu0"].copy(),
                     beta_sum_init=z["beta_sum0"].copy())
    st.ID_prob, st.GT_prob = z["ID_prob0"].copy(), z["GT_prob0"].copy()
    O.vireo_update_theta(st, AD, DP)
    rel_close(st.beta_mu, z["beta_mu1"], 1e-13, "beta_mu")
    rel_close(st.beta_sum, z["beta_sum1"], 1e-13, "beta_sum")
    O.vireo_update_gt(st, AD, DP)
    rel_close(st.GT_prob, z["GT_prob2"], 1e-10, "GT_prob")
    ll = O.vireo_update_id(st, AD, DP)
    rel_close(ll, z["logLik_ID3"], 1e-13, "logLik_ID")
    rel_close(st.ID_prob, z["ID_prob3"], 1e-10, "ID_prob")
    assert abs(O.vireo_elbo(st, ll) - float(z["ELBO3"])) < 1e-8
    assert abs(O.vireo_elbo(st, None, AD, DP) - float(z["ELBO3_none"])) < 1e-8


def test_binom_const(cellsnp, mito):
    z = load_golden("binom_const")
    assert float(O.binom_const(*cellsnp)) == float(z["cellsnp"])
    assert float(O.binom_const(*mito)) == float(z["mito"])
    assert np.array_equal(np.asarray(O.binom_log_terms(*mito)).reshape(-1), z["mito_terms"])


def test_predict_doublet(cellsnp):
    AD, DP = cellsnp
    z = load_golden("doublet_cfg1")
    st = O.vireo_new(AD.shape[1], AD.shape[0], 4, ID_prob_init=z["ID_prob_in"].copy(),
                     GT_prob_init=z["GT_prob_in"].copy(), beta_mu_init=z["beta_mu_in"].copy(),
                     beta_sum_init=z["beta_sum_in"].copy())
    st.ID_prob, st.GT_prob = z["ID_prob_in"].copy(), z["GT_prob_in"].copy()
    dbl, sgl, llr = O.vireo_predict_doublet(st, AD, DP)
    rel_close(dbl, z["doublet_prob"], TIGHT, "doublet_prob")
    rel_close(sgl, z["singlet_prob"], TIGHT, "singlet_prob")
    rel_close(llr, z["LLR"], TIGHT, "LLR")
    rel_close(st.GT_prob, z["GT_prob_out"], TIGHT, "GT_prob")


def _check_wrap(rv, z):
    rel_close(rv["LB_list"], z["LB_list"], 1e-12, "LB_list")
    assert abs(rv["LB_doublet"] - float(z["LB_doublet"])) <= 1e-12 * abs(float(z["LB_doublet"]))
    for key in ("ID_prob", "GT_prob", "doublet_prob", "doublet_LLR", "theta_shapes", "theta_mean", "theta_sum"):
        rel_close(rv[key], z[key], 1e-8, key)


@pytest.mark.parametrize("name", ["wrap_cfg1_n3", "wrap_cfg1_n1", "wrap_cfg1_nodoublet_ase", "wrap_cfg1_extra"])
def test_wrap_cfg1(cellsnp, name):
    z = load_golden(name)
    kw = ast.literal_eval(str(z["kw"]))
    kw.pop("nproc")
    rv, out = _quiet(O.vireo_wrap, *cellsnp, **kw)
    _check_wrap(rv, z)
    if name == "wrap_cfg1_n1":
        assert abs(rv["LB_doublet"] - (-41722.95461102484)) < 1e-7


def test_wrap_with_gt_prior(small):
    pri = load_golden("small_priors")
    cases = {
        "wrap_small_gtgiven": dict(GT_prior=pri["soft"].copy(), learn_GT=False, n_init=5, random_seed=7),
        "wrap_small_gtprior_learn": dict(GT_prior=pri["hard"].copy(), learn_GT=True, n_init=3, random_seed=7),
        "wrap_small_fewer_donors": dict(GT_prior=pri["soft"].copy(), n_donor=2, learn_GT=False, random_seed=7),
        "wrap_small_more_donors": dict(GT_prior=pri["soft"][:, :2, :].copy(), n_donor=3, learn_GT=True, n_init=3,
                                       random_seed=7),
    }
    for name, kw in cases.items():
        rv, _ = _quiet(O.vireo_wrap, *small, **kw)
        _check_wrap(rv, load_golden(name))


def test_bmm_notebook_known_answer(mito):
    """examples/vireoSNP_clones.ipynb:103 prints -190779.74335041404 (any seed)."""
    AD, DP = mito
    z = load_golden("bmm_mito_n50")
    st = O.bmm_new(AD.shape[1], AD.shape[0], 3)
    _quiet(O.bmm_fit, st, AD, DP, min_iter=30, n_init=50, random_seed=0)
    assert repr(float(st.ELBO_iters[-1])) == "-190779.74335041404"
    assert len(st.ELBO_iters) == len(z["ELBO_iters"]) == 62
    rel_close(st.ELBO_iters, z["ELBO_iters"], 1e-13, "ELBO_iters")
    rel_close(st.ELBO_inits, z["ELBO_inits"], 1e-13, "ELBO_inits")
    rel_close(st.ID_prob, z["ID_prob"], TIGHT, "ID_prob")
    rel_close(st.beta_mu, z["beta_mu"], TIGHT, "beta_mu")
    rel_close(st.beta_sum, z["beta_sum"], TIGHT, "beta_sum")


def test_bmm_single_restart(mito):
    AD, DP = mito
    z = load_golden("bmm_mito_single")
    st = O.bmm_new(AD.shape[1], AD.shape[0], 3, ID_prob_init=z["ID_prob_init"].copy())
    _quiet(O.bmm_fit_vb, st, AD, DP, max_iter=100, min_iter=30)
    rel_close(st.ELBO_iters, z["ELBO_iters"], 1e-13, "ELBO_iters")
    rel_close(st.ID_prob, z["ID_prob"], TIGHT, "ID_prob")
    rel_close(st.beta_mu, z["beta_mu"], TIGHT, "beta_mu")


def test_bmm_small():
    z = load_golden("bmm_small_n6")
    AD, DP = csc_from(z, "AD"), csc_from(z, "DP")
    st = O.bmm_new(AD.shape[1], AD.shape[0], 4)
    _quiet(O.bmm_fit, st, AD, DP, min_iter=20, n_init=6, random_seed=2)
    rel_close(st.ELBO_iters, z["ELBO_iters"], 1e-12, "ELBO_iters")
    rel_close(st.ELBO_inits, z["ELBO_inits"], 1e-12, "ELBO_inits")
    rel_close(st.ID_prob, z["ID_prob"], TIGHT, "ID_prob")
    zf = load_golden("bmm_small_fixsum")
    sf = O.bmm_new(AD.shape[1], AD.shape[0], 4, fix_beta_sum=True)
    _quiet(O.bmm_fit, sf, AD, DP, min_iter=20, n_init=2, random_seed=2)
    rel_close(sf.ELBO_iters, zf["ELBO_iters"], 1e-12, "ELBO_iters")
    rel_close(sf.beta_sum, zf["beta_sum"], TIGHT, "beta_sum")


def test_synth_generator_is_deterministic():
    a1, d1, donor1, gt1 = O.synth_counts(200, 150, 3, density=0.05, seed=4)
    a2, d2, donor2, gt2 = O.synth_counts(200, 150, 3, density=0.05, seed=4)
    assert (a1 != a2).nnz == 0 and (d1 != d2).nnz == 0
    assert np.array_equal(donor1, donor2) and np.array_equal(gt1, gt2)
    # pattern(AD) is a subset of pattern(DP), AD <= DP, no explicit zeros
    assert ((d1 - a1).data >= 0).all() and (a1.data > 0).all() and (d1.data > 0).all()
    assert (a1 > d1).nnz == 0 and a1.nnz <= d1.nnz


def test_host_normalize_matches_numpy_bit_for_bit():
    """vireo_b200.normalize takes a shortcut on a short last axis (the genotype axis); the inits it produces must
    be the very arrays the reference's normalize (vireo_base.py:44-56) produces from the same random draws."""
    from vireo_b200.vireo_base import normalize
    rng = np.random.RandomState(0)
    for shape in [(7, 3), (50, 16, 3), (9, 16), (4, 2), (3, 1), (6, 5), (5, 4, 4)]:
        X = rng.rand(*shape)
        assert np.array_equal(normalize(X), X / np.sum(X, axis=-1, keepdims=True)), shape
    X = rng.rand(6, 5, 3)
    assert np.array_equal(normalize(X, axis=1), X / np.sum(X, axis=1, keepdims=True))
    import vireo_b200 as vb
    np.random.seed(3)
    m = vb.Vireo(n_cell=11, n_var=13, n_donor=4)
    np.random.seed(3)
    idp = np.random.rand(11, 4)
    gtp = np.random.rand(13, 4, 3)
    assert np.array_equal(m.ID_prob, idp / idp.sum(1, keepdims=True))
    assert np.array_equal(m.GT_prob, gtp / gtp.sum(2, keepdims=True))
    assert np.array_equal(m.ID_prior, np.ones((11, 4)) / np.sum(np.ones((11, 4)), axis=-1, keepdims=True))
    assert np.array_equal(m.GT_prior, np.ones((13, 4, 3)) / np.sum(np.ones((13, 4, 3)), axis=-1, keepdims=True))
