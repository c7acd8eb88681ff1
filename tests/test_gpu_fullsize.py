"""The BASELINE configurations at their FULL sizes, under the production ("auto") kernel selector.

SURVEY 8d asks for same-inputs parity at full size ("cfg3 parity can use T=3 against the 70 s CPU run"): the CUDA
path and the pinned CPU oracle run freely from the same initial state on the same synthetic matrices and must agree
to the north-star gates -- ID_prob / GT_prob 1e-5 relative, every ELBO entry 1e-6 relative, identical argmax donor
per cell.  Nothing here forces a kernel family: these tests exercise exactly what bench.py and a user get.
"""
import contextlib
import io

import numpy as np
from scipy.sparse import csc_matrix
import pytest

from conftest import rel_close
from oracle import vireo_oracle as O

pytestmark = pytest.mark.gpu

P_TOL = 1e-5
E_TOL = 1e-6


@pytest.fixture(scope="module")
def vb():
    import vireo_b200
    from vireo_b200 import _lib
    _lib.set_path("auto")
    return vireo_b200


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        return fn(*a, **k)


def _family(counts, fmt):
    from vireo_b200 import _lib
    return int(_lib.load().vb_counts_info(counts.handle, 20 + 10 * fmt))


@pytest.fixture(scope="module")
def cfg3():
    return O.synth_counts(100000, 50000, 16, seed=0)


def test_cfg3_full_size_free_running_vs_oracle(vb, cfg3):
    """BASELINE cfg3: 100k cells x 50k SNPs x 16 donors (nnz 1e8), learn GT, T = 3 free-running iterations from the
    reference's random start, against the oracle on the same inputs (about 30 s of CPU)."""
    AD, DP, _, _ = cfg3
    C, V, K = 100000, 50000, 16
    counts = vb.stage(AD, DP)
    np.random.seed(1)
    m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
    o = O.vireo_new(C, V, K, ID_prob_init=m.ID_prob.copy(), GT_prob_init=m.GT_prob.copy())
    o.ID_prob, o.GT_prob = m.ID_prob.copy(), m.GT_prob.copy()
    kw = dict(max_iter=3, min_iter=3, delay_fit_theta=1, verbose=False)       # theta is updated in 2 of the 3
    # _fit_VB: the loop without the binomial constant (the oracle's scipy restatement of get_binom_coeff alone takes
    # about a minute at 1e8 nnz; the constant is checked against the reference at fixture size in test_binom_const)
    elbo = _quiet(m._fit_VB, counts, None, **kw)
    assert _family(counts, 0) == 1 and int(_engine_info(vb, counts, 60)) == 0   # the window-segment kernels served it
    elbo_o = _quiet(O.vireo_fit_vb, o, AD, DP, **kw)
    assert len(elbo) == len(elbo_o) == 2
    rel_close(elbo, elbo_o, E_TOL, "ELBO")
    rel_close(m.ID_prob, o.ID_prob, P_TOL, "ID_prob")
    rel_close(m.GT_prob, o.GT_prob, P_TOL, "GT_prob")
    rel_close(m.beta_mu, o.beta_mu, P_TOL, "beta_mu")
    rel_close(m.beta_sum, o.beta_sum, P_TOL, "beta_sum")
    assert np.array_equal(m.ID_prob.argmax(1), o.ID_prob.argmax(1))


def _engine_info(vb, counts, what):
    from vireo_b200 import _lib
    return _lib.load().vb_counts_info(counts.handle, what)


def test_cfg3_full_size_invariants_and_doublet(vb, cfg3):
    """Size-independent properties at cfg3 (count conservation, normalisation, monotone ELBO, planted donors, additivity
    of the cell pass over a split of the SNPs), then the doublet pass at the headline donor count (136 columns = 9 chunks
    of 16 through the window-segment kernel) against the row-kernel implementation of the same pass."""
    from vireo_b200 import _lib
    AD, DP, donor, _ = cfg3
    C, V, K = 100000, 50000, 16
    counts = vb.stage(AD, DP)
    np.random.seed(1)
    m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
    m.update_theta_size(counts, None)
    tot_dp, tot_ad = float(DP.data.sum()), float(AD.data.sum())
    assert abs((m.beta_sum.sum() - 150.0) - tot_dp) <= 1e-9 * tot_dp
    assert abs(((m.beta_mu * m.beta_sum).sum() - 75.0) - tot_ad) <= 1e-9 * tot_ad
    np.random.seed(1)
    m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
    m.fit(counts, None, max_iter=12, min_iter=12, delay_fit_theta=3, verbose=False)
    assert np.abs(m.ID_prob.sum(1) - 1).max() < 1e-12 and np.abs(m.GT_prob.sum(2) - 1).max() < 1e-12
    assert (np.diff(m.ELBO_) > -1e-6 * np.abs(m.ELBO_[1:])).all()
    conf = np.zeros((K, K), int)
    np.add.at(conf, (donor, m.ID_prob.argmax(1)), 1)
    assert conf.max(1).sum() >= 0.95 * C
    ll_all = m.update_ID_prob(counts, None)
    half = V // 2
    lo = vb.Vireo(n_cell=C, n_var=half, n_donor=K, GT_prob_init=m.GT_prob[:half], ID_prob_init=m.ID_prob,
                  beta_mu_init=m.beta_mu, beta_sum_init=m.beta_sum)
    hi = vb.Vireo(n_cell=C, n_var=V - half, n_donor=K, GT_prob_init=m.GT_prob[half:], ID_prob_init=m.ID_prob,
                  beta_mu_init=m.beta_mu, beta_sum_init=m.beta_sum)
    ADr, DPr = AD.tocsr(), DP.tocsr()
    ll_lo = lo.update_ID_prob(ADr[:half].tocsc(), DPr[:half].tocsc())
    ll_hi = hi.update_ID_prob(ADr[half:].tocsc(), DPr[half:].tocsc())
    assert np.max(np.abs(ll_all - (ll_lo + ll_hi))) <= 1e-9 * np.max(np.abs(ll_all))
    del ADr, DPr, lo, hi

    # doublet pass: segment chunks (auto) vs row kernels on the same state
    state = (m.ID_prob.copy(), m.GT_prob.copy())
    res = {}
    for path in ("auto", "rows"):
        _lib.set_path(path)
        mm = vb.Vireo(n_cell=C, n_var=V, n_donor=K, ID_prob_init=state[0], GT_prob_init=state[1],
                      beta_mu_init=m.beta_mu.copy(), beta_sum_init=m.beta_sum.copy())
        mm.ID_prob, mm.GT_prob = state[0].copy(), state[1].copy()
        res[path] = vb.predict_doublet(mm, counts, None) + (mm.GT_prob,)
    _lib.set_path("auto")
    dbl, sgl, llr, gt = res["auto"]
    assert dbl.shape == (C, 120) and np.abs(dbl.sum(1) + sgl.sum(1) - 1).max() < 1e-12
    rel_close(dbl, res["rows"][0], 1e-9, "doublet_prob, segment chunks vs row kernels")
    rel_close(sgl, res["rows"][1], 1e-9, "singlet_prob")
    assert np.max(np.abs(llr - res["rows"][2])) < 1e-9
    rel_close(gt, res["rows"][3], 1e-9, "GT_prob after the doublet pass")


def test_fixed_point_family_at_the_benchmark_shape(vb, cfg3):
    """cfg3, 20 free-running iterations: the opt-in fixed-point family against the FP64 segment family meets the
    north-star gate here (probabilities 1e-5, ELBO 1e-6, identical donor per cell); one update differs by < 2e-6."""
    from vireo_b200 import _lib
    AD, DP, _, _ = cfg3
    C, V, K = 100000, 50000, 16
    counts = vb.stage(AD, DP)
    out = {}
    try:
        for path in ("seg", "seg32"):
            _lib.set_path(path)
            np.random.seed(1)
            m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
            m.fit(counts, None, max_iter=20, min_iter=20, delay_fit_theta=3, verbose=False)
            out[path] = m
        a, b = out["seg"], out["seg32"]
        rel_close(b.ELBO_, a.ELBO_, E_TOL, "ELBO")
        rel_close(b.ID_prob, a.ID_prob, P_TOL, "ID_prob")
        rel_close(b.GT_prob, a.GT_prob, P_TOL, "GT_prob")
        assert np.array_equal(a.ID_prob.argmax(1), b.ID_prob.argmax(1))
        state = (a.ID_prob.copy(), a.GT_prob.copy(), a.beta_mu.copy(), a.beta_sum.copy())
        res = {}
        for path in ("seg", "seg32"):
            _lib.set_path(path)
            m = vb.Vireo(n_cell=C, n_var=V, n_donor=K, ID_prob_init=state[0].copy(), GT_prob_init=state[1].copy(),
                         beta_mu_init=state[2].copy(), beta_sum_init=state[3].copy())
            m.ID_prob, m.GT_prob = state[0].copy(), state[1].copy()
            m.update_GT_prob(counts, None)
            ll = m.update_ID_prob(counts, None)
            res[path] = (ll, m.ID_prob.copy(), m.GT_prob.copy())
        assert np.max(np.abs(res["seg32"][0] - res["seg"][0])) < 2e-6
        rel_close(res["seg32"][1], res["seg"][1], 2e-6, "ID_prob, one update")
        rel_close(res["seg32"][2], res["seg"][2], 2e-6, "GT_prob, one update")
    finally:
        _lib.set_path("auto")
        vb.clear_cache()


def test_cfg4_gt_given_full_size_vs_oracle(vb):
    """BASELINE cfg4 as specified: 50k cells x 20k SNPs x 8 donors with known donor genotypes -- learn_GT = False
    (reference vireo_wrap.py:48-50), GT_prior = GT_prob_init = 0.98 on the true genotype and 0.01 elsewhere -- five
    free-running iterations at full size against the oracle."""
    C, V, K = 50000, 20000, 8
    AD, DP, donor, GT = O.synth_counts(C, V, K, seed=0)
    prior = np.full((V, K, 3), 0.01)
    np.put_along_axis(prior, GT[:, :, None], 0.98, axis=2)
    np.random.seed(1)
    m = vb.Vireo(n_cell=C, n_var=V, n_donor=K, learn_GT=False, GT_prob_init=prior.copy())
    m.set_prior(GT_prior=prior.copy())
    o = O.vireo_new(C, V, K, learn_GT=False, GT_prob_init=prior.copy(), ID_prob_init=m.ID_prob.copy())
    O.vireo_set_prior(o, GT_prior=prior.copy())
    o.ID_prob = m.ID_prob.copy()
    counts = vb.stage(AD, DP)
    kw = dict(max_iter=5, min_iter=5, delay_fit_theta=0, verbose=False)
    elbo = _quiet(m._fit_VB, counts, None, **kw)
    assert _family(counts, 2) == 1                      # 64-byte-row segment kernels (n_donor <= 8)
    elbo_o = _quiet(O.vireo_fit_vb, o, AD, DP, **kw)
    assert len(elbo) == len(elbo_o) == 4
    rel_close(elbo, elbo_o, E_TOL, "ELBO")
    rel_close(m.ID_prob, o.ID_prob, P_TOL, "ID_prob")
    rel_close(m.GT_prob, o.GT_prob, P_TOL, "GT_prob (fixed)")
    rel_close(m.beta_mu, o.beta_mu, P_TOL, "beta_mu")
    rel_close(m.beta_sum, o.beta_sum, P_TOL, "beta_sum")
    assert np.array_equal(m.ID_prob.argmax(1), o.ID_prob.argmax(1))
    assert (m.ID_prob.argmax(1) == donor).mean() > 0.99
    # the whole wrapper in this mode: one restart, the final fit, the doublet pass
    rv = _quiet(vb.vireo_wrap, counts, None, GT_prior=prior.copy(), n_donor=K, learn_GT=False, n_init=3, random_seed=1)
    assert len(rv["LB_list"]) == 1 and rv["ID_prob"].shape == (C, K) and rv["doublet_prob"].shape == (C, 28)
    assert (rv["ID_prob"].argmax(1) == donor).mean() > 0.97
    vb.clear_cache()


def test_cfg5_clone_mode_full_restarts_vs_oracle(vb):
    """BASELINE cfg5 as specified: BinomMixtureVB on 2k cells x 300 mito SNPs x 6 clones, n_init = 50, min_iter = 30
    (reference bmm_model.py:204-263; the oracle needs about 20 s)."""
    AD, DP, _ = O.synth_clones(2000, 300, 6, seed=0)
    m = vb.BinomMixtureVB(n_var=300, n_cell=2000, n_donor=6)
    o = O.bmm_new(2000, 300, 6)
    _quiet(m.fit, AD, DP, min_iter=30, n_init=50, random_seed=1)
    _quiet(O.bmm_fit, o, AD, DP, min_iter=30, n_init=50, random_seed=1)
    assert len(m.ELBO_inits) == 50 and len(m.ELBO_iters) == len(o.ELBO_iters)
    rel_close(m.ELBO_inits, o.ELBO_inits, E_TOL, "ELBO_inits")
    assert int(np.argmax(m.ELBO_inits)) == int(np.argmax(o.ELBO_inits))
    rel_close(m.ELBO_iters, o.ELBO_iters, E_TOL, "ELBO_iters")
    rel_close(m.ID_prob, o.ID_prob, P_TOL, "ID_prob")
    rel_close(m.beta_mu, o.beta_mu, P_TOL, "beta_mu")
    assert np.array_equal(m.ID_prob.argmax(1), o.ID_prob.argmax(1))
    vb.clear_cache()


def test_cfg2_full_size_graph_replay_and_fused_tails_equal_plain_launches(vb):
    """BASELINE cfg2 (10k x 5k x 4, n_init = 1): small matrices replay captured CUDA graphs of whole iterations, and
    inside every fit loop the theta and ELBO steps run in the tails of the two sparse passes.  Both must be bit-identical
    to five plain launches per iteration, including a fit that stops early and a warm start."""
    from vireo_b200 import _lib
    AD, DP, _, _ = O.synth_counts(10000, 5000, 4, seed=0)
    counts = vb.stage(AD, DP)
    out = {}
    for graphs, fuse in ((1, 1), (0, 1), (1, 0), (0, 0)):
        _lib.load().vb_set_graphs(graphs)
        _lib.load().vb_set_fuse(fuse)
        np.random.seed(1)
        m = vb.Vireo(n_cell=10000, n_var=5000, n_donor=4)
        m.fit(counts, None, max_iter=20, min_iter=20, delay_fit_theta=3, verbose=False)
        m.fit(counts, None, max_iter=200, min_iter=5, verbose=False)       # warm start, stops on the convergence rule
        out[graphs, fuse] = (m.ELBO_.copy(), m.ID_prob.copy(), m.GT_prob.copy())
    _lib.load().vb_set_graphs(1)
    _lib.load().vb_set_fuse(1)
    assert len(out[0, 0][0]) > 19
    for key in ((1, 1), (0, 1), (1, 0)):
        assert len(out[key][0]) == len(out[0, 0][0])
        for a, b in zip(out[key], out[0, 0]):
            assert np.array_equal(a, b), key
    vb.clear_cache()


def test_fused_tails_equal_plain_launches_on_the_segment_kernels(vb):
    """The same identity for the window-segment kernels (forced on a K = 16 problem), for the ASE and fixed-GT flag
    combinations, for the clone model and for the sharded-fit loop on one rank."""
    from vireo_b200 import _lib
    AD, DP, _, GT = O.synth_counts(3000, 2000, 16, density=0.05, seed=2)
    prior = np.full((2000, 16, 3), 0.01)
    np.put_along_axis(prior, GT[:, :, None], 0.98, axis=2)
    out = {}
    try:
        for path in ("seg", "rows"):
            for fuse in (1, 0):
                _lib.set_path(path)
                _lib.load().vb_set_fuse(fuse)
                res = []
                for kw in (dict(), dict(ASE_mode=True), dict(learn_GT=False, GT_prob_init=prior.copy())):
                    np.random.seed(1)
                    m = vb.Vireo(n_cell=3000, n_var=2000, n_donor=16, **kw)
                    m.fit(AD, DP, max_iter=8, min_iter=3, delay_fit_theta=2, verbose=False)
                    res += [m.ELBO_.copy(), m.ID_prob.copy(), m.GT_prob.copy(), m.beta_mu.copy()]
                np.random.seed(1)
                s = vb.Vireo(n_cell=3000, n_var=2000, n_donor=16)
                vb.fit_cell_sharded(s, AD, DP, max_iter=30, min_iter=3, delay_fit_theta=2, verbose=False, poll_every=4)
                res += [s.ELBO_.copy(), s.ID_prob.copy(), s.GT_prob.copy()]
                b = vb.BinomMixtureVB(n_cell=3000, n_var=2000, n_donor=5)
                b.fit(AD, DP, n_init=3, max_iter=12, max_iter_pre=6, min_iter=2, random_seed=1, verbose=False)
                res += [b.ELBO_iters.copy(), b.ID_prob.copy()]
                out[path, fuse] = res
    finally:
        _lib.set_path("auto")
        _lib.load().vb_set_fuse(1)
    for path in ("seg", "rows"):
        for a, b in zip(out[path, 1], out[path, 0]):
            assert a.shape == b.shape and np.array_equal(a, b), path
    vb.clear_cache()


@pytest.mark.parametrize("shape", ["poisson", "heavy_tail", "tiny", "big_counts"])
def test_segment_format_holds_every_pair_exactly_once(vb, shape):
    """The window-segment formats (vb_seg.cu) re-encode the staged counts as lock-step record streams whose records
    carry a shared-memory ring row, not a table row.  vb_seg_verify decodes them on the host with the kernel's own rules
    (windows held at each super-step, ring addressing, count codes) and compares the multiset of (owner, gather row,
    count) with the staged matrices: every pair exactly once, stream or residual list, for both passes and all three
    table kinds, on shapes that stress the schedule (empty stretches, very uneven rows, counts without a code)."""
    import ctypes
    from vireo_b200 import _lib
    rng = np.random.default_rng(5)
    if shape == "poisson":
        AD, DP, _, _ = O.synth_counts(3000, 2000, 16, density=0.05, seed=2)
    elif shape == "heavy_tail":           # a few SNPs covered in most cells, most SNPs in a handful; some empty cells
        C, V = 2500, 1500
        pv = np.minimum(1.0, 0.6 / (1 + np.arange(V)) ** 0.9 * 20)
        mask = rng.random((V, C)) < pv[:, None]
        mask[:, rng.choice(C, 40, replace=False)] = False
        dp = np.where(mask, rng.integers(1, 6, size=(V, C)), 0)
        ad = rng.binomial(dp, 0.4)
        AD, DP = csc_matrix(ad), csc_matrix(dp)
    elif shape == "tiny":
        dp = np.zeros((7, 5), dtype=np.int64); dp[0, 0] = 3; dp[6, 4] = 1; dp[3, 2] = 40
        ad = np.zeros_like(dp); ad[0, 0] = 1; ad[3, 2] = 33
        AD, DP = csc_matrix(ad), csc_matrix(dp)
    else:                                 # counts with and without a 5-significant-bit code, and above 65535 (wide storage)
        C, V = 600, 400
        mask = rng.random((V, C)) < 0.08
        dp = np.where(mask, rng.choice([1, 2, 31, 33, 48, 62, 63, 100, 992, 1000, 70000], size=(V, C)), 0)
        ad = rng.binomial(dp, 0.5)
        AD, DP = csc_matrix(ad), csc_matrix(dp)
    counts = vb.stage(AD, DP)
    n_pairs = int((AD.toarray() > 0).sum() + ((DP - AD).toarray() > 0).sum())
    lib = _lib.load()
    for kind in (0, 1, 2):
        for ori in (0, 1):
            out = (ctypes.c_int64 * 4)()
            _lib.check(lib.vb_seg_verify(counts.handle, kind, ori, out))
            assert out[0] == n_pairs, (kind, ori, list(out), n_pairs)
            assert out[1] == 0, "format kind %d pass %d: %d errors %s" % (kind, ori, out[1], list(out))
    # the row-imbalance diagnostic (per mille of longest row / mean row over the built formats)
    skew = int(lib.vb_counts_info(counts.handle, 62))
    n_virtual = int(lib.vb_counts_info(counts.handle, 63))
    if shape == "poisson":
        assert 1000 <= skew < 2500 and n_virtual == 0, (skew, n_virtual)
    elif shape == "heavy_tail":      # rows far above the mean are cut into parts (FP64 formats), the fixed-point one keeps them
        assert skew > 4000 and n_virtual > 0, (skew, n_virtual)


def test_row_split_cell_pass_vs_oracle(vb):
    """Few cells, many SNPs (2000 x 20000 x 16): the cell pass of the window-segment family cuts the table rows into
    ranges that run side by side and adds their partial sums in a finish kernel (vb_counts_info 61 = launches in use).
    Free-running fits against the oracle: learn-GT from the reference's random start, and GT-given."""
    from vireo_b200 import _lib
    C, V, K = 2000, 20000, 16
    AD, DP, donor, GT = O.synth_counts(C, V, K, density=0.02, seed=11)
    kw = dict(max_iter=8, min_iter=8, delay_fit_theta=2, verbose=False)
    _lib.set_path("seg")
    try:
        counts = vb.stage(AD, DP)
        np.random.seed(3)
        m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
        o = O.vireo_new(C, V, K, ID_prob_init=m.ID_prob.copy(), GT_prob_init=m.GT_prob.copy())
        o.ID_prob, o.GT_prob = m.ID_prob.copy(), m.GT_prob.copy()
        elbo = _quiet(m._fit_VB, counts, None, **kw)
        assert int(_engine_info(vb, counts, 61)) >= 2, "row-split cell pass not in use"
        elbo_o = _quiet(O.vireo_fit_vb, o, AD, DP, **kw)
        rel_close(elbo, elbo_o, E_TOL, "ELBO")
        rel_close(m.ID_prob, o.ID_prob, P_TOL, "ID_prob")
        rel_close(m.GT_prob, o.GT_prob, P_TOL, "GT_prob")
        assert np.array_equal(m.ID_prob.argmax(1), o.ID_prob.argmax(1))

        prior = np.full((V, K, 3), 0.01)
        np.put_along_axis(prior, np.asarray(GT)[:, :, None], 0.98, axis=2)
        np.random.seed(4)
        m = vb.Vireo(n_cell=C, n_var=V, n_donor=K, learn_GT=False, GT_prob_init=prior.copy())
        m.set_prior(GT_prior=prior.copy())
        o = O.vireo_new(C, V, K, learn_GT=False, GT_prob_init=prior.copy(), ID_prob_init=m.ID_prob.copy())
        O.vireo_set_prior(o, GT_prior=prior.copy())
        o.ID_prob = m.ID_prob.copy()
        elbo = _quiet(m._fit_VB, counts, None, **kw)
        elbo_o = _quiet(O.vireo_fit_vb, o, AD, DP, **kw)
        rel_close(elbo, elbo_o, E_TOL, "ELBO (GT given)")
        rel_close(m.ID_prob, o.ID_prob, P_TOL, "ID_prob (GT given)")
        assert np.array_equal(m.ID_prob.argmax(1), o.ID_prob.argmax(1))
        assert (m.ID_prob.argmax(1) == donor).mean() > 0.99

        # restarts batched in grid.y go through the same split (one block of partial sums per restart): a batch of three
        # equals the three fits done one by one, bit for bit
        from vireo_b200 import _engine
        def trio():
            np.random.seed(7)
            return [vb.Vireo(n_cell=C, n_var=V, n_donor=K) for _ in range(3)]
        together, alone = trio(), trio()
        tr_b = _engine.vireo_fit_models(counts, together, 6, 6, 1e-2, 2, False)
        tr_a = [_engine.vireo_fit_models(counts, [m1], 6, 6, 1e-2, 2, False)[0] for m1 in alone]
        for mb, ma, eb, ea in zip(together, alone, tr_b, tr_a):
            assert np.array_equal(eb, ea) and np.array_equal(mb.ID_prob, ma.ID_prob) and np.array_equal(mb.GT_prob, ma.GT_prob)
    finally:
        _lib.set_path("auto")
        vb.clear_cache()


@pytest.mark.parametrize("K", [6, 16])
def test_heavy_tailed_rows_vs_oracle(vb, K):
    """Coverage of real data is heavy-tailed: a few SNPs are seen in most cells, most SNPs in a handful.  A warp task
    of the window-segment kernels is as long as its longest row, so the builder cuts rows above twice the mean into
    parts (virtual owners of a second format whose plain sums are folded into the row's residual sums before the pass).
    Free-running fit against the oracle with such rows present, 8-column and 16-column tables."""
    from vireo_b200 import _lib
    rng = np.random.default_rng(21)
    C, V = 3000, 2500
    pv = np.minimum(1.0, 30.0 / (1 + np.arange(V)) ** 0.9)            # SNP 0 in every cell, SNP 2499 in 2.6 % of them
    mask = rng.random((V, C)) < pv[:, None]
    dp = np.where(mask, rng.integers(1, 5, size=(V, C)), 0)
    donor = rng.integers(0, K, C)
    gt = rng.integers(0, 3, size=(V, K))
    ad = rng.binomial(dp, np.array([0.01, 0.5, 0.99])[gt[:, donor]])
    AD, DP = csc_matrix(ad), csc_matrix(dp)
    _lib.set_path("seg")
    try:
        counts = vb.stage(AD, DP)
        np.random.seed(5)
        m = vb.Vireo(n_cell=C, n_var=V, n_donor=K)
        o = O.vireo_new(C, V, K, ID_prob_init=m.ID_prob.copy(), GT_prob_init=m.GT_prob.copy())
        o.ID_prob, o.GT_prob = m.ID_prob.copy(), m.GT_prob.copy()
        kw = dict(max_iter=8, min_iter=8, delay_fit_theta=2, verbose=False)
        elbo = _quiet(m._fit_VB, counts, None, **kw)
        assert int(_engine_info(vb, counts, 63)) > 0, "no row was cut into parts"
        elbo_o = _quiet(O.vireo_fit_vb, o, AD, DP, **kw)
        rel_close(elbo, elbo_o, E_TOL, "ELBO")
        rel_close(m.ID_prob, o.ID_prob, P_TOL, "ID_prob")
        rel_close(m.GT_prob, o.GT_prob, P_TOL, "GT_prob")
        # same donor per cell; cells whose two best donors tie to round-off (few informative reads) may go either way
        diff = np.flatnonzero(m.ID_prob.argmax(1) != o.ID_prob.argmax(1))
        top2 = np.sort(o.ID_prob[diff], axis=1)[:, -2:]
        assert np.all(top2[:, 1] - top2[:, 0] < 1e-9 * top2[:, 1]), (len(diff), top2[:5])
        o.ID_prob, o.GT_prob = m.ID_prob.copy(), m.GT_prob.copy()       # the doublet pass goes through the same formats
        o.beta_mu, o.beta_sum = m.beta_mu.copy(), m.beta_sum.copy()
        dbl, sgl, llr = vb.predict_doublet(m, counts, None)
        dbl_o, sgl_o, llr_o = O.vireo_predict_doublet(o, AD, DP)
        rel_close(sgl, sgl_o, P_TOL, "singlet_prob")
        rel_close(dbl, dbl_o, P_TOL, "doublet_prob")
        assert np.max(np.abs(llr - llr_o)) < 1e-8
    finally:
        _lib.set_path("auto")
        vb.clear_cache()
