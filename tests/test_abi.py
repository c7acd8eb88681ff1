"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/vireo_b200.h declares,
its scalar math agrees with scipy, and the product path refuses to run without a GPU (no CPU fallback).
No compute entry point is called here."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy.special import betaln, binom, digamma

from conftest import ROOT
from vireo_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vireo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 13, names
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), "libvireo_b200.so lacks %s" % name
        assert name in _lib.SIGNATURES, "ctypes binding lacks %s" % name
    assert lib.vb_version().decode().endswith("sm_100a")


def test_struct_layouts_match_header():
    # field order/size as in the header: 14 int32 + 1 double + 21 pointers + vb_ws_sizes; 8 int32 + 1 double +
    # 17 pointers + vb_ws_sizes
    assert ctypes.sizeof(_lib.WsSizes) == 9 * 8
    assert ctypes.sizeof(_lib.VireoArgs) == 14 * 4 + 8 + 21 * 8 + 9 * 8
    assert ctypes.sizeof(_lib.BmmArgs) == 8 * 4 + 8 + 17 * 8 + 9 * 8
    assert ctypes.sizeof(_lib.DoubletWs) == 3 * 8
    assert _lib.VireoArgs.ws.offset == 14 * 4 + 8 + 21 * 8


def test_single_rank_comm_needs_no_nccl():
    """vb_comm with one rank: created and destroyed without NCCL and without a GPU (the sharded loop then runs without
    its exchange step); bad ranks are rejected."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.vb_comm_create(0, 1, 0, None, ctypes.byref(h)) == 0 and h.value
    lib.vb_comm_destroy(h)
    assert lib.vb_comm_create(0, 2, 2, None, ctypes.byref(h)) != 0
    assert lib.vb_comm_create(0, 2, 0, None, ctypes.byref(h)) != 0 and b"unique id" in lib.vb_last_error()


def test_cache_fingerprints_cover_the_full_buffers():
    """the staged-matrix cache and the prior cache must notice ANY in-place edit -- of values, indices or
    index pointers, anywhere in the buffer (a strided sample misses most positions)."""
    from scipy.sparse import random as sprandom
    from vireo_b200 import _engine
    rng = np.random.default_rng(0)
    M = sprandom(3000, 2000, density=0.05, format="csc", random_state=1, data_rvs=lambda n: rng.integers(1, 9, n))
    M.data = M.data.astype(np.int64)
    fp0 = _engine._fingerprint(M)
    assert _engine._fingerprint(M) == fp0
    M.data[1] += 3                                   # the advisor's example: an entry off any sample grid
    fp1 = _engine._fingerprint(M)
    assert fp1 != fp0
    M.indices[12345] ^= 1
    fp2 = _engine._fingerprint(M)
    assert fp2 != fp1
    M.indptr[7] += 1
    assert _engine._fingerprint(M) != fp2
    big = rng.random(5_000_000)                      # large enough for the threaded path
    c0 = _engine.checksum(big)
    big[4_999_999] += 1e-9
    c1 = _engine.checksum(big)
    big[17], big[3_000_017] = big[3_000_017], big[17]          # a swap across the thread chunks
    assert c0 != c1 and _engine.checksum(big) != c1
    prior = np.full((400, 8, 3), 1 / 3)
    p0 = _engine._fp_array(prior)
    prior[137, 0, :] = [0.5, 0.25, 0.25]
    assert _engine._fp_array(prior) != p0
    odd = np.arange(13, dtype=np.uint8)              # buffers that are not a multiple of 8 bytes
    assert _engine.checksum(odd) != _engine.checksum(odd[::-1].copy())


def test_library_is_built_for_sm100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_device_digamma_formula_matches_scipy():
    lib = _lib.load()
    xs = np.concatenate([np.logspace(-6, 9, 400), [0.5, 1.0, 1.4616321449683623, 2.0, 9.999, 10.0, 49.5, 25.0]])
    for x in xs:
        got, want = lib.vb_host_digamma(float(x)), float(digamma(x))
        assert abs(got - want) <= 2e-15 * max(1.0, abs(want)), (x, got, want)
    assert np.isnan(lib.vb_host_digamma(0.0)) and np.isnan(lib.vb_host_digamma(-1.0))


def test_device_beta_kl_matches_reference_formula():
    lib = _lib.load()

    def ref(p1, p2, q1, q2):   # vireoSNP/utils/vireo_base.py:96-127
        def cross(a1, a2, b1, b2):
            return betaln(b1, b2) - (b1 - 1) * digamma(a1) - (b2 - 1) * digamma(a2) + (b1 + b2 - 2) * digamma(a1 + a2)
        return cross(p1, p2, q1, q2) - cross(p1, p2, p1, p2)

    rng = np.random.default_rng(0)
    for _ in range(200):
        p1, p2 = np.exp(rng.uniform(-1, 12, 2))
        q1, q2 = rng.choice([0.5, 1.0, 25.0, 49.5]), rng.choice([0.5, 1.0, 25.0, 49.5])
        got, want = lib.vb_host_beta_kl(p1, p2, q1, q2), ref(p1, p2, q1, q2)
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (p1, p2, q1, q2, got, want)


def test_device_binom_term_matches_get_binom_coeff():
    lib = _lib.load()
    rng = np.random.default_rng(1)
    d = np.concatenate([rng.integers(1, 60, 300), rng.integers(60, 5000, 100), [1, 1, 2, 900, 1200, 85197]])
    a = (d * rng.random(d.size)).astype(np.int64)
    with np.errstate(over="ignore", divide="ignore"):
        want = np.log(binom(d, a))
    want[want > 700] = 700
    want = want.astype(np.float32)
    for ai, di, w in zip(a, d, want):
        got = lib.vb_host_binom_term(int(ai), int(di))
        assert got == w or abs(got - w) <= 1.2e-7 * abs(w), (ai, di, got, w)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vireo_b200
    m = vireo_b200.Vireo(n_cell=5, n_var=4, n_donor=2)
    with pytest.raises(vireo_b200.VireoB200Error):
        m.fit(np.ones((4, 5)), np.ones((4, 5)))
    with pytest.raises(vireo_b200.VireoB200Error):
        vireo_b200.BinomMixtureVB(n_cell=5, n_var=4, n_donor=2).fit(np.ones((4, 5)), np.ones((4, 5)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vireo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f


def test_api_surface_mirrors_reference():
    """Names and signatures a vireoSNP user relies on (reference vireoSNP/__init__.py:12-15, doc/API.rst:59-75)."""
    import inspect
    import vireo_b200 as vb
    assert vb.vireo_flock is vb.vireo_wrap
    sig = inspect.signature(vb.Vireo.__init__)
    assert list(sig.parameters)[1:] == ["n_cell", "n_var", "n_donor", "n_GT", "learn_GT", "learn_theta", "ASE_mode",
                                        "fix_beta_sum", "beta_mu_init", "beta_sum_init", "ID_prob_init", "GT_prob_init"]
    sig = inspect.signature(vb.Vireo.fit)
    assert list(sig.parameters)[1:] == ["AD", "DP", "max_iter", "min_iter", "epsilon_conv", "delay_fit_theta",
                                        "verbose", "n_inits", "nproc"]
    assert sig.parameters["max_iter"].default == 200 and sig.parameters["min_iter"].default == 5
    sig = inspect.signature(vb.vireo_wrap)
    assert list(sig.parameters) == ["AD", "DP", "GT_prior", "n_donor", "learn_GT", "n_init", "random_seed",
                                    "check_doublet", "max_iter_init", "delay_fit_theta", "n_extra_donor",
                                    "extra_donor_mode", "check_ambient", "nproc", "kwargs"]
    assert sig.parameters["n_init"].default == 20 and sig.parameters["nproc"].default == 4
    sig = inspect.signature(vb.BinomMixtureVB.fit)
    assert list(sig.parameters)[1:] == ["AD", "DP", "n_init", "max_iter", "max_iter_pre", "random_seed", "kwargs"]
    for name in ("update_theta_size", "update_ID_prob", "update_GT_prob", "get_ELBO", "_fit_VB", "set_initial",
                 "set_prior", "theta_s1", "theta_s2", "digamma1_", "digamma2_", "digammas_"):
        assert hasattr(vb.Vireo, name), name


def test_host_state_setup_matches_reference_semantics():
    """Constructor RNG order (Q4) and the in-place GT_prior clipping (Q5) -- pure host logic."""
    import vireo_b200 as vb
    from oracle import vireo_oracle as O
    np.random.seed(3)
    m = vb.Vireo(n_cell=7, n_var=5, n_donor=3)
    np.random.seed(3)
    o = O.vireo_new(7, 5, 3)
    assert np.array_equal(m.ID_prob, o.ID_prob) and np.array_equal(m.GT_prob, o.GT_prob)
    assert np.array_equal(m.theta_s1_prior, o.theta_s1_prior) and np.array_equal(m.ID_prior, o.ID_prior)
    prior = np.zeros((5, 3, 3)); prior[:, :, 0] = 1.0
    keep = prior.copy()
    m.set_prior(GT_prior=prior)
    assert prior.min() == 1e-5 and prior.max() == 1 - 1e-5      # caller's array was clipped in place
    o2 = O.vireo_new(7, 5, 3, ID_prob_init=m.ID_prob, GT_prob_init=m.GT_prob)
    O.vireo_set_prior(o2, GT_prior=keep)
    assert np.array_equal(m.GT_prior, o2.GT_prior)
    np.random.seed(5)
    b = vb.BinomMixtureVB(n_cell=6, n_var=4, n_donor=2)
    np.random.seed(5)
    ob = O.bmm_new(6, 4, 2)
    assert np.array_equal(b.ID_prob, ob.ID_prob) and np.array_equal(b.beta_sum, ob.beta_sum)
    assert np.array_equal(b.theta_s1_prior, ob.theta_s1_prior)


def test_segment_record_count_codes():
    """The records of the window-segment format carry a count as the upper 16 bits of the double 2 * count
    (vb_seg.cu): exact for every count with at most 5 significant bits, refused (residual list) otherwise; those
    bits share their upper 7 for every count, which is what lets the kernel form a shared-memory address from the
    whole record with one shift-add.  Fixed-point tables carry the integer count up to 31."""
    lib = _lib.load()
    for c in list(range(1, 300)) + [496, 512, 992, 1000, 1 << 20, 31 << 18, (1 << 24) - 1, 1 << 24]:
        code = lib.vb_host_seg_count_code(c, 0)
        sig = len(bin(c).rstrip("0")) - 2            # significant bits
        if sig <= 5 and c < (1 << 24):
            top16 = int(np.float64(2.0 * c).view(np.uint64) >> np.uint64(48))
            assert code == top16, (c, hex(code), hex(top16))
            assert np.uint64(code << 48).view(np.float64) == 2.0 * c
            assert code >> 9 == 32 and code >> 10 == 16
        else:
            assert code == 0xffffffff, c
        assert lib.vb_host_seg_count_code(c, 1) == (c if c <= 31 else 0xffffffff)
    assert lib.vb_host_seg_count_code(0, 0) == 0xffffffff
