"""Host-side engine: stages AD/DP into HBM once, runs batches of restarts through libvireo_b200.

PyTorch is used only as a container for device memory (tensors supply ``data_ptr()``), for streams and
for ``torch.distributed``; all arithmetic of the EM path happens in the hand-written CUDA kernels of
``csrc/``.  There is no CPU fallback: without the shared library or without a CUDA device every
compute entry point raises.
"""
import ctypes as C
import os
import weakref

import numpy as np
from scipy import sparse as sp

from . import _lib
from ._lib import VireoB200Error  # noqa: F401  (re-export)

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise VireoB200Error("vireo_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return t


def default_device():
    env = os.environ.get("VIREO_B200_DEVICE")
    if env is not None:
        return int(env)
    return require_cuda().cuda.current_device()


def _stream(device):
    return C.c_void_p(torch().cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------------------------
# staged count matrices
# ---------------------------------------------------------------------------------------------

_DTYPE_CODE = {np.dtype(np.int32): _lib.VB_I32, np.dtype(np.int64): _lib.VB_I64,
               np.dtype(np.float32): _lib.VB_F32, np.dtype(np.float64): _lib.VB_F64}


def _canonical_csc(M):
    """csc_matrix/csc_array with sorted, duplicate-free indices; the caller's object is never mutated."""
    if isinstance(M, np.ndarray) or isinstance(M, np.matrix):
        M = sp.csc_matrix(np.asarray(M))
    elif not sp.issparse(M):
        raise TypeError("AD/DP must be scipy.sparse matrices or numpy arrays, got %r" % type(M))
    elif M.format != "csc":
        M = M.tocsc()
    if not M.has_canonical_format:
        M = M.copy()
        M.sum_duplicates()
    return M


def _typed(a, codes):
    a = np.ascontiguousarray(a)
    if a.dtype not in codes:
        a = a.astype(np.float64 if a.dtype.kind == "f" else np.int64)
    return a


class StagedCounts:
    """AD and DP resident in HBM in both orientations (cell-major and SNP-major), staged once.

    Takes the place of the (AD, DP) pair wherever the API accepts one: ``model.fit(staged, None)``
    or simply ``model.fit(AD, DP)`` -- the latter stages on first use and caches by object identity.
    """

    def __init__(self, AD, DP, device=None):
        require_cuda()
        lib = _lib.load()
        self.device = default_device() if device is None else int(device)
        AD = _canonical_csc(AD)
        DP = _canonical_csc(DP)
        if AD.shape != DP.shape:
            raise ValueError("AD %r and DP %r differ in shape" % (AD.shape, DP.shape))
        self.n_var, self.n_cell = int(DP.shape[0]), int(DP.shape[1])
        idx_t = np.int64 if (DP.indptr.dtype == np.int64 or AD.indptr.dtype == np.int64 or
                             DP.indices.dtype == np.int64 or AD.indices.dtype == np.int64) else np.int32
        arrs = [np.ascontiguousarray(x, dtype=idx_t) for x in (DP.indptr, DP.indices, AD.indptr, AD.indices)]
        dp_data = _typed(DP.data, _DTYPE_CODE)
        ad_data = _typed(AD.data, _DTYPE_CODE)
        if ad_data.dtype != dp_data.dtype:
            common = np.float64 if "f" in (ad_data.dtype.kind, dp_data.dtype.kind) else np.int64
            ad_data, dp_data = ad_data.astype(common), dp_data.astype(common)
        handle = C.c_void_p()
        ptr = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        with torch().cuda.device(self.device):
            _lib.check(lib.vb_counts_create(
                self.device, self.n_cell, self.n_var,
                ptr(arrs[0]), _DTYPE_CODE[np.dtype(idx_t)], ptr(arrs[1]), _DTYPE_CODE[np.dtype(idx_t)],
                ptr(dp_data), _DTYPE_CODE[dp_data.dtype], int(DP.nnz),
                ptr(arrs[2]), ptr(arrs[3]), ptr(ad_data), int(AD.nnz),
                _stream(self.device), C.byref(handle)))
        self._h = handle
        self.nnz = int(lib.vb_counts_info(handle, 2))
        self.wide = bool(lib.vb_counts_info(handle, 3))
        self.bytes = int(lib.vb_counts_info(handle, 5))
        self._binom = None
        self._finalizer = weakref.finalize(self, lib.vb_counts_destroy, handle)

    # (n_var, n_cell), like the matrices it replaces
    @property
    def shape(self):
        return (self.n_var, self.n_cell)

    @property
    def handle(self):
        if self._h is None:
            raise VireoB200Error("StagedCounts was closed")
        return self._h

    def close(self):
        if self._h is not None:
            self._finalizer()
            self._h = None

    def binom_const(self):
        """float32 sum over nnz of min(log C(dp, ad), 700): the constant ``Vireo.fit`` adds to the ELBO
        (reference vireoSNP/utils/vireo_base.py:7-22, vireo_model.py:313), computed on the device."""
        if self._binom is None:
            t = torch()
            scratch = t.empty(1024, dtype=t.float64, device=self.device)
            out = C.c_double()
            _lib.check(_lib.load().vb_binom_const(self.handle, C.c_void_p(scratch.data_ptr()), C.byref(out),
                                                  _stream(self.device)))
            self._binom = np.float32(out.value)
        return self._binom


_CACHE = {}
_CACHE_MAX = 4


def _fingerprint(M):
    if isinstance(M, np.ndarray):
        return ("dense", M.shape, M.ctypes.data, float(M.ravel()[:: max(1, M.size // 1024)].sum()))
    data = M.data
    step = max(1, data.size // 1024)
    index = getattr(M, "indices", None)
    if index is None:
        index = getattr(M, "row", data)
    return (M.format, M.shape, int(M.nnz), data.ctypes.data, index.ctypes.data,
            float(data[::step].sum()) if data.size else 0.0)


def stage(AD, DP=None, device=None):
    """Return the StagedCounts for (AD, DP), staging on first use.  Cached by object identity plus a
    cheap fingerprint, so repeated ``fit`` calls on the same matrices upload nothing."""
    if isinstance(AD, StagedCounts):
        return AD
    dev = default_device() if device is None else int(device)
    key = (id(AD), id(DP), dev)
    fp = (_fingerprint(AD), _fingerprint(DP))
    hit = _CACHE.get(key)
    if hit is not None and hit[0] == fp and hit[1]._h is not None:
        return hit[1]
    staged = StagedCounts(AD, DP, dev)
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    _CACHE[key] = (fp, staged)
    try:   # drop the entry when either matrix is garbage collected (ids get recycled)
        weakref.finalize(AD, _CACHE.pop, key, None)
        weakref.finalize(DP, _CACHE.pop, key, None)
    except TypeError:
        pass
    return staged


def clear_cache():
    for _, s in list(_CACHE.values()):
        s.close()
    _CACHE.clear()
    _PRIORS.clear()


# ---------------------------------------------------------------------------------------------
# helpers shared by the model classes
# ---------------------------------------------------------------------------------------------

def _dev(arr, device):
    """float64 C-contiguous host array -> device tensor."""
    t = torch()
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if not a.flags.writeable:
        a = a.copy()
    return t.from_numpy(a).to("cuda:%d" % device)


_PINNED_UP = {}


def _stage_up(shape, device, fill):
    """Host -> device through a reused pinned buffer: `fill(view)` writes the float64 payload straight into the
    pinned staging area (one host pass instead of a temporary plus a pageable copy), then one DMA."""
    t = torch()
    n = int(np.prod(shape))
    buf = _PINNED_UP.get(n)
    if buf is None:
        if len(_PINNED_UP) > 16:
            _PINNED_UP.clear()
        buf = t.empty(max(n, 1), dtype=t.float64, pin_memory=True)
        _PINNED_UP[n] = buf
    fill(buf.numpy()[:n].reshape(shape))
    out = t.empty(max(n, 1), dtype=t.float64, device="cuda:%d" % device)
    out[:n].copy_(buf[:n], non_blocking=True)
    t.cuda.current_stream(device).synchronize()      # the staging buffer is reused by the next upload
    return out


def _zeros(n, device, dtype=None):
    t = torch()
    return t.zeros(int(max(n, 1)), dtype=dtype or t.float64, device="cuda:%d" % device)


def _ptr(tensor):
    return C.c_void_p(tensor.data_ptr())


def _log_prior_pair(prior, device):
    """Device tensors (log prior as used inside the softmax, log of the row-normalised prior as
    scipy.stats.entropy uses); the prior is uploaded once and both logs are taken by a kernel."""
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    n_col = prior.shape[-1]
    n_row = prior.size // max(n_col, 1)
    src = _dev(prior.reshape(-1), device)
    raw, norm = _zeros(prior.size, device), _zeros(prior.size, device)
    with torch().cuda.device(device):
        _lib.check(_lib.load().vb_log_prior(_ptr(src), n_row, n_col, _ptr(raw), _ptr(norm), _stream(device)))
    return raw, norm


_PINNED = {}


def _to_host(tensor):
    """Device tensor -> fresh numpy array, through a reused pinned staging buffer (pageable
    device-to-host copies run at a fraction of the link rate)."""
    t = torch()
    n = tensor.numel()
    key = (n, tensor.dtype)
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) > 16:
            _PINNED.clear()
        buf = t.empty(n, dtype=tensor.dtype, pin_memory=True)
        _PINNED[key] = buf
    buf.copy_(tensor.reshape(-1), non_blocking=True)
    t.cuda.current_stream(tensor.device).synchronize()
    return buf.numpy().copy().reshape(tuple(tensor.shape))


def _compress_rows(a):
    """(n, K) array with identical rows -> (1, K)."""
    if a.shape[0] > 1 and (a == a[:1]).all():
        return a[:1]
    return a


def replay_convergence(elbo, last, max_iter, min_iter, eps, bmm, verbose):
    """Re-run the reference's convergence rule on the host over the downloaded trace, only to print the
    warnings it would have printed, in order (vireo_model.py:266-274 / bmm_model.py:190-199)."""
    if not verbose:
        return
    for it in range(last + 1):
        if it > min_iter:
            if bmm:
                if elbo[it] - elbo[it - 1] < -1e-6:
                    print("Warning: ELBO decreases %.8f to %.8f!\n" % (elbo[it - 1], elbo[it]))
                elif it == max_iter - 1:
                    print("Warning: VB did not converge!\n")
                elif elbo[it] - elbo[it - 1] < eps:
                    break
            else:
                if elbo[it] < elbo[it - 1] - 1e-6:
                    print("Warning: Lower bound decreases!\n")
                elif it == max_iter - 1:
                    print("Warning: VB did not converge!\n")
                elif elbo[it] - elbo[it - 1] < eps:
                    break


# ---------------------------------------------------------------------------------------------
# Vireo batches
# ---------------------------------------------------------------------------------------------

_PRIORS = {}


def _fp_array(a):
    a = np.asarray(a)
    flat = a.reshape(-1) if a.flags.c_contiguous else a.ravel()
    step = max(1, flat.size // 2048)
    return (a.shape, a.dtype.str, a.ctypes.data, float(flat[::step].sum()) if flat.size else 0.0,
            float(flat[0]) if flat.size else 0.0, float(flat[-1]) if flat.size else 0.0)


def _vireo_priors(m0, dev, C_, V, K, G, T):
    """Device-side priors of a Vireo model: logs of the donor and genotype priors (both flavours, see
    ``_log_prior_pair``) and the theta prior.  They are constants of the model like the count matrices, so they are
    cached the same way (object identity plus a sampled fingerprint): repeated fits of one model upload its state only."""
    arrs = (m0.ID_prior, m0.GT_prior, m0.theta_s1_prior, m0.theta_s2_prior)
    key = tuple(id(a) for a in arrs) + (dev, C_, V, K, G, T)
    fp = tuple(_fp_array(a) for a in arrs)
    hit = _PRIORS.get(key)
    if hit is not None and hit[0] == fp:
        return hit[1]
    out = {}
    id_prior = np.asarray(m0.ID_prior, dtype=np.float64)
    if id_prior.ndim == 1:
        id_prior = id_prior[None, :]
    id_prior = _compress_rows(id_prior)
    if id_prior.shape[0] not in (1, C_) or id_prior.shape[1] != K:
        raise ValueError("ID_prior shape %r does not broadcast to (%d, %d)" % (id_prior.shape, C_, K))
    out["id_rows"] = id_prior.shape[0]
    out["lidp"], out["lidp_kl"] = _log_prior_pair(id_prior, dev)
    gt_prior = np.asarray(m0.GT_prior, dtype=np.float64)
    flat = np.broadcast_to(gt_prior, (V, K, G)).reshape(V * K, G)
    if V * K > 1 and (flat == flat[:1]).all():
        # one genotype prior for every (SNP, donor) -- the default: logs of one row, replicated on the device
        raw, norm = _log_prior_pair(flat[:1], dev)
        out["lgtp"] = raw.view(1, G).expand(V * K, G).contiguous().view(-1)
        out["lgtp_kl"] = norm.view(1, G).expand(V * K, G).contiguous().view(-1)
    else:
        out["lgtp"], out["lgtp_kl"] = _log_prior_pair(flat, dev)
    s1p = np.asarray(m0.theta_s1_prior, dtype=np.float64).reshape(-1, G)
    s2p = np.asarray(m0.theta_s2_prior, dtype=np.float64).reshape(-1, G)
    if s1p.shape[0] not in (1, T):
        raise ValueError("theta prior has %d rows, expected 1 or %d" % (s1p.shape[0], T))
    out["thp_rows"] = s1p.shape[0]
    out["s1p"], out["s2p"] = _dev(s1p, dev), _dev(s2p, dev)
    if len(_PRIORS) >= 8:
        _PRIORS.pop(next(iter(_PRIORS)))
    _PRIORS[key] = (fp, out)
    return out


class VireoBatch:
    """Device state of B Vireo restarts that share shapes, flags and priors."""

    def __init__(self, counts, models):
        self.counts = counts
        self.models = list(models)
        m0 = self.models[0]
        dev = counts.device
        self.dev = dev
        B = len(self.models)
        C_, V, K, G = counts.n_cell, counts.n_var, int(m0.n_donor), int(m0.n_GT)
        if (m0.n_cell, m0.n_var) != (C_, V):
            raise ValueError("model is (%d cells, %d variants) but the count matrices are (%d variants, %d cells)"
                             % (m0.n_cell, m0.n_var, V, C_))
        if G > _lib.MAX_GT or K > _lib.MAX_DONOR:
            raise VireoB200Error("n_GT <= %d and n_donor <= %d supported" % (_lib.MAX_GT, _lib.MAX_DONOR))
        self.B, self.C, self.V, self.K, self.G = B, C_, V, K, G
        self.ase = bool(m0.ASE_mode)
        T = V if self.ase else 1
        self.T = T
        t = torch()

        def stack(name, shape):
            def fill(out):
                for i, m in enumerate(self.models):
                    out[i] = np.broadcast_to(np.asarray(getattr(m, name), dtype=np.float64), shape)
            return _stage_up((B,) + shape, dev, fill)

        self.id_prob = stack("ID_prob", (C_, K))
        self.gt_prob = stack("GT_prob", (V, K, G))
        self.beta_mu = stack("beta_mu", (T, G))
        self.beta_sum = stack("beta_sum", (T, G))

        pri = _vireo_priors(m0, dev, C_, V, K, G, T)
        self.id_rows, self.thp_rows = pri["id_rows"], pri["thp_rows"]
        self.lidp, self.lidp_kl, self.lgtp, self.lgtp_kl = pri["lidp"], pri["lidp_kl"], pri["lgtp"], pri["lgtp_kl"]
        self.s1p, self.s2p = pri["s1p"], pri["s2p"]

        ws = _lib.WsSizes()
        _lib.check(_lib.load().vb_vireo_ws_sizes(counts.handle, K, G, B, int(self.ase), C.byref(ws)))
        self.S12 = _zeros(2 * ws.S, dev)          # S1 | S2 in one allocation: the cell-sharded fit all-reduces both at once
        self.S1, self.S2 = self.S12[:ws.S], self.S12[ws.S:]
        self.W = _zeros(ws.W, dev)
        self.loglik = _zeros(ws.loglik, dev)
        self.ab = _zeros(ws.ab, dev)
        self.part = _zeros(ws.part, dev)
        self.scal = _zeros(ws.scal, dev)
        self.ctrl = _zeros(ws.ctrl, dev, t.int32)
        self.rpad = _zeros(ws.rpad, dev)
        self.heavy = _zeros(ws.heavy, dev)
        self.elbo = None

    def args(self, max_iter=1, min_iter=0, eps=1e-2, delay=0, poll_every=0):
        m0 = self.models[0]
        t = torch()
        if self.elbo is None or self.elbo.numel() != self.B * max_iter:
            self.elbo = _zeros(self.B * max_iter, self.dev)
        a = _lib.VireoArgs()
        a.n_donor, a.n_gt, a.n_batch = self.K, self.G, self.B
        a.ase_mode, a.learn_gt = int(self.ase), int(bool(m0.learn_GT))
        a.learn_theta, a.fix_beta_sum = int(bool(m0.learn_theta)), int(bool(m0.fix_beta_sum))
        a.id_prior_rows, a.theta_prior_rows = self.id_rows, self.thp_rows
        a.max_iter, a.min_iter, a.delay_fit_theta = int(max_iter), int(min_iter), int(delay)
        a.poll_every, a.epsilon_conv = int(poll_every), float(eps)
        for name in ("id_prob", "gt_prob", "beta_mu", "beta_sum", "S1", "S2", "W", "loglik", "ab", "part",
                     "scal", "ctrl", "elbo", "rpad", "heavy"):
            setattr(a, name, getattr(self, name).data_ptr())
        a.log_id_prior, a.log_id_prior_kl = self.lidp.data_ptr(), self.lidp_kl.data_ptr()
        a.log_gt_prior, a.log_gt_prior_kl = self.lgtp.data_ptr(), self.lgtp_kl.data_ptr()
        a.s1_prior, a.s2_prior = self.s1p.data_ptr(), self.s2p.data_ptr()
        del t
        return a

    # -- device runs -------------------------------------------------------------------------
    def run_fit(self, max_iter, min_iter, eps, delay, poll_every=0):
        """max_iter EM iterations (or fewer if every restart converges) on the device; no host copies."""
        a = self.args(max_iter, min_iter, eps, delay, poll_every)
        with torch().cuda.device(self.dev):
            _lib.check(_lib.load().vb_vireo_fit(self.counts.handle, C.byref(a), _stream(self.dev)))
        self.max_iter = max_iter

    def run_step(self, phases):
        a = self.args()
        with torch().cuda.device(self.dev):
            _lib.check(_lib.load().vb_vireo_step(self.counts.handle, C.byref(a), int(phases), _stream(self.dev)))

    # -- results -----------------------------------------------------------------------------
    def traces(self):
        """[(elbo values incl. the one the reference drops, last iteration index)] per restart."""
        ctrl = self.ctrl.cpu().numpy().reshape(self.B, _lib.CTRL_N)
        elbo = self.elbo.cpu().numpy().reshape(self.B, self.max_iter)
        return [(elbo[b], int(ctrl[b, 2])) for b in range(self.B)]

    def download(self, what=("ID_prob", "GT_prob", "theta")):
        host = {}
        B, C_, V, K, G, T = self.B, self.C, self.V, self.K, self.G, self.T
        if "ID_prob" in what:
            host["ID_prob"] = _to_host(self.id_prob).reshape(B, C_, K)
        if "GT_prob" in what:
            host["GT_prob"] = _to_host(self.gt_prob).reshape(B, V, K, G)
        if "theta" in what:
            host["beta_mu"] = _to_host(self.beta_mu).reshape(B, T, G)
            host["beta_sum"] = _to_host(self.beta_sum).reshape(B, T, G)
        for b, m in enumerate(self.models):
            if "ID_prob" in host:
                m.ID_prob = host["ID_prob"][b].copy() if self.B > 1 else host["ID_prob"][b]
            if "GT_prob" in host:
                m.GT_prob = host["GT_prob"][b].copy() if self.B > 1 else host["GT_prob"][b]
            if "beta_mu" in host:
                m.beta_mu = host["beta_mu"][b].copy()
                m.beta_sum = host["beta_sum"][b].copy()

    def scalars(self):
        return self.scal.cpu().numpy().reshape(self.B, _lib.SCAL_N)

    def loglik_host(self):
        return self.loglik.cpu().numpy().reshape(self.B, self.C, self.K)


def vireo_fit_models(counts, models, max_iter, min_iter, epsilon_conv, delay_fit_theta, verbose):
    """Fit a list of same-shaped Vireo models as one device batch; returns the ELBO[:it] trace of each
    (without the binomial constant) and writes the fitted state back into the model objects."""
    batch = VireoBatch(counts, models)
    batch.run_fit(max_iter, min_iter, epsilon_conv, delay_fit_theta)
    what = ["ID_prob", "theta"] + (["GT_prob"] if models[0].learn_GT else [])
    batch.download(what)
    out = []
    for (elbo, last), m in zip(batch.traces(), models):
        replay_convergence(elbo, last, max_iter, min_iter, epsilon_conv, False, verbose)
        out.append(elbo[:last].copy())
    return out


# ---------------------------------------------------------------------------------------------
# BinomMixtureVB batches
# ---------------------------------------------------------------------------------------------

class BmmBatch:
    """Device state of B binomial-mixture restarts: states is a list of dicts with ID_prob, beta_mu, beta_sum."""

    def __init__(self, counts, model, states):
        self.counts = counts
        dev = counts.device
        self.dev = dev
        B = len(states)
        C_, V, K = counts.n_cell, counts.n_var, int(model.n_donor)
        if (model.n_cell, model.n_var) != (C_, V):
            raise ValueError("model is (%d cells, %d variants) but the count matrices are (%d variants, %d cells)"
                             % (model.n_cell, model.n_var, V, C_))
        if K > _lib.MAX_DONOR:
            raise VireoB200Error("n_donor <= %d supported" % _lib.MAX_DONOR)
        self.B, self.C, self.V, self.K = B, C_, V, K
        self.fix_beta_sum = bool(model.fix_beta_sum)
        t = torch()

        def stack(name, shape):
            out = np.empty((B,) + shape, dtype=np.float64)
            for i, s in enumerate(states):
                out[i] = np.broadcast_to(np.asarray(s[name], dtype=np.float64), shape)
            return _dev(out, dev)

        self.id_prob = stack("ID_prob", (C_, K))
        self.beta_mu = stack("beta_mu", (V, K))
        self.beta_sum = stack("beta_sum", (V, K))
        id_prior = np.asarray(model.ID_prior, dtype=np.float64)
        if id_prior.ndim == 1:
            id_prior = id_prior[None, :]
        id_prior = _compress_rows(id_prior)
        self.id_rows = id_prior.shape[0]
        self.lidp, self.lidp_kl = _log_prior_pair(id_prior, dev)
        self.s1p = _dev(np.broadcast_to(np.asarray(model.theta_s1_prior, dtype=np.float64), (V, K)), dev)
        self.s2p = _dev(np.broadcast_to(np.asarray(model.theta_s2_prior, dtype=np.float64), (V, K)), dev)
        ws = _lib.WsSizes()
        _lib.check(_lib.load().vb_bmm_ws_sizes(counts.handle, K, B, C.byref(ws)))
        self.S1, self.S2 = _zeros(ws.S, dev), _zeros(ws.S, dev)
        self.W = _zeros(ws.W, dev)
        self.loglik = _zeros(ws.loglik, dev)
        self.part = _zeros(ws.part, dev)
        self.scal = _zeros(ws.scal, dev)
        self.ctrl = _zeros(ws.ctrl, dev, t.int32)
        self.rpad = _zeros(ws.rpad, dev)
        self.heavy = _zeros(ws.heavy, dev)
        self.elbo = None

    def args(self, max_iter=1, min_iter=0, eps=1e-2, poll_every=0):
        if self.elbo is None or self.elbo.numel() != self.B * max_iter:
            self.elbo = _zeros(self.B * max_iter, self.dev)
        a = _lib.BmmArgs()
        a.n_donor, a.n_batch, a.fix_beta_sum, a.id_prior_rows = self.K, self.B, int(self.fix_beta_sum), self.id_rows
        a.max_iter, a.min_iter, a.poll_every, a.epsilon_conv = int(max_iter), int(min_iter), int(poll_every), float(eps)
        for name in ("id_prob", "beta_mu", "beta_sum", "S1", "S2", "W", "loglik", "part", "scal", "ctrl",
                     "elbo", "rpad", "heavy"):
            setattr(a, name, getattr(self, name).data_ptr())
        a.log_id_prior, a.log_id_prior_kl = self.lidp.data_ptr(), self.lidp_kl.data_ptr()
        a.s1_prior, a.s2_prior = self.s1p.data_ptr(), self.s2p.data_ptr()
        return a

    def run_fit(self, max_iter, min_iter, eps, poll_every=0):
        a = self.args(max_iter, min_iter, eps, poll_every)
        with torch().cuda.device(self.dev):
            _lib.check(_lib.load().vb_bmm_fit(self.counts.handle, C.byref(a), _stream(self.dev)))
        self.max_iter = max_iter

    def run_step(self, phases):
        a = self.args()
        with torch().cuda.device(self.dev):
            _lib.check(_lib.load().vb_bmm_step(self.counts.handle, C.byref(a), int(phases), _stream(self.dev)))

    def traces(self):
        ctrl = self.ctrl.cpu().numpy().reshape(self.B, _lib.CTRL_N)
        elbo = self.elbo.cpu().numpy().reshape(self.B, self.max_iter)
        return [(elbo[b], int(ctrl[b, 2])) for b in range(self.B)]

    def download(self):
        return (self.id_prob.cpu().numpy(), self.beta_mu.cpu().numpy(), self.beta_sum.cpu().numpy())

    def scalars(self):
        return self.scal.cpu().numpy().reshape(self.B, _lib.SCAL_N)

    def loglik_host(self):
        return self.loglik.cpu().numpy().reshape(self.B, self.C, self.K)


# ---------------------------------------------------------------------------------------------
# doublet pass
# ---------------------------------------------------------------------------------------------

def doublet_pass(counts, GT_prob, beta_mu, beta_sum, log_prior_both, ase, want_loglik=False):
    """logLik over singlet + donor-pair columns, softmax with the doublet prior, LLR -- all on the device."""
    dev = counts.device
    V, K, G = GT_prob.shape
    K2 = K + K * (K - 1) // 2
    lp = np.asarray(log_prior_both, dtype=np.float64)
    lp = _compress_rows(lp)
    gt, mu, sm, lpd = _dev(GT_prob, dev), _dev(beta_mu, dev), _dev(beta_sum, dev), _dev(lp, dev)
    W = _zeros(2 * V * K2, dev)
    ll, pr, llr = _zeros(counts.n_cell * K2, dev), _zeros(counts.n_cell * K2, dev), _zeros(counts.n_cell, dev)
    with torch().cuda.device(dev):
        _lib.check(_lib.load().vb_vireo_doublet(counts.handle, K, G, int(bool(ase)), _ptr(gt), _ptr(mu), _ptr(sm),
                                                _ptr(lpd), lp.shape[0], _ptr(W), _ptr(ll), _ptr(pr),
                                                _ptr(llr), _stream(dev)))
    C_ = counts.n_cell
    # the log-likelihoods stay on the device (109 MB at cfg3; no caller reads them): posterior and LLR only
    return (None if not want_loglik else _to_host(ll).reshape(C_, K2), _to_host(pr).reshape(C_, K2), _to_host(llr))
