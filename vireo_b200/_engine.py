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

    Takes the place of the (AD, DP) pair wherever the API accepts one: ``model.fit(staged, None)``.
    ``model.fit(AD, DP)`` with scipy matrices also works: it stages on first use and re-uses the copy as long as a
    checksum over the FULL contents of both matrices still matches (see ``stage``); passing the handle skips that
    check and is the fast path for repeated calls.
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
        _lib.check(lib.vb_counts_create(
            self.device, self.n_cell, self.n_var,
            ptr(arrs[0]), _DTYPE_CODE[np.dtype(idx_t)], ptr(arrs[1]), _DTYPE_CODE[np.dtype(idx_t)],
            ptr(dp_data), _DTYPE_CODE[dp_data.dtype], int(DP.nnz),
            ptr(arrs[2]), ptr(arrs[3]), ptr(ad_data), int(AD.nnz),
            _stream(self.device), C.byref(handle)))
        self._adopt(handle, np.asarray(DP.indptr, dtype=np.int64).copy())

    def _adopt(self, handle, indptr):
        lib = _lib.load()
        self._h = handle
        self.indptr = indptr                       # nnz offsets of the cells (host copy: shard bounds are cut by nnz)
        self.nnz = int(lib.vb_counts_info(handle, 2))
        self.wide = bool(lib.vb_counts_info(handle, 3))
        self.bytes = int(lib.vb_counts_info(handle, 5))
        self._binom = None
        self._pool = {}                            # workspaces of finished batches, re-used by the next one
        self._shards = {}                          # (world, rank) -> (StagedCounts of this rank's cells, c0, c1, bounds)
        self._warned = False
        self._warned_skew = False
        self._finalizer = weakref.finalize(self, lib.vb_counts_destroy, handle)

    @classmethod
    def _from_handle(cls, handle, device, n_var, n_cell, indptr):
        self = cls.__new__(cls)
        self.device, self.n_var, self.n_cell = device, n_var, n_cell
        self._adopt(handle, indptr)
        return self

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
        shards, self._shards = self._shards, {}
        for sh in shards.values():
            if sh[0] is not self:                  # on one rank the "shard" is the handle itself
                sh[0].close()
        self._pool = {}
        if self._h is not None:
            self._finalizer()
            self._h = None

    def slice_cells(self, c0, c1):
        """StagedCounts of the cells [c0, c1), cut on the device from the resident arrays (no host traffic)."""
        out = C.c_void_p()
        _lib.check(_lib.load().vb_counts_slice(self.handle, int(c0), int(c1), _stream(self.device), C.byref(out)))
        return StagedCounts._from_handle(out, self.device, self.n_var, int(c1 - c0),
                                         (self.indptr[c0:c1 + 1] - self.indptr[c0]).copy())

    def check_family(self):
        """Say so (once) when the automatic selector had to fall back to the row kernels because building the
        window-segment formats FAILED -- the fit is then several times slower on a large matrix."""
        if not self._warned and int(_lib.load().vb_counts_info(self.handle, 60)) == 2:
            self._warned = True
            note = _lib.load().vb_counts_note(self.handle)
            import warnings
            warnings.warn("vireo_b200: the window-segment formats could not be built (%s); falling back to the row "
                          "kernels, which are several times slower on large matrices"
                          % (note.decode() if note else "?"), RuntimeWarning, stacklevel=3)
        if not self._warned_skew:
            skew = int(_lib.load().vb_counts_info(self.handle, 62))
            if skew > 8000:
                self._warned_skew = True
                import warnings
                warnings.warn("vireo_b200: one row of the count matrices carries %.0f times the mean number of entries; "
                              "the window-segment kernels are as slow as their longest row (DESIGN.md, section 8)"
                              % (skew / 1000.0), RuntimeWarning, stacklevel=3)

    def binom_const(self):
        """float32 sum over nnz of min(log C(dp, ad), 700): the constant ``Vireo.fit`` adds to the ELBO
        (reference vireoSNP/utils/vireo_base.py:7-22, vireo_model.py:313), computed on the device."""
        if self._binom is None:
            t = torch()
            scratch = t.empty(1024, dtype=t.float64, device="cuda:%d" % self.device)
            out = C.c_double()
            _lib.check(_lib.load().vb_binom_const(self.handle, C.c_void_p(scratch.data_ptr()), C.byref(out),
                                                  _stream(self.device)))
            self._binom = np.float32(out.value)
        return self._binom


_CACHE = {}
_CACHE_MAX = 4
_HASH_POOL = None


def checksum(a):
    """64-bit sum over the FULL buffer of an array (modular arithmetic on its bytes viewed as uint64/uint8): any
    single-element edit changes it.  Large buffers are summed by a few threads (numpy releases the GIL)."""
    a = np.asarray(a)
    if a.size == 0:
        return 0
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    raw = a.reshape(-1).view(np.uint8)
    n8 = raw.size // 8 * 8
    words = raw[:n8].view(np.uint64)
    tail = int(raw[n8:].astype(np.uint64).sum()) if n8 < raw.size else 0
    if words.size < (1 << 21):
        total = int(np.add.reduce(words)) if words.size else 0
    else:
        global _HASH_POOL
        if _HASH_POOL is None:
            from concurrent.futures import ThreadPoolExecutor
            _HASH_POOL = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
        parts = np.array_split(words, 16)
        # position-weighted per part, so that moving a value between parts is seen as well
        total = sum((i + 1) * int(x) for i, x in enumerate(_HASH_POOL.map(np.add.reduce, parts)))
    return (total + tail) & 0xFFFFFFFFFFFFFFFF


def _fingerprint(M):
    """Identity of the CONTENTS of a count matrix: shape, layout and full checksums of every buffer (values, indices
    and index pointers) -- an in-place edit anywhere invalidates a cached device copy."""
    if isinstance(M, np.ndarray):
        return ("dense", M.shape, M.dtype.str, checksum(M))
    parts = [M.format, M.shape, int(M.nnz), M.data.dtype.str, checksum(M.data)]
    for name in ("indices", "indptr", "row", "col", "offsets"):
        arr = getattr(M, name, None)
        if arr is not None:
            parts.append(checksum(arr))
    return tuple(parts)


def stage(AD, DP=None, device=None):
    """Return the StagedCounts for (AD, DP), staging on first use.

    A staged copy is re-used for the same two matrix objects only if a checksum over the full contents of both
    still matches, so an in-place edit of ``.data``, ``.indices`` or ``.indptr`` between two calls is always picked
    up (the reference reads the live matrices on every call).  The check reads both matrices once (tens of
    milliseconds at 1e8 nnz); callers that fit the same matrices repeatedly should stage once and pass the handle:
    ``counts = vireo_b200.stage(AD, DP); model.fit(counts, None)``."""
    if isinstance(AD, StagedCounts):
        return AD
    dev = default_device() if device is None else int(device)
    key = (id(AD), id(DP), dev)
    fp = (_fingerprint(AD), _fingerprint(DP))
    hit = _CACHE.get(key)
    if hit is not None and hit[0] == fp and hit[1]._h is not None:
        return hit[1]
    staged = StagedCounts(AD, DP, dev)          # a stale copy is simply dropped (freed when nothing refers to it)
    if len(_CACHE) >= _CACHE_MAX and key not in _CACHE:
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


def _upload_into(dst, arr, shape):
    """Host array (broadcast to `shape`) -> the device view `dst`, ONE asynchronous copy and no staging pass on the
    host: arrays this package handed out live in pinned memory (see ``_to_host``) and go by DMA; anything else is
    staged by the driver, and ``cudaMemcpyAsync`` returns once a pageable source has been consumed."""
    t = torch()
    a = np.asarray(arr, dtype=np.float64)
    if a.shape != tuple(shape):
        a = np.broadcast_to(a, shape)
    if not (a.flags.c_contiguous and a.flags.writeable):
        a = np.array(a, dtype=np.float64, order="C")
    dst.copy_(t.from_numpy(a).reshape(-1), non_blocking=True)


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


_PIN_LIMIT = int(os.environ.get("VIREO_B200_PIN_LIMIT_MB", "64")) << 20


def _to_host(tensor):
    """Device tensor -> numpy array that LIVES in pinned memory: one DMA, no second host copy.  The pinned block
    comes from PyTorch's caching host allocator and returns to it when the array is garbage collected."""
    t = torch()
    n = tensor.numel()
    if n * tensor.element_size() > _PIN_LIMIT:
        # page-locking a fresh block costs more than the copy it speeds up; blocks this large are rarely recycled
        return tensor.cpu().numpy()
    buf = t.empty(max(n, 1), dtype=tensor.dtype, pin_memory=True)
    buf[:n].copy_(tensor.reshape(-1), non_blocking=True)
    t.cuda.current_stream(tensor.device).synchronize()
    return buf.numpy()[:n].reshape(tuple(tensor.shape))


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
    return (a.shape, a.dtype.str, checksum(a))


def _vireo_priors(m0, dev, C_, V, K, G, T, rows=None):
    """Device-side priors of a Vireo model: logs of the donor and genotype priors (both flavours, see
    ``_log_prior_pair``) and the theta prior.  They are constants of the model like the count matrices, so the
    device copies are re-used while a checksum over the FULL contents of every prior array still matches
    (an in-place edit such as ``model.GT_prior[i, 0, :] = ...`` is always picked up).
    ``rows = (c0, c1, n_cell_full)``: the batch covers the cells [c0, c1) of a model over n_cell_full cells (cell-sharded
    fit): the priors are those of the full model, a per-cell ID prior is sliced on the device."""
    if rows is not None:
        c0, c1, C_full = rows
        out = dict(_vireo_priors(m0, dev, C_full, V, K, G, T))
        if out["id_rows"] != 1:
            out["lidp"], out["lidp_kl"] = out["lidp"][c0 * K:c1 * K], out["lidp_kl"][c0 * K:c1 * K]
            out["id_rows"] = c1 - c0
        return out
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


_WS_FIELDS = ("S", "W", "loglik", "ab", "part", "scal", "ctrl", "rpad", "heavy")


def _acquire_ws(counts, kind, K, G, B, ase, sizes_fn):
    """Workspaces (and the state buffer) of a batch on `counts`: taken from the handle's pool when a finished batch of
    the same geometry left them there, so that repeated fits allocate nothing.  The pool key includes the kernel
    family the sizes were computed for."""
    t = torch()
    dev = counts.device
    ws = _lib.WsSizes()
    _lib.check(sizes_fn(ws))
    counts.check_family()
    key = (kind, K, G, B, ase) + tuple(int(getattr(ws, f)) for f in _WS_FIELDS)
    free = counts._pool.get(key)
    if free:
        return key, ws, free.pop()
    bufs = {"S12": _zeros(2 * ws.S, dev), "W": _zeros(ws.W, dev), "loglik": _zeros(ws.loglik, dev),
            "ab": _zeros(ws.ab, dev), "part": _zeros(ws.part, dev), "scal": _zeros(ws.scal, dev),
            "ctrl": _zeros(ws.ctrl, dev, t.int32), "rpad": _zeros(ws.rpad, dev), "heavy": _zeros(ws.heavy, dev)}
    return key, ws, bufs


def _release_ws(counts, key, bufs):
    if counts._h is None:
        return
    free = counts._pool.setdefault(key, [])
    if len(free) < 2:
        free.append(bufs)


class VireoBatch:
    """Device state of B Vireo restarts that share shapes, flags and priors."""

    def __init__(self, counts, models, rows=None):
        """``rows = (c0, c1)``: `counts` holds the cells [c0, c1) of the models' matrices (cell-sharded fit): the batch
        keeps those rows of ID_prob and of a per-cell ID prior; everything else is the full model's."""
        self.counts = counts
        self.models = list(models)
        m0 = self.models[0]
        dev = counts.device
        self.dev = dev
        B = len(self.models)
        C_, V, K, G = counts.n_cell, counts.n_var, int(m0.n_donor), int(m0.n_GT)
        self.rows = None if rows is None else (int(rows[0]), int(rows[1]), int(m0.n_cell))
        if rows is not None and (B != 1 or rows[1] - rows[0] != C_ or m0.n_var != V):
            raise ValueError("cell range %r does not match the staged shard (%d cells)" % (rows, C_))
        if rows is None and (m0.n_cell, m0.n_var) != (C_, V):
            raise ValueError("model is (%d cells, %d variants) but the count matrices are (%d variants, %d cells)"
                             % (m0.n_cell, m0.n_var, V, C_))
        if G > _lib.MAX_GT or K > _lib.MAX_DONOR:
            raise VireoB200Error("n_GT <= %d and n_donor <= %d supported" % (_lib.MAX_GT, _lib.MAX_DONOR))
        self.B, self.C, self.V, self.K, self.G = B, C_, V, K, G
        self.ase = bool(m0.ASE_mode)
        T = V if self.ase else 1
        self.T = T
        t = torch()
        lib = _lib.load()

        with t.cuda.device(dev):
            self._key, self.ws, bufs = _acquire_ws(
                counts, "vireo", K, G, B, int(self.ase),
                lambda ws: lib.vb_vireo_ws_sizes(counts.handle, K, G, B, int(self.ase), _stream(dev), C.byref(ws)))
        self._bufs = bufs
        # state of the batch in ONE allocation: ID_prob | GT_prob | beta_mu | beta_sum (one download brings all of it)
        n_id, n_gt, n_th = B * C_ * K, B * V * K * G, B * T * G
        if "state" not in bufs or bufs["state"].numel() != n_id + n_gt + 2 * n_th:
            bufs["state"] = t.empty(max(n_id + n_gt + 2 * n_th, 1), dtype=t.float64, device="cuda:%d" % dev)
        st = bufs["state"]
        self.state = st
        self.id_prob, self.gt_prob = st[:n_id], st[n_id:n_id + n_gt]
        self.beta_mu, self.beta_sum = st[n_id + n_gt:n_id + n_gt + n_th], st[n_id + n_gt + n_th:n_id + n_gt + 2 * n_th]
        for i, m in enumerate(self.models):
            idp = m.ID_prob if rows is None else np.asarray(m.ID_prob)[rows[0]:rows[1]]
            _upload_into(self.id_prob[i * C_ * K:(i + 1) * C_ * K], idp, (C_, K))
            _upload_into(self.gt_prob[i * V * K * G:(i + 1) * V * K * G], m.GT_prob, (V, K, G))
            _upload_into(self.beta_mu[i * T * G:(i + 1) * T * G], m.beta_mu, (T, G))
            _upload_into(self.beta_sum[i * T * G:(i + 1) * T * G], m.beta_sum, (T, G))

        pri = _vireo_priors(m0, dev, C_, V, K, G, T, self.rows)
        self.id_rows, self.thp_rows = pri["id_rows"], pri["thp_rows"]
        self.lidp, self.lidp_kl, self.lgtp, self.lgtp_kl = pri["lidp"], pri["lidp_kl"], pri["lgtp"], pri["lgtp_kl"]
        self.s1p, self.s2p = pri["s1p"], pri["s2p"]

        self.S12 = bufs["S12"]                    # S1 | S2 in one allocation
        half = self.S12.numel() // 2
        self.S1, self.S2 = self.S12[:half], self.S12[half:]
        for name in ("W", "loglik", "ab", "part", "scal", "ctrl", "rpad", "heavy"):
            setattr(self, name, bufs[name])
        self.elbo = bufs.get("elbo")

    def close(self):
        """Hand the workspaces back to the pool of the staged matrices (also done when the batch is collected)."""
        if self._bufs is not None:
            self._bufs["elbo"] = self.elbo
            _release_ws(self.counts, self._key, self._bufs)
            self._bufs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def args(self, max_iter=1, min_iter=0, eps=1e-2, delay=0, poll_every=0):
        m0 = self.models[0]
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
        a.ws = self.ws
        return a

    # -- device runs -------------------------------------------------------------------------
    def run_fit(self, max_iter, min_iter, eps, delay, poll_every=0):
        """max_iter EM iterations (or fewer if every restart converges) on the device; no host copies."""
        a = self.args(max_iter, min_iter, eps, delay, poll_every)
        _lib.check(_lib.load().vb_vireo_fit(self.counts.handle, C.byref(a), _stream(self.dev)))
        self.max_iter = max_iter

    def run_step(self, phases):
        a = self.args()
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
        if "ID_prob" in what and "GT_prob" in what and "theta" in what:
            full = _to_host(self.state)                         # one DMA for the whole state
            n_id, n_gt, n_th = B * C_ * K, B * V * K * G, B * T * G
            host["ID_prob"] = full[:n_id].reshape(B, C_, K)
            host["GT_prob"] = full[n_id:n_id + n_gt].reshape(B, V, K, G)
            host["beta_mu"] = full[n_id + n_gt:n_id + n_gt + n_th].reshape(B, T, G)
            host["beta_sum"] = full[n_id + n_gt + n_th:n_id + n_gt + 2 * n_th].reshape(B, T, G)
        else:
            if "ID_prob" in what:
                host["ID_prob"] = _to_host(self.id_prob).reshape(B, C_, K)
            if "GT_prob" in what:
                host["GT_prob"] = _to_host(self.gt_prob).reshape(B, V, K, G)
            if "theta" in what:
                th = _to_host(self.state[B * C_ * K + B * V * K * G:])
                host["beta_mu"], host["beta_sum"] = th[:B * T * G].reshape(B, T, G), th[B * T * G:].reshape(B, T, G)
        for b, m in enumerate(self.models):
            if "ID_prob" in host:
                m.ID_prob = host["ID_prob"][b]
            if "GT_prob" in host:
                m.GT_prob = host["GT_prob"][b]
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
    batch.close()
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
        _lib.check(_lib.load().vb_bmm_ws_sizes(counts.handle, K, B, _stream(dev), C.byref(ws)))
        counts.check_family()
        self.ws = ws
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
        a.ws = self.ws
        return a

    def run_fit(self, max_iter, min_iter, eps, poll_every=0):
        a = self.args(max_iter, min_iter, eps, poll_every)
        _lib.check(_lib.load().vb_bmm_fit(self.counts.handle, C.byref(a), _stream(self.dev)))
        self.max_iter = max_iter

    def run_step(self, phases):
        a = self.args()
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

def doublet_pass(counts, GT_prob, beta_mu, beta_sum, log_prior_both, ase, want_loglik=False, keep_device=False):
    """logLik over singlet + donor-pair columns, softmax with the doublet prior, LLR -- all on the device.
    Returns (loglik or None, posterior over all K2 columns, LLR) as host arrays, or the device tensors (posterior, LLR)
    when ``keep_device`` (the sharded wrapper gathers them across ranks before the download)."""
    dev = counts.device
    lib = _lib.load()
    V, K, G = GT_prob.shape
    K2 = K + K * (K - 1) // 2
    lp = np.asarray(log_prior_both, dtype=np.float64)
    lp = _compress_rows(lp)
    gt, mu, sm, lpd = _dev(GT_prob, dev), _dev(beta_mu, dev), _dev(beta_sum, dev), _dev(lp, dev)
    ws = _lib.DoubletWs()
    _lib.check(lib.vb_doublet_ws_sizes(counts.handle, K, G, int(bool(ase)), _stream(dev), C.byref(ws)))
    W, heavy, ab2 = _zeros(ws.W, dev), _zeros(ws.heavy, dev), _zeros(ws.ab2, dev)
    ll, pr, llr = _zeros(counts.n_cell * K2, dev), _zeros(counts.n_cell * K2, dev), _zeros(counts.n_cell, dev)
    _lib.check(lib.vb_vireo_doublet(counts.handle, K, G, int(bool(ase)), _ptr(gt), _ptr(mu), _ptr(sm),
                                    _ptr(lpd), lp.shape[0], _ptr(W), _ptr(heavy), _ptr(ab2), C.byref(ws),
                                    _ptr(ll), _ptr(pr), _ptr(llr), _stream(dev)))
    C_ = counts.n_cell
    if keep_device:
        return pr, llr
    # the log-likelihoods stay on the device (109 MB at cfg3; no caller reads them): posterior and LLR only
    return (None if not want_loglik else _to_host(ll).reshape(C_, K2), _to_host(pr).reshape(C_, K2), _to_host(llr))
