"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for the rendezvous (NCCL over NVLink on the
B200 box, gloo in the CPU tests), and an NCCL communicator of the library's own (``vb_comm``) for the collectives
the fits enqueue themselves.

Two things are spread over the ranks:

* **restarts** -- the ``n_init`` warm-up fits of ``vireo_wrap`` / ``BinomMixtureVB.fit`` never interact until model
  selection (reference vireoSNP/utils/vireo_wrap.py:85-92, bmm_model.py:242-254; the reference parallelises them with
  ``multiprocessing.Pool``, vireo_wrap.py:74-83).  Every rank draws ALL initial states from the numpy RNG in the
  reference's order and keeps restarts ``i % world == rank``; ONE all-gather exchanges the final ELBOs, every rank takes
  the same argmax, and the owner broadcasts the winner's state.
* **cells** of the one fit restart sharding cannot spread (the final fit, ``sharded.py``).

Sharding is **opt-in**: nothing here looks at ``torch.distributed`` until ``enable()`` has been called (or the
environment says ``VIREO_B200_DIST=1``) -- a process group that exists for other reasons, e.g. one sample per rank
under torchrun, is left alone.  ``enable()`` also pins the calling process to ``cuda:LOCAL_RANK`` unless told not to.
Before any model selection the ranks compare a fingerprint of the problem they were given (shapes, nnz, n_init, seed,
checksum of the counts) and raise if they differ, instead of selecting across different data or hanging.
"""
import ctypes as C
import os

import numpy as np

_GROUP = {"on": os.environ.get("VIREO_B200_DIST", "0") == "1", "group": None}
_COMMS = {}


def enable(group=None, set_device=True):
    """Shard restarts and the final fit over the ranks of ``group`` (default: the default process group, which the
    caller must have initialised).  With ``set_device`` the process is pinned to ``cuda:LOCAL_RANK`` first."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("vireo_b200.dist.enable(): torch.distributed is not initialised")
    _GROUP["on"], _GROUP["group"] = True, group
    if set_device:
        import torch
        if torch.cuda.is_available() and "LOCAL_RANK" in os.environ:
            torch.cuda.set_device(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())


def disable():
    _GROUP["on"], _GROUP["group"] = False, None
    for c in _COMMS.values():
        c.close()
    _COMMS.clear()


def _dist():
    if not _GROUP["on"]:
        return None
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    d = _dist()
    return (d.get_rank(_GROUP["group"]), d.get_world_size(_GROUP["group"])) if d else (0, 1)


def _src(rank_in_group):
    """global rank of a group rank (torch's broadcast takes global ranks)"""
    d = _dist()
    g = _GROUP["group"]
    if g is None:
        return rank_in_group
    return d.get_global_rank(g, rank_in_group)


def shard_restarts(n_init):
    """Indices of the restarts this rank fits (round-robin)."""
    rank, ws = world()
    return [i for i in range(n_init) if i % ws == rank]


def _comm_device(device):
    import torch
    d = _dist()
    if d and d.get_backend(_GROUP["group"]) == "nccl":
        return torch.device("cuda", device if device is not None else torch.cuda.current_device())
    return torch.device("cpu")


def check_same_problem(*items, device=None):
    """Raise on every rank unless all ranks pass the same ``items`` (numbers: shapes, nnz, n_init, seed, checksums).
    One all-gather of a few doubles; a no-op on a single rank."""
    d = _dist()
    if d is None:
        return
    import torch
    rank, ws = world()
    mine = np.array([float(x if x is not None else -1.0) for x in items], dtype=np.float64)
    dev = _comm_device(device)
    recv = [torch.empty(mine.size, dtype=torch.float64, device=dev) for _ in range(ws)]
    d.all_gather(recv, torch.from_numpy(mine).to(dev), group=_GROUP["group"])
    got = np.stack([r.cpu().numpy() for r in recv])
    if not (got == got[0]).all():
        bad = [r for r in range(ws) if not (got[r] == got[0]).all()]
        raise RuntimeError("vireo_b200.dist: ranks %s were given a different problem than rank 0 (shapes / nnz / n_init / "
                           "seed / data checksum differ: %s vs %s); restart sharding needs every rank to pass the same "
                           "matrices and arguments -- call vireo_b200.dist.disable() to fit independent problems"
                           % (bad, got[bad[0]].tolist(), got[0].tolist()))


def allgather_elbo(final, device=None):
    """``final`` holds this rank's values at its own restart indices (anything elsewhere); returns the
    complete vector on every rank via one all_gather of ceil(n / world) doubles per rank."""
    d = _dist()
    if d is None:
        return np.asarray(final, dtype=np.float64)
    import torch
    rank, ws = world()
    n = len(final)
    per = (n + ws - 1) // ws
    mine = np.full(per, -np.inf)
    own = np.asarray(final, dtype=np.float64)[rank::ws]
    mine[:own.size] = own
    dev = _comm_device(device)
    send = torch.from_numpy(mine).to(dev)
    recv = [torch.empty(per, dtype=torch.float64, device=dev) for _ in range(ws)]
    d.all_gather(recv, send, group=_GROUP["group"])
    out = np.empty(n)
    for r in range(ws):
        vals = recv[r].cpu().numpy()
        idx = np.arange(r, n, ws)
        out[idx] = vals[:idx.size]
    return out


def broadcast_state(state, keys, owner, device=None):
    """Broadcast the dict of numpy arrays / scalars ``state`` (only meaningful on ``owner``) to every rank."""
    d = _dist()
    if d is None:
        return state
    import torch
    rank, _ = world()
    meta = [None]
    if rank == owner:
        meta = [[(k, np.asarray(state[k]).shape, str(np.asarray(state[k]).dtype)) for k in keys]]
    d.broadcast_object_list(meta, src=_src(owner), group=_GROUP["group"])
    dev = _comm_device(device)
    out = {}
    for k, shape, dtype in meta[0]:
        if rank == owner:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(state[k]).reshape(-1))).to(dev)
        else:
            t = torch.empty(int(np.prod(shape)) if len(shape) else 1, dtype=getattr(torch, dtype), device=dev)
        d.broadcast(t, src=_src(owner), group=_GROUP["group"])
        arr = t.cpu().numpy().reshape(shape)
        out[k] = arr if arr.ndim else arr[()]
    return out


def gather_restarts(final, results, keys, device=None):
    """Model selection across ranks.  ``results[i]`` holds restart i's state on its owner.  Returns
    (all final ELBOs, index of the best restart, the best restart's state on every rank)."""
    rank, ws = world()
    final = allgather_elbo(final, device)
    best = int(np.argmax(final))
    owner = best % ws
    state = results.get(best) if rank == owner else None
    state = broadcast_state(state, keys, owner, device)
    return final, best, state


class Comm:
    """``vb_comm`` of the library (an NCCL communicator bound at run time) over the ranks of the enabled group;
    on a single rank a communicator without NCCL.  The 128-byte NCCL id travels through ``torch.distributed``."""

    def __init__(self, device):
        from . import _lib
        lib = _lib.load()
        d = _dist()
        rank, ws = world()
        ident = C.create_string_buffer(_lib.COMM_ID_BYTES)
        if ws > 1:
            box = [None]
            if rank == 0:
                _lib.check(lib.vb_comm_unique_id(ident))
                box = [ident.raw]
            d.broadcast_object_list(box, src=_src(0), group=_GROUP["group"])
            ident = C.create_string_buffer(box[0], _lib.COMM_ID_BYTES)
        handle = C.c_void_p()
        _lib.check(lib.vb_comm_create(int(device), ws, rank, ident, C.byref(handle)))
        self.handle, self.device, self.rank, self.world = handle, int(device), rank, ws

    def close(self):
        if self.handle is not None:
            from . import _lib
            _lib.load().vb_comm_destroy(self.handle)
            self.handle = None


def comm(device):
    """The (cached) library communicator of this process for ``device``."""
    rank, ws = world()
    key = (int(device), rank, ws)
    c = _COMMS.get(key)
    if c is None or c.handle is None:
        c = Comm(device)
        _COMMS[key] = c
    return c
