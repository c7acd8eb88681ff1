"""Multi-GPU restart sharding: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
B200 box, gloo in the CPU tests) for the plumbing.

The restarts of ``vireo_wrap`` / ``BinomMixtureVB.fit`` never interact until model selection
(reference vireoSNP/utils/vireo_wrap.py:85-92, bmm_model.py:242-254; the reference parallelises them
with ``multiprocessing.Pool``, vireo_wrap.py:74-83).  So: every rank draws ALL initial states from the
numpy RNG in the reference's order and keeps restarts ``i % world == rank``; after the warm-up fits ONE
all-gather exchanges the final ELBOs, every rank takes the same argmax, and the owner broadcasts the
winner's state so that all ranks continue (and return) identically.  No collective touches the EM data
path.  Without an initialised process group everything degenerates to a single rank.
"""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_restarts(n_init):
    """Indices of the restarts this rank fits (round-robin)."""
    rank, ws = world()
    return [i for i in range(n_init) if i % ws == rank]


def _comm_device(device):
    import torch
    d = _dist()
    if d and d.get_backend() == "nccl":
        return torch.device("cuda", device if device is not None else torch.cuda.current_device())
    return torch.device("cpu")


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
    d.all_gather(recv, send)
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
    d.broadcast_object_list(meta, src=owner)
    dev = _comm_device(device)
    out = {}
    for k, shape, dtype in meta[0]:
        if rank == owner:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(state[k]).reshape(-1))).to(dev)
        else:
            t = torch.empty(int(np.prod(shape)) if len(shape) else 1, dtype=getattr(torch, dtype), device=dev)
        d.broadcast(t, src=owner)
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
