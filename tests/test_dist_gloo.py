"""Restart sharding across ranks (vireo_b200/dist.py) on CPU: world_size 2, gloo backend.

The N>1 path of the product has no data-path collective: every rank fits restarts i % world == rank, then ONE
all-gather of final ELBOs and one broadcast of the winner's state (reference: multiprocessing.Pool over restarts,
vireoSNP/utils/vireo_wrap.py:74-92).  These tests run that host logic with two real processes.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_init, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vireo_b200 import dist as vd
    # sharding is opt-in: an initialised process group alone changes nothing (one sample per rank under
    # torchrun must not be mistaken for restart sharding)
    assert vd.world() == (0, 1) and vd.shard_restarts(3) == [0, 1, 2]
    vd.enable(set_device=False)
    assert vd.world() == (rank, world)
    # every rank must have been handed the same problem; a mismatch raises on every rank instead of selecting a
    # winner across different data
    vd.check_same_problem(100, 50, 1234, n_init, 7)
    try:
        vd.check_same_problem(100, 50, 1234 + rank, n_init, 7)
        mismatch = "not raised"
    except RuntimeError as exc:
        mismatch = "raised" if "different problem" in str(exc) else str(exc)
    assert mismatch == "raised", mismatch
    mine = vd.shard_restarts(n_init)
    assert mine == [i for i in range(n_init) if i % world == rank]
    # every rank "fits" its own restarts: ELBO and state are functions of the restart index
    rng = np.random.RandomState(123)
    elbo_all = rng.randn(n_init) * 100 - 5000
    final = np.full(n_init, np.nan)
    results = {}
    for i in mine:
        final[i] = elbo_all[i]
        results[i] = {"ID_prob": np.full((5, 3), float(i)), "GT_prob": np.arange(24, dtype=np.float64).reshape(4, 2, 3) + i,
                      "beta_mu": np.array([[0.1, 0.5, 0.9]]) * (i + 1), "n_iter": np.int64(10 + i)}
    got_final, best, state = vd.gather_restarts(final, results, ["ID_prob", "GT_prob", "beta_mu", "n_iter"])
    np.save(os.path.join(out_dir, "final_%d.npy" % rank), got_final)
    np.save(os.path.join(out_dir, "best_%d.npy" % rank), np.array([best]))
    np.savez(os.path.join(out_dir, "state_%d.npz" % rank), **{k: np.asarray(v) for k, v in state.items()})
    np.save(os.path.join(out_dir, "expect_%d.npy" % rank), elbo_all)
    vd.disable()
    assert vd.world() == (0, 1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_init", [1, 5, 8])
def test_restart_sharding_world2(tmp_path, n_init):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_init, str(tmp_path)), nprocs=world, join=True)
    expect = np.load(tmp_path / "expect_0.npy")
    best = int(np.argmax(expect))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("final_%d.npy" % r)), expect)      # one all-gather, every rank complete
        assert int(np.load(tmp_path / ("best_%d.npy" % r))[0]) == best               # identical model selection
        st = np.load(tmp_path / ("state_%d.npz" % r))
        assert np.array_equal(st["ID_prob"], np.full((5, 3), float(best)))           # winner's state on every rank
        assert np.array_equal(st["GT_prob"], np.arange(24, dtype=np.float64).reshape(4, 2, 3) + best)
        assert np.allclose(st["beta_mu"], np.array([[0.1, 0.5, 0.9]]) * (best + 1))
        assert int(st["n_iter"]) == 10 + best


def test_single_process_degenerates():
    sys.path.insert(0, ROOT)
    from vireo_b200 import dist as vd
    assert vd.world() == (0, 1)
    assert vd.shard_restarts(4) == [0, 1, 2, 3]
    final = np.array([1.0, 3.0, 2.0])
    out_final, best, state = vd.gather_restarts(final, {1: {"x": np.ones(2)}}, ["x"])
    assert np.array_equal(out_final, final) and best == 1 and np.array_equal(state["x"], np.ones(2))


def test_cell_shard_bounds():
    """Host logic of the cell-sharded fit (vireo_b200/sharded.py): nnz-balanced contiguous shards, identical on every
    rank (a pure function of the index pointer)."""
    sys.path.insert(0, ROOT)
    from vireo_b200.sharded import cell_shards
    rng = np.random.RandomState(0)
    nnz = rng.randint(0, 50, size=1000)
    indptr = np.concatenate([[0], np.cumsum(nnz)])
    for n in (1, 2, 3, 8):
        b = cell_shards(indptr, n)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == n + 1 and (np.diff(b) >= 0).all()
        per = np.diff(indptr[b])
        assert per.sum() == indptr[-1] and per.max() - per.min() <= 2 * nnz.max()
    assert list(cell_shards(np.zeros(5, dtype=np.int64), 2)) in ([0, 0, 4], [0, 4, 4])     # no reads at all
