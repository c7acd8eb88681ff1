import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, so a plain `pytest tests/` works anywhere."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def csc_from(npz, prefix):
    from scipy.sparse import csc_matrix
    return csc_matrix((npz[prefix + "_data"], npz[prefix + "_indices"], npz[prefix + "_indptr"]),
                      shape=tuple(npz[prefix + "_shape"]))


@pytest.fixture(scope="session")
def cellsnp():
    z = load_golden("fixture_cellsnp")
    return csc_from(z, "AD"), csc_from(z, "DP")


@pytest.fixture(scope="session")
def mito():
    z = load_golden("fixture_mito")
    return csc_from(z, "AD"), csc_from(z, "DP")


@pytest.fixture(scope="session")
def small():
    z = load_golden("vireo_small_default")
    return csc_from(z, "AD"), csc_from(z, "DP")


@pytest.fixture(scope="session")
def edge():
    z = load_golden("vireo_edge")
    return csc_from(z, "AD"), csc_from(z, "DP")


def rel_close(x, ref, rtol, what=""):
    """|x - ref| <= rtol*|ref| + 1e-300 elementwise (the north-star parity gate)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (what, x.shape, ref.shape)
    err = np.abs(x - ref)
    bound = rtol * np.abs(ref) + 1e-300
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err / bound), err.shape)
        raise AssertionError("%s: %d/%d entries off; worst at %s: got %r want %r (rel %.3e > %.1e)" % (
            what, bad.sum(), bad.size, i, x[i], ref[i], err[i] / max(abs(ref[i]), 1e-300), rtol))
