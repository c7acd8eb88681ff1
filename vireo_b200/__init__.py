"""vireo_b200 -- the variational-EM inner loop of vireoSNP on NVIDIA B200 (sm_100a).

Drop-in for the hot path of ``vireoSNP`` v0.5.9: same names, signatures and result conventions as
``vireoSNP.Vireo``, ``vireoSNP.vireo_wrap`` (alias ``vireo_flock``) and ``vireoSNP.BinomMixtureVB``
(reference vireoSNP/__init__.py:12-15), with the per-iteration work in hand-written CUDA behind a
C ABI (include/vireo_b200.h).  Importing the package never needs a GPU; calling a compute entry point
without one raises (there is no CPU fallback).
"""
from .version import __version__

from . import dist
from . import vireo_base as base
from . import vireo_model as model
from ._engine import StagedCounts, VireoB200Error, clear_cache, stage
from .bmm_model import BinomMixtureVB
from .vireo_base import get_binom_coeff, loglik_amplify, match, normalize, optimal_match
from .sharded import fit_cell_sharded, predict_doublet_sharded
from .vireo_doublet import predict_doublet
from .vireo_model import Vireo
from .vireo_wrap import vireo_flock, vireo_wrap

__all__ = ["__version__", "Vireo", "vireo_wrap", "vireo_flock", "BinomMixtureVB", "predict_doublet",
           "fit_cell_sharded", "predict_doublet_sharded", "dist", "StagedCounts", "stage", "clear_cache", "VireoB200Error", "normalize", "loglik_amplify",
           "get_binom_coeff", "match", "optimal_match", "base", "model"]
