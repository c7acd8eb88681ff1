"""CPU-side checks of bench.py's bookkeeping: the algorithmic-byte accounting behind `roofline`, the
bounded CPU sample of the reference arm, and the committed DRAM-traffic file `roofline.traffic` is read from."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey_formulas():
    w = bench.WORKLOADS["cfg3"]
    N = 99996203
    alg = bench.algorithmic_bytes(w, N)
    C, V, K, G = w["C"], w["V"], w["K"], 3
    assert alg["iter"] == 2 * N * 8 + 16 * C * K + 24 * V * K * G + 32 * V * K + 4 * (C + V + 2)
    assert alg["k_cell"] == N * 8 + 16 * V * K + 8 * C * K + 4 * (C + 1)           # B_ID
    assert abs(alg["iter"] / 1e9 - 1.709) < 0.002 and abs(alg["k_cell"] / 1e9 - 0.826) < 0.002
    wide = bench.algorithmic_bytes(w, N, wide=True)
    assert wide["k_cell"] - alg["k_cell"] == 4 * N                                  # 12-byte records


def test_cpu_sample_is_bounded_and_scaled():
    from scipy.sparse import random as sprand
    w = dict(C=4000, V=300, K=16)
    DP = sprand(300, 4000, density=0.05, format="csc", random_state=0)
    AD = DP.copy()
    ADs, DPs, frac, n_cells = bench.cpu_sample(AD, DP, w, budget_s=1e-4, n_steps=3, workers=1)
    assert 500 <= n_cells < 4000 and DPs.shape == (300, n_cells) and 0 < frac < 1
    assert abs(frac - DPs.nnz / DP.nnz) < 1e-12
    ADf, DPf, frac, n_cells = bench.cpu_sample(AD, DP, w, budget_s=1e9, n_steps=1, workers=1)
    assert frac == 1.0 and n_cells == 4000 and DPf is DP


def test_committed_traffic_file_has_both_passes():
    d = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    for key in ("cfg3", "cfg3_fixed32"):
        assert set(d[key]) == {"k_snp", "k_cell"}
        # DRAM bytes per pass stay below the algorithmic bytes: the record format is denser than 8 B per nnz
        assert all(1e8 < v < 0.826e9 for v in d[key].values())
