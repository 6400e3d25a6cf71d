"""Pins the CPU oracle to the reference's own published results.

The reference has no tests; what pins this path is (a) the clean/noisy pair it ships under
``Document/Presentation`` and (b) the per-file rows of ``Code/results_summary/*/all_results.json``.
``tests/golden/make_golden.py`` extracted both into ``tests/golden``.
"""
import numpy as np
import pytest

import oracle
from oracle.search import score_candidate
from tests.golden_util import load_p257_090, published_rows

STOI_TOL = 5e-7     # observed 1.1e-7 with the soxr-HQ-like resampler design (see golden_util); 1.2e-5 with scipy's default
SNR_TOL = 2e-4      # dB (observed 3e-5)


@pytest.fixture(scope="module")
def pair():
    return load_p257_090()


def test_baseline_row(pair):
    c, n = pair
    row = published_rows("p257_090")[0]
    assert len(c) == 30176
    assert abs(oracle.stoi(c, n, 16000) - row["stoi_noisy"]) < STOI_TOL
    assert abs(oracle.global_snr(c, n) - row["snr_noisy"]) < SNR_TOL


ROWS = published_rows("p257_090")


@pytest.mark.parametrize("row", ROWS, ids=[f"run{r['run']}-{r['alg']}-{r['criterion']}" for r in ROWS])
def test_published_row(pair, row):
    c, n = pair
    fn = oracle.ALGORITHMS[row["alg"]]
    p = row["params"]
    kw = {"clean_audio": c} if p["noise_method"] == "true_noise" else {}
    sc = score_candidate(c, fn(n, 16000, **kw, **p), 16000)
    assert abs(sc["stoi"] - row["stoi"]) < STOI_TOL
    assert abs(sc["snr"] - row["snr"]) < SNR_TOL


ROWS_135 = published_rows("p257_135")


@pytest.mark.parametrize("row", ROWS_135, ids=[f"p257_135-run{r['run']}-{r['alg']}-{r['criterion']}" for r in ROWS_135])
def test_published_row_second_stem(row):
    """The reference's other shipped pair (``Document/Presentation/wiener_p257_135``): its reproducible rows.

    ``reproducible`` in ``published_rows.json`` was decided by ``make_golden.py`` with this oracle; what keeps that
    from being circular is SURVEY.md 8c's independent scratch restatement, which reproduces the same 22 of 38 rows
    (the other 16 come from older code / grids, see ``test_q_semantics_of_stale_row``), and the fact that a row
    matching to 5e-7 STOI and 2e-4 dB by accident is not a plausible failure mode."""
    from tests.golden_util import load_pair
    c, n = load_pair("p257_135")
    p = row["params"]
    kw = {"clean_audio": c} if p["noise_method"] == "true_noise" else {}
    sc = score_candidate(c, oracle.ALGORITHMS[row["alg"]](n, 16000, **kw, **p), 16000)
    assert abs(oracle.stoi(c, n, 16000) - row["stoi_noisy"]) < STOI_TOL
    assert abs(sc["stoi"] - row["stoi"]) < STOI_TOL
    assert abs(sc["snr"] - row["snr"]) < SNR_TOL


def test_rows_cover_all_algorithms_and_methods():
    assert len(ROWS) == 17 and len(ROWS_135) == 5
    algs = {r["alg"] for r in ROWS}
    methods = {r["params"]["noise_method"] for r in ROWS}
    shapes = {(r["params"]["n_fft"], r["params"]["hop_length"]) for r in ROWS}
    assert algs == {"spectralSubtractor", "mmse", "wiener", "omlsa"}
    assert methods == {"percentile", "min_tracking", "true_noise"}
    assert shapes == {(512, 128), (512, 256), (1024, 128), (1024, 256)}


def test_shipped_winner_waveforms(pair):
    """The three SS outputs shipped for p257_090 are run 29's winners; PCM16 + resampler residual."""
    import os
    from tests.golden_util import GOLDEN
    c, n = pair
    wavs = np.load(os.path.join(GOLDEN, "p257_090_winner_wavs.npz"))
    rows = {r["criterion"]: r for r in published_rows("p257_090") if r["run"] == 29
            and r["alg"] == "spectralSubtractor"}
    checked = 0
    for crit, row in rows.items():
        sc = score_candidate(c, oracle.spectral_subtraction(n, 16000, clean_audio=c, **row["params"]), 16000)
        w = wavs[crit].astype(np.float64) / 32768.0
        m = min(len(w), len(sc["enhanced"]))
        rel = np.sqrt(np.mean((w[:m] - sc["enhanced"][:m]) ** 2) / np.mean(w[:m] ** 2))
        assert rel < 0.01
        checked += 1
    assert checked >= 2


def test_q_semantics_of_stale_row(pair):
    """Run 22's OMLSA STOI winner (published at q=0.3) was produced when q meant the absence
    prior; the committed code reproduces it at q=0.7 - evidence that non-reproducible rows are
    stale results, not oracle errors."""
    c, n = pair
    row = [r for r in published_rows("p257_090", reproducible=False)
           if r["run"] == 22 and r["alg"] == "omlsa" and r["criterion"] == "stoi"][0]
    p = dict(row["params"])
    p["q"] = round(1.0 - p["q"], 6)
    sc = score_candidate(c, oracle.advanced_mmse(n, 16000, **p), 16000)
    assert abs(sc["stoi"] - row["stoi"]) < STOI_TOL
    assert abs(sc["snr"] - row["snr"]) < SNR_TOL
