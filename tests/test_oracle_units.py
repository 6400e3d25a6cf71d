"""Unit checks of the oracle's third-party restatements and host-side selection logic."""
import importlib.util
import os

import numpy as np
import pytest

import oracle
from oracle import intelligibility as ostoi
from oracle.search import grid_points, select_best
from classical_speech_enhancement_b200 import parameter_ranges as pr
from classical_speech_enhancement_b200.synth import make_pair


def test_grids_match_reference_file_when_present():
    path = "/root/reference/Code/parameter_ranges.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location("ref_parameter_ranges", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name in ("param_ranges_ss", "param_ranges_mmse", "param_ranges_wiener", "param_ranges_omlsa"):
        a, b = getattr(pr, name), getattr(ref, name)
        assert list(a.keys()) == list(b.keys())
        assert a == b


def test_grid_sizes_and_order():
    sizes = [len(grid_points(g)) for g in (pr.param_ranges_ss, pr.param_ranges_mmse,
                                           pr.param_ranges_wiener, pr.param_ranges_omlsa)]
    assert sizes == [720, 1920, 192, 6912]
    pts = grid_points(pr.param_ranges_wiener)
    assert pts[0]["noise_method"] == "percentile" and pts[1]["noise_method"] == "min_tracking"
    assert pts[2]["noise_percentile"] == 20.0 and pts[0]["alpha"] == pts[63]["alpha"] == 0.90


def test_stft_istft_roundtrip_and_shapes():
    rng = np.random.default_rng(0)
    for L, n_fft, hop in ((4000, 512, 128), (4097, 1024, 256), (3000, 256, 64)):
        y = rng.standard_normal(L)
        S = oracle.stft(y, n_fft, hop)
        assert S.shape == (n_fft // 2 + 1, 1 + L // hop)
        assert np.max(np.abs(oracle.istft(S, hop, L) - y)) < 1e-12


def test_resampler_closed_form():
    """scipy's resample_poly with the Octave window equals
    y10[m] = sum_j x[j] * h5[8m + 290 - 5j] - the form the kernel uses."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(2003)
    h, p, q = ostoi.resample_window(10000, 16000)
    assert (p, q, len(h)) == (5, 8, 581)
    h5 = 5 * h / np.sum(h)
    ref = ostoi.resample_to_10k(x, 16000)
    n_out = -(-len(x) * 5 // 8)
    assert len(ref) == n_out
    got = np.zeros(n_out)
    for m in range(n_out):
        j = np.arange(len(x))
        k = 8 * m + 290 - 5 * j
        ok = (k >= 0) & (k <= 580)
        got[m] = np.sum(x[j[ok]] * h5[k[ok]])
    assert np.max(np.abs(got - ref)) < 1e-13


def test_band_edges():
    edges, _ = ostoi.third_octave_bands()
    assert edges == [(7, 9), (9, 11), (11, 14), (14, 17), (17, 22), (22, 27), (27, 34), (34, 43),
                     (43, 55), (55, 69), (69, 87), (87, 109), (109, 138), (138, 174), (174, 219)]


def test_stoi_identity_and_short():
    c, n = make_pair(3, 32000)
    assert abs(oracle.stoi(c, c, 16000) - 1.0) < 1e-9
    assert 0.3 < oracle.stoi(c, n, 16000) < 1.0
    assert oracle.stoi(c[:3000], n[:3000], 16000) == 1e-5


def test_selection_hysteresis_is_not_argmax():
    pts = [{"i": i} for i in range(4)]
    sc = [{"stoi": 0.5, "pesq": 2.0, "snr": 1.0}, {"stoi": 0.5000008, "pesq": 2.00005, "snr": 2.0},
          None, {"stoi": 0.5000009, "pesq": 2.00008, "snr": 3.0}]
    best = select_best(pts, sc)
    assert best["stoi"]["index"] == 0 and best["pesq"]["index"] == 0 and best["balance"]["index"] == 0
    sc[3] = {"stoi": 0.5000011, "pesq": 2.0011, "snr": 3.0}
    best = select_best(pts, sc)
    assert best["stoi"]["index"] == 3 and best["pesq"]["index"] == 3
    assert select_best(pts, [None] * 4)["stoi"]["index"] is None


def test_dead_parameters_give_identical_outputs():
    c, n = make_pair(2, 12000)
    a = oracle.wiener_filter(n, 16000, 512, 128, 0.95, 0.05, 10.0, "min_tracking")
    b = oracle.wiener_filter(n, 16000, 512, 128, 0.95, 0.05, 20.0, "min_tracking")
    assert np.array_equal(a, b)
    kw = dict(n_fft=512, hop_length=128, alpha=0.9, ksi_min=0.01, q=0.4, gain_floor=0.1,
              noise_percentile=10.0, noise_method="percentile")
    assert np.array_equal(oracle.advanced_mmse(n, 16000, noise_mu=0.92, **kw),
                          oracle.advanced_mmse(n, 16000, noise_mu=0.98, **kw))


def test_alignment_lag_detects_shift():
    c, n = make_pair(5, 40000)
    assert oracle.alignment_lag(c, n, 16000) == 0
    shifted = np.concatenate([np.zeros(37), n])[:len(n)]       # delayed by 37 -> lag -37
    assert oracle.alignment_lag(c, shifted, 16000) == -37
    assert oracle.alignment_lag(c[:200], n[:200], 16000) is None


def test_vectorised_selection_equals_the_sequential_scan():
    """sweep.select_all scans all utterances at once; it must reproduce grid.select_best (the reference's
    hysteresis scan, speech_enhancement_comparison.py:186-216) exactly, including ties, near-ties below the
    tolerances, invalid candidates and skipped (PESQ None) candidates."""
    import numpy as np
    from classical_speech_enhancement_b200 import parameter_ranges as pr
    from classical_speech_enhancement_b200.grid import grid_points as gp, select_best as sb, select_best_batch
    pts = gp(pr.param_ranges_wiener)
    rng = np.random.default_rng(7)
    U, P = 12, len(pts)
    stoi = np.round(rng.uniform(0.5, 0.9, (U, P)), 2).astype(np.float32) + (rng.integers(0, 3, (U, P)) * 5e-7).astype(np.float32)
    snr = rng.normal(5, 3, (U, P)).astype(np.float32)
    snr[0, 3] = np.inf
    valid = rng.random((U, P)) > 0.1
    valid[1] = False                                                     # an utterance without any valid candidate
    pesq = [[None if rng.random() < 0.05 else float(np.round(rng.uniform(1, 3), 2)) for _ in range(P)] for _ in range(U)]
    pq = np.array([[np.nan if v is None else v for v in row] for row in pesq])
    batch = select_best_batch(pts, stoi.astype(np.float64), pq, snr.astype(np.float64), valid)
    for u in range(U):
        assert batch[u] == sb(pts, [float(v) for v in stoi[u]], pesq[u], [float(v) for v in snr[u]], valid[u])
    assert batch[1]["stoi"]["index"] is None
    nop = select_best_batch(pts, stoi.astype(np.float64), None, snr.astype(np.float64), valid)
    for u in range(U):
        assert nop[u] == sb(pts, [float(v) for v in stoi[u]], [0.0] * P, [float(v) for v in snr[u]], valid[u])
