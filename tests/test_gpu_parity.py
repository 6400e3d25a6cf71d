"""Parity of the sm_100a build against the oracle, through the C ABI, on a real GPU.

Tolerances are the north_star's: enhanced waveforms within 1e-4 of the oracle relative to the
waveform peak (fp32), STOI within 1e-4, identical grid argmax per algorithm.  Inputs are rounded
to float32 first and the SAME rounded values are given to the float64 oracle, so the comparison
measures the arithmetic, not the input quantisation.
"""
import numpy as np
import pytest

import oracle
from oracle.noise import noise_psd
from oracle.search import score_candidate
from classical_speech_enhancement_b200 import grid
from classical_speech_enhancement_b200 import parameter_ranges as pr
from classical_speech_enhancement_b200.synth import make_batch, make_pair
from tests.golden_util import load_p257_090, published_rows

pytestmark = pytest.mark.gpu
TOL_WAVE = 1e-4
TOL_STOI = 1e-4
TOL_SNR_DB = 1e-3


def f32(x):
    return np.asarray(x).astype(np.float32).astype(np.float64)


@pytest.fixture(scope="module")
def lib():
    from classical_speech_enhancement_b200 import _lib
    return _lib.load()


def engine_for(clean, noisy, **kw):
    from classical_speech_enhancement_b200.engine import SweepEngine
    return SweepEngine(np.atleast_2d(clean), np.atleast_2d(noisy), **kw)


def test_library_is_the_cuda_build(lib):
    import torch
    assert torch.cuda.is_available()
    assert lib.path.endswith("libcse_sm100a.so") and lib.real_bits == 32
    assert torch.cuda.get_device_capability()[0] >= 10


@pytest.mark.parametrize("n_fft,hop", [(512, 128), (512, 256), (1024, 128), (1024, 256), (256, 64), (2048, 512)])
def test_stft_and_noise_psds(n_fft, hop):
    c, n = make_pair(11, 48000)
    c, n = f32(c), f32(n)
    eng = engine_for(c, n)
    nb = n_fft // 2 + 1
    Y = eng.be.to_host(eng.stft(n_fft, hop))
    ref = oracle.stft(n, n_fft, hop)
    got = (Y[0, :, :nb, 0] + 1j * Y[0, :, :nb, 1]).T
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-6
    for method, pct in (("percentile", 10.0), ("percentile", 20.0), ("min_tracking", 10.0), ("true_noise", 10.0)):
        for eps in (1e-10, 1e-12):
            N = eng.noise_psd_host(method, n_fft, hop, pct, eps)[0]
            Nref = noise_psd(n, method, n_fft, hop, percentile=pct, clean=c, eps=eps)
            assert N.shape == Nref.shape
            assert np.abs(N - Nref).max() / Nref.max() < 5e-6, (method, pct, eps)


def _oracle_fn(name):
    return oracle.ALGORITHMS[name]


CONFIG_POINTS = {
    "spectralSubtractor": [dict(alpha=1.0, beta=0.001), dict(alpha=5.0, beta=0.15), dict(alpha=0.5, beta=0.05)],
    "wiener": [dict(alpha=0.90, gain_floor=0.01), dict(alpha=0.98, gain_floor=0.1)],
    "mmse": [dict(alpha=0.98, ksi_min=0.0001, gain_min=0.001, gain_max=1.0), dict(alpha=0.90, ksi_min=0.15, gain_min=0.2, gain_max=1.0)],
    "omlsa": [dict(alpha=0.7, ksi_min=0.001, gain_floor=0.05, noise_mu=0.92, q=0.3),
              dict(alpha=0.95, ksi_min=0.05, gain_floor=0.2, noise_mu=0.98, q=0.5)],
}


@pytest.mark.parametrize("alg", list(CONFIG_POINTS))
@pytest.mark.parametrize("method", ["percentile", "min_tracking", "true_noise"])
def test_enhanced_waveforms_all_shapes(alg, method):
    """Configs 1-4 of BASELINE.json at full length (3 s): waveform, STOI, SNR, lag per candidate."""
    c, n = make_pair(21, 48000)
    c, n = f32(c), f32(n)
    eng = engine_for(c, n)
    pts = []
    for n_fft, hop in ((512, 128), (512, 256), (1024, 128), (1024, 256)):
        for base in CONFIG_POINTS[alg]:
            pts.append(dict(base, n_fft=n_fft, hop_length=hop, noise_percentile=20.0, noise_method=method))
    wav = eng.enhance(alg, pts)[0]
    sc = eng.sweep(alg, pts)[0]
    for i, p in enumerate(pts):
        kw = {"clean_audio": c} if method == "true_noise" else {}
        ref = _oracle_fn(alg)(n, 16000, **kw, **p)
        assert np.abs(wav[i] - ref).max() / np.abs(ref).max() < TOL_WAVE, p
        rs = score_candidate(c, ref, 16000)
        assert sc[i]["flags"] & 1 and sc[i]["lag"] == oracle.alignment_lag(c, ref, 16000)
        assert abs(sc[i]["stoi"] - rs["stoi"]) < TOL_STOI, p
        assert abs(sc[i]["snr"] - rs["snr"]) < TOL_SNR_DB, p


def test_config2_wiener_percentile_sweep_argmax():
    """BASELINE config 2: Wiener grid, percentile noise, 1 pair: scores and selection vs the oracle."""
    from classical_speech_enhancement_b200.sweep import select_all
    c, n = make_pair(0, 48000)
    c, n = f32(c), f32(n)
    ranges = dict(pr.param_ranges_wiener, noise_method=["percentile"])
    pts, scores, best = oracle.sweep_one_pair(c, n, 16000, oracle.wiener_filter, ranges)
    assert len(pts) == 96
    eng = engine_for(c, n)
    sc = eng.sweep("wiener", pts)
    ref_stoi = np.array([s["stoi"] for s in scores])
    ref_snr = np.array([s["snr"] for s in scores])
    assert np.abs(sc[0]["stoi"] - ref_stoi).max() < TOL_STOI
    assert np.abs(sc[0]["snr"] - ref_snr).max() < TOL_SNR_DB
    sel = select_all({"wiener": sc}, {"wiener": pts})["wiener"][0]
    assert sel["stoi"]["index"] == best["stoi"]["index"]


def test_config3_mmse_mintracking_batch():
    """BASELINE config 3 (reduced): MMSE min_tracking slice over a batch; batch == singles bit for bit."""
    clean, noisy = make_batch(4, 48000, first=40)
    clean, noisy = f32(clean), f32(noisy)
    ranges = dict(pr.param_ranges_mmse, noise_method=["min_tracking"], ksi_min=[0.001, 0.1], gain_min=[0.01, 0.2])
    pts = grid.grid_points(ranges)
    eng = engine_for(clean, noisy, chunk_items=37)
    sc = eng.sweep("mmse", pts)
    assert eng.last_unique == len(pts) // 2            # noise_percentile is dead under min_tracking
    single = engine_for(clean[2], noisy[2]).sweep("mmse", pts)
    assert np.array_equal(sc[2], single[0])
    rng = np.random.default_rng(0)
    for i in rng.choice(len(pts), 6, replace=False):
        ref = score_candidate(clean[1], oracle.mmse(noisy[1], 16000, **pts[i]), 16000)
        assert abs(sc[1, i]["stoi"] - ref["stoi"]) < TOL_STOI and abs(sc[1, i]["snr"] - ref["snr"]) < TOL_SNR_DB


@pytest.mark.parametrize("stem,n_rows", [("p257_090", 17), ("p257_135", 5)])
def test_published_rows_on_device(stem, n_rows):
    """The reference's own published per-file results (tests/golden) reproduced by the CUDA path: both shipped
    pairs, every row the committed reference code regenerates (fp32 device arithmetic + float32-rounded inputs
    against numbers the reference computed in float64: 1e-5 STOI / 1e-3 dB)."""
    from tests.golden_util import load_pair
    c, n = load_pair(stem)
    c, n = f32(c), f32(n)
    eng = engine_for(c, n)
    base = eng.baseline()[0]
    rows = published_rows(stem)
    assert len(rows) == n_rows
    assert abs(base["stoi"] - rows[0]["stoi_noisy"]) < 1e-5 and abs(base["snr"] - rows[0]["snr_noisy"]) < 1e-3
    for r in rows:
        sc = eng.sweep(r["alg"], [r["params"]])[0, 0]
        assert abs(sc["stoi"] - r["stoi"]) < 1e-5, r
        assert abs(sc["snr"] - r["snr"]) < 1e-3, r


def test_size_independent_properties_full_grid_shapes():
    """Properties that need no oracle, at the benchmark's sizes."""
    clean, noisy = make_batch(3, 48000, first=100)
    eng = engine_for(clean, noisy)
    # (1) analysis/synthesis round trip: alpha=0 leaves the power untouched -> output == input
    for n_fft, hop in ((512, 128), (512, 256), (1024, 128), (1024, 256)):
        p = dict(alpha=0.0, beta=0.0, n_fft=n_fft, hop_length=hop, noise_percentile=10.0, noise_method="min_tracking")
        out = eng.enhance("spectralSubtractor", [p])[:, 0]
        assert np.abs(out - noisy.astype(np.float32)).max() < 2e-6
    # (2) STOI(clean, clean) == 1, SNR infinite; scoring the noisy signal == baseline
    same = eng.score_waveforms(clean[:, None, :], finalize=False)[:, 0]
    assert np.all(np.abs(same["stoi"] - 1.0) < 1e-5) and np.all(same["flags"] & 4)
    # (3) dead parameters give bit-identical scores; gain floor 1.0 makes Wiener the identity
    pts = [dict(alpha=0.95, gain_floor=1.0, n_fft=512, hop_length=128, noise_percentile=pc, noise_method="min_tracking")
           for pc in (10.0, 20.0)]
    sc = eng.sweep("wiener", pts)
    assert np.array_equal(sc[:, 0], sc[:, 1])
    base = eng.baseline()
    assert np.all(np.abs(sc[:, 0]["stoi"] - base["stoi"]) < 1e-5) and np.all(sc[:, 0]["lag"] == 0)
    # (4) a delayed candidate is re-aligned: its score equals the undelayed one's up to the lost tail
    delayed = np.concatenate([np.zeros((3, 29)), noisy], axis=1)[:, :48000]
    d = eng.score_waveforms(delayed[:, None, :], finalize=True)[:, 0]
    assert np.all(d["lag"] == -29) and np.all(np.abs(d["stoi"] - base["stoi"]) < 2e-3)


def test_ragged_lengths_and_edge_cases():
    for L in (32000, 41237, 64000, 9000):
        c, n = make_pair(7, L)
        c, n = f32(c), f32(n)
        eng = engine_for(c, n)
        p = dict(alpha=0.95, gain_floor=0.05, n_fft=1024, hop_length=256, noise_percentile=10.0, noise_method="percentile")
        sc = eng.sweep("wiener", [p])[0, 0]
        ref = score_candidate(c, oracle.wiener_filter(n, 16000, **p), 16000)
        assert abs(sc["stoi"] - ref["stoi"]) < TOL_STOI and abs(sc["snr"] - ref["snr"]) < TOL_SNR_DB
    c, n = make_pair(8, 20000)
    eng = engine_for(c, n)
    bad = n.copy()
    bad[15000] = np.nan
    sc = eng.score_waveforms(np.stack([n, bad])[None], finalize=True)[0]
    assert sc[0]["flags"] & 1 and not sc[1]["flags"] & 1
    with pytest.raises(ValueError):
        eng.sweep("wiener", [dict(alpha=0.9, gain_floor=0.1, n_fft=512, hop_length=128, noise_percentile=10.0,
                                  noise_method="bogus")])
    from classical_speech_enhancement_b200._lib import CseError
    with pytest.raises(CseError):
        eng.sweep("wiener", [dict(alpha=0.9, gain_floor=0.1, n_fft=500, hop_length=128, noise_percentile=10.0,
                                  noise_method="percentile")])


def test_drop_in_entry_points_on_device():
    from classical_speech_enhancement_b200.evaluation_metrics import calculate_snr, calculate_stoi
    from classical_speech_enhancement_b200.speech_enhancement_comparison import optimize_parameters
    from classical_speech_enhancement_b200.wiener_filter import wiener_filter
    c, n = make_pair(9, 40000)
    c, n = f32(c), f32(n)
    out = wiener_filter(n, 16000, 512, 128, 0.95, 0.05, 10.0, "min_tracking")
    ref = oracle.wiener_filter(n, 16000, 512, 128, 0.95, 0.05, 10.0, "min_tracking")
    assert out.dtype == np.float64 and np.abs(out - ref).max() / np.abs(ref).max() < TOL_WAVE
    assert abs(calculate_stoi(c, n, 16000) - oracle.stoi(c, n, 16000)) < TOL_STOI
    assert abs(calculate_snr(c, n) - oracle.global_snr(c, n)) < TOL_SNR_DB
    ranges = {"alpha": [0.9, 0.98], "gain_floor": [0.01, 0.1], "n_fft": [512], "hop_length": [128],
              "noise_percentile": [10.0], "noise_method": ["percentile", "min_tracking"]}
    res = optimize_parameters(c, n, 16000, wiener_filter, ranges, pesq_scorer=lambda a, b, sr: 2.0, verbose=False)
    _, _, best = oracle.sweep_one_pair(c, n, 16000, oracle.wiener_filter, ranges, pesq_fn=lambda a, b, sr: 2.0)
    assert res["stoi"]["params"] == best["stoi"]["params"]


@pytest.mark.parametrize("n_fft,hop", [(256, 64), (256, 128), (2048, 256), (2048, 1024), (1024, 512), (512, 100)])
def test_wider_operating_range(n_fft, hop):
    """n_fft 256-2048 (north_star) and hops up to n_fft/2, including a non-power-of-two hop."""
    c, n = make_pair(13, 40000)
    c, n = f32(c), f32(n)
    eng = engine_for(c, n)
    for alg, extra in (("wiener", dict(alpha=0.95, gain_floor=0.05)),
                       ("omlsa", dict(alpha=0.9, ksi_min=0.01, gain_floor=0.1, noise_mu=0.95, q=0.4))):
        p = dict(extra, n_fft=n_fft, hop_length=hop, noise_percentile=20.0, noise_method="min_tracking")
        wav = eng.enhance(alg, [p])[0, 0]
        ref = oracle.ALGORITHMS[alg](n, 16000, **p)
        assert np.abs(wav - ref).max() / np.abs(ref).max() < TOL_WAVE, (alg, n_fft, hop)
        sc = eng.sweep(alg, [p])[0, 0]
        rs = score_candidate(c, ref, 16000)
        assert abs(sc["stoi"] - rs["stoi"]) < TOL_STOI and abs(sc["snr"] - rs["snr"]) < TOL_SNR_DB


def test_long_utterance_10s():
    c, n = make_pair(17, 160000)
    c, n = f32(c), f32(n)
    eng = engine_for(c, n)
    p = dict(alpha=0.98, ksi_min=0.001, gain_min=0.05, gain_max=1.0, n_fft=512, hop_length=128,
             noise_percentile=10.0, noise_method="percentile")
    sc = eng.sweep("mmse", [p])[0, 0]
    rs = score_candidate(c, oracle.mmse(n, 16000, **p), 16000)
    assert abs(sc["stoi"] - rs["stoi"]) < TOL_STOI and abs(sc["snr"] - rs["snr"]) < TOL_SNR_DB


def test_variable_length_pairs_and_full_grid_selection():
    """Bucketed variable-length sweep over the FULL Wiener grid: every score vs the oracle for one pair."""
    from classical_speech_enhancement_b200.sweep import sweep_pairs
    pairs = [tuple(f32(x) for x in make_pair(50, 32000)), tuple(f32(x) for x in make_pair(51, 36000)),
             tuple(f32(x) for x in make_pair(52, 32000))]
    out = sweep_pairs(pairs, grids=(("wiener", pr.param_ranges_wiener),))
    sc = out["scores"]["wiener"]
    assert sc.shape == (3, 192) and out["nominal"] == 576 and out["unique"] == 432
    pts, scores, best = oracle.sweep_one_pair(pairs[1][0], pairs[1][1], 16000, oracle.wiener_filter, pr.param_ranges_wiener)
    assert np.abs(sc[1]["stoi"] - np.array([s["stoi"] for s in scores])).max() < TOL_STOI
    assert out["selection"]["wiener"][1]["stoi"]["index"] == best["stoi"]["index"]


def test_full_ss_grid_scores_and_argmax():
    """Full spectral-subtraction grid (720 nominal / 540 unique points) for one pair: every score and the
    sequential STOI selection against the oracle."""
    from classical_speech_enhancement_b200.sweep import select_all
    c, n = make_pair(60, 32000)
    c, n = f32(c), f32(n)
    pts = grid.grid_points(pr.param_ranges_ss)
    seen, scores = {}, []
    for p in pts:                                   # the oracle is deterministic: reuse duplicates of dead parameters
        key = (p["alpha"], p["beta"], p["n_fft"], p["hop_length"], p["noise_method"],
               p["noise_percentile"] if p["noise_method"] == "percentile" else None)
        if key not in seen:
            seen[key] = score_candidate(c, oracle.spectral_subtraction(n, 16000, **p), 16000)
        scores.append(seen[key])
    best = oracle.select_best(pts, [dict(s, pesq=0.0) for s in scores])
    eng = engine_for(c, n)
    sc = eng.sweep("spectralSubtractor", pts)
    assert eng.last_unique == 540 == len(seen)
    assert np.abs(sc[0]["stoi"] - np.array([s["stoi"] for s in scores])).max() < TOL_STOI
    assert np.abs(sc[0]["snr"] - np.array([s["snr"] for s in scores])).max() < TOL_SNR_DB
    sel = select_all({"spectralSubtractor": sc}, {"spectralSubtractor": pts})["spectralSubtractor"][0]
    assert sel["stoi"]["index"] == best["stoi"]["index"]


def test_random_points_of_the_big_grids():
    """Random samples of the MMSE and Log-MMSE grids (the full grids cost CPU-hours in the oracle)."""
    c, n = make_pair(61, 48000)
    c, n = f32(c), f32(n)
    eng = engine_for(c, n)
    rng = np.random.default_rng(5)
    for alg, ranges, fn in (("mmse", pr.param_ranges_mmse, oracle.mmse), ("omlsa", pr.param_ranges_omlsa, oracle.advanced_mmse)):
        pts = grid.grid_points(ranges)
        sc = eng.sweep(alg, pts)[0]
        for i in rng.choice(len(pts), 24, replace=False):
            ref = score_candidate(c, fn(n, 16000, **pts[i]), 16000)
            assert abs(sc[i]["stoi"] - ref["stoi"]) < TOL_STOI and abs(sc[i]["snr"] - ref["snr"]) < TOL_SNR_DB, (alg, pts[i])


def test_recycled_pinned_result_buffers_give_the_same_tables():
    """Throughput drivers may ask for score tables in recycled pinned staging buffers: same values, and the
    next sweep of the same size overwrites the previous result (documented aliasing)."""
    from classical_speech_enhancement_b200 import engine as eng_mod
    clean, noisy = make_batch(2, 32000)
    pts = grid.grid_points(dict(alpha=[0.9, 0.98], gain_floor=[0.05, 0.2], n_fft=[512], hop_length=[128],
                           noise_percentile=[10.0], noise_method=["percentile", "min_tracking"]))
    eng = engine_for(f32(clean), f32(noisy))
    ref = eng.sweep("wiener", pts).copy()
    eng_mod.reuse_result_buffers(True)
    try:
        a = eng.sweep("wiener", pts)
        assert np.array_equal(a, ref)
        b = eng.sweep("wiener", pts)
        assert np.array_equal(b, ref) and np.shares_memory(a, b)
    finally:
        eng_mod.reuse_result_buffers(False)
