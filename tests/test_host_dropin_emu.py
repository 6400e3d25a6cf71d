"""Host-side drop-in layer (grid planning, dedupe, selection, reference-named entry points)
exercised end to end without a GPU: the engine runs on the thread-emulated kernels."""
import os

import numpy as np
import pytest

import oracle
from oracle.noise import noise_psd
from classical_speech_enhancement_b200 import grid
from classical_speech_enhancement_b200 import parameter_ranges as pr
from classical_speech_enhancement_b200.synth import make_pair
from tests.emu_util import use_emulated_runtime, use_product_runtime


@pytest.fixture(scope="module", autouse=True)
def emulated():
    use_emulated_runtime()
    yield
    use_product_runtime()


def f32(x):
    return x.astype(np.float32).astype(np.float64)


def test_plan_dedupes_dead_parameters():
    for ranges, alg, nominal, unique in ((pr.param_ranges_ss, 0, 720, 540), (pr.param_ranges_mmse, 2, 1920, 1440),
                                         (pr.param_ranges_wiener, 1, 192, 144), (pr.param_ranges_omlsa, 3, 6912, 2880)):
        pts = grid.grid_points(ranges)
        groups = grid.plan(alg, pts, lambda n_fft, hop: 376)
        assert len(pts) == nominal
        assert sum(len(g["rows"]) for g in groups.values()) == unique
        assert sorted(i for g in groups.values() for m in g["members"] for i in m) == list(range(nominal))
        assert len(groups) == 12          # 4 STFT shapes x (2 percentile PSDs + 1 min_tracking PSD)


def test_per_call_functions_match_oracle():
    from classical_speech_enhancement_b200.advanced_mmse import advanced_mmse
    from classical_speech_enhancement_b200.mmse import mmse
    from classical_speech_enhancement_b200.noise_estimation import noise_estimation
    from classical_speech_enhancement_b200.spectral_subtractor import spectral_subtraction
    from classical_speech_enhancement_b200.wiener_filter import wiener_filter
    c, n = make_pair(2, 5000)
    c, n = f32(c), f32(n)
    kw = dict(n_fft=512, hop_length=128, noise_percentile=10.0)
    pairs = [
        (spectral_subtraction, oracle.spectral_subtraction, dict(alpha=2.0, beta=0.05, noise_method="true_noise", clean_audio=c)),
        (wiener_filter, oracle.wiener_filter, dict(alpha=0.95, gain_floor=0.02, noise_method="percentile")),
        (mmse, oracle.mmse, dict(alpha=0.98, ksi_min=0.01, gain_min=0.1, gain_max=1.0, noise_method="min_tracking")),
        (advanced_mmse, oracle.advanced_mmse, dict(alpha=0.8, ksi_min=0.005, q=0.5, noise_mu=0.98, gain_floor=0.2,
                                                   noise_method="min_tracking")),
    ]
    for ours, ref, extra in pairs:
        a = ours(n, 16000, **kw, **extra)
        b = ref(n, 16000, **kw, **extra)
        assert a.dtype == np.float64 and a.shape == b.shape
        assert np.abs(a - b).max() / np.abs(b).max() < 1e-5
    stereo = np.stack([n, n], axis=1)
    assert np.allclose(wiener_filter(stereo, 16000, 512, 128, 0.95, 0.02, 10.0, "percentile"),
                       wiener_filter(n, 16000, 512, 128, 0.95, 0.02, 10.0, "percentile"))
    N = noise_estimation(n, 16000, method="min_tracking", n_fft=512, hop_length=128, percentile=10.0, eps=1e-10)
    ref = noise_psd(n, "min_tracking", 512, 128, eps=1e-10)
    assert N.shape == ref.shape and np.abs(N - ref).max() / ref.max() < 2e-6
    N = noise_estimation(n, 16000, method="percentile", n_fft=512, hop_length=128, percentile=20.0, eps=1e-10)
    assert N.shape == (257, 1)
    with pytest.raises(ValueError):
        noise_estimation(n, 16000, method="nope", n_fft=512, hop_length=128)
    with pytest.raises(ValueError):
        spectral_subtraction(n, 16000, 1.0, 0.01, 512, 128, 10.0, "true_noise")


def test_metrics_match_oracle():
    from classical_speech_enhancement_b200.evaluation_metrics import (calculate_combined_speech_score,
                                                                      calculate_snr, calculate_stoi)
    c, n = make_pair(3, 12000)
    c, n = f32(c), f32(n)
    assert abs(calculate_stoi(c, n, 16000) - oracle.stoi(c, n, 16000)) < 1e-5
    assert abs(calculate_snr(c, n) - oracle.global_snr(c, n)) < 1e-4
    assert calculate_snr(c, c) == float("inf")
    assert calculate_combined_speech_score(0.8, None) == 0.4
    assert abs(calculate_combined_speech_score(0.8, 2.25) - 0.65) < 1e-12


SMALL_WIENER = {"alpha": [0.90, 0.98], "gain_floor": [0.01, 0.1], "n_fft": [512], "hop_length": [128, 256],
                "noise_percentile": [10.0, 20.0], "noise_method": ["percentile", "min_tracking"]}


def test_optimize_parameters_matches_oracle_sweep():
    """Selections identical to the oracle's sequential scan, with an injected PESQ scorer."""
    from classical_speech_enhancement_b200.speech_enhancement_comparison import optimize_parameters
    from classical_speech_enhancement_b200.wiener_filter import wiener_filter
    c, n = make_pair(5, 11000)
    c, n = f32(c), f32(n)

    def fake_pesq(clean, deg, sr):          # deterministic stand-in, sensitive to the waveform
        return 1.0 + 3.0 * float(np.clip(np.corrcoef(clean, deg)[0, 1], 0, 1))

    pts, scores, best = oracle.sweep_one_pair(c, n, 16000, oracle.wiener_filter, SMALL_WIENER, pesq_fn=fake_pesq)
    out = optimize_parameters(c, n, 16000, wiener_filter, SMALL_WIENER, pesq_scorer=fake_pesq, verbose=False)
    for crit in ("stoi", "pesq", "balance"):
        assert out[crit]["params"] == best[crit]["params"], crit
        assert abs(out[crit]["score"] - best[crit]["score"]) < 1e-4
        assert abs(out[crit]["snr"] - best[crit]["snr"]) < 1e-3
    assert abs(out["baseline"]["stoi"] - oracle.stoi(c, n, 16000)) < 1e-5
    ref_wave = [s for s, p in zip(scores, pts) if p == best["stoi"]["params"]][0]
    assert set(out) == {"stoi", "pesq", "balance", "baseline", "improvements"}
    assert set(out["balance"]) == {"score", "params", "enhanced", "stoi", "pesq", "snr"}
    assert out["stoi"]["enhanced"].shape == c.shape and np.abs(out["stoi"]["enhanced"]).max() <= 1.0
    assert ref_wave is not None


def test_run_algorithm_on_pair_row_schema(tmp_path):
    from classical_speech_enhancement_b200.speech_enhancement_comparison import run_algorithm_on_pair
    from classical_speech_enhancement_b200.spectral_subtractor import spectral_subtraction
    c, n = make_pair(6, 10000)
    ranges = {"alpha": [1.0, 3.0], "beta": [0.01], "n_fft": [512], "hop_length": [128],
              "noise_percentile": [10.0], "noise_method": ["true_noise", "min_tracking"]}
    row = run_algorithm_on_pair("spectralSubtractor", spectral_subtraction, ranges, f32(c), f32(n), 16000,
                                str(tmp_path), "synth_006", pesq_scorer=lambda a, b, sr: 2.0, verbose=False)
    assert list(row)[:6] == ["alg", "stem", "sr", "stoi_noisy", "pesq_noisy", "snr_noisy"]
    assert row["best_params_stoi"]["noise_method"] == "true_noise"
    assert sorted(p.name for p in tmp_path.iterdir()) == [
        "synth_006_spectralSubtractor_optimized_balanced.wav", "synth_006_spectralSubtractor_optimized_pesq.wav",
        "synth_006_spectralSubtractor_optimized_stoi.wav"]


def test_sweep_pairs_buckets_by_length_and_result_files(tmp_path):
    from classical_speech_enhancement_b200.results_io import write_results
    from classical_speech_enhancement_b200.speech_enhancement_comparison import run_algorithm_on_pair
    from classical_speech_enhancement_b200.sweep import sweep_dataset, sweep_pairs
    from classical_speech_enhancement_b200.wiener_filter import wiener_filter
    grid = (("wiener", {"alpha": [0.95], "gain_floor": [0.02, 0.1], "n_fft": [256], "hop_length": [128],
                        "noise_percentile": [10.0], "noise_method": ["percentile", "min_tracking"]}),)
    pairs = [make_pair(0, 9000), make_pair(1, 10000), make_pair(2, 9000)]
    out = sweep_pairs(pairs, grids=grid)
    assert out["scores"]["wiener"].shape == (3, 4) and out["nominal"] == 12
    single = sweep_dataset(pairs[1][0][None], pairs[1][1][None], grids=grid)
    assert np.array_equal(out["scores"]["wiener"][1], single["scores"]["wiener"][0])
    c, n = pairs[0]
    row = run_algorithm_on_pair("wiener", wiener_filter, grid[0][1], f32(c), f32(n), 16000, None, "synth_000",
                                pesq_scorer=lambda a, b, sr: 1.5, verbose=False)
    summary = write_results([row], ["wiener"], str(tmp_path))
    assert summary["wiener"]["count"] == 1 and abs(summary["wiener"]["pesq_noisy_mean"] - 1.5) < 1e-12
    lines = (tmp_path / "all_results.csv").read_text().splitlines()
    assert lines[0].startswith("stem,alg,stoi_noisy") and lines[1].startswith("synth_000,wiener,")
    import json
    assert json.loads((tmp_path / "all_results.json").read_text())[0]["best_params_stoi"]["n_fft"] == 256


def _fake_pesq(clean, deg, sr):          # deterministic stand-in, sensitive to the waveform
    return 1.0 + 3.0 * float(np.clip(np.corrcoef(clean, deg)[0, 1], 0, 1))


def test_optimize_parameters_accepts_the_reference_wrapper_closure_and_foreign_callables():
    """The reference passes ``algorithm_wrapper`` - a closure around the entry point
    (``speech_enhancement_comparison.py:282-294``) - not the entry point itself."""
    import warnings
    from classical_speech_enhancement_b200.speech_enhancement_comparison import _resolve_algorithm, optimize_parameters
    from classical_speech_enhancement_b200.wiener_filter import wiener_filter
    c, n = make_pair(5, 11000)
    c, n = f32(c), f32(n)
    alg_fn = wiener_filter

    def algorithm_wrapper(noisy_audio, sr, **params):
        if params.get("noise_method") == "true_noise":
            return alg_fn(noisy_audio, sr, clean_audio=c, **params)
        return alg_fn(noisy_audio, sr, **params)

    assert _resolve_algorithm(algorithm_wrapper) == "wiener"
    import functools
    assert _resolve_algorithm(functools.partial(wiener_filter)) == "wiener"
    ranges = dict(SMALL_WIENER, hop_length=[128], noise_percentile=[10.0], noise_method=["percentile", "true_noise"])
    direct = optimize_parameters(c, n, 16000, wiener_filter, ranges, pesq_scorer=_fake_pesq, pesq_workers=0, verbose=False)
    wrapped = optimize_parameters(c, n, 16000, algorithm_wrapper, ranges, pesq_scorer=_fake_pesq, pesq_workers=2, verbose=False)

    def foreign(noisy_audio, sr, **params):            # nothing to unwrap: executed one grid point at a time
        kw = {"clean_audio": c} if params["noise_method"] == "true_noise" else {}
        return oracle.wiener_filter(noisy_audio, sr, **kw, **params)

    assert _resolve_algorithm(foreign) is None
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        loop = optimize_parameters(c, n, 16000, foreign, ranges, pesq_scorer=_fake_pesq, pesq_workers=0, verbose=False)
    assert any("one grid point at a time" in str(x.message) for x in w)
    for crit in ("stoi", "pesq", "balance"):
        assert wrapped[crit]["params"] == direct[crit]["params"] and wrapped[crit]["score"] == direct[crit]["score"]
        assert loop[crit]["params"] == direct[crit]["params"] and abs(loop[crit]["score"] - direct[crit]["score"]) < 1e-5
        assert np.abs(loop[crit]["enhanced"] - direct[crit]["enhanced"]).max() < 1e-5


def test_dataset_sweep_with_pesq_pool_equals_oracle_selection():
    """PESQ-complete selection (8f-1): every candidate's finalized waveform goes chunk-wise to a host process pool
    (stand-in scorer) and the three winners equal ``oracle.sweep_one_pair(pesq_fn=...)`` for every utterance."""
    from classical_speech_enhancement_b200.sweep import select_all, sweep_dataset
    from classical_speech_enhancement_b200.synth import make_batch
    clean, noisy = make_batch(2, 9000, first=30)
    clean, noisy = f32(clean), f32(noisy)
    ranges = {"alpha": [0.90, 0.98], "gain_floor": [0.01, 0.1], "n_fft": [256], "hop_length": [128],
              "noise_percentile": [10.0, 20.0], "noise_method": ["percentile", "min_tracking"]}
    out = sweep_dataset(clean, noisy, grids=(("wiener", ranges),), pesq_scorer=_fake_pesq, pesq_workers=2,
                        chunk_items=5)
    assert out["pesq"]["wiener"].shape == (2, 16) and not np.isnan(out["pesq"]["wiener"]).any()
    host = select_all(out["scores"], out["points"], pesq=out["pesq"])
    for u in range(2):
        pts, scores, best = oracle.sweep_one_pair(clean[u], noisy[u], 16000, oracle.wiener_filter, ranges, pesq_fn=_fake_pesq)
        assert np.abs(out["pesq"]["wiener"][u] - np.array([s["pesq"] for s in scores])).max() < 1e-4
        for crit in ("stoi", "pesq", "balance"):
            assert out["selection"]["wiener"][u][crit]["params"] == best[crit]["params"], (u, crit)
            assert out["selection"]["wiener"][u][crit]["index"] == host["wiener"][u][crit]["index"]
    # dead parameters: one PESQ evaluation serves both noise_percentile values under min_tracking
    pq = out["pesq"]["wiener"][0].reshape(2, 2, 1, 1, 2, 2)
    assert np.array_equal(pq[..., 0, 1], pq[..., 1, 1])


def test_run_dataset_rows_wavs_and_resume(tmp_path):
    """8f-2: the reference's batch loop (``main``, ``:441-471``) as bucketed device sweeps - rows, winner WAVs and
    result files equal to per-pair ``run_algorithm_on_pair``; all_results.json is the resume record."""
    import json
    from scipy.io import wavfile
    from classical_speech_enhancement_b200 import results_io
    from classical_speech_enhancement_b200.dataset import find_pairs, processed_stems, run_dataset
    from classical_speech_enhancement_b200.spectral_subtractor import spectral_subtraction
    from classical_speech_enhancement_b200.speech_enhancement_comparison import run_algorithm_on_pair, write_wav_pcm16
    from classical_speech_enhancement_b200.wiener_filter import wiener_filter
    shape = {"n_fft": [256], "hop_length": [128], "noise_percentile": [10.0], "noise_method": ["percentile", "min_tracking"]}
    algorithms = [("spectralSubtractor", spectral_subtraction, dict({"alpha": [1.0, 3.0], "beta": [0.01]}, **shape)),
                  ("wiener", wiener_filter, dict({"alpha": [0.95], "gain_floor": [0.02, 0.1]}, **shape))]
    data = tmp_path / "data"
    data.mkdir()
    raw = {}
    for u, L in ((0, 9000), (1, 10000), (2, 9000)):
        c, n = make_pair(u, L)
        stem = f"p{u:03d}_001"
        write_wav_pcm16(str(data / f"{stem}_clean.wav"), c, 16000)
        write_wav_pcm16(str(data / f"{stem}_noisy.wav"), n, 16000)
        raw[stem] = None
    pairs = sorted(find_pairs(str(data)), key=lambda p: p["stem"])
    assert [p["stem"] for p in pairs] == sorted(raw)
    out_dirs = {a[0]: str(tmp_path / f"results_{a[0]}") for a in algorithms}
    rows, summary = run_dataset(pairs, out_dirs, str(tmp_path / "summary"), algorithms=algorithms, pesq_scorer=_fake_pesq,
                                pesq_workers=0, verbose=False)
    assert len(rows) == 6 and summary["wiener"]["count"] == 3
    assert processed_stems(out_dirs.values()) == set(raw)
    # per-pair reference-shaped path on the same prepared signals
    from classical_speech_enhancement_b200.dataset import _prepare
    for p in pairs:
        c, n = _prepare(p, 16000)
        for name, fn, ranges in algorithms:
            ref = run_algorithm_on_pair(name, fn, ranges, c, n, 16000, str(tmp_path / "single"), p["stem"],
                                        pesq_scorer=_fake_pesq, pesq_workers=0, verbose=False)
            got = next(r for r in rows if r["stem"] == p["stem"] and r["alg"] == name)
            assert list(got) == list(ref)                                     # same keys, same order
            for k in ref:
                if isinstance(ref[k], float):
                    assert abs(got[k] - ref[k]) < 1e-6, (p["stem"], name, k)
                else:
                    assert got[k] == ref[k], (p["stem"], name, k)
            for tag in ("stoi", "pesq", "balanced"):
                a = wavfile.read(str(tmp_path / f"results_{name}" / f"{p['stem']}_{name}_optimized_{tag}.wav"))[1]
                b = wavfile.read(str(tmp_path / "single" / f"{p['stem']}_{name}_optimized_{tag}.wav"))[1]
                assert np.abs(a.astype(int) - b.astype(int)).max() <= 1
    # the files statistics.py consumes
    saved = json.loads((tmp_path / "summary" / "all_results.json").read_text())
    assert saved == json.loads(json.dumps(rows))
    for col in ("alg", "stoi_noisy", "pesq_noisy", "stoi_stoiopt", "pesq_stoiopt", "stoi_pesqopt", "pesq_pesqopt",
                "stoi_balopt", "pesq_balopt", "snr_balopt", "best_params_stoi", "best_params_pesq", "best_params_balanced"):
        assert col in saved[0]
    assert saved[0]["best_params_stoi"]["noise_method"] in ("percentile", "min_tracking")
    header = (tmp_path / "summary" / "all_results.csv").read_text().splitlines()[0]
    assert header == "stem,alg,stoi_noisy,pesq_noisy,stoi_stoiopt,pesq_stoiopt,stoi_pesqopt,pesq_pesqopt,stoi_balopt,pesq_balopt,snr_balopt"
    assert list(summary["wiener"]) == ["count"] + [f"{c}_mean" for c in results_io.REPORTED]
    # resume: nothing left to do; a new pair is appended, existing rows untouched
    rows2, _ = run_dataset(pairs, out_dirs, str(tmp_path / "summary"), algorithms=algorithms, pesq_scorer=_fake_pesq,
                           pesq_workers=0, verbose=False)
    assert rows2 == saved
    assert run_dataset(pairs, out_dirs, str(tmp_path / "summary"), algorithms=algorithms, resume=True,
                       pesq_scorer=_fake_pesq, pesq_workers=0, verbose=False)[0] == saved
    # without PESQ only the stoi winner exists and the PESQ columns are None
    with pytest.warns(UserWarning):
        rows3, _ = run_dataset(pairs[:1], {k: str(tmp_path / "nopesq" / k) for k in out_dirs}, str(tmp_path / "summary_nopesq"),
                               algorithms=algorithms[1:], pesq_scorer=None and "auto" or "auto", verbose=False)
    assert rows3[0]["pesq_noisy"] is None and rows3[0]["stoi_pesqopt"] is None and rows3[0]["stoi_stoiopt"] is not None
    assert sorted(os.listdir(tmp_path / "nopesq" / "wiener")) == [f"{pairs[0]['stem']}_wiener_optimized_stoi.wav"]


def test_front_end_resampler_and_threaded_prepare():
    """``resample_to`` (``speech_enhancement_comparison.py:23-27``): the 48 -> 16 kHz decimation by overlap-add FFT
    convolution gives the polyphase sums of ``resample_poly`` with the same filter and equals the test-side
    resampler the oracle was pinned with; dtype, length and short-input behaviour as ``librosa.resample``; the
    dataset run's threaded front end returns what the serial one does."""
    from scipy.signal import resample_poly
    from classical_speech_enhancement_b200 import speech_enhancement_comparison as sec
    from classical_speech_enhancement_b200.dataset import _prepare
    from tests.golden_util import soxr_hq_like_48k_to_16k
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "p257_135_48k.npz"))
    x = d["clean"].astype(np.float64) / 32768.0
    h = sec._soxr_hq_like_fir(1, 3, 48000)
    y = sec.resample_to(x, 48000, 16000)
    assert y.dtype == np.float64 and len(y) == -(-len(x) // 3)
    assert np.abs(y - resample_poly(x, 1, 3, window=h)).max() < 1e-14
    assert np.abs(y - soxr_hq_like_48k_to_16k(x)).max() < 6e-8       # that one rounds to float32 as librosa does
    assert sec.resample_to(x.astype(np.float32), 48000, 16000).dtype == np.float32
    short = x[:100]                                              # shorter than the filter: the polyphase path
    assert np.array_equal(sec.resample_to(short, 48000, 16000), resample_poly(short, 1, 3, window=h))
    assert sec.resample_to(x, 16000, 16000) is x
    up = sec.resample_to(x[:4000], 16000, 48000)                 # interpolation keeps the polyphase path
    assert len(up) == 12000
    pairs = [{"stem": f"s{i}", "clean": d["clean"][i * 1000:].astype(np.float32) / 32768, "noisy": d["noisy"][i * 1000:].astype(np.float32) / 32768,
              "sr": 48000} for i in range(4)]
    from concurrent.futures import ThreadPoolExecutor
    serial = [_prepare(p, 16000) for p in pairs]
    with ThreadPoolExecutor(4) as ex:
        threaded = list(ex.map(lambda p: _prepare(p, 16000), pairs))
    for a, b in zip(serial, threaded):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_single_pair_scores_all_groups_in_one_launch():
    """A one-pair engine enhances every noise-PSD group into its slice of one buffer and aligns / scores the whole
    grid with one launch each; scores equal the per-group launches bit for bit, and the two-pair batch (per-group
    path) gives the same rows."""
    from classical_speech_enhancement_b200 import engine as eng_mod
    from classical_speech_enhancement_b200.engine import SweepEngine
    ranges = {"alpha": [0.9, 0.98], "gain_floor": [0.05], "n_fft": [256], "hop_length": [128],
              "noise_percentile": [10.0], "noise_method": ["percentile", "min_tracking", "true_noise"]}
    pts = grid.grid_points(ranges)                               # three noise-PSD groups of two candidates
    c0, n0 = make_pair(5, 14000)
    c1, n1 = make_pair(6, 14000)
    one = SweepEngine(c0[None], n0[None])
    fused = one.sweep("wiener", pts)
    launches_fused = one.launches
    eng_mod._runtime["fuse_single"] = False
    try:
        per = SweepEngine(c0[None], n0[None])
        per_group = per.sweep("wiener", pts)
    finally:
        eng_mod._runtime["fuse_single"] = True
    assert fused.tobytes() == per_group.tobytes() and launches_fused < per.launches
    both = SweepEngine(np.stack([c0, c1]), np.stack([n0, n1])).sweep("wiener", pts)
    assert both[0].tobytes() == fused[0].tobytes()
    assert (fused["flags"] & 1).all() and fused["snr"].std() > 0 and fused["stoi"].std() > 0
