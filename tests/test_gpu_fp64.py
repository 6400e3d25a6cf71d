"""The -DCSE_FP64 build (same symbols, double buffers) against the float64 oracle: north_star
asks max relative error <= 1e-10 for enhanced waveforms; selection must then be exact."""
import numpy as np
import pytest

import oracle
from oracle.search import score_candidate
from classical_speech_enhancement_b200 import parameter_ranges as pr
from classical_speech_enhancement_b200.synth import make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib64():
    from classical_speech_enhancement_b200 import _lib
    lib = _lib.load(fp64=True)
    assert lib.real_bits == 64
    return lib


def test_fp64_waveforms_and_scores(lib64):
    from classical_speech_enhancement_b200.engine import SweepEngine
    c, n = make_pair(31, 48000)
    eng = SweepEngine(c[None], n[None], lib=lib64)
    cases = [
        ("spectralSubtractor", oracle.spectral_subtraction, dict(alpha=2.0, beta=0.05)),
        ("wiener", oracle.wiener_filter, dict(alpha=0.95, gain_floor=0.02)),
        ("mmse", oracle.mmse, dict(alpha=0.98, ksi_min=0.001, gain_min=0.05, gain_max=1.0)),
        ("omlsa", oracle.advanced_mmse, dict(alpha=0.9, ksi_min=0.01, gain_floor=0.1, noise_mu=0.95, q=0.4)),
    ]
    for name, fn, base in cases:
        for method in ("percentile", "min_tracking", "true_noise"):
            for n_fft, hop in ((512, 128), (1024, 256)):
                p = dict(base, n_fft=n_fft, hop_length=hop, noise_percentile=10.0, noise_method=method)
                wav = eng.enhance(name, [p])[0, 0]
                sc = eng.sweep(name, [p])[0, 0]
                kw = {"clean_audio": c} if method == "true_noise" else {}
                ref = fn(n, 16000, **kw, **p)
                assert np.abs(wav - ref).max() / np.abs(ref).max() < 1e-10, (name, method, n_fft)
                rs = score_candidate(c, ref, 16000)
                assert abs(sc["stoi"] - rs["stoi"]) < 1e-9 and abs(sc["snr"] - rs["snr"]) < 1e-8


def test_fp64_selection_is_exact(lib64):
    from classical_speech_enhancement_b200.engine import SweepEngine
    from classical_speech_enhancement_b200.sweep import select_all
    c, n = make_pair(32, 32000)
    ranges = dict(pr.param_ranges_wiener, n_fft=[512], hop_length=[128, 256])
    pts, scores, best = oracle.sweep_one_pair(c, n, 16000, oracle.wiener_filter, ranges)
    eng = SweepEngine(c[None], n[None], lib=lib64)
    sc = eng.sweep("wiener", pts)
    assert np.abs(sc[0]["stoi"] - np.array([s["stoi"] for s in scores])).max() < 1e-9
    sel = select_all({"wiener": sc}, {"wiener": pts})["wiener"][0]
    assert sel["stoi"]["index"] == best["stoi"]["index"]
