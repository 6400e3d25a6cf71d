"""Kernel logic and fp32 error budget, checked WITHOUT a GPU.

The CUDA sources are compiled by g++ against tests/emu/cuda_emu.h (one OS thread per CUDA
thread) and driven through the same C ABI as the product.  This is test infrastructure: the
product package never loads the emulation library.  The `-m gpu` tests repeat these
comparisons on the real sm_100a build.
"""
import numpy as np
import pytest
from scipy import special as sp

import oracle
from oracle.noise import noise_psd
from oracle.search import score_candidate
from classical_speech_enhancement_b200.synth import make_pair
from tests.emu_util import Emu, ptr

TOL_WAVE = 1e-4      # north_star: max error relative to the waveform peak, fp32
TOL_STOI = 1e-4


@pytest.fixture(scope="module")
def emu():
    return Emu(False)


@pytest.fixture(scope="module")
def pair():
    c, n = make_pair(0, 6000)
    c32, n32 = c.astype(np.float32), n.astype(np.float32)
    return c32, n32, c32.astype(np.float64), n32.astype(np.float64)


def test_symbols_and_geometry(emu):
    assert emu.lib.abi_version() == 1 and emu.lib.real_bits == 32
    assert emu.lib.bins_padded(512) == 264 and emu.lib.bins_padded(1024) == 520
    assert emu.lib.num_frames(48000, 128) == 376


def test_special_functions(emu):
    v = np.concatenate([np.logspace(-12, 0, 400), np.linspace(1, 80, 3000)])
    out = np.zeros_like(v)
    emu.lib.debug_special(0, ptr(v), ptr(out), len(v))
    ref = (1 + v) * sp.i0e(v / 2) + v * sp.i1e(v / 2)
    assert np.max(np.abs(out - ref) / ref) < 1e-6
    emu.lib.debug_special(1, ptr(v), ptr(out), len(v))
    assert np.max(np.abs(out - sp.exp1(v))) < 5e-6 * 28          # absolute: E1 enters through exp(E1/2)
    assert np.max(np.abs(out - sp.exp1(v))[v > 1]) < 2e-7


@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 256), (256, 64), (2048, 512)])
def test_stft_psd(emu, pair, n_fft, hop):
    _, n32, _, n64 = pair
    Y, P = emu.stft_psd(n32, n_fft, hop)
    ref = oracle.stft(n64, n_fft, hop)
    nb = n_fft // 2 + 1
    assert Y.shape == (1, 1 + len(n32) // hop, emu.lib.bins_padded(n_fft))
    assert np.abs(Y[0, :, :nb].T - ref).max() / np.abs(ref).max() < 1e-6
    assert np.abs(P[0, :, :nb].T - np.abs(ref) ** 2).max() / (np.abs(ref) ** 2).max() < 1e-6
    assert np.all(Y[0, :, nb:] == 0)


@pytest.mark.parametrize("L", [6000, 3000, 600])
def test_noise_estimators(emu, L):
    c, n = make_pair(1, L)
    n32 = n.astype(np.float32)
    n64 = n32.astype(np.float64)
    for n_fft, hop in ((512, 128), (256, 64)):
        nb = n_fft // 2 + 1
        _, P = emu.stft_psd(n32, n_fft, hop)
        for eps in (1e-10, 1e-12):
            for pct in (10.0, 20.0):
                N = emu.noise_percentile(P, n_fft, pct, eps)
                ref = noise_psd(n64, "percentile", n_fft, hop, percentile=pct, eps=eps)[:, 0]
                assert np.abs(N[0, :nb] - ref).max() / ref.max() < 2e-6
            if P.shape[1] >= 5:
                N = emu.noise_mintrack(P, n_fft, eps)
                ref = noise_psd(n64, "min_tracking", n_fft, hop, eps=eps)
                assert np.abs(N[0, :, :nb].T - ref).max() / ref.max() < 2e-6


CASES = [
    (0, oracle.spectral_subtraction, (3.0, 0.1), dict(alpha=3.0, beta=0.1)),
    (1, oracle.wiener_filter, (0.98, 0.01), dict(alpha=0.98, gain_floor=0.01)),
    (2, oracle.mmse, (0.98, 0.001, 0.05, 1.0, 0.98), dict(alpha=0.98, ksi_min=0.001, gain_min=0.05, gain_max=1.0)),
    (3, oracle.advanced_mmse, (0.9, 0.01, 0.1, 0.95, 0.4, 80.0),
     dict(alpha=0.9, ksi_min=0.01, gain_floor=0.1, noise_mu=0.95, q=0.4)),
]


@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 256), (256, 128)])
@pytest.mark.parametrize("method", ["percentile", "min_tracking", "true_noise"])
def test_enhance_all_algorithms(emu, pair, n_fft, hop, method):
    c32, n32, c64, n64 = pair
    L = len(n32)
    Y, _ = emu.stft_psd(n32, n_fft, hop)
    for alg, fn, row, kw in CASES:
        eps = 1e-12 if alg == 2 else 1e-10
        No = noise_psd(n64, method, n_fft, hop, percentile=10.0, clean=c64, eps=eps)
        row = list(row)
        if alg in (2, 3) and (method == "true_noise" or No.shape[1] == 1):
            row[4 if alg == 2 else 3] = -1.0
        out = emu.enhance(alg, Y, emu.layout_noise(No, n_fft), L, n_fft, hop, [row])
        ref = fn(n64, 16000, n_fft=n_fft, hop_length=hop, noise_percentile=10.0, noise_method=method,
                 clean_audio=c64 if method == "true_noise" else None, **kw)
        assert np.abs(out[0, 0] - ref).max() / np.abs(ref).max() < TOL_WAVE / 10


def test_score_and_finalize_edge_cases(emu):
    L = 20000
    c, n = make_pair(3, L)
    c32, n32 = c.astype(np.float32), n.astype(np.float32)
    c64, n64 = c32.astype(np.float64), n32.astype(np.float64)
    clean, cache = emu.prepare_clean(c32)
    delayed = np.concatenate([np.zeros(37), n64])[:L]
    advanced = np.concatenate([n64[11:], np.zeros(11)])
    nan_head = n64.copy()
    nan_head[100] = np.nan          # reference: all-NaN correlation -> lag -1600 drops it -> scored
    nan_tail = n64.copy()
    nan_tail[15000] = np.nan        # reference: finalize_enhanced returns None -> skipped
    cands = [n64, delayed, advanced, 4.0 * n64, nan_head, nan_tail]
    wav = np.stack(cands)[None].astype(np.float32)
    sc = emu.score(wav, clean, cache, finalize=True)[0]
    import warnings
    for i in range(len(cands)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = score_candidate(c64, wav[0, i].astype(np.float64), 16000)
            lag = oracle.alignment_lag(c64, wav[0, i].astype(np.float64), 16000)
        if ref is None:
            assert sc[i]["flags"] & 1 == 0
            continue
        assert sc[i]["flags"] & 1
        assert sc[i]["lag"] == lag
        assert abs(sc[i]["stoi"] - ref["stoi"]) < TOL_STOI / 10
        assert abs(sc[i]["snr"] - ref["snr"]) < 1e-4
    assert [int(x) for x in sc["lag"][:5]] == [0, -37, 11, 0, -1600]
    base = emu.score(wav[:, :1], clean, cache, finalize=False)[0, 0]
    assert abs(base["stoi"] - oracle.stoi(c64, n64, 16000)) < TOL_STOI / 10
    assert abs(base["snr"] - oracle.global_snr(c64, n64)) < 1e-4


def test_short_utterance_stoi_flag(emu):
    c, n = make_pair(4, 5000)
    clean, cache = emu.prepare_clean(c.astype(np.float32))
    sc = emu.score(n.astype(np.float32)[None, None], clean, cache)[0, 0]
    assert sc["flags"] & 8 and abs(sc["stoi"] - 1e-5) < 1e-9


def test_sweep_equals_enhance_then_score(emu):
    L = 12000
    cs, ns = zip(*[make_pair(u, L) for u in (6, 7)])
    c32 = np.stack(cs).astype(np.float32)
    n32 = np.stack(ns).astype(np.float32)
    clean, cache = emu.prepare_clean(c32)
    Y, P = emu.stft_psd(n32, 512, 128)
    N = emu.noise_mintrack(P, 512, 1e-10)
    rows = [(0.9, 0.01), (0.95, 0.05), (0.98, 0.1), (0.98, 0.02), (0.9, 0.2)]
    wav = emu.enhance(1, Y, N, L, 512, 128, rows)
    a = emu.score(wav, clean, cache)
    b = emu.sweep(1, Y, N, L, 512, 128, rows, clean, cache, chunk=3)
    assert np.array_equal(a, b)
    ref = oracle.wiener_filter(n32[1].astype(np.float64), 16000, 512, 128, 0.98, 0.1, 10.0, "min_tracking")
    sc = score_candidate(c32[1].astype(np.float64), ref, 16000)
    assert abs(b[1, 2]["stoi"] - sc["stoi"]) < TOL_STOI / 10 and abs(b[1, 2]["snr"] - sc["snr"]) < 1e-3


def test_fp64_build_reaches_1e10():
    e = Emu(True)
    c, n = make_pair(0, 6000)
    Y, _ = e.stft_psd(n, 512, 128)
    worst = 0.0
    for alg, fn, row, kw in CASES:
        eps = 1e-12 if alg == 2 else 1e-10
        No = noise_psd(n, "min_tracking", 512, 128, eps=eps)
        out = e.enhance(alg, Y, e.layout_noise(No, 512), 6000, 512, 128, [row])
        ref = fn(n, 16000, n_fft=512, hop_length=128, noise_percentile=10.0, noise_method="min_tracking", **kw)
        worst = max(worst, np.abs(out[0, 0] - ref).max() / np.abs(ref).max())
    assert worst < 1e-10
