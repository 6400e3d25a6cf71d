"""Deterministic synthetic clean/noisy utterance pairs for benchmarks and tests.

The reference ships no dataset (``Code/data`` is git-ignored) and there is no
network, so throughput and parity runs use speech-like synthetic pairs
(BASELINE.json ``north_star``): a harmonic, vibrato'd,
syllable-modulated source with a few pauses, plus coloured AR(1) noise at a
global SNR drawn from U(0, 15) dB.  Utterance ``u`` depends only on
``seed = 1000 + u`` so every rank / test regenerates identical data.
"""
import numpy as np
from scipy.signal import lfilter

SR = 16000


def make_pair(u, length=48000, sr=SR):
    """Return (clean, noisy) float64 arrays of ``length`` samples, sample-aligned."""
    rng = np.random.default_rng(1000 + int(u))
    t = np.arange(length) / sr
    f0 = rng.uniform(90.0, 250.0)
    vib = 1.0 + 0.10 * np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 2 * np.pi))
    phase = 2 * np.pi * np.cumsum(f0 * vib) / sr
    n_form = int(rng.integers(2, 4))
    f_c = rng.uniform(300.0, 3400.0, n_form)
    f_bw = rng.uniform(150.0, 500.0, n_form)
    clean = np.zeros(length)
    for k in range(1, 26):
        fk = k * f0
        if fk >= 0.45 * sr:
            break
        amp = (1.0 / k) * (1.0 + 4.0 * np.sum(np.exp(-0.5 * ((fk - f_c) / f_bw) ** 2)))
        clean += amp * np.sin(k * phase + rng.uniform(0, 2 * np.pi))
    f_am = rng.uniform(2.5, 5.0)
    clean *= (0.5 + 0.5 * np.sin(2 * np.pi * f_am * t + rng.uniform(0, 2 * np.pi))) ** 2
    for _ in range(int(rng.integers(2, 4))):
        dur = int(rng.uniform(0.150, 0.300) * sr)
        if length > dur + 1:
            s = int(rng.integers(0, length - dur))
            clean[s:s + dur] *= 1e-3
    clean *= 0.3 / (np.max(np.abs(clean)) + 1e-12)
    a = rng.uniform(0.5, 0.95)
    w = rng.standard_normal(length)
    noise = lfilter([1.0], [1.0, -a], w)          # AR(1): n[i] = a*n[i-1] + w[i]
    snr_db = rng.uniform(0.0, 15.0)
    scale = np.sqrt(np.sum(clean ** 2) / (np.sum(noise ** 2) * 10 ** (snr_db / 10)))
    noisy = clean + scale * noise
    return clean, noisy


def make_batch(n_utts, length=48000, first=0):
    """Stacked (clean[U, L], noisy[U, L]) float64."""
    pairs = [make_pair(first + u, length) for u in range(n_utts)]
    return np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
