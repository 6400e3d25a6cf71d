"""Oracle noise-PSD estimators (test infrastructure; see oracle/__init__.py).

Follows ``Code/noise_estimation.py``: dispatcher ``:158-212``, short-signal
fallback ``:194-195,226-232``, percentile estimator ``:20-56``, minimum tracking
``:64-99``, oracle ("true_noise") ``:115-155``.
"""
import math

import numpy as np
from scipy.ndimage import minimum_filter1d

from .spectral import stft

METHODS = ("percentile", "min_tracking", "true_noise")


def quiet_frame_count(n_frames, percentile, min_frames=10, max_fraction=0.30):
    """(k, effective percentile) of ``noise_estimation.py:29-41``."""
    if n_frames < 30:
        min_frames = max(2, n_frames // 4)
        target = max(3, int(n_frames * 0.15))
        percentile = min(50.0, 100.0 * target / n_frames)
    k = max(min_frames, int(np.ceil(n_frames * (percentile / 100.0))))
    k = min(k, max(1, int(np.ceil(n_frames * max_fraction))))
    return min(k, n_frames), percentile


def percentile_psd(P, percentile, eps):
    """``noise_estimation.py:20-56`` -> (n_bins, 1)."""
    n_frames = P.shape[1]
    k, pct = quiet_frame_count(n_frames, percentile)
    frame_energy = np.mean(np.log(np.maximum(P, eps)), axis=0)          # :44
    quiet = np.argsort(frame_energy)[:k]                                 # :47
    psd = np.percentile(P[:, quiet], pct, axis=1, keepdims=True)         # :50
    psd = np.maximum(psd, 0.02 * np.median(P, axis=1, keepdims=True))    # :53-54
    return np.maximum(psd, eps)                                          # :56


def min_tracking_window(n_frames, window_size=50):
    """``noise_estimation.py:97-99``."""
    w = min(max(3, window_size), n_frames)
    return w if w % 2 == 1 else w + 1


def min_tracking_psd(P, eps):
    """``noise_estimation.py:64-95`` -> (n_bins, n_frames)."""
    n_frames = P.shape[1]
    a = max(0.8, min(0.95, 1 - 5 / n_frames))                            # :73-75
    S = np.zeros_like(P)
    S[:, 0] = P[:, 0]
    for t in range(1, n_frames):                                         # :81-82
        S[:, t] = a * S[:, t - 1] + (1 - a) * P[:, t]
    minima = minimum_filter1d(S, size=min_tracking_window(n_frames), axis=1, mode="nearest")
    psd = np.maximum(minima, 0.01 * np.median(P, axis=1, keepdims=True))  # :93-94
    return np.maximum(psd, eps)


def true_noise_psd(P, noisy, clean, n_fft, hop, eps):
    """``noise_estimation.py:115-155`` -> (n_bins, n_frames)."""
    if clean is None or noisy is None:
        raise ValueError("TrueNoiseEstimator requires clean_audio and noisy_audio")
    m = min(len(clean), len(noisy))
    d = np.asarray(noisy[:m], dtype=np.float64) - np.asarray(clean[:m], dtype=np.float64)
    N = np.maximum(np.abs(stft(d, n_fft, hop)) ** 2, eps)
    nf = P.shape[1]
    if N.shape[1] > nf:
        N = N[:, :nf]
    elif N.shape[1] < nf:
        N = np.pad(N, ((0, 0), (0, nf - N.shape[1])), mode="edge")
    return N


def noise_psd(y, method, n_fft, hop, percentile=20.0, clean=None, eps=1e-10):
    """``noise_estimation.py:158-212``: returns (n_bins, 1) or (n_bins, n_frames)."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim > 1:
        y = np.mean(y, axis=1)
    P = np.abs(stft(y, n_fft, hop)) ** 2
    n_frames = P.shape[1]
    if n_frames < 5:                                                      # :194-195
        if n_frames < 2:
            out = np.mean(P, axis=1, keepdims=True)
        else:
            out = np.percentile(P, 25, axis=1, keepdims=True)
        return np.maximum(out, eps)
    if method == "percentile":
        return percentile_psd(P, percentile, eps)
    if method == "min_tracking":
        return min_tracking_psd(P, eps)
    if method == "true_noise":
        return true_noise_psd(P, y, clean, n_fft, hop, eps)
    raise ValueError(f"Unbekannte Methode: {method}")
