"""Oracle scalar metrics (test infrastructure; see oracle/__init__.py).

``global_snr`` <- ``Code/evaluation_metrics.py:39-58`` (a *global* SNR, not a
segmental one); ``combined_score`` <- ``:104-114``.
"""
import numpy as np


def global_snr(clean, processed):
    clean = np.asarray(clean)
    processed = np.asarray(processed)
    m = min(len(clean), len(processed))
    clean, processed = clean[:m], processed[:m]
    p_signal = np.sum(clean ** 2)
    p_noise = np.sum((clean - processed) ** 2)
    if p_noise == 0:
        return float("inf")
    return float(10 * np.log10(p_signal / (p_noise + 1e-10)))


def combined_score(stoi, pesq):
    if stoi is None:
        stoi = 0
    if pesq is None:
        pesq = 0
    return 0.5 * stoi + 0.5 * (max(0, pesq) / 4.5)
