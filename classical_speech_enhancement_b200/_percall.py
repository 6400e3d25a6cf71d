"""Shared plumbing of the per-call drop-in functions (batch of one utterance, one grid point)."""
import numpy as np

from .engine import SweepEngine


def mono64(x, rule):
    """Mono fold exactly as each reference entry point does it."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim > 1:
        if rule == "short_axis":     # spectral_subtractor.py:12-14, advanced_mmse.py:25-28
            x = x.mean(axis=0) if x.shape[0] < x.shape[1] else x.mean(axis=1)
        else:                        # wiener_filter.py:24-25, mmse.py:12-13, noise_estimation.py:179-180
            x = np.mean(x, axis=1)
    return x


def one_shot_engine(noisy, clean_audio, noise_method):
    clean = None
    if noise_method == "true_noise":
        if clean_audio is None:
            raise ValueError("TrueNoiseEstimator requires clean_audio and noisy_audio")
        clean = np.asarray(clean_audio, dtype=np.float64)
        if clean.ndim != 1 or len(clean) != len(noisy):
            raise NotImplementedError("true_noise needs a 1-D clean_audio of the noisy signal's length")
    return SweepEngine(None if clean is None else clean[None, :], noisy[None, :], prepare_scoring=False)


def run_one(alg_name, noisy, point, clean_audio):
    eng = one_shot_engine(noisy, clean_audio, point["noise_method"])
    return eng.enhance(alg_name, [point])[0, 0].astype(np.float64)
