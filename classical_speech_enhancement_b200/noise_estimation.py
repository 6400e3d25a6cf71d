"""Drop-in for ``noise_estimation`` of the reference's ``Code/noise_estimation.py:158-212``."""
from typing import Any, Dict, Optional

import numpy as np

from ._percall import mono64, one_shot_engine


def noise_estimation(y, sr, method="percentile", n_fft=1024, hop_length=256, win_length: Optional[int] = None,
                     estimator_params: Optional[Dict[str, Any]] = None, window="hann", center=True,
                     pad_mode="reflect", **kwargs):
    """Noise PSD, shape (n_bins, 1) (``percentile``; any method when fewer than 5 frames) or
    (n_bins, n_frames) (``min_tracking``, ``true_noise``), float64.  ``percentile``,
    ``clean_audio`` and ``eps`` travel in ``**kwargs`` / ``estimator_params`` as in the reference."""
    full = {**(estimator_params or {}), **kwargs}
    if method not in ("percentile", "min_tracking", "true_noise"):
        raise ValueError(f"Unbekannte Methode: {method}")
    if (win_length or n_fft) != n_fft or window != "hann" or not center or pad_mode != "reflect":
        raise NotImplementedError("the reference only uses win_length=n_fft, window='hann', center=True, reflect")
    unsupported = set(full) - {"percentile", "clean_audio", "eps"}
    if unsupported:
        raise NotImplementedError(f"estimator parameters not used by the reference's callers: {sorted(unsupported)}")
    y = mono64(y, "axis1")
    eng = one_shot_engine(y, full.get("clean_audio"), method)
    return eng.noise_psd_host(method, int(n_fft), int(hop_length), float(full.get("percentile", 20.0)),
                              float(full.get("eps", 1e-10)))[0]
