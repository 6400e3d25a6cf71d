"""Oracle per-candidate post-processing (test infrastructure; see oracle/__init__.py).

Follows ``Code/speech_enhancement_comparison.py``: ``to_mono`` ``:14-21``,
``match_length`` ``:29-36``, ``align_to_reference`` ``:38-69``,
``finalize_enhanced`` ``:92-106``.
"""
import numpy as np
from scipy.signal import correlate


def to_mono(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return x
    return np.mean(x, axis=1) if x.shape[0] >= x.shape[1] else np.mean(x, axis=0)


def match_length(x, L):
    x = np.asarray(x, dtype=np.float64)
    if len(x) > L:
        return x[:L]
    if len(x) < L:
        return np.pad(x, (0, L - len(x)))
    return x


def alignment_lag(ref, sig, sr, max_shift_s=0.10, corr_seconds=2.0):
    """Lag chosen by ``align_to_reference`` (``:44-60``); ``None`` when the
    function returns its input untouched (window shorter than 256 samples)."""
    ref = np.asarray(ref, dtype=np.float64)
    sig = np.asarray(sig, dtype=np.float64)
    N = int(min(len(ref), len(sig), corr_seconds * sr))
    if N < 256:
        return None
    r0 = ref[:N] - np.mean(ref[:N])
    s0 = sig[:N] - np.mean(sig[:N])
    c = correlate(r0, s0, mode="full", method="auto")
    lags = np.arange(-N + 1, N)
    max_lag = int(max_shift_s * sr)
    keep = (lags >= -max_lag) & (lags <= max_lag)
    if not np.any(keep):
        return None
    return int(lags[keep][np.argmax(c[keep])])


def align_to_reference(ref, sig, sr, max_shift_s=0.10, corr_seconds=2.0):
    sig = np.asarray(sig, dtype=np.float64)
    lag = alignment_lag(ref, sig, sr, max_shift_s, corr_seconds)
    if lag is None or lag == 0:
        return sig
    if lag > 0:
        return np.pad(sig, (lag, 0))                                      # :62-63
    return sig[-lag:]                                                     # :64-65


def finalize_enhanced(enhanced, clean_ref, sr, do_align=True):
    enhanced = to_mono(enhanced)
    if do_align:
        enhanced = align_to_reference(clean_ref, enhanced, sr, 0.10, 2.0)
    enhanced = match_length(enhanced, len(clean_ref))
    if not np.all(np.isfinite(enhanced)):
        return None
    return np.clip(enhanced, -1.0, 1.0)
