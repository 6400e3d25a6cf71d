"""Drop-in for the reference's ``Code/evaluation_metrics.py`` (STOI / SNR on the device).

``calculate_pesq`` stays on the host ``pesq`` C extension exactly as in the reference (BASELINE
north_star); that package is not installable in this image, in which case the function prints the
failure and returns ``None`` - the reference's own behaviour when ``pesq.pesq`` raises (``:25-27``).
"""
import numpy as np

from .engine import SweepEngine
from .grid import combined_score as _combined


def _score_pair(clean, test):
    clean = np.asarray(clean, dtype=np.float64)
    test = np.asarray(test, dtype=np.float64)
    m = min(len(clean), len(test))
    eng = SweepEngine(clean[None, :m], test[None, :m])
    return eng.baseline()[0]


def calculate_stoi(clean_reference, test_audio, sr):
    """``pystoi.stoi(clean, test, sr, extended=False)`` (``Code/evaluation_metrics.py:30-36``)."""
    try:
        if sr != 16000:
            raise NotImplementedError("device STOI is built for 16 kHz input")
        return float(_score_pair(clean_reference, test_audio)["stoi"])
    except Exception as e:
        print(f"STOI calculation failed: {e}")
        return None


def calculate_snr(clean, processed):
    """Global SNR in dB, ``inf`` for a zero residual (``Code/evaluation_metrics.py:39-58``)."""
    try:
        sc = _score_pair(clean, processed)
        return float("inf") if sc["flags"] & 4 else float(sc["snr"])
    except Exception as e:
        print(f"SNR calculation failed: {e}")
        return None


def calculate_pesq(clean_reference, test_audio, sr):
    """``pesq.pesq(16000, ref, deg, 'wb')`` on the host (``Code/evaluation_metrics.py:9-27``)."""
    try:
        import pesq
        m = min(len(clean_reference), len(test_audio))
        if sr != 16000:
            raise NotImplementedError("resample to 16 kHz first (prepare_pair does)")
        return pesq.pesq(16000, np.asarray(clean_reference[:m]), np.asarray(test_audio[:m]), "wb")
    except Exception as e:
        print(f"PESQ calculation failed: {e}")
        return None


def calculate_combined_speech_score(stoi, pesq):
    """0.5*STOI + 0.5*max(0, PESQ)/4.5 (``Code/evaluation_metrics.py:104-114``)."""
    return _combined(stoi, pesq)
