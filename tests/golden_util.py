"""Helpers shared by the golden-vector tests and tests/golden/make_golden.py."""
import json
import os

import numpy as np
from scipy.signal import resample_poly

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def soxr_hq_like_48k_to_16k(x):
    """Oracle-side restatement of ``librosa.resample(x, orig_sr=48000, target_sr=16000)`` (soxr_hq: pass band to
    0.913 of the 8 kHz Nyquist frequency, stop band from 8 kHz, linear phase; float32 in, float32 out).  soxr is
    not installable here; this Kaiser design with a 140 dB stop band reproduces the reference's published rows to
    1.1e-7 STOI / 3e-5 dB (scipy's default ``resample_poly`` filter: 1.2e-5 / 3.4e-3)."""
    from scipy.signal import firwin, kaiserord
    numtaps, beta = kaiserord(140.0, (8000.0 - 0.913 * 8000.0) / 24000.0)
    h = firwin(numtaps | 1, 0.5 * (0.913 * 8000.0 + 8000.0), window=("kaiser", beta), fs=48000.0)
    return resample_poly(np.asarray(x, dtype=np.float64), 1, 3, window=h).astype(np.float32).astype(np.float64)


def prepare_48k_pair(clean_i16, noisy_i16):
    """The reference's ``prepare_pair`` (``Code/speech_enhancement_comparison.py:71-90``) for a
    48 kHz PCM16 pair: int16/32768 -> float32 (``librosa.load``) -> 16 kHz -> common length ->
    cross-correlation alignment."""
    from oracle.postprocess import align_to_reference, match_length
    c = soxr_hq_like_48k_to_16k(clean_i16.astype(np.float32) / 32768.0)
    n = soxr_hq_like_48k_to_16k(noisy_i16.astype(np.float32) / 32768.0)
    L = min(len(c), len(n))
    c, n = c[:L], n[:L]
    n = match_length(align_to_reference(c, n, 16000), len(c))
    return c, n


def load_pair(stem):
    z = np.load(os.path.join(GOLDEN, f"{stem}_48k.npz"))
    return prepare_48k_pair(z["clean"], z["noisy"])


def load_p257_090():
    return load_pair("p257_090")


def published_rows(stem=None, reproducible=True):
    rows = json.load(open(os.path.join(GOLDEN, "published_rows.json")))
    return [r for r in rows if (stem is None or r["stem"] == stem)
            and (reproducible is None or r["reproducible"] == reproducible)]
