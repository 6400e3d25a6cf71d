"""Helpers shared by the golden-vector tests and tests/golden/make_golden.py."""
import json
import os

import numpy as np
from scipy.signal import resample_poly

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def prepare_48k_pair(clean_i16, noisy_i16):
    """The reference's ``prepare_pair`` (``Code/speech_enhancement_comparison.py:71-90``) for a
    48 kHz PCM16 pair: int16/32768 -> float32 (``librosa.load``) -> 16 kHz -> common length ->
    cross-correlation alignment.  ``scipy.signal.resample_poly(x, 1, 3)`` stands in for librosa's
    soxr_hq resampler (not installable here); that substitution is the whole residual
    (<= 1.2e-5 STOI, <= 4e-3 dB SNR) against the published rows."""
    from oracle.postprocess import align_to_reference, match_length
    c = resample_poly((clean_i16.astype(np.float32) / 32768.0).astype(np.float64), 1, 3)
    n = resample_poly((noisy_i16.astype(np.float32) / 32768.0).astype(np.float64), 1, 3)
    L = min(len(c), len(n))
    c, n = c[:L], n[:L]
    n = match_length(align_to_reference(c, n, 16000), len(c))
    return c, n


def load_p257_090():
    z = np.load(os.path.join(GOLDEN, "p257_090_48k.npz"))
    return prepare_48k_pair(z["clean"], z["noisy"])


def published_rows(stem=None, reproducible=True):
    rows = json.load(open(os.path.join(GOLDEN, "published_rows.json")))
    return [r for r in rows if (stem is None or r["stem"] == stem)
            and (reproducible is None or r["reproducible"] == reproducible)]
