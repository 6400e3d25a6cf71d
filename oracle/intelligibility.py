"""Oracle STOI (test infrastructure; see oracle/__init__.py).

Restates ``pystoi.stoi(x, y, fs_sig, extended=False)`` of pystoi 0.4.1, the
function ``Code/evaluation_metrics.py:30-36`` calls.  pystoi is not vendored in
the reference and not installable here; the algorithm below is the published one
(Taal et al. 2011; C. H. Taal's MATLAB code as ported by pystoi) with pystoi's
implementation details: Octave-compatible 16k->10k resampler, ``hanning(N+2)[1:-1]``
windows, ``range(0, len - N, hop)`` framing (the last full frame is dropped),
EPS = ``np.finfo(float).eps`` placement, and silent-frame removal keyed on the
clean signal only.
"""
import numpy as np
from scipy.signal import resample_poly

FS = 10000
N_FRAME = 256
NFFT = 512
NUMBAND = 15
MINFREQ = 150
N_SEG = 30
BETA = -15.0
DYN_RANGE = 40
EPS = np.finfo(float).eps


def resample_window(p, q):
    """Kaiser-windowed sinc of Octave's ``resample`` for the ratio p/q (reduced)."""
    g = np.gcd(p, q)
    p, q = p // g, q // g
    fc = 1.0 / (2 * max(p, q))
    roll = fc / 10
    rej_db = 60.0
    half = int(np.ceil((rej_db - 8) / (28.714 * roll)))
    t = np.arange(-half, half + 1)
    ideal = 2 * p * fc * np.sinc(2 * fc * t)
    beta = 0.1102 * (rej_db - 8.7)
    return np.kaiser(2 * half + 1, beta) * ideal, p, q


def resample_to_10k(x, fs_sig):
    h, p, q = resample_window(FS, fs_sig)
    return resample_poly(x, p, q, window=h / np.sum(h))


def third_octave_bands(fs=FS, nfft=NFFT, num_bands=NUMBAND, min_freq=MINFREQ):
    """Band edges as [lo, hi) FFT-bin index pairs, and the 0/1 band matrix."""
    f = np.linspace(0, fs, nfft + 1)[: nfft // 2 + 1]
    k = np.arange(num_bands, dtype=float)
    f_lo = min_freq * 2.0 ** ((2 * k - 1) / 6)
    f_hi = min_freq * 2.0 ** ((2 * k + 1) / 6)
    obm = np.zeros((num_bands, len(f)))
    edges = []
    for i in range(num_bands):
        lo = int(np.argmin((f - f_lo[i]) ** 2))
        hi = int(np.argmin((f - f_hi[i]) ** 2))
        obm[i, lo:hi] = 1
        edges.append((lo, hi))
    return edges, obm


def _frames(x, n, hop):
    w = np.hanning(n + 2)[1:-1]
    starts = range(0, len(x) - n, hop)
    if len(starts) == 0:
        return np.zeros((0, n))
    return np.array([w * x[i:i + n] for i in starts])


def vad_mask(x10):
    """Boolean keep-mask over the 256/128 frames of the *clean* 10 kHz signal."""
    fr = _frames(x10, N_FRAME, N_FRAME // 2)
    e = 20 * np.log10(np.linalg.norm(fr, axis=1) + EPS)
    return (np.max(e) - DYN_RANGE - e) < 0


def remove_silent_frames(x, y):
    hop = N_FRAME // 2
    xf = _frames(x, N_FRAME, hop)
    yf = _frames(y, N_FRAME, hop)
    e = 20 * np.log10(np.linalg.norm(xf, axis=1) + EPS)
    mask = (np.max(e) - DYN_RANGE - e) < 0
    xf, yf = xf[mask], yf[mask]
    n_sil = (len(xf) - 1) * hop + N_FRAME
    xs = np.zeros(n_sil)
    ys = np.zeros(n_sil)
    for i in range(xf.shape[0]):
        xs[i * hop:i * hop + N_FRAME] += xf[i]
        ys[i * hop:i * hop + N_FRAME] += yf[i]
    return xs, ys


def band_envelopes(sig, obm):
    spec = np.array([np.fft.rfft(f, n=NFFT) for f in _frames(sig, N_FRAME, N_FRAME // 2)]).T
    if spec.ndim < 2:
        return np.zeros((NUMBAND, 0))
    return np.sqrt(obm @ (np.abs(spec) ** 2))


def stoi(x, y, fs_sig, extended=False):
    """x = clean, y = processed; returns d in [-1, 1] (1e-5 if < 30 frames)."""
    if extended:
        raise NotImplementedError("the reference only calls extended=False")
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.shape != y.shape:
        raise Exception("x and y should have the same length")
    if fs_sig != FS:
        x = resample_to_10k(x, fs_sig)
        y = resample_to_10k(y, fs_sig)
    x, y = remove_silent_frames(x, y)
    _, obm = third_octave_bands()
    x_tob = band_envelopes(x, obm)
    y_tob = band_envelopes(y, obm)
    if x_tob.shape[1] < N_SEG:
        return 1e-5
    xs = np.array([x_tob[:, m - N_SEG:m] for m in range(N_SEG, x_tob.shape[1] + 1)])
    ys = np.array([y_tob[:, m - N_SEG:m] for m in range(N_SEG, x_tob.shape[1] + 1)])
    alpha = np.linalg.norm(xs, axis=2, keepdims=True) / (np.linalg.norm(ys, axis=2, keepdims=True) + EPS)
    yp = np.minimum(ys * alpha, xs * (1 + 10 ** (-BETA / 20)))
    yp = yp - np.mean(yp, axis=2, keepdims=True)
    xs = xs - np.mean(xs, axis=2, keepdims=True)
    yp = yp / (np.linalg.norm(yp, axis=2, keepdims=True) + EPS)
    xs = xs / (np.linalg.norm(xs, axis=2, keepdims=True) + EPS)
    J, M = xs.shape[0], xs.shape[1]
    return float(np.sum(yp * xs) / (J * M))
