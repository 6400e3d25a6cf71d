"""Oracle STFT / ISTFT (test infrastructure; see oracle/__init__.py).

Restates ``librosa.stft`` / ``librosa.istft`` (0.11.0) exactly as the reference
calls them (``Code/spectral_subtractor.py:19-26,55-62``,
``Code/wiener_filter.py:30-35,87-94``, ``Code/mmse.py:19-29,111-118``,
``Code/advanced_mmse.py:33-39,128-135``): ``window="hann"``, ``win_length=n_fft``,
``center=True``, ``pad_mode="reflect"``, ``length=len(input)``.
"""
import numpy as np


def hann_periodic(n_fft):
    """scipy.signal.get_window('hann', n_fft, fftbins=True)."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)


def stft(y, n_fft, hop):
    """(n_fft//2+1, 1+len(y)//hop) complex128; frame t = reflect-padded samples
    [t*hop, t*hop+n_fft) times the periodic Hann window, unnormalised rFFT."""
    y = np.asarray(y, dtype=np.float64)
    if len(y) <= n_fft // 2:
        # np.pad(mode='reflect') needs more than pad samples for a single reflection;
        # librosa raises for such inputs; numpy reflects repeatedly. Keep numpy's rule.
        pass
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[idx] * hann_periodic(n_fft)[None, :]
    return np.fft.rfft(frames, axis=1).T


def istft(S, hop, length):
    """Weighted overlap-add inverse of :func:`stft`; divides by the window
    sum-of-squares where it exceeds ``np.finfo(float64).tiny`` (librosa's rule),
    drops n_fft//2 leading samples and truncates / zero-pads to ``length``."""
    n_bins, n_frames = S.shape
    n_fft = 2 * (n_bins - 1)
    pad = n_fft // 2
    n_frames = min(n_frames, int(np.ceil((length + 2 * pad) / hop)))
    w = hann_periodic(n_fft)
    frames = np.fft.irfft(S[:, :n_frames], n=n_fft, axis=0) * w[:, None]
    total = n_fft + hop * (n_frames - 1)
    y = np.zeros(total)
    wss = np.zeros(total)
    w2 = w * w
    for t in range(n_frames):
        y[t * hop:t * hop + n_fft] += frames[:, t]
        wss[t * hop:t * hop + n_fft] += w2
    y = y[pad:]
    wss = wss[pad:]
    if len(y) >= length:
        y, wss = y[:length], wss[:length]
    else:
        y = np.pad(y, (0, length - len(y)))
        wss = np.pad(wss, (0, length - len(wss)))
    nz = wss > np.finfo(np.float64).tiny
    y[nz] /= wss[nz]
    return y
