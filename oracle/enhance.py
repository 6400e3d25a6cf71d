"""Oracle enhancement algorithms (test infrastructure; see oracle/__init__.py).

One generic frame march (`_march`) with a per-algorithm gain rule restates the
four reference entry points:

* ``spectral_subtraction``  <- ``Code/spectral_subtractor.py:6-65``
* ``wiener_filter``         <- ``Code/wiener_filter.py:7-95``
* ``mmse``                  <- ``Code/mmse.py:6-120``
* ``advanced_mmse``         <- ``Code/advanced_mmse.py:7-136``

Signatures and keyword names are the reference's so that the parameter dicts of
``Code/parameter_ranges.py`` can be splatted into them unchanged.
"""
import numpy as np
from scipy.special import expn, i0, i1

from .noise import noise_psd
from .spectral import istft, stft


def _mono64(x, rule):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim > 1:
        if rule == "short_axis":        # spectral_subtractor.py:12-14, advanced_mmse.py:25-28
            x = x.mean(axis=0) if x.shape[0] < x.shape[1] else x.mean(axis=1)
        else:                           # wiener_filter.py:24-25, mmse.py:12-13
            x = np.mean(x, axis=1)
    return x


def _smooth_noise(N, mu):
    """First-order recursive smoothing of a time-varying noise PSD
    (``mmse.py:48-54``, ``advanced_mmse.py:60-66``)."""
    mu = float(np.clip(mu, 0.0, 0.9999))
    out = np.empty_like(N)
    out[:, 0] = N[:, 0]
    for t in range(1, N.shape[1]):
        out[:, t] = mu * out[:, t - 1] + (1.0 - mu) * N[:, t]
    return out


def _march(P, N, eps, alpha, first_xi, next_xi_floor, gain_rule, g_init):
    """Decision-directed frame march shared by Wiener / MMSE / Log-MMSE.

    State per bin: previous clipped gain and previous a-posteriori SNR.
    """
    n_bins, n_frames = P.shape
    adaptive = N.ndim == 2 and N.shape[1] > 1
    G = np.zeros((n_bins, n_frames))
    g_prev = np.full((n_bins, 1), g_init, dtype=np.float64)
    gam_prev = np.ones((n_bins, 1))
    for t in range(n_frames):
        Nt = np.maximum(N[:, t:t + 1] if adaptive else N, eps)
        gam = np.maximum(P[:, t:t + 1] / Nt, eps)
        if t == 0:
            xi = first_xi(gam)
        else:
            xi = alpha * (g_prev ** 2) * gam_prev + (1.0 - alpha) * np.maximum(gam - 1.0, 0.0)
            xi = np.maximum(xi, next_xi_floor)
        g = gain_rule(xi, gam)
        G[:, t:t + 1] = g
        g_prev, gam_prev = g, gam
    return G


def spectral_subtraction(noisy_audio, sr, alpha, beta, n_fft, hop_length,
                         noise_percentile, noise_method, clean_audio=None):
    y = _mono64(noisy_audio, "short_axis")
    L = len(y)
    eps = 1e-10
    Y = stft(y, n_fft, hop_length)
    P = np.abs(Y) ** 2
    N = np.maximum(noise_psd(y, noise_method, n_fft, hop_length, percentile=noise_percentile,
                             clean=clean_audio, eps=eps), eps)           # :28-37
    Pc = np.maximum(P - alpha * N, beta * N)                             # :44,48
    S = np.sqrt(Pc) * np.exp(1j * np.angle(Y))                           # :51-53
    return istft(S, hop_length, L)


def wiener_filter(noisy_audio, sr, n_fft, hop_length, alpha, gain_floor,
                  noise_percentile, noise_method, clean_audio=None):
    y = _mono64(noisy_audio, "axis1")
    L = len(y)
    eps = 1e-10
    Y = stft(y, n_fft, hop_length)
    P = np.abs(Y) ** 2
    N = np.maximum(noise_psd(y, noise_method, n_fft, hop_length, percentile=noise_percentile,
                             clean=clean_audio, eps=eps), eps)           # :40-47

    def first(gam):                                                      # :63-67,72
        return np.maximum(np.maximum(gam - 1.0, 0.0), 1e-10)

    def rule(xi, gam):                                                   # :75-78
        return np.clip(xi / (1.0 + xi), gain_floor, 1.0)

    G = _march(P, N, eps, alpha, first, 1e-10, rule, 1.0)
    return istft(Y * G, hop_length, L)


def mmse(noisy_audio, sr, alpha, ksi_min, gain_min, gain_max, n_fft, hop_length,
         noise_percentile, noise_method, noise_mu=0.98, clean_audio=None,
         log=True, log_every=50):
    y = _mono64(noisy_audio, "axis1")
    L = len(y)
    eps = 1e-12                                                          # :17
    Y = stft(y, n_fft, hop_length)
    P = np.abs(Y) ** 2
    N = noise_psd(y, noise_method, n_fft, hop_length, percentile=noise_percentile,
                  clean=clean_audio, eps=eps)                            # :35-46 (no floor here)
    if noise_method != "true_noise" and N.ndim == 2 and N.shape[1] > 1:  # :48-54
        N = _smooth_noise(N, noise_mu)

    def first(gam):                                                      # :78-80
        return np.maximum(gam - 1.0, ksi_min)

    def rule(xi, gam):                                                   # :88-100
        v = np.clip((xi * gam) / (1.0 + xi), eps, 80.0)
        x = 0.5 * v
        A = (np.sqrt(np.pi) / 2.0) * (np.sqrt(v) / (gam + eps))
        g = A * np.exp(-x) * ((1.0 + v) * i0(x) + v * i1(x))
        g = np.nan_to_num(g, nan=gain_min, posinf=gain_max, neginf=gain_min)
        return np.clip(g, gain_min, gain_max)

    G = _march(P, N, eps, alpha, first, ksi_min, rule, 1.0)
    return istft(Y * G, hop_length, L)


def advanced_mmse(noisy_audio, sr, n_fft, hop_length, alpha, ksi_min, q, noise_mu,
                  gain_floor, noise_percentile, noise_method, clean_audio=None, v_max=80.0):
    y = _mono64(noisy_audio, "short_axis")
    L = len(y)
    eps = 1e-10
    Y = stft(y, n_fft, hop_length)
    P = np.abs(Y) ** 2
    N = np.maximum(noise_psd(y, noise_method, n_fft, hop_length, percentile=noise_percentile,
                             clean=clean_audio, eps=eps), eps)           # :43-51
    if noise_method != "true_noise" and N.ndim == 2 and N.shape[1] > 1:  # :60-66
        N = _smooth_noise(N, noise_mu)
    qv = float(np.clip(q, 1e-3, 1 - 1e-3))                               # :72

    def first(gam):                                                      # :92-93
        return np.maximum(gam - 1.0, ksi_min)

    def rule(xi, gam):                                                   # :101-113
        v = np.clip((xi * gam) / (1.0 + xi), 1e-12, v_max)
        g_lsa = (xi / (1.0 + xi)) * np.exp(0.5 * expn(1, v))
        g_lsa = np.nan_to_num(g_lsa, nan=gain_floor, posinf=1.0, neginf=gain_floor)
        lam = (1.0 / (1.0 + xi)) * np.exp(v)
        p = np.clip(1.0 / (1.0 + (1.0 - qv) / (qv * lam + eps)), 0.0, 1.0)
        return np.clip((g_lsa ** p) * (gain_floor ** (1.0 - p)), gain_floor, 1.0)

    G = _march(P, N, eps, alpha, first, ksi_min, rule, gain_floor)
    return istft(Y * G, hop_length, L)


#: reference algorithm table (``Code/speech_enhancement_comparison.py:395-401``)
ALGORITHMS = {
    "spectralSubtractor": spectral_subtraction,
    "mmse": mmse,
    "wiener": wiener_filter,
    "omlsa": advanced_mmse,
}
