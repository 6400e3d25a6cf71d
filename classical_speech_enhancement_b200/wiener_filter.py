"""Drop-in for the reference's ``Code/wiener_filter.py`` (same name, signature, result)."""
from ._percall import mono64, run_one


def wiener_filter(noisy_audio, sr, n_fft, hop_length, alpha, gain_floor, noise_percentile, noise_method,
                  clean_audio=None):
    """Decision-directed Wiener filter (``Code/wiener_filter.py:7-95``) on the sm_100a kernels."""
    y = mono64(noisy_audio, "axis1")
    point = dict(alpha=alpha, gain_floor=gain_floor, n_fft=n_fft, hop_length=hop_length,
                 noise_percentile=noise_percentile, noise_method=noise_method)
    return run_one("wiener", y, point, clean_audio)


wiener_filter.__cse_algorithm__ = "wiener"
