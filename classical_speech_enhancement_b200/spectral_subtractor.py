"""Drop-in for the reference's ``Code/spectral_subtractor.py`` (same name, signature, result)."""
from ._percall import mono64, run_one


def spectral_subtraction(noisy_audio, sr, alpha, beta, n_fft, hop_length, noise_percentile, noise_method,
                         clean_audio=None):
    """Power spectral subtraction with over-subtraction ``alpha`` and floor ``beta``
    (``Code/spectral_subtractor.py:6-65``), computed by the sm_100a kernels."""
    y = mono64(noisy_audio, "short_axis")
    point = dict(alpha=alpha, beta=beta, n_fft=n_fft, hop_length=hop_length,
                 noise_percentile=noise_percentile, noise_method=noise_method)
    return run_one("spectralSubtractor", y, point, clean_audio)


spectral_subtraction.__cse_algorithm__ = "spectralSubtractor"
