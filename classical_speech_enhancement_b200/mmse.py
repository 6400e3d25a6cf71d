"""Drop-in for the reference's ``Code/mmse.py`` (same name, signature, result)."""
from ._percall import mono64, run_one


def mmse(noisy_audio, sr, alpha, ksi_min, gain_min, gain_max, n_fft, hop_length, noise_percentile, noise_method,
         noise_mu=0.98, clean_audio=None, log=True, log_every=50):
    """Ephraim-Malah MMSE-STSA (``Code/mmse.py:6-120``) on the sm_100a kernels.
    ``log`` / ``log_every`` are accepted and ignored, as in the reference (``:8``)."""
    y = mono64(noisy_audio, "axis1")
    point = dict(alpha=alpha, ksi_min=ksi_min, gain_min=gain_min, gain_max=gain_max, n_fft=n_fft,
                 hop_length=hop_length, noise_percentile=noise_percentile, noise_method=noise_method,
                 noise_mu=noise_mu)
    return run_one("mmse", y, point, clean_audio)


mmse.__cse_algorithm__ = "mmse"
