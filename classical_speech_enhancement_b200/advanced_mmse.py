"""Drop-in for the reference's ``Code/advanced_mmse.py`` (same name, signature, result)."""
from ._percall import mono64, run_one


def advanced_mmse(noisy_audio, sr, n_fft, hop_length, alpha, ksi_min, q, noise_mu, gain_floor, noise_percentile,
                  noise_method, clean_audio=None, v_max=80.0):
    """Log-MMSE with speech-presence probability (``Code/advanced_mmse.py:7-136``) on the
    sm_100a kernels."""
    y = mono64(noisy_audio, "short_axis")
    point = dict(alpha=alpha, ksi_min=ksi_min, gain_floor=gain_floor, noise_mu=noise_mu, q=q, n_fft=n_fft,
                 hop_length=hop_length, noise_percentile=noise_percentile, noise_method=noise_method, v_max=v_max)
    return run_one("omlsa", y, point, clean_audio)


advanced_mmse.__cse_algorithm__ = "omlsa"
