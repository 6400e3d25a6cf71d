"""CPU oracle for the enhancement-and-scoring sweep.  TEST INFRASTRUCTURE ONLY.

This package is a float64 numpy/scipy restatement of the hot path of
Katja39/Classical_Speech_Enhancement (reference tree: ``/root/reference``).  It
exists so that the CUDA path can be checked against something; it is *not* part
of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``classical_speech_enhancement_b200`` never does, and fails
loudly if its CUDA library is missing.

Why a restatement: the reference imports librosa, pystoi, pesq and soundfile at
module top level and none of them is installed (no network), so nothing on the
path can be imported here.  The third-party arithmetic that the reference
relies on is restated from the published algorithms of the versions the
reference's report pins (librosa 0.11.0, pystoi 0.4.1, SciPy 1.16.2, NumPy
2.3.4); numpy/scipy themselves are called directly where the reference calls
them (``scipy.special.i0/i1/expn``, ``scipy.ndimage.minimum_filter1d``,
``scipy.signal.correlate``, ``scipy.signal.resample_poly``, ``np.percentile``).

Parity status: STOI / SNR / waveforms are PINNED by the reference's published
per-file results and shipped audio (``tests/test_oracle_pinning.py``,
``tests/golden/``).  PESQ is UNPINNED (package absent): the selection logic is
tested with injected PESQ values only.

Each function cites the reference ``file:line`` it follows.
"""

from .spectral import stft, istft, hann_periodic                      # noqa: F401
from .noise import noise_psd                                           # noqa: F401
from .enhance import (spectral_subtraction, wiener_filter, mmse,       # noqa: F401
                      advanced_mmse, ALGORITHMS)
from .postprocess import align_to_reference, finalize_enhanced, alignment_lag  # noqa: F401
from .intelligibility import stoi                                                 # noqa: F401
from .metrics import global_snr, combined_score                        # noqa: F401
from .search import grid_points, select_best, sweep_one_pair           # noqa: F401
