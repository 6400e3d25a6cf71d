#!/usr/bin/env python
"""Reduced sweep for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): 2 utterances, all four
algorithms, n_fft 512 and 1024, hops 128 and 256, the three noise methods, one odd (ragged) length, one
utterance with a NaN sample, the clean-side caches, the score expansion and the selection kernel.

    PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool memcheck python tools/sanitize_sweep.py

(The caching allocator is disabled so that every buffer is its own cudaMalloc and out-of-bounds accesses
cannot land in a neighbouring tensor.)  Prints the number of candidates processed; the verdict is the
sanitizer's own summary line.
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")


def main():
    from classical_speech_enhancement_b200.engine import SweepEngine
    from classical_speech_enhancement_b200.sweep import run_engine_device, select_winners_device
    from classical_speech_enhancement_b200.synth import make_batch
    shape = {"n_fft": [512, 1024], "hop_length": [128, 256], "noise_percentile": [10.0],
             "noise_method": ["percentile", "min_tracking", "true_noise"]}
    grids = (("spectralSubtractor", dict({"alpha": [1.0, 4.0], "beta": [0.01]}, **shape)),
             ("mmse", dict({"alpha": [0.98], "ksi_min": [0.001, 0.1], "gain_min": [0.05], "gain_max": [1.0]}, **shape)),
             ("wiener", dict({"alpha": [0.95], "gain_floor": [0.01, 0.1]}, **shape)),
             ("omlsa", dict({"alpha": [0.9], "ksi_min": [0.01], "gain_floor": [0.1], "noise_mu": [0.92, 0.98], "q": [0.3]}, **shape)))
    total = 0
    for L, nan_at in ((12000, None), (9001, 5000)):          # even length; odd length with a non-finite sample
        clean, noisy = make_batch(2, L, first=900)
        if nan_at is not None:
            noisy[1, nan_at] = np.nan
        eng = SweepEngine(clean, noisy, chunk_items=37)
        items = run_engine_device(eng, grids)
        wins = select_winners_device(eng, items)
        for name, pts, buf, pl in items:
            sc = eng.table_to_host(eng.be.view_bytes_as(buf, np.uint8), pl, eng.U)
            w = eng.winners_to_host(wins[name])
            total += sc.size
            assert w.shape == (2, 3)
        eng.baseline()
        eng.enhance("wiener", [dict(alpha=0.95, gain_floor=0.05, n_fft=256, hop_length=64, noise_percentile=20.0,
                                    noise_method="percentile")])
    print(f"sanitize_sweep: {total} utterance-configs processed")


if __name__ == "__main__":
    main()
