"""Device time of the per-utterance (amortised) stages of one step: clean-side scoring caches, STFTs,
noise PSDs, score expansion + D2H.  Usage: python tools/phase_times.py [--utts 824]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_shard  # noqa: E402
from classical_speech_enhancement_b200 import sweep as sw  # noqa: E402
from classical_speech_enhancement_b200.engine import SweepEngine  # noqa: E402
from classical_speech_enhancement_b200.grid import ALGORITHM_IDS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=824)
a = ap.parse_args()
clean, noisy = make_shard(0, a.utts, 48000)
eng = SweepEngine(clean, noisy)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for rep in range(2):
    t_reset = timed(eng.reset)
    t_stft = timed(lambda: [eng.stft(n, h) for n in (512, 1024) for h in (128, 256)])

    def all_noise():
        for name, ranges in sw.DEFAULT_GRIDS:
            for g in eng._plan(ALGORITHM_IDS[name], sw.cached_points(name, ranges))["groups"]:
                eng.noise(g["key"])
    t_noise = timed(all_noise)
    print(f"rep {rep}: prepare_scoring {t_reset:.1f} ms, 4 STFTs {t_stft:.1f} ms, noise PSDs {t_noise:.1f} ms "
          f"({len(eng._noise)} keys)")
