"""Where the device time of ONE pair goes (fresh engine per repetition, all four full grids): engine set-up, the groups'
enhance launches (incl. the spectrograms / noise PSDs / a-posteriori SNRs they need), alignment, STOI - against the same
work at the large-batch rates.  python tools/pair_timeline.py"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from classical_speech_enhancement_b200.engine import SweepEngine  # noqa: E402
from classical_speech_enhancement_b200.sweep import DEFAULT_GRIDS, cached_points  # noqa: E402
from classical_speech_enhancement_b200.synth import make_pair  # noqa: E402

warnings.filterwarnings("ignore")


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


tot = {}
reps = 6
for rep in range(reps + 2):
    c, n = make_pair(100 + rep, 48000)
    torch.cuda.synchronize()
    marks = [("start", ev())]
    eng = SweepEngine(c[None].astype(np.float32), n[None].astype(np.float32))
    marks.append(("engine: H2D, clean-side caches", ev()))
    for name, ranges in DEFAULT_GRIDS:
        table, pl = eng.sweep_device(name, cached_points(name, ranges))
        marks.append((f"{name}: {pl['unique']} unique candidates in {len(pl['groups'])} groups", ev()))
        eng.select_device(table, pl["n_points"])
        marks.append((f"{name}: selection", ev()))
    torch.cuda.synchronize()
    if rep >= 2:
        for (_, a), (k, b) in zip(marks, marks[1:]):
            tot[k] = tot.get(k, 0.0) + a.elapsed_time(b) / reps
for k, v in tot.items():
    print(f"{k:70s} {v:7.3f} ms")
print(f"{'total':70s} {sum(tot.values()):7.3f} ms  ({9744 / sum(tot.values()):.0f} k configs/s; large-batch rate: 5.12 ms per pair)")
