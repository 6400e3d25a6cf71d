"""Latency of the reference-shaped per-pair entry point: optimize_parameters(clean, noisy, sr, algorithm, grid) for one
3 s pair and each algorithm's full grid (what the reference's main loop calls 4 x per pair), plus a cProfile of the
host side.  python tools/dropin_latency.py [--profile]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from classical_speech_enhancement_b200.speech_enhancement_comparison import algorithms_table, optimize_parameters  # noqa: E402
from classical_speech_enhancement_b200.synth import make_pair  # noqa: E402

warnings.filterwarnings("ignore")
c, n = make_pair(5, 48000)
c, n = c.astype(np.float32).astype(np.float64), n.astype(np.float32).astype(np.float64)
algs = algorithms_table()
for name, fn, ranges in algs:                       # warm-up: plans, tables, kernel attributes
    optimize_parameters(c, n, 16000, fn, ranges, pesq_scorer=None, verbose=False)
torch.cuda.synchronize()
for name, fn, ranges in algs:
    t = time.perf_counter()
    for _ in range(5):
        optimize_parameters(c, n, 16000, fn, ranges, pesq_scorer=None, verbose=False)
    torch.cuda.synchronize()
    pts = int(np.prod([len(v) for v in ranges.values()]))
    print(f"{name:20s} {pts:5d} grid points: {1e3 * (time.perf_counter() - t) / 5:7.2f} ms per optimize_parameters call")
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    for name, fn, ranges in algs:
        optimize_parameters(c, n, 16000, fn, ranges, pesq_scorer=None, verbose=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
