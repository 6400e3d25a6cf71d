"""Run ONE (algorithm, n_fft, hop, noise_method) group of the sweep on U utterances - the unit
to put under ncu.  Usage: python tools/profile_group.py --alg omlsa --n-fft 1024 --hop 128
--method min_tracking --utts 16 [--reps 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import config_bytes, make_shard  # noqa: E402
from classical_speech_enhancement_b200 import sweep as sw  # noqa: E402
from classical_speech_enhancement_b200.engine import SweepEngine  # noqa: E402
from classical_speech_enhancement_b200.grid import grid_points  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--alg", default="omlsa")
ap.add_argument("--n-fft", type=int, default=1024)
ap.add_argument("--hop", type=int, default=128)
ap.add_argument("--method", default="min_tracking")
ap.add_argument("--utts", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--chunk", type=int, default=14208)
ap.add_argument("--length", type=int, default=48000)
a = ap.parse_args()
ranges = dict(dict(sw.DEFAULT_GRIDS)[a.alg])
ranges.update(n_fft=[a.n_fft], hop_length=[a.hop], noise_method=[a.method], noise_percentile=[10.0])
pts = grid_points(ranges)
clean, noisy = make_shard(0, a.utts, a.length)
eng = SweepEngine(clean, noisy, chunk_items=a.chunk)
eng.sweep(a.alg, pts)
torch.cuda.synchronize()
for rep in range(a.reps):
    eng.enable_timing(True)
    eng.sweep(a.alg, pts)
    torch.cuda.synchronize()
    for (kind, alg, n_fft, hop, method), (items, ms) in sorted(eng.timing_summary().items()):
        _, eb, sb = config_bytes(n_fft, hop, method, a.length)
        by = items * (eb if kind == "enhance" else (sb if kind == "stoi" else 4 * min(a.length, 32000)))
        print(f"rep {rep} {kind:8s} items {items:6d} {ms:8.3f} ms {1e3 * ms / items:7.3f} us/item "
              f"{by / (ms * 1e-3) / 1e9:8.1f} algGB/s")
