"""Per-kernel-family device time of one full-grid sweep (CUDA events around every chunk launch).
Usage: python tools/kernel_breakdown.py [--utts 32] [--chunk 1184]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import config_bytes, make_shard  # noqa: E402
from classical_speech_enhancement_b200 import sweep as sw  # noqa: E402
from classical_speech_enhancement_b200.engine import SweepEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=32)
ap.add_argument("--chunk", type=int, default=14208)
ap.add_argument("--length", type=int, default=48000)
a = ap.parse_args()
clean, noisy = make_shard(0, a.utts, a.length)
eng = SweepEngine(clean, noisy, chunk_items=a.chunk)
sw.run_engine(eng)
torch.cuda.synchronize()
eng.reset()
eng.enable_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sw.run_engine(eng)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1)
names = {0: "ss", 1: "wiener", 2: "mmse", 3: "omlsa"}
rows = []
for (kind, alg, n_fft, hop, method), (items, ms) in eng.timing_summary().items():
    _, eb, sb = config_bytes(n_fft, hop, method, a.length)
    by = items * (eb if kind == "enhance" else (sb if kind == "stoi" else 4 * min(a.length, 32000)))
    rows.append((ms, kind, names[alg], n_fft, hop, method, items, 1e3 * ms / items, by / (ms * 1e-3) / 1e9))
rows.sort(reverse=True)
print(f"total step {total:.1f} ms for {a.utts} utterances ({a.utts * 9744 / total * 1e3:.0f} nominal configs/s)")
print(f"{'ms':>9} {'share':>6} kind    alg    n_fft hop method       items   us/item  algGB/s")
for ms, kind, alg, n_fft, hop, method, items, us, gbs in rows:
    print(f"{ms:9.2f} {ms / total:6.1%} {kind:7s} {alg:6s} {n_fft:5d} {hop:3d} {method:12s} {items:7d} {us:8.2f} {gbs:8.1f}")
acc = {}
for r in rows:
    acc[r[1]] = acc.get(r[1], 0) + r[0]
print({k: f"{v:.1f} ms ({v / total:.1%})" for k, v in acc.items()})
