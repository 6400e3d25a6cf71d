"""Throughput of the variable-length path (one engine per distinct length, sweep.sweep_pairs) against the equal-length
batch, full grids: python tools/ragged_probe.py [--pairs 24] [--profile]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from classical_speech_enhancement_b200.sweep import sweep_dataset, sweep_pairs  # noqa: E402
from classical_speech_enhancement_b200.synth import make_pair  # noqa: E402

warnings.filterwarnings("ignore")
n = int(sys.argv[sys.argv.index("--pairs") + 1]) if "--pairs" in sys.argv else 24
rng = np.random.default_rng(1)
lengths = [int(v) for v in rng.integers(32000, 64000, n)]
ragged = [tuple(x.astype(np.float32) for x in make_pair(i, L)) for i, L in enumerate(lengths)]
equal = [tuple(x.astype(np.float32) for x in make_pair(i, 48000)) for i in range(n)]
sweep_pairs(ragged[:2], tables=False)
torch.cuda.synchronize()
for name, data in (("equal lengths, one batch", equal), ("distinct lengths, one engine each", ragged)):
    if name.startswith("equal"):
        t = time.perf_counter()
        sweep_dataset(np.stack([p[0] for p in data]), np.stack([p[1] for p in data]), tables=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print(f"{name:44s}: {1e3 * dt:8.1f} ms for {n} pairs x 9744 grid points = {n * 9744 / dt / 1e3:8.1f} k configs/s")
        continue
    ref = None
    for k in (1, 2, 3, 4, 6, 8, 12):
        sweep_pairs(data, tables=False, in_flight=k)                # the allocator pools of the k streams, warm
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = sweep_pairs(data, tables=False, in_flight=k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        same = ref is None or all(np.array_equal(ref[a], out["winners"][a]) for a in ref)
        ref = ref or out["winners"]
        print(f"{name + f', {k} in flight':44s}: {1e3 * dt:8.1f} ms for {n} pairs x 9744 grid points = {n * 9744 / dt / 1e3:8.1f} k configs/s"
              f"  winners as with 1 in flight: {same}")
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    sweep_pairs(ragged, tables=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(45)
