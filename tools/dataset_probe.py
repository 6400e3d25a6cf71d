"""Dataset-level run over a ragged synthetic corpus (no PESQ): seconds per pair with one bucket at a time and with
several in flight: python tools/dataset_probe.py [--pairs 48] [--files]
(--files: the corpus as 48 kHz PCM16 WAV files on disk, so that read + resample + pair alignment are inside the time)"""
import os
import sys
import tempfile
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from classical_speech_enhancement_b200.dataset import run_dataset  # noqa: E402
from classical_speech_enhancement_b200.synth import make_pair  # noqa: E402

warnings.filterwarnings("ignore")
n = int(sys.argv[sys.argv.index("--pairs") + 1]) if "--pairs" in sys.argv else 48
rng = np.random.default_rng(2)
pairs = []
for i, L in enumerate(int(v) for v in rng.integers(32000, 64000, n)):
    c, x = make_pair(i, L)
    pairs.append({"stem": f"p{i:03d}_001", "clean": c, "noisy": x, "prepared": True})
files = "--files" in sys.argv
corpus = tempfile.TemporaryDirectory()
if files:
    from classical_speech_enhancement_b200.dataset import find_pairs
    from classical_speech_enhancement_b200.speech_enhancement_comparison import write_wav_pcm16
    for i, L in enumerate(int(v) for v in rng.integers(3 * 32000, 3 * 64000, n)):
        c, x = make_pair(i, L)
        write_wav_pcm16(os.path.join(corpus.name, f"p{i:03d}_001_clean.wav"), c, 48000)
        write_wav_pcm16(os.path.join(corpus.name, f"p{i:03d}_001_noisy.wav"), x, 48000)
    pairs = sorted(find_pairs(corpus.name), key=lambda p: p["stem"])
for k in (8, 1, 8):                                             # first pass warms allocator pools and plans
    with tempfile.TemporaryDirectory() as d:
        out_dirs = {a: os.path.join(d, f"results_{a}") for a in ("spectralSubtractor", "mmse", "wiener", "omlsa")}
        t = time.perf_counter()
        rows, _ = run_dataset(pairs, out_dirs, os.path.join(d, "summary"), pesq_scorer=None, verbose=False, in_flight=k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        print(f"{k} buckets in flight: {dt:6.2f} s for {n} pairs x 9744 grid points, 3 winner WAVs x 4 algorithms each = "
              f"{1e3 * dt / n:6.1f} ms per pair, {n * 9744 / dt / 1e3:7.1f} k configs/s, {len(rows)} rows")
