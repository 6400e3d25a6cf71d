"""Experiment: does running the scoring kernels of chunk k on a second stream, concurrently with the enhance kernel of
chunk k+1 (two alternating waveform / score workspaces), beat the serial chunk loop?  One (algorithm, shape, method)
group over U utterances.  python tools/overlap_probe.py [--utts 48] [--alg omlsa] [--n-fft 1024] [--hop 128]"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import make_shard  # noqa: E402
from classical_speech_enhancement_b200 import _lib, sweep as sw  # noqa: E402
from classical_speech_enhancement_b200.engine import SR, SweepEngine  # noqa: E402
from classical_speech_enhancement_b200.grid import ALGORITHM_IDS, grid_points  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=48)
ap.add_argument("--alg", default="omlsa")
ap.add_argument("--n-fft", type=int, default=1024)
ap.add_argument("--hop", type=int, default=128)
ap.add_argument("--method", default="min_tracking")
ap.add_argument("--chunk", type=int, default=3552)
a = ap.parse_args()
ranges = dict(dict(sw.DEFAULT_GRIDS)[a.alg])
ranges.update(n_fft=[a.n_fft], hop_length=[a.hop], noise_method=[a.method], noise_percentile=[10.0])
pts = grid_points(ranges)
clean, noisy = make_shard(0, a.utts, 48000)
eng = SweepEngine(clean, noisy, chunk_items=a.chunk)
be, lib = eng.be, eng.lib
alg = ALGORITHM_IDS[a.alg]
pl = eng._plan(alg, pts)
g = pl["groups"][0]
key = g["key"]
Y = eng.stft(key[0], key[1])
N, tv = eng.noise(key)
params = be.from_host(g["params_host"])
n_rows = g["n_rows"]
total = eng.U * n_rows
rec = lib.score_dtype.itemsize
scores = be.empty((total * rec,), np.uint8)
chunk = min(a.chunk, total)
wavs = [be.empty((chunk * eng.L,), np.float32) for _ in range(2)]
nbytes = lib.score_workspace_bytes(chunk, eng.L, SR)
wss = [be.empty((nbytes,), np.uint8) for _ in range(2)]
main = torch.cuda.current_stream()
side = torch.cuda.Stream()


def sptr(s):
    return ctypes.c_void_p(s.cuda_stream)


def run(overlap):
    done = [None, None]
    for k, i0 in enumerate(range(0, total, chunk)):
        n = min(chunk, total - i0)
        b = k % 2 if overlap else 0
        if overlap and done[b] is not None:
            main.wait_event(done[b])                    # the scoring of chunk k-2 has released this buffer pair
        lib.enhance_items(be.ptr(eng.tables), alg, be.ptr(Y), be.ptr(N), int(tv), eng.L, key[0], key[1], be.ptr(params), n_rows,
                          i0, n, be.ptr(wavs[b]), sptr(main))
        s = side if overlap else main
        if overlap:
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
        sargs = (be.ptr(eng.tables), be.ptr(wavs[b]), i0, n, n_rows, eng.L, SR, be.ptr(eng.clean), be.ptr(eng.cache), 1,
                 be.ptr(scores), be.ptr(wss[b]), nbytes, sptr(s))
        lib.align_items(*sargs)
        lib.stoi_items(*sargs)
        if overlap:
            done[b] = torch.cuda.Event()
            done[b].record(side)
    if overlap:
        main.wait_stream(side)


ref = None
for mode in (False, True, False, True):
    run(mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run(mode)
    e1.record()
    torch.cuda.synchronize()
    host = scores.cpu().numpy().copy()
    if ref is None:
        ref = host
    print(f"{'overlap' if mode else 'serial '}: {e0.elapsed_time(e1) / 3:8.2f} ms for {total} candidates "
          f"({1e3 * e0.elapsed_time(e1) / 3 / total:.3f} us/candidate), scores identical: {np.array_equal(host, ref)}")
