"""PESQ for EVERY candidate, as the reference's selection needs it (``Code/speech_enhancement_comparison.py:178``:
two of its three winners - ``pesq`` and ``balance`` - depend on ``calculate_pesq`` of each grid point).

PESQ stays on the host (BASELINE north_star: the reference's ``pesq`` C extension).  The device produces the
candidates' waveforms chunk by chunk; each chunk is copied out on a side stream while the next chunk is being
computed, finalized (shift by the device-estimated lag, length-match, clip: ``finalize_enhanced`` ``:92-106``)
and scored by a pool of host processes.  The resulting [U][P] table of doubles (NaN where ``calculate_pesq``
returned None or the candidate was invalid) feeds ``cse_select_best``.

Any ``scorer(clean, wav, sr) -> float | None`` can stand in for ``calculate_pesq`` (the ``pesq`` package is not
installable in every image; PESQ parity is unpinned until it is - DESIGN.md section 2).
"""
import os

import numpy as np

from .engine import finalize_host

_worker_state = {}


def _score_job(job, scorer=None, sr=None):
    """Runs in a pool process (scorer inherited through the fork) or inline (scorer passed): -> [(u, grid indices, value | None)]."""
    u, idx_lists, clean, wavs, lags, flags = job
    if scorer is None:
        scorer, sr = _worker_state["scorer"], _worker_state["sr"]
    clean = np.asarray(clean, dtype=np.float64)
    out = []
    for j, idx in enumerate(idx_lists):
        if flags is not None and not int(flags[j]) & 1:
            out.append((u, idx, None))              # invalid candidate: the reference skips it before scoring (:171-173)
            continue
        w = np.asarray(wavs[j], dtype=np.float64)
        if lags is not None:
            w = finalize_host(w, int(lags[j]), len(clean))
        try:
            v = scorer(clean, w, sr)
        except Exception as e:                      # calculate_pesq prints and returns None (evaluation_metrics.py:25-27)
            print(f"PESQ calculation failed: {e}")
            v = None
        out.append((u, idx, None if v is None else float(v)))
    return out


class PesqPool:
    """``workers`` host processes scoring candidate waveforms; ``workers=0`` scores inline (tests, scorers that
    keep state).  The pool forks BEFORE any job is queued, with the scorer in the children's globals, so
    closures and lambdas work; the children never touch CUDA."""

    def __init__(self, scorer, sr, workers=None):
        if workers is None:
            workers = max(1, min((os.cpu_count() or 2) - 1, 32))
        self.workers = int(workers)
        self.scorer, self.sr = scorer, sr
        self.results = {}
        self.pending = []
        self.pool = None
        if self.workers > 0:
            import multiprocessing as mp
            _worker_state.update(scorer=scorer, sr=sr)      # the children take their copy at the fork below
            self.pool = mp.get_context("fork").Pool(self.workers)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def submit(self, u, idx_lists, clean, wavs, lags=None, flags=None):
        """Queue the candidates of utterance ``u``: ``wavs[j]`` is scored once and its value stored for the grid
        index (or list of indices - duplicates through dead parameters) ``idx_lists[j]``.  With ``lags`` the
        waveforms are raw device output and are finalized in the worker; without, they are final already."""
        job = (int(u), [([int(i)] if np.isscalar(i) else [int(k) for k in i]) for i in idx_lists], np.asarray(clean),
               [np.asarray(w) for w in wavs], None if lags is None else np.asarray(lags),
               None if flags is None else np.asarray(flags))
        if self.pool is None:
            self._store(_score_job(job, self.scorer, self.sr))
        else:
            self.pending.append(self.pool.apply_async(_score_job, (job,)))

    def _store(self, rows):
        for u, idx, v in rows:
            for i in idx:
                self.results[(u, i)] = v

    def drain(self):
        for p in self.pending:
            self._store(p.get())
        self.pending = []

    def table(self, n_utts, n_points):
        """float64 [n_utts, n_points]; NaN = not scored / scorer returned None (candidate skipped, ``:180-181``)."""
        self.drain()
        out = np.full((n_utts, n_points), np.nan)
        for (u, i), v in self.results.items():
            if v is not None:
                out[u, i] = v
        return out

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
