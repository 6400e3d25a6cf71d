"""Dataset-level sweep: every algorithm x every grid point x every utterance pair.

``sweep_dataset`` is the batched form of the reference's double loop
``for pair: for algorithm: run_algorithm_on_pair`` (``Code/speech_enhancement_comparison.py:441-458``):
it takes HOST arrays, moves them to the device once, runs the device sweep and returns the
score tables (and, optionally, the reference's three winners per (utterance, algorithm)).
"""
import numpy as np

from .engine import DEFAULT_CHUNK_ITEMS, SweepEngine
from .grid import grid_points, select_best, select_best_batch
from .parameter_ranges import (param_ranges_mmse, param_ranges_omlsa, param_ranges_ss,
                               param_ranges_wiener)

#: the reference's algorithm order (``speech_enhancement_comparison.py:395-401``)
DEFAULT_GRIDS = (("spectralSubtractor", param_ranges_ss), ("mmse", param_ranges_mmse),
                 ("wiener", param_ranges_wiener), ("omlsa", param_ranges_omlsa))


def nominal_and_unique(grids=DEFAULT_GRIDS):
    from .grid import ALGORITHM_IDS, plan
    nominal = unique = 0
    for name, ranges in grids:
        pts = grid_points(ranges)
        nominal += len(pts)
        unique += sum(len(g["rows"]) for g in plan(ALGORITHM_IDS[name], pts, lambda a, b: 376).values())
    return nominal, unique


_points_cache = {}


def cached_points(name, ranges):
    """grid_points(ranges), cached per (algorithm, ranges object) so that engines can reuse their launch plans."""
    key = (name, id(ranges))
    hit = _points_cache.get(key)
    if hit is None or hit[0] is not ranges:
        hit = (ranges, grid_points(ranges))
        _points_cache[key] = hit
    return hit[1]


def run_engine_device(engine, grids=DEFAULT_GRIDS, u_pad=None):
    """Enqueue every algorithm's sweep; returns [(name, points, device buffer, plan)] - nothing synchronised."""
    out = []
    for name, ranges in grids:
        pts = cached_points(name, ranges)
        buf, pl = engine.sweep_device(name, pts, u_pad=u_pad)
        out.append((name, pts, buf, pl))
    return out


def run_engine(engine, grids=DEFAULT_GRIDS):
    """Device part + one device->host copy per algorithm: {alg: structured scores [U, n_points]} (+ points)."""
    scores, points, unique = {}, {}, 0
    for name, pts, buf, pl in run_engine_device(engine, grids):
        scores[name] = engine.table_to_host(engine.be.view_bytes_as(buf, np.uint8), pl, engine.U)
        points[name] = pts
        unique += pl["unique"]
    return scores, points, unique


def select_all(scores, points, pesq=None):
    """The reference's three winners per (utterance, algorithm), by its sequential scan (vectorised over
    utterances).  ``pesq[alg][u][i]`` may be injected (None entries = candidate skipped); otherwise PESQ is
    0.0 (see speech_enhancement_comparison)."""
    out = {}
    for name, sc in scores.items():
        valid = (sc["flags"] & 1) != 0
        snr = np.where((sc["flags"] & 4) != 0, np.inf, sc["snr"].astype(np.float64))
        pq = None
        if pesq is not None:
            pq = np.array([[np.nan if v is None else float(v) for v in row] for row in pesq[name]], dtype=np.float64)
        out[name] = select_best_batch(points[name], sc["stoi"].astype(np.float64), pq, snr, valid)
    return out


def sweep_dataset(clean, noisy, grids=DEFAULT_GRIDS, sr=16000, select=True, chunk_items=DEFAULT_CHUNK_ITEMS, engine_kwargs=None):
    """clean, noisy: host arrays [U, L] (equal-length, 16 kHz, pair-aligned).

    Returns ``{"scores", "points", "nominal", "unique", "selection", "engine"}``."""
    eng = SweepEngine(clean, noisy, sr=sr, chunk_items=chunk_items, **(engine_kwargs or {}))
    scores, points, unique = run_engine(eng, grids)
    nominal = sum(len(p) for p in points.values()) * eng.U
    return {"scores": scores, "points": points, "nominal": nominal, "unique": unique * eng.U,
            "selection": select_all(scores, points) if select else None, "engine": eng}


def sweep_pairs(pairs, grids=DEFAULT_GRIDS, sr=16000, select=True, chunk_items=DEFAULT_CHUNK_ITEMS, engine_kwargs=None):
    """Variable-length form of :func:`sweep_dataset`: ``pairs`` is a list of (clean, noisy) 1-D arrays
    (each pair equal length, pair-aligned, 16 kHz).  Pairs are bucketed by length - one engine (and
    one set of cached spectrograms) per distinct length - and results are returned in input order."""
    by_len = {}
    for i, (c, n) in enumerate(pairs):
        c = np.asarray(c)
        n = np.asarray(n)
        if c.ndim != 1 or c.shape != n.shape:
            raise ValueError(f"pair {i}: clean and noisy must be 1-D arrays of equal length")
        by_len.setdefault(len(c), []).append(i)
    scores, points, nominal, unique = None, None, 0, 0
    for L, idx in sorted(by_len.items()):
        out = sweep_dataset(np.stack([pairs[i][0] for i in idx]), np.stack([pairs[i][1] for i in idx]), grids=grids,
                            sr=sr, select=False, chunk_items=chunk_items, engine_kwargs=engine_kwargs)
        if scores is None:
            points = out["points"]
            scores = {name: np.zeros((len(pairs), sc.shape[1]), dtype=sc.dtype) for name, sc in out["scores"].items()}
        for name, sc in out["scores"].items():
            scores[name][idx] = sc
        nominal += out["nominal"]
        unique += out["unique"]
    return {"scores": scores, "points": points, "nominal": nominal, "unique": unique,
            "selection": select_all(scores, points) if select else None}
