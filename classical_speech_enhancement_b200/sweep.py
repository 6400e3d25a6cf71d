"""Dataset-level sweep: every algorithm x every grid point x every utterance pair.

``sweep_dataset`` is the batched form of the reference's double loop
``for pair: for algorithm: run_algorithm_on_pair`` (``Code/speech_enhancement_comparison.py:441-458``):
it takes HOST arrays, moves them to the device once, runs the device sweep and returns the
score tables (and, optionally, the reference's three winners per (utterance, algorithm)).
"""
import warnings

import numpy as np

from .engine import DEFAULT_CHUNK_ITEMS, GridPoints, SweepEngine
from .grid import best_from_winners, grid_points, select_best_batch
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


def _ranges_signature(ranges):
    return tuple((k, tuple(v)) for k, v in ranges.items())


def cached_points(name, ranges):
    """grid_points(ranges) as a :class:`GridPoints`, cached by the CONTENT of the range dict (bounded), so that engines
    reuse launch plans and the per-pair entry points do not re-enumerate and re-hash a grid for every pair."""
    key = (name, _ranges_signature(ranges))
    hit = _points_cache.get(key)
    if hit is None:
        if len(_points_cache) >= 64:
            _points_cache.pop(next(iter(_points_cache)))
        hit = _points_cache[key] = GridPoints(grid_points(ranges))
    return hit


def run_engine_device(engine, grids=DEFAULT_GRIDS, u_pad=None):
    """Enqueue every algorithm's sweep; returns [(name, points, device buffer, plan)] - nothing synchronised."""
    out = []
    for name, ranges in grids:
        pts = cached_points(name, ranges)
        buf, pl = engine.sweep_device(name, pts, u_pad=u_pad)
        out.append((name, pts, buf, pl))
    return out


def run_engine_device_with_pesq(engine, pesq_scorer, grids=DEFAULT_GRIDS, u_pad=None, pesq_workers=None):
    """:func:`run_engine_device` + host PESQ of EVERY candidate (``speech_enhancement_comparison.py:178``): each
    chunk's waveforms are copied out on a side stream while the next chunk runs and scored by a process pool.
    Returns (items, {alg: float64 [U, P] PESQ table, NaN = skipped})."""
    from .pesq_pool import PesqPool
    out, tables = [], {}
    for name, ranges in grids:
        pts = cached_points(name, ranges)
        with PesqPool(pesq_scorer, engine.sr, workers=pesq_workers) as pool:
            def sink(g, i0, wav, recs, pool=pool):
                n_rows, members = g["n_rows"], g["members"]
                j = 0
                while j < len(recs):                                   # runs of items of one utterance
                    u = (i0 + j) // n_rows
                    j1 = min(len(recs), (u + 1) * n_rows - i0)
                    rows = [(i0 + k) % n_rows for k in range(j, j1)]
                    pool.submit(u, [members[r] for r in rows], engine.clean_host[u], wav[j:j1], recs["lag"][j:j1],
                                recs["flags"][j:j1])
                    j = j1
            buf, pl = engine.sweep_device(name, pts, u_pad=u_pad, chunk_sink=sink)
            tables[name] = pool.table(engine.U, len(pts))
        out.append((name, pts, buf, pl))
    return out, tables


def run_engine(engine, grids=DEFAULT_GRIDS):
    """Device part + one device->host copy per algorithm: {alg: structured scores [U, n_points]} (+ points)."""
    scores, points, unique = {}, {}, 0
    for name, pts, buf, pl in run_engine_device(engine, grids):
        scores[name] = engine.table_to_host(engine.be.view_bytes_as(buf, np.uint8, tag=(name, "table")), pl, engine.U)
        points[name] = pts
        unique += pl["unique"]
    return scores, points, unique


def _pesq_array(pesq_rows):
    """[[float | None]] -> float64 [U, P] with NaN where ``calculate_pesq`` returned None (candidate skipped)."""
    if isinstance(pesq_rows, np.ndarray) and pesq_rows.dtype == np.float64:
        return pesq_rows
    return np.array([[np.nan if v is None else float(v) for v in row] for row in pesq_rows], dtype=np.float64)


def _mark_pesq_unavailable(best):
    for c in ("pesq", "balance"):
        best[c] = {"index": None, "score": None, "params": {}, "unavailable": "PESQ was not computed"}
    return best


def select_all(scores, points, pesq=None):
    """HOST restatement of the selection (vectorised over utterances): the reference's three winners per
    (utterance, algorithm) by its sequential scan.  The product path selects on the device
    (:func:`select_winners_device`); this scan is its cross-check and serves callers that only hold host tables.
    ``pesq[alg][u][i]`` may be injected (None / NaN = candidate skipped); without it the ``pesq`` / ``balance``
    entries are marked unavailable (a constant PESQ would make them the first valid grid point / the STOI
    winner - not something to write into result files)."""
    out = {}
    for name, sc in scores.items():
        valid = (sc["flags"] & 1) != 0
        snr = np.where((sc["flags"] & 4) != 0, np.inf, sc["snr"].astype(np.float64))
        pq = _pesq_array(pesq[name]) if pesq is not None else None
        sel = select_best_batch(points[name], sc["stoi"].astype(np.float64), pq, snr, valid)
        if pq is None:
            warnings.warn("selection without PESQ: only the 'stoi' winner is available", stacklevel=2)
            sel = [_mark_pesq_unavailable(b) for b in sel]
        out[name] = sel
    return out


def select_winners_device(engine, items, pesq=None):
    """Enqueue ``cse_select_best`` for every algorithm's nominal device table: {alg: device ``cse_winner_t``
    [U][3]}.  Nothing is synchronised."""
    return {name: engine.select_device(buf, pl["n_points"], None if pesq is None else _pesq_array(pesq[name]))
            for name, _pts, buf, pl in items}


def selection_from_winners(points, winners, pesq_available):
    """{alg: host winners [U, 3]} -> {alg: [per-utterance dict as grid.select_best returns]}."""
    return {name: [best_from_winners(points[name], w[u], pesq_available) for u in range(w.shape[0])]
            for name, w in winners.items()}


PESQ_CHUNK_ITEMS = 592        # candidates per launch when their waveforms go to the host PESQ pool (2 x 114 MB in flight)


def sweep_dataset(clean, noisy, grids=DEFAULT_GRIDS, sr=16000, select=True, chunk_items=DEFAULT_CHUNK_ITEMS,
                  engine_kwargs=None, tables=True, pesq=None, pesq_scorer=None, pesq_workers=None):
    """clean, noisy: host arrays [U, L] (equal-length, 16 kHz, pair-aligned).

    The device computes every candidate's scores AND the reference's three winners per (utterance, algorithm);
    ``tables=False`` returns only the winners (48-byte records) instead of also copying the per-point score
    tables (16 B x points x utterances) to the host.  ``pesq`` = {alg: [U][P]} host-side PESQ values for the
    ``pesq`` / ``balance`` winners; or ``pesq_scorer(clean, wav, sr)`` (e.g. ``calculate_pesq``) to have every
    candidate scored by a host process pool while the sweep runs (:mod:`.pesq_pool`; costs what PESQ costs).

    Returns ``{"scores", "points", "nominal", "unique", "winners", "selection", "engine"}``."""
    if pesq_scorer is not None:
        if pesq is not None:
            raise ValueError("give either a PESQ table or a scorer, not both")
        eng = SweepEngine(clean, noisy, sr=sr, chunk_items=min(chunk_items, PESQ_CHUNK_ITEMS), **(engine_kwargs or {}))
        items, pesq = run_engine_device_with_pesq(eng, pesq_scorer, grids, pesq_workers=pesq_workers)
    else:
        eng = SweepEngine(clean, noisy, sr=sr, chunk_items=chunk_items, **(engine_kwargs or {}))
        items = run_engine_device(eng, grids)
    points = {name: pts for name, pts, _, _ in items}
    unique = sum(pl["unique"] for _, _, _, pl in items)
    winners = selection = None
    if select:
        dev = select_winners_device(eng, items, pesq)
        winners = {name: eng.winners_to_host(w, tag=(name, "winners")).copy() for name, w in dev.items()}
        if pesq is None:
            warnings.warn("selection without PESQ: only the 'stoi' winner is available", stacklevel=2)
        selection = selection_from_winners(points, winners, pesq is not None)
    scores = None
    if tables:
        scores = {name: eng.table_to_host(eng.be.view_bytes_as(buf, np.uint8, tag=(name, "table")), pl, eng.U)
                  for name, _, buf, pl in items}
    nominal = sum(len(p) for p in points.values()) * eng.U
    return {"scores": scores, "points": points, "nominal": nominal, "unique": unique * eng.U,
            "winners": winners, "selection": selection, "pesq": pesq, "engine": eng}


def sweep_pairs(pairs, grids=DEFAULT_GRIDS, sr=16000, select=True, chunk_items=DEFAULT_CHUNK_ITEMS, engine_kwargs=None,
                tables=True, in_flight=12):
    """Variable-length form of :func:`sweep_dataset`: ``pairs`` is a list of (clean, noisy) 1-D arrays
    (each pair equal length, pair-aligned, 16 kHz).  Pairs are bucketed by length - one engine (and
    one set of cached spectrograms) per distinct length - and results are returned in input order.

    A real corpus has almost as many lengths as utterances, and a bucket of one utterance launches grids far smaller
    than the GPU (a 144-candidate group on 148 SMs).  Up to ``in_flight`` buckets are therefore enqueued - each on its
    own CUDA stream, from this one host thread, nothing synchronised - before the oldest one's results are read back,
    so that the small kernels of neighbouring buckets run side by side.  Results do not depend on ``in_flight``."""
    by_len = {}
    for i, (c, n) in enumerate(pairs):
        c = np.asarray(c)
        n = np.asarray(n)
        if c.ndim != 1 or c.shape != n.shape:
            raise ValueError(f"pair {i}: clean and noisy must be 1-D arrays of equal length")
        by_len.setdefault(len(c), []).append(i)
    if not by_len:
        raise ValueError("sweep_pairs needs at least one (clean, noisy) pair")
    # longest first: every later bucket fits the blocks the caching allocator already holds (ascending order would
    # ask for a slightly larger block each time - a cudaMalloc, and its device-wide synchronisation, per bucket)
    buckets = [idx for _, idx in sorted(by_len.items(), reverse=True)]
    try:
        import torch
        use_streams = in_flight > 1 and len(buckets) > 1 and torch.cuda.is_available() and not _engine_runtime_is_emulated()
    except ImportError:
        use_streams = False

    import contextlib

    def on(stream):
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def enqueue(idx, stream):
        with on(stream):
            eng = SweepEngine(np.stack([pairs[i][0] for i in idx]), np.stack([pairs[i][1] for i in idx]), sr=sr,
                              chunk_items=chunk_items, **(engine_kwargs or {}))
            items = run_engine_device(eng, grids)
            dev_w = select_winners_device(eng, items) if select else None
        return eng, items, dev_w, stream

    def collect(job):
        eng, items, dev_w, stream = job
        with on(stream):
            w = {name: eng.winners_to_host(v).copy() for name, v in dev_w.items()} if select else None
            sc = ({name: eng.table_to_host(eng.be.view_bytes_as(buf, np.uint8), pl, eng.U).copy() for name, _, buf, pl in items}
                  if tables else None)
        return {"winners": w, "scores": sc, "points": {name: pts for name, pts, _, _ in items},
                "nominal": sum(pl["n_points"] for _, _, _, pl in items) * eng.U,
                "unique": sum(pl["unique"] for _, _, _, pl in items) * eng.U}

    outs, pending = [], []
    streams = _bucket_streams(int(in_flight)) if use_streams else [None]
    for k, idx in enumerate(buckets):
        if len(pending) >= len(streams):
            outs.append(collect(pending.pop(0)))
        pending.append(enqueue(idx, streams[k % len(streams)]))
    while pending:
        outs.append(collect(pending.pop(0)))

    scores, winners, points, nominal, unique = None, None, None, 0, 0
    for idx, out in zip(buckets, outs):
        points = out["points"]
        if tables:
            if scores is None:
                scores = {name: np.zeros((len(pairs), sc.shape[1]), dtype=sc.dtype) for name, sc in out["scores"].items()}
            for name, sc in out["scores"].items():
                scores[name][idx] = sc
        if select:
            if winners is None:
                winners = {name: np.zeros((len(pairs), 3), dtype=w.dtype) for name, w in out["winners"].items()}
            for name, w in out["winners"].items():
                winners[name][idx] = w
        nominal += out["nominal"]
        unique += out["unique"]
    return {"scores": scores, "points": points, "nominal": nominal, "unique": unique, "winners": winners,
            "selection": selection_from_winners(points, winners, False) if select else None}


_BUCKET_STREAMS = {}


def _bucket_streams(n):
    """The same streams on every call: the caching allocator keeps one pool of blocks per stream, so fresh streams
    would mean fresh cudaMallocs for every bucket."""
    import torch
    pool = _BUCKET_STREAMS.setdefault(torch.cuda.current_device(), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream())
    return pool[:n]


def _engine_runtime_is_emulated():
    from . import engine
    return engine._runtime.get("backend_factory") is not None
