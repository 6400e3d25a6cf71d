"""Oracle grid search and selection (test infrastructure; see oracle/__init__.py).

Follows ``Code/speech_enhancement_comparison.py``: grid enumeration ``:149-150``,
the per-candidate body ``:156-220``, the three running-best updates with
hysteresis ``:186-216``, the result dict ``:237-252`` and the ``true_noise``
routing of ``run_algorithm_on_pair.algorithm_wrapper`` ``:282-292``.
"""
from itertools import product

import numpy as np

from .metrics import combined_score, global_snr
from .postprocess import finalize_enhanced
from .intelligibility import stoi

TOL = {"stoi": 1e-6, "pesq": 1e-3, "balance": 1e-5}


def grid_points(param_ranges):
    """Parameter dicts in the reference's order: dict insertion order, last key fastest."""
    names = list(param_ranges.keys())
    return [dict(zip(names, values)) for values in product(*param_ranges.values())]


def score_candidate(clean, enhanced, sr, pesq_fn=None):
    """finalize -> clip -> STOI / PESQ / SNR for one candidate (``:171-184``).
    Returns None when the reference would ``continue``."""
    if enhanced is None or len(enhanced) == 0:
        return None
    enhanced = finalize_enhanced(np.asarray(enhanced, dtype=np.float64), clean, sr, do_align=True)
    if enhanced is None:
        return None
    enhanced = np.clip(enhanced, -1.0, 1.0)
    m = min(len(clean), len(enhanced))
    s = stoi(clean[:m], enhanced[:m], sr, extended=False)
    p = pesq_fn(clean, enhanced, sr) if pesq_fn is not None else 0.0
    if s is None or p is None:
        return None
    return {"stoi": s, "pesq": p, "snr": global_snr(clean, enhanced), "enhanced": enhanced}


def select_best(points, scores):
    """The sequential three-way best-of (``:124-147,186-216``).

    ``scores[i]`` is None (candidate skipped) or a dict with stoi/pesq/snr.
    Returns ``{criterion: {"index", "score", "params", other metrics}}``; index is
    None when no candidate was valid (the reference then raises ValueError, ``:233-235``).
    """
    best = {c: {"index": None, "score": -1, "params": {}} for c in TOL}
    for i, sc in enumerate(scores):
        if sc is None:
            continue
        vals = {"stoi": sc["stoi"], "pesq": sc["pesq"],
                "balance": combined_score(sc["stoi"], sc["pesq"])}
        for c in TOL:
            if vals[c] > best[c]["score"] + TOL[c]:
                best[c] = {"index": i, "score": vals[c], "params": dict(points[i]),
                           "stoi": sc["stoi"], "pesq": sc["pesq"], "snr": sc["snr"]}
    return best


def sweep_one_pair(clean, noisy, sr, algorithm, param_ranges, pesq_fn=None, points=None,
                   keep_waveforms=False):
    """The reference's loop for one (pair, algorithm): every candidate recomputes its
    STFT and noise PSD (no cross-candidate caching), exactly as the reference does.
    Returns (points, scores, best)."""
    if points is None:
        points = grid_points(param_ranges)
    scores = []
    for pd in points:
        try:
            if pd.get("noise_method") == "true_noise":
                enh = algorithm(noisy, sr, clean_audio=clean, **pd)
            else:
                enh = algorithm(noisy, sr, **pd)
            sc = score_candidate(clean, enh, sr, pesq_fn)
            if sc is not None and not keep_waveforms:
                sc.pop("enhanced")
            scores.append(sc)
        except Exception:                                                 # :218-220
            scores.append(None)
    return points, scores, select_best(points, scores)
