"""Host-side grid logic: enumeration, grouping, dead-parameter dedupe, param packing, selection.

Mirrors ``Code/speech_enhancement_comparison.py``: candidates are the ``itertools.product`` of
the range dict in insertion order, last key fastest (``:149-150``); the three winners come from
a sequential scan with hysteresis (``:186-216``) and therefore stay on the host, in grid order.
"""
from collections import OrderedDict
from itertools import product

from ._lib import ALG_MMSE, ALG_OMLSA, ALG_SS, ALG_WIENER

#: reference algorithm names (``speech_enhancement_comparison.py:395-401``) -> C-ABI ids
ALGORITHM_IDS = {"spectralSubtractor": ALG_SS, "wiener": ALG_WIENER, "mmse": ALG_MMSE, "omlsa": ALG_OMLSA}
NOISE_METHODS = ("percentile", "min_tracking", "true_noise")
TOL = {"stoi": 1e-6, "pesq": 1e-3, "balance": 1e-5}      # :186, :196, :206


def alg_eps(alg):
    """Epsilon each reference entry point hands to ``noise_estimation`` (``mmse.py:17`` vs the rest)."""
    return 1e-12 if alg == ALG_MMSE else 1e-10


def grid_points(param_ranges):
    names = list(param_ranges.keys())
    return [dict(zip(names, values)) for values in product(*param_ranges.values())]


def noise_key(point, eps):
    """Cache key of the noise PSD a point needs.  ``min_tracking`` and ``true_noise`` ignore
    ``noise_percentile`` (``Code/noise_estimation.py:60-95,115-155``)."""
    method = point["noise_method"]
    if method not in NOISE_METHODS:
        raise ValueError(f"Unbekannte Methode: {method}")
    pct = float(point["noise_percentile"]) if method == "percentile" else None
    return (int(point["n_fft"]), int(point["hop_length"]), method, pct, eps)


def param_row(alg, point, noise_time_varying):
    """The ``cse_params.v`` row of one grid point (layout documented in include/cse.h).

    ``noise_mu`` only acts when the noise PSD is time-varying and the method is not
    ``true_noise`` (``mmse.py:48``, ``advanced_mmse.py:60``); otherwise it is packed as -1 so
    that candidates differing only in that dead parameter collapse onto one device candidate.
    """
    smooth = noise_time_varying and point["noise_method"] != "true_noise"
    if alg == ALG_SS:
        return (float(point["alpha"]), float(point["beta"]))
    if alg == ALG_WIENER:
        return (float(point["alpha"]), float(point["gain_floor"]))
    if alg == ALG_MMSE:
        mu = float(point.get("noise_mu", 0.98)) if smooth else -1.0
        return (float(point["alpha"]), float(point["ksi_min"]), float(point["gain_min"]),
                float(point["gain_max"]), mu)
    if alg == ALG_OMLSA:
        mu = float(point["noise_mu"]) if smooth else -1.0
        return (float(point["alpha"]), float(point["ksi_min"]), float(point["gain_floor"]), mu,
                float(point["q"]), float(point.get("v_max", 80.0)))
    raise ValueError(f"unknown algorithm id {alg}")


def plan(alg, points, n_frames_of, split_mu=False):
    """Group grid points by the noise PSD they share and dedupe identical device candidates.

    Returns an ordered dict ``noise_key -> {"rows": [unique param rows], "members": [[grid
    indices sharing row r], ...]}``.  ``n_frames_of(n_fft, hop)`` tells whether the PSD is
    time-varying (the short-signal rule of ``noise_estimation.py:194-195`` makes it static).

    ``split_mu``: the group key additionally carries the effective ``noise_mu`` (None where it is dead), and the
    rows carry -1 in its place: all candidates of a group then share the smoothed PSD and with it the whole
    candidate-invariant front of the gain rule (``cse_gamma``).
    """
    groups = OrderedDict()
    eps = alg_eps(alg)
    for i, pt in enumerate(points):
        key = noise_key(pt, eps)
        n_frames = n_frames_of(key[0], key[1])
        tv = key[2] != "percentile" and n_frames >= 5
        row = param_row(alg, pt, tv)
        if split_mu and alg != ALG_SS:
            mu_slot = {ALG_MMSE: 4, ALG_OMLSA: 3}.get(alg)
            mu = row[mu_slot] if mu_slot is not None else -1.0
            key = key + (mu if mu >= 0 else None,)
            if mu_slot is not None:
                row = row[:mu_slot] + (-1.0,) + row[mu_slot + 1:]
        g = groups.setdefault(key, {"rows": [], "members": [], "index": {}, "time_varying": tv})
        r = g["index"].get(row)
        if r is None:
            r = len(g["rows"])
            g["index"][row] = r
            g["rows"].append(row)
            g["members"].append([])
        g["members"][r].append(i)
    for g in groups.values():
        del g["index"]
    return groups


def combined_score(stoi, pesq):
    """``calculate_combined_speech_score`` (``Code/evaluation_metrics.py:104-114``)."""
    if stoi is None:
        stoi = 0
    if pesq is None:
        pesq = 0
    return 0.5 * stoi + 0.5 * (max(0, pesq) / 4.5)


def select_best(points, stoi, pesq, snr, valid):
    """Sequential best-of with hysteresis over candidates in grid order; ``pesq[i]`` may be None
    (candidate skipped, ``:180-181``).  Returns ``{criterion: dict}`` with index None if no
    candidate qualified."""
    best = {c: {"index": None, "score": -1, "params": {}} for c in TOL}
    for i in range(len(points)):
        if not valid[i] or pesq[i] is None or stoi[i] is None:
            continue
        vals = {"stoi": stoi[i], "pesq": pesq[i], "balance": combined_score(stoi[i], pesq[i])}
        for c in TOL:
            if vals[c] > best[c]["score"] + TOL[c]:
                best[c] = {"index": i, "score": vals[c], "params": dict(points[i]),
                           "stoi": stoi[i], "pesq": pesq[i], "snr": snr[i]}
    return best


CRITERIA = tuple(TOL)      # record order of cse_winner_t rows: stoi, pesq, balance


def best_from_winners(points, wrow, pesq_available=True):
    """One utterance's ``cse_winner_t[3]`` (device selection, ``cse_select_best``) -> the dict ``select_best``
    returns.  Without PESQ the ``pesq`` / ``balance`` entries are marked unavailable instead of carrying the
    meaningless "first valid point" a constant PESQ of 0.0 would select."""
    best = {}
    for k, c in enumerate(CRITERIA):
        w = wrow[k]
        i = int(w["index"])
        if c != "stoi" and not pesq_available:
            best[c] = {"index": None, "score": None, "params": {}, "unavailable": "PESQ was not computed"}
        elif i < 0:
            best[c] = {"index": None, "score": -1, "params": {}}
        else:
            best[c] = {"index": i, "score": float(w["score"]), "params": dict(points[i]), "stoi": float(w["stoi"]),
                       "pesq": float(w["pesq"]), "snr": float(w["snr"]), "lag": int(w["lag"])}
    return best


def select_best_batch(points, stoi, pesq, snr, valid):
    """``select_best`` for U utterances at once: the scan is sequential over the grid (hysteresis makes
    it order-dependent) but independent per utterance, so every step is one vector operation over the
    utterances.  ``stoi``, ``snr``: float arrays [U, P]; ``valid``: bool [U, P]; ``pesq``: float array
    [U, P] with NaN where the reference's ``calculate_pesq`` returned None (candidate skipped), or None
    for "PESQ = 0.0 everywhere".  Returns a list of U dicts identical to ``select_best``'s."""
    import numpy as np
    stoi = np.asarray(stoi, dtype=np.float64)
    snr = np.asarray(snr, dtype=np.float64)
    U, P = stoi.shape
    pq = np.zeros((U, P)) if pesq is None else np.asarray(pesq, dtype=np.float64)
    ok = np.asarray(valid, dtype=bool) & ~np.isnan(pq)
    vals = {"stoi": stoi, "pesq": pq, "balance": 0.5 * stoi + 0.5 * (np.maximum(0, pq) / 4.5)}
    best_score = {c: np.full(U, -1.0) for c in TOL}
    best_index = {c: np.full(U, -1, dtype=np.int64) for c in TOL}
    for i in range(P):
        oki = ok[:, i]
        if not oki.any():
            continue
        for c in TOL:
            upd = oki & (vals[c][:, i] > best_score[c] + TOL[c])
            if upd.any():
                best_score[c] = np.where(upd, vals[c][:, i], best_score[c])
                best_index[c][upd] = i
    out = []
    for u in range(U):
        best = {}
        for c in TOL:
            i = int(best_index[c][u])
            if i < 0:
                best[c] = {"index": None, "score": -1, "params": {}}
            else:
                best[c] = {"index": i, "score": float(vals[c][u, i]), "params": dict(points[i]),
                           "stoi": float(stoi[u, i]), "pesq": float(pq[u, i]), "snr": float(snr[u, i])}
        out.append(best)
    return out
