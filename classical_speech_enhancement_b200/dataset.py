"""Dataset-level run: the reference's ``main()`` loop (``Code/speech_enhancement_comparison.py:375-473``) as
batched device work.

For every prepared pair and every algorithm the reference calls ``run_algorithm_on_pair`` (``:278-338``): grid
search, three winners, three PCM16 WAVs, one result row; rows accumulate in ``all_results.json`` (rewritten
after every pair, which is also its resume record, ``:431-458``) and end in ``summary_means.json`` /
``all_results.csv`` (``:341-373,460-471``).  Here the pairs are bucketed by length, each bucket is ONE device
sweep per algorithm (scores + selection on the device), and only the winners - at most three distinct grid
points per (utterance, algorithm) - are re-materialised, by one sparse ``cse_enhance_list`` launch per noise-PSD
group, finalized with the lag the device estimated and written out.  Same files, same row keys, same resume
semantics.
"""
import json
import os
import re
import warnings

import numpy as np

from . import results_io
from .engine import SweepEngine, finalize_host
from .evaluation_metrics import calculate_pesq
from .grid import best_from_winners
from .sweep import PESQ_CHUNK_ITEMS, cached_points, run_engine_device, run_engine_device_with_pesq, select_winners_device

WAV_TAGS = (("stoi", "stoi"), ("pesq", "pesq"), ("balance", "balanced"))     # criterion -> file suffix (:300-302)


def find_pairs(data_dir):
    """``*_clean.wav`` + ``*_noisy.wav`` / ``*_noise.wav`` pairs of a directory (``_find_pairs``, ``:254-267``)."""
    wavs = [f for f in os.listdir(data_dir) if f.lower().endswith(".wav")]
    pairs = []
    for cf in wavs:
        if "_clean" not in cf.lower():
            continue
        stem = re.sub(r"(?i)_clean\.wav$", "", cf)
        named = [c for c in (f"{stem}_noisy.wav", f"{stem}_noise.wav") if c in wavs]
        loose = [f for f in wavs if f.lower().startswith(stem.lower()) and f.lower() != cf.lower()
                 and ("noise" in f.lower() or "noisy" in f.lower())]
        noisy = named[0] if named else (loose[0] if len(loose) == 1 else None)
        if noisy:
            pairs.append({"stem": stem, "clean": os.path.join(data_dir, cf), "noisy": os.path.join(data_dir, noisy)})
    return pairs


def load_wav(path):
    """(samples float32 in [-1, 1) shaped like ``librosa.load(sr=None, mono=False)``: (n,) or (channels, n), sr)."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(1 << (8 * x.dtype.itemsize - 1))
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    return (x.T if x.ndim == 2 else x).astype(np.float32), int(sr)


def processed_stems(out_dirs):
    """Stems that already have winner WAVs (the reference's ``--resume`` test, ``:406-415``)."""
    done = set()
    for d in out_dirs:
        if os.path.isdir(d):
            for name in os.listdir(d):
                if "_stoi.wav" in name:
                    parts = name.split("_")
                    if len(parts) >= 2:
                        done.add("_".join(parts[:2]))
    return done


def _prepare(pair, target_sr):
    from .speech_enhancement_comparison import prepare_pair
    c, n = pair["clean"], pair["noisy"]
    if isinstance(c, str):
        c, sr_c = load_wav(c)
    else:
        sr_c = int(pair.get("sr_clean", pair.get("sr", target_sr)))
    if isinstance(n, str):
        n, sr_n = load_wav(n)
    else:
        sr_n = int(pair.get("sr_noisy", pair.get("sr", target_sr)))
    if pair.get("prepared"):
        return np.asarray(c, dtype=np.float64), np.asarray(n, dtype=np.float64)
    clean, noisy, _ = prepare_pair(c, sr_c, n, sr_n, target_sr=target_sr, do_align=True)
    return clean, noisy


def run_dataset(pairs, out_dirs, summary_dir, *, algorithms=None, resume=False, start_from="", pesq_scorer="auto",
                pesq_workers=None, target_sr=16000, write_wavs=True, verbose=True):
    """The reference's batch run over ``pairs`` ([{"stem", "clean", "noisy"}], paths or arrays; arrays already at
    16 kHz and pair-aligned may carry ``"prepared": True``).

    ``out_dirs``: {algorithm name: directory of its winner WAVs} (``results_<alg>`` in the reference);
    ``summary_dir`` receives ``all_results.json`` (also the resume record: (stem, algorithm) rows found there are
    skipped, ``:451-453``), ``summary_means.json`` and ``all_results.csv``.  ``resume=True`` additionally skips
    stems whose winner WAVs exist (``--resume``), ``start_from`` skips everything before that stem (``--start-from``).
    Returns (all_results, summary)."""
    from .speech_enhancement_comparison import algorithms_table, write_wav_pcm16
    algorithms = algorithms or algorithms_table()
    os.makedirs(summary_dir, exist_ok=True)
    json_path = os.path.join(summary_dir, "all_results.json")
    all_results = []
    if os.path.exists(json_path):
        with open(json_path, "r", encoding="utf-8") as f:
            all_results = json.load(f)
    have = {(r.get("stem"), r.get("alg")) for r in all_results}
    pairs = list(pairs)
    if resume:
        done = processed_stems(out_dirs.values())
        pairs = [p for p in pairs if p["stem"] not in done]
    if start_from:
        idx = next((i for i, p in enumerate(pairs) if p["stem"] == start_from), 0)
        pairs = pairs[idx:]

    scorer = pesq_scorer
    if scorer == "auto":
        try:
            import pesq  # noqa: F401
            scorer = calculate_pesq
        except ImportError:
            warnings.warn("pesq is not installed: rows carry PESQ = None and only the 'stoi' winner of every "
                          "(utterance, algorithm) is produced")
            scorer = None

    todo = [p for p in pairs if any((p["stem"], a[0]) not in have for a in algorithms)]
    prepared = {p["stem"]: _prepare(p, target_sr) for p in todo}
    buckets = {}
    for p in todo:
        buckets.setdefault(len(prepared[p["stem"]][0]), []).append(p["stem"])

    for L, stems in sorted(buckets.items()):
        clean = np.stack([prepared[s][0] for s in stems])
        noisy = np.stack([prepared[s][1] for s in stems])
        chunk = PESQ_CHUNK_ITEMS if scorer is not None else None
        eng = SweepEngine(clean, noisy, sr=target_sr, **({"chunk_items": chunk} if chunk else {}))
        base = eng.baseline()
        base_pesq = [None] * len(stems)
        if scorer is not None:
            from .pesq_pool import PesqPool
            with PesqPool(scorer, target_sr, workers=pesq_workers) as pool:
                for u in range(len(stems)):
                    pool.submit(u, [0], clean[u], [noisy[u]])
                tab = pool.table(len(stems), 1)
            base_pesq = [None if np.isnan(v) else float(v) for v in tab[:, 0]]
        rows_by_stem = {s: [] for s in stems}
        for alg_name, _fn, ranges, *_ in algorithms:
            need = [u for u, s in enumerate(stems) if (s, alg_name) not in have]
            if not need:
                continue
            grids = ((alg_name, ranges),)
            if scorer is not None:
                items, pesq = run_engine_device_with_pesq(eng, scorer, grids, pesq_workers=pesq_workers)
            else:
                items, pesq = run_engine_device(eng, grids), None
            pts = cached_points(alg_name, ranges)
            win = eng.winners_to_host(select_winners_device(eng, items, pesq)[alg_name]).copy()
            best = [best_from_winners(pts, win[u], pesq_available=scorer is not None) for u in range(len(stems))]
            wanted = sorted({(u, b[c]["index"]) for u in need for b in (best[u],) for c, _ in WAV_TAGS if b[c]["index"] is not None})
            raw = eng.enhance_list(alg_name, pts, wanted) if (write_wavs and wanted) else {}
            for u in need:
                stem, b = stems[u], best[u]
                if b["stoi"]["index"] is None:
                    if verbose:
                        print(f" {stem} / {alg_name}: no valid parameters found")     # the reference raises here (:233-235)
                    continue
                if write_wavs:
                    os.makedirs(out_dirs[alg_name], exist_ok=True)
                    for crit, tag in WAV_TAGS:
                        if b[crit]["index"] is None:
                            continue
                        wav = finalize_host(raw[(u, b[crit]["index"])], b[crit]["lag"], L)
                        write_wav_pcm16(os.path.join(out_dirs[alg_name], f"{stem}_{alg_name}_optimized_{tag}.wav"), wav, target_sr)
                baseline = {"stoi": float(base[u]["stoi"]) or 0, "pesq": base_pesq[u] if scorer is not None else None,
                            "snr": (float("inf") if base[u]["flags"] & 4 else float(base[u]["snr"])) or 0}
                rows_by_stem[stem].append(results_io.result_row(alg_name, stem, target_sr, baseline, b))
        for s in stems:                                       # the reference appends and rewrites the JSON per pair (:455-458)
            all_results.extend(rows_by_stem[s])
            have.update((s, r["alg"]) for r in rows_by_stem[s])
        with open(json_path, "w", encoding="utf-8") as f:
            json.dump(all_results, f, indent=2, ensure_ascii=False)
        if verbose:
            print(f"bucket of {len(stems)} pairs x {L} samples done ({len(all_results)} rows so far)")

    summary = results_io.write_results(all_results, [a[0] for a in algorithms], summary_dir)
    return all_results, summary
