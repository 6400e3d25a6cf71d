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
import contextlib
import json
import os
import re
import time
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
                pesq_workers=None, target_sr=16000, write_wavs=True, verbose=True, in_flight=12, checkpoint_seconds=2.0,
                prep_workers=None):
    """The reference's batch run over ``pairs`` ([{"stem", "clean", "noisy"}], paths or arrays; arrays already at
    16 kHz and pair-aligned may carry ``"prepared": True``).

    ``out_dirs``: {algorithm name: directory of its winner WAVs} (``results_<alg>`` in the reference);
    ``summary_dir`` receives ``all_results.json`` (also the resume record: (stem, algorithm) rows found there are
    skipped, ``:451-453``), ``summary_means.json`` and ``all_results.csv``.  ``resume=True`` additionally skips
    stems whose winner WAVs exist (``--resume``), ``start_from`` skips everything before that stem (``--start-from``).
    ``in_flight``: length buckets enqueued side by side (PESQ-free runs); ``checkpoint_seconds``: how often at most
    ``all_results.json`` is rewritten while running; ``prep_workers``: host threads of the front end (read, resample,
    align).
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
    n_before = len(all_results)
    input_order = {p["stem"]: i for i, p in enumerate(pairs)}
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
    # front end on the host (north_star): WAV read, mono fold, resampling, pair alignment - numpy / scipy / pocketfft
    # release the GIL, so a thread pool spreads the pairs over the host cores
    if prep_workers is None:
        prep_workers = max(1, min(8, (os.cpu_count() or 2) // 2))
    if prep_workers > 1 and len(todo) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(prep_workers) as ex:
            prepared = dict(zip((p["stem"] for p in todo), ex.map(lambda p: _prepare(p, target_sr), todo)))
    else:
        prepared = {p["stem"]: _prepare(p, target_sr) for p in todo}
    buckets = {}
    for p in todo:
        buckets.setdefault(len(prepared[p["stem"]][0]), []).append(p["stem"])

    streams = [None]
    if scorer is None and in_flight > 1 and len(buckets) > 1:
        from .sweep import _bucket_streams, _engine_runtime_is_emulated
        if not _engine_runtime_is_emulated():
            streams = _bucket_streams(int(in_flight))

    def on(stream):
        if stream is None:
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(stream)

    def start(L, stems, stream):
        """Everything of a bucket that only ENQUEUES device work: engine, baseline scores, the sweeps and the
        selection of every algorithm still needed (with PESQ the sweep itself waits for the pool, chunk by chunk)."""
        with on(stream):
            clean = np.stack([prepared[s][0] for s in stems])
            noisy = np.stack([prepared[s][1] for s in stems])
            chunk = PESQ_CHUNK_ITEMS if scorer is not None else None
            eng = SweepEngine(clean, noisy, sr=target_sr, **({"chunk_items": chunk} if chunk else {}))
            base = eng.baseline_device()
            base_pesq = [None] * len(stems)
            if scorer is not None:
                from .pesq_pool import PesqPool
                with PesqPool(scorer, target_sr, workers=pesq_workers) as pool:
                    for u in range(len(stems)):
                        pool.submit(u, [0], clean[u], [noisy[u]])
                    tab = pool.table(len(stems), 1)
                base_pesq = [None if np.isnan(v) else float(v) for v in tab[:, 0]]
            swept = []
            for alg_name, _fn, ranges, *_ in algorithms:
                need = [u for u, s in enumerate(stems) if (s, alg_name) not in have]
                if not need:
                    continue
                grids = ((alg_name, ranges),)
                if scorer is not None:
                    items, pesq = run_engine_device_with_pesq(eng, scorer, grids, pesq_workers=pesq_workers)
                else:
                    items, pesq = run_engine_device(eng, grids), None
                swept.append((alg_name, ranges, need, select_winners_device(eng, items, pesq)[alg_name]))
        return L, stems, stream, eng, base, base_pesq, swept

    def finish(job):
        """Winners to the host, their waveforms re-materialised and written, the rows appended."""
        L, stems, stream, eng, base, base_pesq, swept = job
        rows_by_stem = {s: [] for s in stems}
        with on(stream):
            base = eng.scores_to_host(base, 1)[:, 0]
            for alg_name, ranges, need, dev_win in swept:
                pts = cached_points(alg_name, ranges)
                win = eng.winners_to_host(dev_win).copy()
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
        for s in stems:
            all_results.extend(rows_by_stem[s])
            have.update((s, r["alg"]) for r in rows_by_stem[s])
        checkpoint()
        if verbose:
            print(f"bucket of {len(stems)} pairs x {L} samples done ({len(all_results)} rows so far)")

    last_write = [time.monotonic()]

    def checkpoint(force=False):
        """The reference rewrites the whole JSON after every pair (:455-458) - quadratic in the corpus and, at a few
        milliseconds of device work per pair, the dominant cost.  Here: at most every ``checkpoint_seconds``, and at
        the end; an interrupted run loses at most that much work."""
        if not force and time.monotonic() - last_write[0] < checkpoint_seconds:
            return
        tmp = json_path + ".part"
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump(all_results, f, indent=2, ensure_ascii=False)
        os.replace(tmp, json_path)
        last_write[0] = time.monotonic()

    # longest bucket first (the caching allocator's blocks then fit every later bucket); without PESQ up to
    # ``in_flight`` buckets are enqueued, each on its own stream, before the oldest is read back (sweep.sweep_pairs)
    pending = []
    try:
        for k, (L, stems) in enumerate(sorted(buckets.items(), reverse=True)):
            if len(pending) >= len(streams):
                finish(pending.pop(0))
            pending.append(start(L, stems, streams[k % len(streams)]))
        while pending:
            finish(pending.pop(0))
    finally:
        # rows of this run in the order the pairs were given (the reference's loop order), whatever the bucket order
        all_results[n_before:] = sorted(all_results[n_before:], key=lambda r: input_order.get(r["stem"], -1))
        checkpoint(force=True)

    summary = results_io.write_results(all_results, [a[0] for a in algorithms], summary_dir)
    return all_results, summary
