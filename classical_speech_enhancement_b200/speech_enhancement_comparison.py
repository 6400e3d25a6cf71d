"""Drop-in for the hot path of the reference's ``Code/speech_enhancement_comparison.py``:
``optimize_parameters`` (``:109-252``) and ``run_algorithm_on_pair`` (``:278-338``) with their
helpers.  The grid of one (pair, algorithm) is executed as ONE batched device sweep
(enhance + finalize + STOI + SNR for every grid point) instead of the reference's sequential
Python loop; the three winners are then chosen on the host by the reference's sequential
hysteresis scan, in grid order, and only the winners' waveforms are re-materialised.

PESQ: the reference scores every candidate with the ``pesq`` C extension (``:178``).  When that
package is importable this module does the same - the candidates' finalized waveforms are handed slab-wise
to a host process pool (:mod:`.pesq_pool`) while the device computes the next slab - for exact parity of
all three selections; any ``scorer(clean, wav, sr)`` can be injected in its place.  When it is not (this image), ``pesq_scorer`` must be injected, or PESQ is taken as 0.0
for every candidate with a warning: the ``stoi`` winner is then exact, the ``pesq``/``balance``
winners are not comparable with the reference.
"""
import inspect
import os
import warnings
from typing import Any, Dict

import numpy as np
from scipy.signal import correlate

from .engine import SweepEngine, finalize_host
from .evaluation_metrics import calculate_combined_speech_score, calculate_pesq
from .grid import ALGORITHM_IDS, best_from_winners, grid_points  # noqa: F401
from .parameter_ranges import (param_ranges_mmse, param_ranges_omlsa, param_ranges_ss,  # noqa: F401
                               param_ranges_wiener)


def to_mono(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return x
    return np.mean(x, axis=1) if x.shape[0] >= x.shape[1] else np.mean(x, axis=0)


def match_length(x, L):
    x = np.asarray(x, dtype=np.float64)
    if len(x) > L:
        return x[:L]
    if len(x) < L:
        return np.pad(x, (0, L - len(x)))
    return x


def align_to_reference(ref, sig, sr, max_shift_s=0.10, corr_seconds=2.0):
    """Host-side pair alignment (front end, outside the timed path; BASELINE north_star keeps
    load / resample / pair alignment on the host).  Same arithmetic as the reference ``:38-69``."""
    ref = np.asarray(ref, dtype=np.float64)
    sig = np.asarray(sig, dtype=np.float64)
    N = int(min(len(ref), len(sig), corr_seconds * sr))
    if N < 256:
        return sig
    c = correlate(ref[:N] - np.mean(ref[:N]), sig[:N] - np.mean(sig[:N]), mode="full", method="auto")
    lags = np.arange(-N + 1, N)
    keep = np.abs(lags) <= int(max_shift_s * sr)
    if not np.any(keep):
        return sig
    lag = int(lags[keep][np.argmax(c[keep])])
    if lag > 0:
        return np.pad(sig, (lag, 0))
    if lag < 0:
        return sig[-lag:]
    return sig


_RESAMPLER_FIRS = {}


def _soxr_hq_like_fir(up, down, sr_in):
    """Anti-alias / anti-image FIR of the rational resampler, at the rate ``sr_in * up``: the specification of
    soxr's HQ recipe (what ``librosa.resample`` uses by default) - pass band to 0.913 of the lower Nyquist
    frequency, stop band from that Nyquist frequency on, linear phase, Kaiser window - with a 140 dB stop band.
    soxr itself is not installable here; on the reference's own 48 kHz files this design reproduces its published
    per-file results to 1.1e-7 STOI / 3e-5 dB SNR (``tests/test_oracle_pinning.py``; scipy's default
    ``resample_poly`` filter was off by 1.2e-5 / 3.4e-3)."""
    key = (up, down, sr_in)
    h = _RESAMPLER_FIRS.get(key)
    if h is None:
        from scipy.signal import firwin, kaiserord
        rate = float(sr_in) * up
        nyq = 0.5 * min(sr_in, sr_in * up / down)
        f_pass, f_stop = 0.913 * nyq, nyq
        numtaps, beta = kaiserord(140.0, (f_stop - f_pass) / (0.5 * rate))
        h = firwin(numtaps | 1, 0.5 * (f_pass + f_stop), window=("kaiser", beta), fs=rate) * up
        _RESAMPLER_FIRS[key] = h
    return h


def resample_to(x, sr_in, sr_out):
    """``librosa.resample(x, orig_sr=sr_in, target_sr=sr_out)`` (``:23-27``; soxr_hq, output length
    ceil(n * sr_out / sr_in), float32 in -> float32 out) as a polyphase FIR resampler with a soxr-HQ-like filter."""
    if sr_in == sr_out:
        return x
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(int(sr_in), int(sr_out))
    up, down = int(sr_out) // g, int(sr_in) // g
    x = np.asarray(x)
    h = _soxr_hq_like_fir(up, down, int(sr_in))
    if up == 1 and x.ndim == 1 and len(x) >= len(h):
        # pure decimation (48 -> 16 kHz, the reference's corpus): the same sums as resample_poly - output m is
        # (x * h)[m * down + (len(h) - 1) / 2] - by overlap-add FFT convolution, 4 x faster for the 637-tap filter
        from scipy.signal import oaconvolve
        y = oaconvolve(x.astype(np.float64), h)[(len(h) - 1) // 2::down][:-(-len(x) // down)]
    else:
        y = resample_poly(x.astype(np.float64), up, down, axis=-1, window=h)
    return y.astype(np.float32) if x.dtype == np.float32 else y


def prepare_pair(clean, sr_c, noisy, sr_n, target_sr=16000, do_align=True):
    clean = resample_to(to_mono(clean), sr_c, target_sr)
    noisy = resample_to(to_mono(noisy), sr_n, target_sr)
    L = min(len(clean), len(noisy))
    clean, noisy = clean[:L], noisy[:L]
    if do_align:
        noisy = match_length(align_to_reference(clean, noisy, target_sr, 0.10, 2.0), len(clean))
    return clean, noisy, target_sr


def finalize_enhanced(enhanced, clean_ref, sr, do_align=True):
    """Align / length-match / finite-check / clip one waveform (``:92-106``); the lag comes from
    the device alignment kernel."""
    enhanced = to_mono(enhanced)
    clean_ref = np.asarray(clean_ref, dtype=np.float64)
    if len(enhanced) != len(clean_ref) or not do_align:
        if do_align:
            enhanced = align_to_reference(clean_ref, enhanced, sr, 0.10, 2.0)
        enhanced = match_length(enhanced, len(clean_ref))
        return None if not np.all(np.isfinite(enhanced)) else np.clip(enhanced, -1.0, 1.0)
    eng = SweepEngine(clean_ref[None, :], enhanced[None, :])
    sc = eng.score_waveforms(enhanced[None, None, :].astype(eng.real), finalize=True)[0, 0]
    if not sc["flags"] & 1:
        return None
    return finalize_host(enhanced, int(sc["lag"]), len(clean_ref))


_last_engine = {"key": None, "engine": None}


def _engine_for_pair(clean, noisy, sr):
    """The reference's ``main`` calls ``run_algorithm_on_pair`` four times per pair (``:447-455``), one algorithm each:
    the engine of the most recent pair - waveforms on the device, clean-side scoring caches, STFTs, noise PSDs - is kept
    and reused when the SAME samples come back (content checksum, not object identity), so only the first of the
    four calls pays for them."""
    import zlib
    from . import engine as _engine
    key = (sr, clean.shape, noisy.shape, zlib.crc32(clean.tobytes()), zlib.crc32(noisy.tobytes()),
           id(_engine._runtime["lib"]), id(_engine._runtime["backend_factory"]), bool(_engine._runtime.get("gamma")))
    if _last_engine["key"] != key:
        _last_engine["engine"] = None                      # release the previous pair's buffers first
        _last_engine["engine"] = SweepEngine(clean[None, :], noisy[None, :], sr=sr)
        _last_engine["key"] = key
    return _last_engine["engine"]


def _resolve_algorithm(algorithm_function, _depth=0):
    """Which of the four device algorithms a callable stands for, or None.

    The reference hands ``optimize_parameters`` a closure (``algorithm_wrapper``, ``:282-294``) around the real
    entry point, so the callable is unwrapped the ways Python offers: the ``__cse_algorithm__`` tag this
    package's entry points carry, ``functools.wraps`` / ``partial`` chains, and the free variables of a closure."""
    if algorithm_function is None or _depth > 4:
        return None
    name = getattr(algorithm_function, "__cse_algorithm__", None)
    if name is not None:
        return name
    for attr in ("__wrapped__", "func"):
        inner = getattr(algorithm_function, attr, None)
        if callable(inner):
            name = _resolve_algorithm(inner, _depth + 1)
            if name is not None:
                return name
    found = {_resolve_algorithm(c.cell_contents, _depth + 1) for c in (getattr(algorithm_function, "__closure__", None) or ())
             if _cell_is_callable(c)}
    found.discard(None)
    return found.pop() if len(found) == 1 else None


def _cell_is_callable(cell):
    try:
        return callable(cell.cell_contents)
    except ValueError:          # empty cell
        return False


def _sweep_foreign_callable(eng, algorithm_function, clean, noisy, sr, points):
    """``optimize_parameters`` for a callable that is none of the device algorithms: the reference's own loop -
    one call per grid point (``:165``) - with finalize / STOI / SNR of the returned waveforms batched on the
    device.  Returns (score records [P], {index: finalized waveform})."""
    sc = np.zeros(len(points), dtype=eng.lib.score_dtype)
    finalized = {}
    slab_idx, slab = [], []

    def flush():
        if slab:
            out = eng.score_waveforms(np.stack(slab)[None], finalize=True)[0]
            for j, i in enumerate(slab_idx):
                sc[i] = out[j]
                if out[j]["flags"] & 1:
                    finalized[i] = finalize_host(slab[j], int(out[j]["lag"]), len(clean))
            slab_idx.clear()
            slab.clear()

    for i, p in enumerate(points):
        try:
            enhanced = algorithm_function(noisy, sr, **p)
            if enhanced is None or len(enhanced) == 0:
                continue
            enhanced = to_mono(np.asarray(enhanced, dtype=np.float64))
        except Exception as e:                                    # the reference prints and skips (:226-228)
            print(f" Warning with params {p}: {e}")
            continue
        if len(enhanced) == len(clean):
            slab_idx.append(i)
            slab.append(enhanced)
            if len(slab) == 64:
                flush()
        else:                                                     # other lengths: host finalize (:92-106), device scoring
            fin = finalize_enhanced(enhanced, clean, sr, do_align=True)
            if fin is not None:
                one = eng.score_waveforms(fin[None, None, :], finalize=False)[0, 0]
                sc[i] = one
                sc[i]["flags"] |= 1
                finalized[i] = fin
    flush()
    return sc, finalized


def optimize_parameters(clean_reference, noisy_audio, sr, algorithm_function, param_ranges, *,
                        pesq_scorer="auto", pesq_workers=None, engine=None, verbose=True) -> Dict[str, Any]:
    """Brute-force grid search for STOI, PESQ and balance winners (reference ``:109-252``).

    ``algorithm_function`` may be one of this package's four entry points, the reference's
    ``algorithm_wrapper`` closure around one (or any ``functools.wraps`` / ``partial`` wrapper): the whole grid
    then runs as one batched device sweep.  Any other callable is executed one grid point at a time, as the
    reference does, with a warning; finalize and scoring still run on the device.

    Returns the reference's dict: ``stoi`` / ``pesq`` / ``balance`` (score, params, enhanced, the
    other metrics, snr), ``baseline`` and ``improvements``."""
    alg_name = _resolve_algorithm(algorithm_function)
    clean = np.asarray(clean_reference, dtype=np.float64)
    noisy = np.asarray(noisy_audio, dtype=np.float64)
    eng = engine if engine is not None else _engine_for_pair(clean, noisy, sr)
    from .sweep import cached_points
    points = cached_points(alg_name or "custom", param_ranges)     # the grid and its launch plan are reused from pair to pair
    if verbose:
        print(f"\n{'=' * 60}\nParameter Optimization\n{'=' * 60}")
        print(f"Testing {len(points)} parameter combinations")

    base = eng.baseline()[0]
    scorer = pesq_scorer
    if scorer == "auto":
        try:
            import pesq  # noqa: F401
            scorer = calculate_pesq
        except ImportError:
            warnings.warn("pesq is not installed: PESQ taken as 0.0 for every candidate; only the "
                          "'stoi' winner is comparable with the reference")
            scorer = None
    baseline_stoi = float(base["stoi"]) or 0
    baseline_pesq = (scorer(clean, noisy, sr) if scorer else 0.0) or 0
    baseline_snr = (float("inf") if base["flags"] & 4 else float(base["snr"])) or 0
    baseline_comp = calculate_combined_speech_score(baseline_stoi, baseline_pesq)

    foreign = None
    if alg_name is None:
        warnings.warn("optimize_parameters: the callable is not one of the device algorithms (or a wrapper around "
                      "one); running it one grid point at a time as the reference does")
        sc, foreign = _sweep_foreign_callable(eng, algorithm_function, clean, noisy, sr, points)
        table_dev = eng.be.from_host(sc.view(np.uint8).reshape(-1))
    else:
        table_dev, _ = eng.sweep_device(alg_name, points)
        sc = eng.table_to_host(eng.be.view_bytes_as(table_dev, np.uint8), {"n_points": len(points)}, 1)[0]
    valid = (sc["flags"] & 1) != 0
    pesq_tab = None
    if scorer is not None:
        # the reference scores every candidate with PESQ (:178): waveforms are re-materialised in slabs on
        # the device and scored by a host process pool while the next slab is computed
        from .pesq_pool import PesqPool
        with PesqPool(scorer, sr, workers=pesq_workers) as pool:
            slab = 64
            for s0 in range(0, len(points), slab):
                idx = [i for i in range(s0, min(s0 + slab, len(points))) if valid[i]]
                if not idx:
                    continue
                if foreign is not None:
                    wavs = [foreign[i] for i in idx]
                else:
                    raw = eng.enhance(alg_name, [points[i] for i in idx])[0]
                    wavs = [finalize_host(raw[j], int(sc["lag"][i]), len(clean)) for j, i in enumerate(idx)]
                pool.submit(0, idx, clean, wavs)
            pesq_tab = pool.table(1, len(points))
    win = eng.winners_to_host(eng.select_device(table_dev, len(points), pesq_tab))[0]
    best = best_from_winners(points, win)

    for key in ("stoi", "pesq", "balance"):
        if best[key]["index"] is None:
            raise ValueError(f"Optimization failed for {key} - no valid parameters found!")
    winners = sorted({best[k]["index"] for k in best})
    if foreign is not None:
        final = {i: foreign[i] for i in winners}
    else:
        wav = eng.enhance(alg_name, [points[i] for i in winners])[0]
        final = {i: finalize_host(wav[j], int(sc["lag"][i]), len(clean)) for j, i in enumerate(winners)}

    def pack(key, others):
        b = best[key]
        d = {"score": b["score"], "params": b["params"], "enhanced": final[b["index"]].copy()}
        for o in others:
            d[o] = b[o]
        d["snr"] = b["snr"]
        return d

    results = {"stoi": pack("stoi", ["pesq"]), "pesq": pack("pesq", ["stoi"]),
               "balance": pack("balance", ["stoi", "pesq"])}
    if verbose:
        print(f"Best STOI: {results['stoi']['score']:.4f} | Best PESQ: {results['pesq']['score']:.2f} | "
              f"Best Balance: {results['balance']['score']:.4f}")
    return {
        "stoi": results["stoi"], "pesq": results["pesq"], "balance": results["balance"],
        "baseline": {"stoi": baseline_stoi, "pesq": baseline_pesq, "snr": baseline_snr, "balance": baseline_comp},
        "improvements": {"stoi": results["stoi"]["score"] - baseline_stoi,
                         "pesq": results["pesq"]["score"] - baseline_pesq,
                         "balance": results["balance"]["score"] - baseline_comp},
    }


def write_wav_pcm16(path, x, sr):
    """``sf.write(path, float32_array, sr)`` stores PCM16 (``:306-312``); libsndfile scales by 0x7FFF."""
    from scipy.io import wavfile
    x = np.asarray(x, dtype=np.float32)
    wavfile.write(path, sr, np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16))


def run_algorithm_on_pair(alg_name, alg_fn, param_ranges, clean, noisy, sr, out_dir, stem, *,
                          pesq_scorer="auto", pesq_workers=None, engine=None, verbose=True):
    """Optimise one algorithm for one pair; save the three winners and return the reference's
    result row (``:278-338``).  ``true_noise`` points get the clean signal routed in, as
    ``algorithm_wrapper`` does (``:282-292``) - here the engine owns it."""
    if "clean_audio" not in inspect.signature(alg_fn).parameters and any(
            m == "true_noise" for m in param_ranges.get("noise_method", [])):
        raise ValueError(f"{alg_name} does not support 'true_noise' (no clean_audio parameter)")
    opt = optimize_parameters(clean, noisy, sr, alg_fn, param_ranges, pesq_scorer=pesq_scorer, pesq_workers=pesq_workers,
                              engine=engine, verbose=verbose)
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        for key, tag in (("stoi", "stoi"), ("pesq", "pesq"), ("balance", "balanced")):
            write_wav_pcm16(os.path.join(out_dir, f"{stem}_{alg_name}_optimized_{tag}.wav"), opt[key]["enhanced"], sr)
    from .results_io import result_row
    return result_row(alg_name, stem, sr, opt["baseline"], opt)


def algorithms_table():
    """(name, function, grid) rows of the reference's ``main`` (``:395-401``)."""
    from .advanced_mmse import advanced_mmse
    from .mmse import mmse
    from .spectral_subtractor import spectral_subtraction
    from .wiener_filter import wiener_filter
    return [("spectralSubtractor", spectral_subtraction, param_ranges_ss),
            ("mmse", mmse, param_ranges_mmse),
            ("wiener", wiener_filter, param_ranges_wiener),
            ("omlsa", advanced_mmse, param_ranges_omlsa)]
