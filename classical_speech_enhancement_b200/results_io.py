"""Result rows and files of a batch run, in the schema ``Code/evaluation/statistics.py`` loads
(``Code/speech_enhancement_comparison.py:314-338`` rows, ``:341-373`` summary, ``:460-471`` CSV): one table of
(selection, metric) pairs drives the row builder, the per-algorithm means and the CSV columns, so the three
cannot drift apart.  Host-side I/O, outside the timed path."""
import json
import os

import numpy as np

#: selection tag in the file schema -> criterion key of the winners dict (None = the unprocessed noisy signal)
SELECTIONS = (("noisy", None), ("stoiopt", "stoi"), ("pesqopt", "pesq"), ("balopt", "balance"))
METRICS = ("stoi", "pesq", "snr")
#: every metric column of a row, reference order: stoi_noisy, pesq_noisy, snr_noisy, stoi_stoiopt, ...
ROW_METRIC_FIELDS = tuple(f"{m}_{tag}" for tag, _ in SELECTIONS for m in METRICS)
#: the subset the reference reports in the CSV and averages in the summary: no SNR except for the balanced winner
REPORTED = tuple(f for f in ROW_METRIC_FIELDS if not f.startswith("snr_") or f == "snr_balopt")
PARAM_FIELDS = (("best_params_stoi", "stoi"), ("best_params_pesq", "pesq"), ("best_params_balanced", "balance"))
CSV_HEADER = ["stem", "alg", *REPORTED]


def result_row(alg_name, stem, sr, baseline, best):
    """One (utterance, algorithm) row.  ``baseline``: {"stoi", "pesq", "snr"} of the noisy signal; ``best``:
    {criterion: {"score", "stoi", "pesq", "snr", "params"}} as ``optimize_parameters`` / ``grid.best_from_winners``
    produce them (an unavailable criterion - no PESQ - yields None metrics and empty parameters)."""
    row = {"alg": alg_name, "stem": stem, "sr": sr}
    for tag, crit in SELECTIONS:
        src = baseline if crit is None else best.get(crit, {})
        for m in METRICS:
            v = src.get(m)
            if crit is not None and v is None and m == crit:        # the criterion's own metric is stored as "score"
                v = src.get("score")
            row[f"{m}_{tag}"] = v
    for field, crit in PARAM_FIELDS:
        row[field] = best.get(crit, {}).get("params", {})
    return row


def _fmt(x, digits=4):
    return "NA" if x is None else f"{x:.{digits}f}"


def _mean(values):
    vals = [v for v in values if v is not None]
    return float(np.mean(vals)) if vals else None


def compute_summary(all_results, alg_names):
    """Per-algorithm row count and the mean of every reported column (None values skipped)."""
    summary = {}
    for alg in alg_names:
        rows = [r for r in all_results if r["alg"] == alg]
        summary[alg] = {"count": len(rows), **{f"{col}_mean": _mean([r.get(col) for r in rows]) for col in REPORTED}}
    return summary


def write_results(all_results, alg_names, summary_dir):
    os.makedirs(summary_dir, exist_ok=True)
    with open(os.path.join(summary_dir, "all_results.json"), "w", encoding="utf-8") as f:
        json.dump(all_results, f, indent=2, ensure_ascii=False)
    summary = compute_summary(all_results, alg_names)
    with open(os.path.join(summary_dir, "summary_means.json"), "w", encoding="utf-8") as f:
        json.dump(summary, f, indent=2, ensure_ascii=False)
    with open(os.path.join(summary_dir, "all_results.csv"), "w", encoding="utf-8") as f:
        f.write(",".join(CSV_HEADER) + "\n")
        for r in all_results:
            f.write(",".join([r["stem"], r["alg"], *(_fmt(r.get(col)) for col in REPORTED)]) + "\n")
    return summary
