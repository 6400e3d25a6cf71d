"""Result files of a batch run, byte-compatible with what the reference's ``main`` writes
(``Code/speech_enhancement_comparison.py:341-373,457-471``): ``all_results.json``,
``summary_means.json`` and ``all_results.csv``, so that ``Code/evaluation/statistics.py`` can consume
them unchanged.  Host-side I/O, outside the timed path."""
import json
import os

import numpy as np

CSV_HEADER = ["stem", "alg", "stoi_noisy", "pesq_noisy", "stoi_stoiopt", "pesq_stoiopt", "stoi_pesqopt",
              "pesq_pesqopt", "stoi_balopt", "pesq_balopt", "snr_balopt"]


def _fmt(x, digits=4):
    return "NA" if x is None else f"{x:.{digits}f}"


def compute_summary(all_results, alg_names):
    def safe_mean(values):
        valid = [v for v in values if v is not None]
        return float(np.mean(valid)) if valid else None

    summary = {}
    for alg in alg_names:
        rows = [r for r in all_results if r["alg"] == alg]
        summary[alg] = {
            "count": len(rows),
            "stoi_noisy_mean": safe_mean([r["stoi_noisy"] for r in rows]),
            "pesq_noisy_mean": safe_mean([r["pesq_noisy"] for r in rows]),
            "stoi_stoiopt_mean": safe_mean([r["stoi_stoiopt"] for r in rows]),
            "pesq_stoiopt_mean": safe_mean([r["pesq_stoiopt"] for r in rows]),
            "stoi_pesqopt_mean": safe_mean([r["stoi_pesqopt"] for r in rows]),
            "pesq_pesqopt_mean": safe_mean([r["pesq_pesqopt"] for r in rows]),
            "stoi_balopt_mean": safe_mean([r.get("stoi_balopt") for r in rows]),
            "pesq_balopt_mean": safe_mean([r.get("pesq_balopt") for r in rows]),
            "snr_balopt_mean": safe_mean([r.get("snr_balopt") for r in rows]),
        }
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
            row = [r["stem"], r["alg"], _fmt(r["stoi_noisy"]), _fmt(r["pesq_noisy"]), _fmt(r["stoi_stoiopt"]),
                   _fmt(r["pesq_stoiopt"]), _fmt(r["stoi_pesqopt"]), _fmt(r["pesq_pesqopt"]),
                   _fmt(r.get("stoi_balopt")), _fmt(r.get("pesq_balopt")), _fmt(r.get("snr_balopt"))]
            f.write(",".join(row) + "\n")
    return summary
