"""The device selection kernel (``cse_select_best``, thread-emulated here) against the sequential host scan
``grid.select_best`` - the restatement of ``Code/speech_enhancement_comparison.py:186-216`` - on adversarial
score tables: exact ties, near-ties inside the hysteresis tolerances, invalid and skipped candidates,
rows without any valid candidate, +inf SNR, grid sizes that are not a multiple of the warp width."""
import numpy as np
import pytest

from classical_speech_enhancement_b200 import parameter_ranges as pr
from classical_speech_enhancement_b200._lib import WINNER_DTYPE
from classical_speech_enhancement_b200.grid import best_from_winners, grid_points, select_best
from tests.emu_util import emu_lib, ptr


def adversarial(U, P, seed, lib):
    rng = np.random.default_rng(seed)
    stoi = np.round(rng.uniform(0.5, 0.9, (U, P)), 2).astype(np.float32) + (rng.integers(0, 4, (U, P)) * 5e-7).astype(np.float32)
    snr = rng.normal(5, 3, (U, P)).astype(np.float32)
    flags = np.where(rng.random((U, P)) > 0.1, 3, 2).astype(np.int32)
    flags[0, min(3, P - 1)] |= 4                                  # +inf SNR
    flags[1] = 2                                                  # no valid candidate at all
    pesq = np.round(rng.uniform(1, 3, (U, P)), 2) + rng.integers(0, 3, (U, P)) * 4e-4
    pesq[rng.random((U, P)) < 0.05] = np.nan                      # calculate_pesq returned None
    table = np.zeros((U, P), dtype=lib.score_dtype)
    table["stoi"], table["snr"], table["flags"] = stoi, snr, flags
    table["lag"] = rng.integers(-5, 6, (U, P))
    return table, pesq


def host_scan(points, row, pesq_row):
    stoi = [float(v) for v in row["stoi"]]
    snr = [float("inf") if f & 4 else float(v) for v, f in zip(row["snr"], row["flags"])]
    pq = [0.0] * len(points) if pesq_row is None else [None if np.isnan(v) else float(v) for v in pesq_row]
    return select_best(points, stoi, pq, snr, (row["flags"] & 1) != 0)


@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("P", [1, 31, 192, 333])
def test_device_selection_equals_sequential_scan(fp64, P):
    lib = emu_lib(fp64)
    U = 5
    points = (grid_points(pr.param_ranges_wiener) * 2)[:P]
    table, pesq = adversarial(U, P, seed=P, lib=lib)
    win = np.zeros((U, 3), dtype=WINNER_DTYPE)
    for use_pesq in (True, False):
        lib.select_best(ptr(table), ptr(pesq) if use_pesq else None, U, P, ptr(win), None)
        for u in range(U):
            ref = host_scan(points, table[u], pesq[u] if use_pesq else None)
            got = best_from_winners(points, win[u])
            for c in ("stoi", "pesq", "balance"):
                assert got[c]["index"] == ref[c]["index"], (u, c)
                if ref[c]["index"] is not None:
                    assert got[c]["score"] == ref[c]["score"] and got[c]["stoi"] == ref[c]["stoi"]
                    assert got[c]["snr"] == ref[c]["snr"] and got[c]["pesq"] == ref[c]["pesq"]
                    assert got[c]["params"] == ref[c]["params"]
    assert best_from_winners(points, win[1])["stoi"]["index"] is None
    marked = best_from_winners(points, win[0], pesq_available=False)
    assert marked["pesq"]["index"] is None and marked["balance"]["unavailable"] and marked["stoi"]["index"] is not None


def test_selection_is_order_dependent_not_argmax():
    lib = emu_lib()
    points = [{"i": i} for i in range(40)]
    table = np.zeros((1, 40), dtype=lib.score_dtype)
    table["flags"] = 1
    table["stoi"][0, :] = 0.5
    table["stoi"][0, 33] = np.float32(0.5) + np.float32(6e-7)     # argmax, but inside the 1e-6 hysteresis of index 0
    win = np.zeros((1, 3), dtype=WINNER_DTYPE)
    lib.select_best(ptr(table), None, 1, 40, ptr(win), None)
    assert win[0, 0]["index"] == 0
    table["stoi"][0, 35] = np.float32(0.5) + np.float32(2e-6)
    lib.select_best(ptr(table), None, 1, 40, ptr(win), None)
    assert win[0, 0]["index"] == 35


def test_pesq_pool_inline_and_forked_scorers_do_not_mix():
    """Two pools alive at once keep their own scorer (inline: on the instance; forked: the children's copy)."""
    from classical_speech_enhancement_b200.pesq_pool import PesqPool
    clean = np.linspace(-0.5, 0.5, 400)
    wavs = [clean * 0.5, clean * 2.5]            # the second one clips in finalize
    a = PesqPool(lambda c, w, sr: 1.0 + float(np.abs(w).max()), 16000, workers=0)
    b = PesqPool(lambda c, w, sr: None if np.abs(w).max() >= 1.0 else 3.0, 16000, workers=2)
    for pool in (a, b):
        pool.submit(0, [0, [1, 2]], clean, wavs, lags=np.array([0, 0]), flags=np.array([1, 1]))
        pool.submit(1, [0], clean, [clean], lags=np.array([3]), flags=np.array([0]))       # invalid candidate: not scored
    ta, tb = a.table(2, 3), b.table(2, 3)
    a.close(); b.close()
    assert np.allclose(ta[0], [1.25, 2.0, 2.0]) and np.isnan(ta[1]).all()
    assert tb[0, 0] == 3.0 and np.isnan(tb[0, 1]) and np.isnan(tb[0, 2]) and np.isnan(tb[1]).all()
