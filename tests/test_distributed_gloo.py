"""N>1 host path on CPU: world_size-2 gloo processes shard the utterances, run the sweep on the
thread-emulated kernels and all_gather the score tables; the result must equal the single-process
sweep bit for bit."""
import os
import socket

import numpy as np

from classical_speech_enhancement_b200.distributed import shard_bounds

GRID = (("wiener", {"alpha": [0.95], "gain_floor": [0.02, 0.1], "n_fft": [256], "hop_length": [128],
                    "noise_percentile": [10.0], "noise_method": ["percentile", "min_tracking"]}),)


def test_shard_bounds():
    assert shard_bounds(824, 8) == [0, 103, 206, 309, 412, 515, 618, 721, 824]
    assert shard_bounds(5, 2) == [0, 3, 5]
    assert shard_bounds(3, 4) == [0, 1, 2, 3, 3]


def _ragged_pairs():
    from classical_speech_enhancement_b200.synth import make_pair
    return [make_pair(40 + i, L) for i, L in enumerate((7000, 9001, 6000))]


def test_deal_by_length():
    from classical_speech_enhancement_b200.distributed import deal_by_length
    assert deal_by_length([5, 9, 7, 9, 1], 2) == [[1, 2, 4], [3, 0]]
    assert deal_by_length([3], 2) == [[0], []]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from classical_speech_enhancement_b200.distributed import sweep_sharded
    from classical_speech_enhancement_b200.synth import make_batch
    from tests.emu_util import use_emulated_runtime
    use_emulated_runtime()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clean, noisy = make_batch(3, 9000)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = sweep_sharded(clean, noisy, grids=GRID, select=True, tables=True)
        lean = sweep_sharded(clean, noisy, grids=GRID, select=True)          # winners only: no table gather
    assert lean["scores"] is None and np.array_equal(lean["winners"]["wiener"], out["winners"]["wiener"])
    rng = np.random.default_rng(3)
    pesq = {"wiener": np.round(rng.uniform(1, 3, (3, 4)), 2)}                 # host-side PESQ table of the whole dataset
    withp = sweep_sharded(clean, noisy, grids=GRID, select=True, pesq=pesq)
    from classical_speech_enhancement_b200.distributed import sweep_pairs_sharded
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ragged = sweep_pairs_sharded(_ragged_pairs(), grids=GRID)
    if rank == 0:
        q.put((out["scores"]["wiener"], out["winners"]["wiener"], withp["winners"]["wiener"],
               [b["stoi"]["index"] for b in out["selection"]["wiener"]], ragged["winners"]["wiener"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process():
    import torch.multiprocessing as mp
    from classical_speech_enhancement_b200.synth import make_batch
    from classical_speech_enhancement_b200.sweep import sweep_dataset
    from tests.emu_util import use_emulated_runtime, use_product_runtime
    from classical_speech_enhancement_b200 import build
    build.build_emu()                      # before the ranks start: they load it instead of compiling it side by side
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    scores, winners, winners_pesq, sel, ragged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    use_emulated_runtime()
    try:
        clean, noisy = make_batch(3, 9000)
        import warnings
        from classical_speech_enhancement_b200.sweep import select_all
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            single = sweep_dataset(clean, noisy, grids=GRID)
            host = select_all(single["scores"], single["points"])
        rng = np.random.default_rng(3)
        pesq = {"wiener": np.round(rng.uniform(1, 3, (3, 4)), 2)}
        single_p = sweep_dataset(clean, noisy, grids=GRID, pesq=pesq)
        from classical_speech_enhancement_b200.sweep import sweep_pairs
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ragged_single = sweep_pairs(_ragged_pairs(), grids=GRID, tables=False)
        host_p = select_all(single_p["scores"], single_p["points"], pesq=pesq)
    finally:
        use_product_runtime()
    assert np.array_equal(scores, single["scores"]["wiener"])
    assert sel == [b["stoi"]["index"] for b in single["selection"]["wiener"]]
    # the winners gathered from the two ranks == the single-process device selection == the host scan
    assert np.array_equal(winners, single["winners"]["wiener"]) and np.array_equal(winners_pesq, single_p["winners"]["wiener"])
    assert single["selection"]["wiener"][0]["pesq"]["index"] is None          # no PESQ -> marked unavailable
    for u in range(3):
        assert single["selection"]["wiener"][u]["stoi"]["index"] == host["wiener"][u]["stoi"]["index"]
        for c in ("stoi", "pesq", "balance"):
            a, b = single_p["selection"]["wiener"][u][c], host_p["wiener"][u][c]
            assert a["index"] == b["index"] and a["score"] == b["score"] and a["params"] == b["params"]
    assert single["nominal"] == 3 * 4 and single["unique"] == 3 * 4
    # variable-length corpus dealt over the two ranks == one process, in input order
    assert np.array_equal(ragged, ragged_single["winners"]["wiener"]) and ragged["index"][:, 0].min() >= 0
