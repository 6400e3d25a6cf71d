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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from classical_speech_enhancement_b200.distributed import sweep_sharded
    from classical_speech_enhancement_b200.synth import make_batch
    from tests.emu_util import use_emulated_runtime
    use_emulated_runtime()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clean, noisy = make_batch(3, 9000)
    out = sweep_sharded(clean, noisy, grids=GRID, select=True)
    if rank == 0:
        q.put((out["scores"]["wiener"], [b["stoi"]["index"] for b in out["selection"]["wiener"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process():
    import torch.multiprocessing as mp
    from classical_speech_enhancement_b200.synth import make_batch
    from classical_speech_enhancement_b200.sweep import sweep_dataset
    from tests.emu_util import use_emulated_runtime, use_product_runtime
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    scores, sel = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    use_emulated_runtime()
    try:
        clean, noisy = make_batch(3, 9000)
        single = sweep_dataset(clean, noisy, grids=GRID)
    finally:
        use_product_runtime()
    assert np.array_equal(scores, single["scores"]["wiener"])
    assert sel == [b["stoi"]["index"] for b in single["selection"]["wiener"]]
    assert single["nominal"] == 3 * 4 and single["unique"] == 3 * 4
