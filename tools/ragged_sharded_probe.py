"""Variable-length corpus over N GPUs (distributed.sweep_pairs_sharded): configs/s, and the gathered winners against
one rank sweeping the whole corpus.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ragged_sharded_probe.py [--pairs 200]"""
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from classical_speech_enhancement_b200.distributed import sweep_pairs_sharded  # noqa: E402
from classical_speech_enhancement_b200.sweep import sweep_pairs  # noqa: E402
from classical_speech_enhancement_b200.synth import make_pair  # noqa: E402

warnings.filterwarnings("ignore")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n = int(sys.argv[sys.argv.index("--pairs") + 1]) if "--pairs" in sys.argv else 200
rng = np.random.default_rng(1)
pairs = [tuple(x.astype(np.float32) for x in make_pair(i, int(L))) for i, L in enumerate(rng.integers(32000, 64000, n))]
sweep_pairs_sharded(pairs)                                     # plans, allocator pools, NCCL communicator
best = None
for _ in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    out = sweep_pairs_sharded(pairs)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    best = float(dt) if best is None else min(best, float(dt))
if rank == 0:
    single = sweep_pairs(pairs, tables=False)
    same = all(np.array_equal(single["winners"][a], out["winners"][a]) for a in single["winners"])
    print(json.dumps({"n_gpus": world, "pairs": n, "distinct_lengths": len({len(p[0]) for p in pairs}), "seconds": round(best, 3),
                      "configs_per_s": round(out["nominal"] / best, 1), "winners_equal_single_rank": same}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
