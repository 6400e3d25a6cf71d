#!/usr/bin/env python
"""Benchmark of the enhancement-and-scoring sweep (BASELINE.json metric: enhanced
utterance-configs/s at 1/2/4/8 B200, plus achieved HBM GB/s against the measured peak).

One "step" = one pass of the whole hot path over the whole synthetic test set: for every
utterance pair, all four algorithms x the full ``parameter_ranges.py`` grids (9744 nominal /
5004 unique grid points): STFTs, noise PSDs, clean-side scoring caches, gain + ISTFT,
alignment, STOI and SNR for every candidate, and the reference's three-way selection scan per
(utterance, algorithm) on the device.  With N GPUs the utterances are sharded in contiguous blocks
(strong scaling: the job is fixed at --utts utterances), every rank selects for its own utterances
and the winners' records are all-gathered over NCCL at the end of the step.  The line also carries
its own checks: oracle parity of sampled table entries, device-vs-host selection, checksums that
must be equal at every N.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU path (oracle port of the reference) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte model behind `roofline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "enhanced utterance-configs/sec"
UNIT = "utterance-configs/s"
SR = 16000


# ------------------------------------------------------------------------------ data
def _pair(args):
    from classical_speech_enhancement_b200.synth import make_pair
    u, L = args
    c, n = make_pair(u, L)
    return c.astype(np.float32), n.astype(np.float32)


def make_shard(first, count, L):
    """Synthetic pairs first..first+count-1 (seed 1000+u each), generated on the host cores."""
    import multiprocessing as mp
    if count == 0:
        return np.zeros((0, L), np.float32), np.zeros((0, L), np.float32)
    procs = max(1, min(count, (os.cpu_count() or 2) - 1, 32))
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            out = pool.map(_pair, [(first + i, L) for i in range(count)], chunksize=4)
    else:
        out = [_pair((first + i, L)) for i in range(count)]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


# ------------------------------------------------------------------------------ byte model
def config_bytes(n_fft, hop, method, L, real_bytes=4):
    """Algorithmic bytes of ONE utterance-config: read Y, read the noise PSD, write the
    enhanced waveform, read it back for scoring, 16 B of scores.  Returns (total, enhance part, score part)."""
    nb, nf = n_fft // 2 + 1, 1 + L // hop
    y = 2 * real_bytes * nb * nf
    n = real_bytes * nb * (1 if method == "percentile" else nf)
    return y + n + 2 * real_bytes * L + 16, y + n + real_bytes * L, real_bytes * L + 16


def grid_mean_bytes(L):
    from classical_speech_enhancement_b200.grid import grid_points
    from classical_speech_enhancement_b200.sweep import DEFAULT_GRIDS
    tot = cnt = 0
    for _, ranges in DEFAULT_GRIDS:
        for p in grid_points(ranges):
            tot += config_bytes(p["n_fft"], p["hop_length"], p["noise_method"], L)[0]
            cnt += 1
    return tot / cnt, cnt


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  One poller (rank 0) covers
    every GPU of the job; it is started before the warm-up so that its NVML start-up is over when the
    timed region begins, and only the samples between mark() and stop() are used."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.indices, self.proc, self.lines, self.first = list(indices), None, [], 0

    def start(self):
        if not self.indices:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--id=" + ",".join(str(i) for i in self.indices),
                                          f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi not sampled on this rank" if not self.indices else "nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.lines = self.lines[self.first:]
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU path (oracle port)
_PAIRS = {}          # utterance id -> (clean, noisy) float64 of float32-rounded samples; filled BEFORE any timing / fork


def _load_pairs(utts, L):
    from classical_speech_enhancement_b200.synth import make_pair
    for u in utts:
        if u not in _PAIRS:
            c, n = make_pair(u, L)
            _PAIRS[u] = (c.astype(np.float32).astype(np.float64), n.astype(np.float32).astype(np.float64))


def _cpu_task(task):
    """One utterance-config through the reference's per-candidate procedure (no caching), no PESQ.  The synthetic
    pair comes from the pre-generated cache: data generation is NOT part of the timed work."""
    import oracle
    from oracle.search import score_candidate
    alg, point, u, L = task
    c, n = _PAIRS[u]
    t = time.perf_counter()
    enh = oracle.ALGORITHMS[alg](n, SR, **point)
    sc = score_candidate(c, enh, SR)
    return time.perf_counter() - t, (sc["stoi"] if sc else None), (sc["snr"] if sc else None)


def cpu_sample(n_tasks, L, seed=0, n_utts=824):
    """Random (algorithm, grid point index, utterance) triples drawn uniformly from the nominal 9744-point grid."""
    from classical_speech_enhancement_b200.grid import grid_points
    from classical_speech_enhancement_b200.sweep import DEFAULT_GRIDS
    allpts = [(name, i, p) for name, ranges in DEFAULT_GRIDS for i, p in enumerate(grid_points(ranges))]
    rng = np.random.default_rng(seed)
    idx = rng.choice(len(allpts), n_tasks, replace=False)
    return [(allpts[i][0], allpts[i][1], allpts[i][2], int(rng.integers(0, n_utts)), L) for i in idx]


def cpu_baseline_single_thread(L, n_tasks=160, n_utts=824):
    """-> (cpu_baseline dict, [(alg, point index, utterance, stoi, snr)] of the sample for the parity check)."""
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(v, "1")
    tasks = cpu_sample(n_tasks, L, seed=0, n_utts=n_utts)
    _load_pairs({t[3] for t in tasks}, L)                      # untimed
    dt, sample = 0.0, []
    for alg, i, point, u, _ in tasks:
        d, stoi, snr = _cpu_task((alg, point, u, L))
        dt += d
        sample.append((alg, i, u, stoi, snr))
    return ({"value": n_tasks / dt, "unit": UNIT, "cores": 1, "kind": "port",
             "sample": f"{n_tasks} random (algorithm, grid point, utterance) configs of the {L}-sample workload, "
                       f"seed 0, fp64 numpy/scipy oracle, no cross-candidate caching, no PESQ, {dt:.1f} s of oracle "
                       "time (synthetic-pair generation excluded)"}, sample)


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be
    imported (librosa/pystoi/pesq absent), so this is the oracle port, on all host cores.  Each step is a bounded
    random sample of the C5 workload (the CPU never executes all 8 M configs); the synthetic pairs are generated
    before the pool forks, outside every timed region."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    cores = os.cpu_count() or 1
    per_step = max(64, 4 * cores)
    L = args.length
    mean_bytes, nominal = grid_mean_bytes(L)
    from classical_speech_enhancement_b200.sweep import nominal_and_unique
    unique_points = nominal_and_unique()[1]         # a property of the grid (dead parameters), identical in both arms' config
    steps = [[(a, p, u, L) for a, _i, p, u, _ in cpu_sample(per_step, L, seed=step, n_utts=args.utts)]
             for step in range(args.warmup + args.steps)]
    _load_pairs({t[2] for st in steps for t in st}, L)
    with mp.get_context("fork").Pool(cores) as pool:
        times = []
        for step, tasks in enumerate(steps):
            t = time.perf_counter()
            pool.map(_cpu_task, tasks, chunksize=1)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = per_step * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, nominal, unique_points),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} random configs of the workload per step on {cores} processes "
                                   "(fp64 numpy/scipy oracle restatement of the reference; PESQ excluded; every config "
                                   "recomputes its STFT / noise PSD as the reference does; data generation untimed)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "configs/s of NOMINAL grid points (the CPU path never dedupes dead parameters); extrapolates a sampled "
                "subset, the full C5 step would take ~" + f"{nominal * args.utts / value / 3600:.0f} h on this host",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, nominal_points, unique_points):
    return {"workload": f"C5: all four algorithms x full parameter_ranges.py grid ({nominal_points} nominal grid "
                        f"points/utterance) x {args.utts} synthetic 16 kHz pairs of {args.length} samples, "
                        "enhance + finalize + STOI + SNR per candidate + the three-way selection scan per (utterance, "
                        "algorithm) (PESQ excluded, host-side)",
            "utterances": args.utts, "length": args.length, "grid_points_nominal": nominal_points,
            "grid_points_unique": unique_points, "chunk_items": args.chunk,
            "cache_hygiene": "inputs larger than L2: waveforms + spectrogram / noise-PSD caches of the shard are "
                             "GBs and every step recomputes them from the raw signals; candidate waveforms are "
                             "rewritten every chunk; winners return through recycled pinned staging buffers",
            "parallelism": f"utterance-sharded x{args.gpus}"}


# ------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=824)
    ap.add_argument("--length", type=int, default=48000)
    ap.add_argument("--chunk", type=int, default=14208)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)          # timing rule: at least three untimed steps (the JSON reports the value used)
    args.steps = max(args.steps, 1)

    import torch
    import torch.distributed as dist
    from classical_speech_enhancement_b200 import sweep as sw
    from classical_speech_enhancement_b200.distributed import gather_device_scores, shard_bounds
    from classical_speech_enhancement_b200.engine import SweepEngine, reuse_result_buffers
    reuse_result_buffers(True)      # each step's score tables are consumed (selection / checks) before the next step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    L = args.length
    b = shard_bounds(args.utts, world)
    clean_h, noisy_h = make_shard(b[rank], b[rank + 1] - b[rank], L)
    clean_pin = torch.from_numpy(clean_h).pin_memory()
    noisy_pin = torch.from_numpy(noisy_h).pin_memory()
    mean_bytes, nominal_points = grid_mean_bytes(L)
    _, unique_points = sw.nominal_and_unique()
    total_configs = args.utts * nominal_points

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: inputs already in HBM
    from classical_speech_enhancement_b200.distributed import gather_winners
    import warnings
    warnings.filterwarnings("ignore", message="selection without PESQ")
    eng = SweepEngine(clean_pin.numpy(), noisy_pin.numpy(), chunk_items=args.chunk)

    u_pad = max(b[r + 1] - b[r] for r in range(world))

    def step_resident(e=None):
        """The whole path: caches dropped, STFTs / noise PSDs / clean caches / sweep of all four grids, the
        selection scan on this rank's device, ONE all_gather of the winners' records (N>1), winners to the host."""
        e = e or eng
        e.reset()
        items = sw.run_engine_device(e, u_pad=u_pad)
        return items, gather_winners(e, sw.select_winners_device(e, items), args.utts, u_pad)

    sampler = ClockSampler(range(world) if rank == 0 else [])
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler.mark()
    eng.enable_timing(True)
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        items, winners = step_resident()
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = eng.launches - launches0
    timing = eng.timing_summary()
    launch_counts = {}
    for tag, _n, _e0, _e1 in (eng._timing or []):
        launch_counts[tag[0]] = launch_counts.get(tag[0], 0) + 1
    eng.enable_timing(False)
    t = torch.tensor([max(dev_ms / 1e3, 0.0), wall], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t[0]), float(t[1])
    step_s = max(dev_s, 1e-9) / args.steps

    # ---------------- untimed: the full per-point tables of the last step (parity check, checksum across N)
    import hashlib
    scores = gather_device_scores(eng, items, args.utts, device=device)
    names = [name for name, _r in sw.DEFAULT_GRIDS]
    sha = hashlib.sha256()
    for name in names:
        sha.update(np.ascontiguousarray(scores[name]).view(np.uint8).reshape(-1).data)
    score_sha256 = sha.hexdigest()
    wsha = hashlib.sha256()
    for name in names:
        wsha.update(np.ascontiguousarray(winners[name]).view(np.uint8).reshape(-1).data)
    # the device selection against the host restatement of the reference's scan, over the whole job
    host_sel = sw.select_all(scores, {name: pts for name, pts, _, _ in items}) if rank == 0 else None
    sel_mismatch = None
    if rank == 0:
        sel_mismatch = sum(int(host_sel[name][u]["stoi"]["index"] != (int(winners[name][u, 0]["index"]) if winners[name][u, 0]["index"] >= 0 else None))
                           for name in names for u in range(args.utts))
    table_for_parity = {name: scores[name].copy() for name in names} if rank == 0 else None
    del scores

    # ---------------- end-to-end arm: host buffers in, winners out, every step
    def step_e2e():
        # the public dataset-level path (sweep.sweep_dataset / distributed.sweep_sharded with tables=False):
        # pinned host waveforms -> device, whole sweep, device selection, winners -> host
        e = SweepEngine(clean_pin.numpy(), noisy_pin.numpy(), chunk_items=args.chunk)
        its = sw.run_engine_device(e, u_pad=u_pad)
        return e, gather_winners(e, sw.select_winners_device(e, its), args.utts, u_pad)

    def step_e2e_tables():
        # same + the full per-point score tables gathered to every rank's host (what round 1 reported as e2e)
        e = SweepEngine(clean_pin.numpy(), noisy_pin.numpy(), chunk_items=args.chunk)
        its = sw.run_engine_device(e, u_pad=u_pad)
        w = gather_winners(e, sw.select_winners_device(e, its), args.utts, u_pad)
        return e, (w, gather_device_scores(e, its, args.utts, device=device))

    del eng, items
    torch.cuda.empty_cache()

    def time_e2e(fn):
        fn()
        barrier()
        n = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        e = None
        for _ in range(n):
            e = None                     # the previous step's engine returns its buffers to the caching allocator first
            e, _out = fn()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = e.h2d_bytes
        del e
        torch.cuda.empty_cache()
        return float(te[0]) / n, h2d

    e2e_s, h2d = time_e2e(step_e2e)
    e2e_tab_s, _ = time_e2e(step_e2e_tables)
    from classical_speech_enhancement_b200._lib import WINNER_DTYPE
    d2h = args.utts * len(names) * 3 * WINNER_DTYPE.itemsize                 # the (gathered) winners, per rank
    d2h_tables = d2h + args.utts * nominal_points * 16

    # ---------------- config 1 latency: one spectral-subtraction call through the drop-in entry point
    c1 = None
    if rank == 0:
        from classical_speech_enhancement_b200.spectral_subtractor import spectral_subtraction
        c1n, c1c = noisy_h[0].astype(np.float64), clean_h[0].astype(np.float64)
        kw = dict(alpha=2.0, beta=0.01, n_fft=512, hop_length=128, noise_percentile=10.0, noise_method="true_noise",
                  clean_audio=c1c)
        for _ in range(3):
            spectral_subtraction(c1n, SR, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            spectral_subtraction(c1n, SR, **kw)
        c1 = {"workload": "C1: spectral_subtraction(noisy, 16000, n_fft=512, hop=128, true_noise), one 3 s pair, host "
                          "array in -> host float64 waveform out (H2D, STFT, oracle noise PSD, gain + ISTFT, D2H)",
              "ms_per_call": 1e3 * (time.perf_counter() - t0) / 20}

    # ---------------- the reference's own call pattern: optimize_parameters per (pair, algorithm), full grids
    pair = None
    if rank == 0 and world == 1:
        import warnings
        from classical_speech_enhancement_b200.speech_enhancement_comparison import algorithms_table, optimize_parameters
        pc, pn = clean_h[1].astype(np.float64), noisy_h[1].astype(np.float64)
        per_alg = {}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                  # "PESQ unavailable: only the stoi winner"
            for name, fn, ranges in algorithms_table():
                for _ in range(2):
                    optimize_parameters(pc, pn, SR, fn, ranges, pesq_scorer=None, verbose=False)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    optimize_parameters(pc, pn, SR, fn, ranges, pesq_scorer=None, verbose=False)
                torch.cuda.synchronize()
                per_alg[name] = round(1e3 * (time.perf_counter() - t0) / 5, 3)
        pair = {"workload": "optimize_parameters(clean, noisy, 16000, algorithm, parameter_ranges grid) for one 3 s pair: host arrays "
                            "in -> best parameters, scores and the winner's waveform out, full grid per call (no PESQ)",
                "ms_per_call": per_alg, "ms_per_pair": round(sum(per_alg.values()), 3),
                "configs_per_s": round(nominal_points / (1e-3 * sum(per_alg.values())), 1)}

    # ---------------- roofline of the dominant kernel (events around every chunk launch, timed steps only)
    names = {0: "ss", 1: "wiener", 2: "mmse", 3: "omlsa"}
    fam, tags = {}, {}
    for (kind, alg, n_fft, hop, method), (items, ms) in timing.items():
        _, eb, sb = config_bytes(n_fft, hop, method, L)
        # byte model per kernel: enhance = Y + noise PSD + waveform write; stoi = waveform read + scores
        # (the "score" stage of SURVEY 8d); align re-reads the 2 s correlation window (an extra read the
        # survey's count-once model does not include - listed for completeness).
        per = eb if kind == "enhance" else (sb if kind == "stoi" else 4 * min(L, 32000))
        by = items * per
        f = fam.setdefault(kind, [0.0, 0.0, 0])
        f[0] += by; f[1] += ms; f[2] += items
        kname = (f"enhance_kernel<{alg}, {n_fft.bit_length() - 1}>" if kind == "enhance"
                 else ("stoi_stream_kernel" if kind == "stoi" else "align_kernel<0>"))
        t = tags.setdefault(kname, [0.0, 0.0, 0, per])
        t[0] += by; t[1] += ms; t[2] += items
    exec_items = sum(v[2] for k, v in fam.items() if k == "enhance")
    exec_bytes = sum(v[0] for k, v in fam.items() if k in ("enhance", "stoi"))
    peaks, traffic_tab = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    roofline = None
    if tags:
        # "the dominant kernel" = the __global__ function with the largest share of the step; the
        # instantiations of the enhance_kernel<ALG, LOG2N> template are one kernel (same source), listed
        # one by one under "kernels".
        famkern = {"enhance": "enhance_kernel<ALG, LOG2N>", "stoi": "stoi_stream_kernel", "align": "align_kernel<0>"}
        domfam = max(fam, key=lambda k: fam[k][1])
        dom = famkern[domfam]
        by, ms, items = fam[domfam]
        launches_dom = max(1, launch_counts.get(domfam, 0))
        ach = by / (ms * 1e-3) / 1e9
        top_inst = max((k for k in tags if k.startswith(dom.split("<")[0])), key=lambda k: tags[k][1])
        tr = traffic_tab.get(top_inst)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": (tr["dram_bytes_per_item"] * items / launches_dom) if tr else None,
                    "algorithmic_bytes_per_launch": by / launches_dom, "avg_launch_ms": ms / launches_dom,
                    "launches": launches_dom, "peak_source": peak_kind,
                    "share_of_step": ms / (dev_ms if dev_ms > 0 else 1.0),
                    "traffic_source": (tr.get("source") + f"; per-candidate DRAM bytes of {top_inst} x candidates per launch") if tr else None,
                    "families": {k: {"ms": v[1], "algorithmic_GBps": v[0] / (v[1] * 1e-3) / 1e9 if v[1] else None,
                                     "items": v[2], "share_of_step": v[1] / (dev_ms if dev_ms > 0 else 1.0)}
                                 for k, v in fam.items()},
                    "kernels": {k: {"ms": v[1], "algorithmic_GBps": v[0] / (v[1] * 1e-3) / 1e9 if v[1] else None,
                                    "us_per_item": 1e3 * v[1] / v[2]} for k, v in sorted(tags.items(), key=lambda kv: -kv[1][1])},
                    "whole_path": {"bytes_per_config": mean_bytes,
                                   "achieved": total_configs / step_s * mean_bytes / 1e9 / world,
                                   "frac": total_configs / step_s * mean_bytes / 1e9 / world / peak,
                                   "note": "per GPU; nominal configs x grid-mean algorithmic bytes (SURVEY 8d)"},
                    "whole_path_executed": {"bytes_per_config": exec_bytes / max(1, exec_items),
                                            "achieved": exec_bytes / args.steps / step_s / 1e9,
                                            "frac": exec_bytes / args.steps / step_s / 1e9 / peak,
                                            "note": "per GPU (this rank); only the candidates the kernels executed "
                                                    "(dead-parameter duplicates are broadcast, not computed) x their own bytes"},
                    "note": "instruction-issue-bound kernels (see DESIGN.md section 6): DRAM traffic is far below the "
                            "algorithmic bytes because Y / noise PSD are shared by all candidates of an utterance and hit in L2"}

    if rank == 0:
        cpu, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            cpu, sample = cpu_baseline_single_thread(L, n_utts=args.utts)
            ds = [abs(float(table_for_parity[a][u, i]["stoi"]) - st) for a, i, u, st, _ in sample if st is not None]
            dn = [abs(float(table_for_parity[a][u, i]["snr"]) - sn) for a, i, u, _, sn in sample if sn is not None and np.isfinite(sn)]
            bad = sum(1 for a, i, u, st, _ in sample if (st is None) != (not table_for_parity[a][u, i]["flags"] & 1))
            parity = {"n": len(ds), "max_abs_dstoi": max(ds), "max_abs_dsnr_db": max(dn), "validity_mismatches": bad,
                      "tolerance": {"stoi": 1e-4, "snr_db": 1e-3},
                      "what": "device score-table entries of the timed job vs the fp64 oracle on the cpu_baseline sample"}
        line = {
            "metric": METRIC, "value": total_configs / step_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, nominal_points, unique_points),
            "unique_value": args.utts * unique_points / step_s,
            "e2e": {"value": total_configs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                    "what": "host waveforms in -> the three winners per (utterance, algorithm) out (device selection)"},
            "e2e_with_tables": {"value": total_configs / e2e_tab_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                "d2h_bytes_per_step": int(d2h_tables), "ms_per_step": 1e3 * e2e_tab_s,
                                "what": "same + every candidate's score record gathered to every rank's host"},
            "gpu_launches": int(launches), "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "score_sha256": score_sha256, "winners_sha256": wsha.hexdigest(),
            "selection_check": {"utterance_algorithms": args.utts * len(names), "device_vs_host_scan_mismatches": sel_mismatch},
            "parity_check": parity, "c1_latency": c1, "pair_latency": pair,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
