"""Multi-GPU sharding of the sweep: one process per GPU, utterances split in contiguous blocks.

Every (utterance, algorithm, grid point) is independent; the only coupling is the
per-(utterance, algorithm) selection scan, so utterances are partitioned across ranks, each rank
runs the complete grid AND the selection scan for its block with its own caches, and ONE collective
at the end gathers the winners' records - on request also the per-point score tables
(``torch.distributed`` all_gather: NCCL over NVLink on GPUs, gloo in the CPU tests).  The reference has no distributed code at all
(SURVEY.md section 5); this is its B200-native replacement for running ``main()`` for a day.
"""
import numpy as np

from .engine import DEFAULT_CHUNK_ITEMS


def shard_bounds(n_utts, world_size):
    """Contiguous, balanced blocks: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(n_utts, world_size)
    b = [0]
    for r in range(world_size):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def local_slice(n_utts, rank, world_size):
    b = shard_bounds(n_utts, world_size)
    return slice(b[rank], b[rank + 1])


def gather_scores(local, n_utts, device=None):
    """all_gather of a structured score table [U_local, C] -> [n_utts, C] on every rank.

    The table travels as raw bytes (scores are already final per point: no reduction, so the
    N-GPU result equals the 1-GPU result bit for bit).  Blocks are padded to the largest shard so
    that one fixed-size all_gather suffices."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    b = shard_bounds(n_utts, world)
    assert local.shape[0] == b[rank + 1] - b[rank]
    C = local.shape[1]
    row_bytes = C * local.dtype.itemsize
    max_rows = max(b[r + 1] - b[r] for r in range(world))
    buf = np.zeros((max_rows, row_bytes), dtype=np.uint8)
    buf[:local.shape[0]] = np.ascontiguousarray(local).view(np.uint8).reshape(local.shape[0], row_bytes)
    send = torch.from_numpy(buf)
    if device is not None:
        send = send.to(device)
    recv = torch.empty((world,) + tuple(send.shape), dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(recv, send) if send.is_cuda else dist.all_gather(list(recv.unbind(0)), send)
    recv = recv.cpu().numpy()
    out = np.zeros((n_utts, C), dtype=local.dtype)
    for r in range(world):
        n = b[r + 1] - b[r]
        out[b[r]:b[r + 1]] = recv[r, :n].reshape(-1).view(local.dtype).reshape(n, C)
    return out


def gather_device_scores(engine, items, n_utts, device=None):
    """all_gather of the per-algorithm nominal device tables [u_pad][points] (no host staging on the
    way in), one device->host copy, and views per rank block.  `items` is the list returned by
    sweep.run_engine_device with u_pad = the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    b = shard_bounds(n_utts, world)
    scores = {}
    for name, pts, buf, pl in items:
        if world == 1:
            scores[name] = engine.table_to_host(engine.be.view_bytes_as(buf, np.uint8, tag=(name, "table")), pl, engine.U)
            continue
        send = buf if isinstance(buf, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(buf))
        recv = torch.empty((world,) + tuple(send.shape), dtype=torch.uint8, device=send.device)
        if send.is_cuda:
            dist.all_gather_into_tensor(recv, send)
        else:
            dist.all_gather(list(recv.unbind(0)), send)
        host = (engine.be.staged_to_host(recv, tag=(name, "gathered")) if recv.is_cuda and hasattr(engine.be, "staged_to_host") else recv.cpu().numpy().reshape(-1)).reshape(world, -1)
        if all(b[r + 1] - b[r] == b[1] - b[0] for r in range(world)):
            scores[name] = host.reshape(-1).view(engine.lib.score_dtype).reshape(n_utts, pl["n_points"])   # zero copy
        else:
            scores[name] = np.concatenate([engine.table_to_host(host[r], pl, b[r + 1] - b[r]) for r in range(world)])
    return scores


def gather_winners(engine, winners_dev, n_utts, u_pad):
    """all_gather of every algorithm's per-rank winners ({alg: device ``cse_winner_t`` [U_local][3]}) in ONE
    collective: 48-byte records, ~0.6 KB per utterance for four algorithms - instead of the per-point
    score tables (156 KB per utterance).  Returns {alg: host array [n_utts, 3]} on every rank."""
    import torch
    import torch.distributed as dist
    names = list(winners_dev)
    wdt = engine.lib.winner_dtype
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return {name: engine.winners_to_host(winners_dev[name], tag=(name, "winners")).copy() for name in names}
    b = shard_bounds(n_utts, world)
    rec = 3 * wdt.itemsize
    is_torch = isinstance(winners_dev[names[0]], torch.Tensor)
    if is_torch:
        send = torch.zeros((len(names), u_pad * rec), dtype=torch.uint8, device=winners_dev[names[0]].device)
        for k, name in enumerate(names):
            send[k, :engine.U * rec] = winners_dev[name].reshape(-1).view(torch.uint8)[:engine.U * rec]
    else:
        host = np.zeros((len(names), u_pad * rec), dtype=np.uint8)
        for k, name in enumerate(names):
            host[k, :engine.U * rec] = np.asarray(winners_dev[name]).reshape(-1).view(np.uint8)[:engine.U * rec]
        send = torch.from_numpy(host)
    recv = torch.empty((world,) + tuple(send.shape), dtype=torch.uint8, device=send.device)
    if send.is_cuda:
        dist.all_gather_into_tensor(recv, send)
    else:
        dist.all_gather(list(recv.unbind(0)), send)
    host = recv.cpu().numpy()
    out = {}
    for k, name in enumerate(names):
        out[name] = np.concatenate([host[r, k, :(b[r + 1] - b[r]) * rec].view(wdt).reshape(-1, 3) for r in range(world)])
    return out


def sweep_sharded(clean, noisy, grids=None, select=True, chunk_items=DEFAULT_CHUNK_ITEMS, device=None, engine_kwargs=None,
                  tables=False, pesq=None, pesq_scorer=None, pesq_workers=None):
    """Each rank sweeps its block of utterances AND runs the selection scan for them on its own device; the
    only exchange is one all_gather of the winners' records (every rank returns the winners of all
    utterances).  ``tables=True`` additionally all-gathers the per-point score tables (128 MB for the
    824-utterance job) - for callers that want every candidate's score on every rank.
    ``clean`` / ``noisy`` are the full host arrays [U, L] (each rank slices its block); ``pesq`` =
    {alg: [U][P]} for the full dataset (each rank slices its rows), or ``pesq_scorer`` to have each rank's
    candidates scored by its own host process pool while its sweep runs (:mod:`.pesq_pool`)."""
    import warnings
    import torch.distributed as dist
    from . import sweep as sw
    from .engine import SweepEngine
    grids = grids or sw.DEFAULT_GRIDS
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n_utts = clean.shape[0]
    if n_utts < world:
        raise ValueError(f"{n_utts} utterances cannot be sharded over {world} ranks")
    b = shard_bounds(n_utts, world)
    u_pad = max(b[r + 1] - b[r] for r in range(world))
    sl = slice(b[rank], b[rank + 1])
    local_pesq = None if pesq is None else {name: sw._pesq_array(pesq[name])[sl] for name in pesq}
    if pesq_scorer is not None:
        eng = SweepEngine(clean[sl], noisy[sl], chunk_items=min(chunk_items, sw.PESQ_CHUNK_ITEMS), **(engine_kwargs or {}))
        items, local_pesq = sw.run_engine_device_with_pesq(eng, pesq_scorer, grids, u_pad=u_pad, pesq_workers=pesq_workers)
        pesq = local_pesq
    else:
        eng = SweepEngine(clean[sl], noisy[sl], chunk_items=chunk_items, **(engine_kwargs or {}))
        items = sw.run_engine_device(eng, grids, u_pad=u_pad)
    points = {name: pts for name, pts, _, _ in items}
    winners = selection = None
    if select:
        winners = gather_winners(eng, sw.select_winners_device(eng, items, local_pesq), n_utts, u_pad)
        if pesq is None:
            warnings.warn("selection without PESQ: only the 'stoi' winner is available", stacklevel=2)
        selection = sw.selection_from_winners(points, winners, pesq is not None)
    scores = gather_device_scores(eng, items, n_utts, device=device) if tables else None
    return {"scores": scores, "points": points, "winners": winners, "selection": selection, "local_engine": eng}


def deal_by_length(lengths, world_size):
    """Variable-length corpus -> per-rank lists of pair indices: longest first, dealt round-robin, so that every
    rank holds the same number of pairs (+-1) and nearly the same number of samples."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[r::world_size] for r in range(world_size)]


def sweep_pairs_sharded(pairs, grids=None, chunk_items=DEFAULT_CHUNK_ITEMS, engine_kwargs=None, in_flight=12):
    """Variable-length form of :func:`sweep_sharded`: ``pairs`` = the whole corpus as [(clean, noisy)] 1-D arrays on
    every rank; each rank takes its deal of :func:`deal_by_length`, runs :func:`sweep.sweep_pairs` (one engine per
    distinct length, several in flight) with the selection on its device, and one all_gather of the 48-byte winner
    records gives every rank the winners of every pair, in input order."""
    import torch
    import torch.distributed as dist
    from . import sweep as sw
    grids = grids or sw.DEFAULT_GRIDS
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    deals = deal_by_length([len(p[0]) for p in pairs], world)
    mine = deals[rank]
    out = sw.sweep_pairs([pairs[i] for i in mine], grids=grids, select=True, chunk_items=chunk_items,
                         engine_kwargs=engine_kwargs, tables=False, in_flight=in_flight) if mine else None
    names = [name for name, _ in grids]
    points = {name: sw.cached_points(name, ranges) for name, ranges in grids}
    from ._lib import WINNER_DTYPE as wdt
    rec = 3 * wdt.itemsize
    n_pad = max(len(d) for d in deals)
    host = np.zeros((len(names), n_pad * rec), dtype=np.uint8)
    for k, name in enumerate(names):
        if mine:
            host[k, :len(mine) * rec] = np.ascontiguousarray(out["winners"][name]).reshape(-1).view(np.uint8)
    if world > 1:
        send = torch.from_numpy(host)
        on_gpu = dist.get_backend() == "nccl"
        if on_gpu:
            send = send.cuda()
        recv = torch.empty((world,) + tuple(send.shape), dtype=torch.uint8, device=send.device)
        if on_gpu:
            dist.all_gather_into_tensor(recv, send)
        else:
            dist.all_gather(list(recv.unbind(0)), send)
        gathered = recv.cpu().numpy()
    else:
        gathered = host[None]
    winners = {name: np.zeros((len(pairs), 3), dtype=wdt) for name in names}
    for r, d in enumerate(deals):
        for k, name in enumerate(names):
            winners[name][d] = gathered[r, k, :len(d) * rec].view(wdt).reshape(-1, 3)
    return {"points": points, "winners": winners, "selection": sw.selection_from_winners(points, winners, False),
            "local_pairs": mine, "nominal": sum(len(points[n]) for n in names) * len(pairs)}
