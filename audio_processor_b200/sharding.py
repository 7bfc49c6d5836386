"""Multi-GPU plumbing: one process per GPU, clips sharded across ranks, NO collective on the data path.

The path partitions by clip: FIR/STFT halos, the silence scan and Whisper's global max are all
intra-clip, so ranks never exchange audio.  NCCL (torch.distributed over NVLink) is used only after the last
kernel, to gather the per-clip segment tables (a few KB) — `gather_segment_tables` — and for the barrier /
max-over-ranks timing of the benchmark.  The same code runs on the `gloo` backend for CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_clips(durations: Sequence[float], world_size: int) -> List[List[int]]:
    """Duration-balanced greedy bin packing (longest first).  Returns clip indices per rank; deterministic."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(durations)), key=lambda i: (-float(durations[i]), i))
    loads = [0.0] * world_size
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += float(durations[i])
    for b in bins:
        b.sort()
    return bins


def pack_tables(tables: Sequence[Sequence[Sequence[int]]], cap: int):
    """[[start_ms, end_ms], ...] per clip -> (counts int32 [n], padded int32 [n, cap, 2]) torch tensors (CPU)."""
    import torch
    n = len(tables)
    counts = torch.zeros(n, dtype=torch.int32)
    padded = torch.zeros((n, cap, 2), dtype=torch.int32)
    for i, t in enumerate(tables):
        k = len(t)
        if k > cap:
            raise ValueError(f"clip {i}: {k} segments exceed cap={cap}")
        counts[i] = k
        if k:
            padded[i, :k] = torch.as_tensor(t, dtype=torch.int32)
    return counts, padded


def gather_segment_tables(clip_ids: Sequence[int], tables: Sequence[Sequence[Sequence[int]]], cap: int, device=None,
                          group=None) -> List[Tuple[int, List[List[int]]]]:
    """all_gather the per-clip kept/nonsilent tables of every rank.  Returns [(clip_id, table), ...] sorted by
    clip id, identical on every rank.  Ranks may hold different numbers of clips."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(zip([int(c) for c in clip_ids], [[list(map(int, r)) for r in t] for t in tables]))
    world = dist.get_world_size(group)
    counts, padded = pack_tables(tables, cap)
    ids = torch.as_tensor(list(clip_ids), dtype=torch.int32)
    n_local = torch.tensor([len(clip_ids)], dtype=torch.int32)
    if device is not None:
        counts, padded, ids, n_local = (t.to(device) for t in (counts, padded, ids, n_local))
    ns = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(ns, n_local, group=group)
    n_max = max(int(x.item()) for x in ns)

    def pad_rows(t, rows):
        if t.shape[0] == rows:
            return t.contiguous()
        extra = torch.zeros((rows - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        return torch.cat([t, extra], 0).contiguous()

    ids_p, counts_p, padded_p = pad_rows(ids, n_max), pad_rows(counts, n_max), pad_rows(padded, n_max)
    g_ids = [torch.zeros_like(ids_p) for _ in range(world)]
    g_counts = [torch.zeros_like(counts_p) for _ in range(world)]
    g_padded = [torch.zeros_like(padded_p) for _ in range(world)]
    dist.all_gather(g_ids, ids_p, group=group)
    dist.all_gather(g_counts, counts_p, group=group)
    dist.all_gather(g_padded, padded_p, group=group)
    out = []
    for r in range(world):
        nr = int(ns[r].item())
        idr, cr, pr = g_ids[r].cpu(), g_counts[r].cpu(), g_padded[r].cpu()
        for j in range(nr):
            k = int(cr[j])
            out.append((int(idr[j]), pr[j, :k].tolist()))
    out.sort(key=lambda x: x[0])
    return out
