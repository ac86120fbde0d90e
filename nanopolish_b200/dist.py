"""Multi-GPU plumbing: reads shard across ranks, results come back with one gather.

Reads are independent units for every kernel of the path (no cross-read term in
calculate_methylation_for_read, scorereads or ABEA), so there is no data-path collective during
compute; per-job results (log-likelihoods / LLRs) return to rank 0 with ONE collective per batch:
a padded gather (NCCL over NVLink on GPUs, gloo in the CPU tests).  torch.distributed is the
plumbing; the payload never touches Python objects.
"""
from __future__ import annotations

import numpy as np


def partition_reads(n_events: np.ndarray, world: int) -> list[np.ndarray]:
    """Split read indices over `world` ranks balanced by summed event count (greedy longest-first),
    keeping each rank's indices ascending so a read's jobs stay together and in input order."""
    order = np.argsort(-n_events.astype(np.int64), kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.empty(n_events.shape[0], np.int32)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += int(n_events[i])
    return [np.flatnonzero(owner == r) for r in range(world)]


def job_owner(job_reads: np.ndarray, parts: list[np.ndarray]) -> np.ndarray:
    """Rank that owns each job (the rank owning its read)."""
    n_reads = int(max(int(p.max()) if p.size else -1 for p in parts)) + 1
    read_owner = np.empty(n_reads, np.int32)
    for r, p in enumerate(parts):
        read_owner[p] = r
    return read_owner[job_reads]


def gather_to_rank0(local, counts: list[int] | None = None, group=None):
    """One padded gather of a 1-D tensor of per-job results to rank 0.

    `local` may be longer than this rank's count (pre-padded to max(counts)); returns on rank 0 the list of
    per-rank tensors trimmed to their counts, elsewhere None.  Exactly one data collective (plus, when
    counts is None, one tiny all_gather of the counts)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if counts is None:
        # one collective and ONE host read-back for all the counts
        c = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        allc = torch.empty(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(allc, c, group=group)
        counts = [int(x) for x in allc.tolist()]
    maxc = max(counts)
    if local.shape[0] != maxc:
        pad = torch.zeros(maxc, dtype=local.dtype, device=local.device)
        pad[:min(local.shape[0], maxc)] = local[:maxc]
        local = pad
    bufs = [torch.empty(maxc, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == 0 else None
    dist.gather(local, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return [b[:n] for b, n in zip(bufs, counts)]


def gather_records_to_rank0(records: np.ndarray, device=None, group=None):
    """The variable-length form: every rank holds a different number of fixed-size records — eventalign's
    nph_ea_record rows, call-methylation's per-site results — as a numpy structured array.  One tiny all_gather of the
    byte counts, then ONE padded gather of the raw bytes to rank 0, which gets the list of per-rank record arrays
    (other ranks: None).  `device`: where the staging tensor lives ("cuda:<local>" under NCCL, None for gloo)."""
    import torch

    raw = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).reshape(-1).copy())
    if device is not None:
        raw = raw.to(device)
    parts = gather_to_rank0(raw, None, group)
    if parts is None:
        return None
    return [p.cpu().numpy().view(records.dtype) for p in parts]


def scatter_results(parts_scores: list, owner: np.ndarray) -> np.ndarray:
    """Rank-0 reassembly: per-rank result vectors (in each rank's local job order) back to global job order."""
    out = np.empty(owner.shape[0], np.float32)
    for r, s in enumerate(parts_scores):
        out[owner == r] = s.cpu().numpy() if hasattr(s, "cpu") else np.asarray(s)
    return out
