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


class ByteGather:
    """The same exchange for a caller that repeats it with page-locked buffers (bench.py's call-methylation e2e: every rank's TSV
    bytes to rank 0): staging on the device and the page-locked landing area on rank 0 are allocated once, the copies are
    asynchronous, and each call is one tiny all_gather of the byte counts plus ONE padded gather.  `device` None: CPU tensors (gloo)."""

    def __init__(self, cap_bytes: int, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.group, self.device, self.cap = group, device, int(cap_bytes)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = device if device is not None else "cpu"
        self.local = torch.empty(self.cap, dtype=torch.uint8, device=dev)
        self.counts_dev = torch.empty(self.world, dtype=torch.int64, device=dev)
        self.bufs = [torch.empty(self.cap, dtype=torch.uint8, device=dev) for _ in range(self.world)] if self.rank == 0 else None
        pin = device is not None and torch.cuda.is_available()
        self.out = torch.empty((self.world, self.cap), dtype=torch.uint8, pin_memory=pin) if self.rank == 0 else None

    def gather(self, host_bytes, n: int):
        """host_bytes: 1-D uint8 torch tensor (page-locked for an asynchronous copy), n valid bytes.  Rank 0 gets the list of per-rank
        numpy views into its landing area (valid until the next call); other ranks None."""
        import torch
        import torch.distributed as dist
        assert n <= self.cap
        self.local[:n].copy_(host_bytes[:n], non_blocking=True)
        c = torch.tensor([n], dtype=torch.int64).to(self.local.device, non_blocking=True)
        dist.all_gather_into_tensor(self.counts_dev, c, group=self.group)
        counts = [int(x) for x in self.counts_dev.tolist()]
        maxc = max(counts)
        dist.gather(self.local[:maxc], [b[:maxc] for b in self.bufs] if self.rank == 0 else None, dst=0, group=self.group)
        if self.rank != 0:
            return None
        for r, k in enumerate(counts):
            self.out[r, :k].copy_(self.bufs[r][:k], non_blocking=True)
        if self.local.is_cuda:
            torch.cuda.current_stream(self.local.device).synchronize()
        return [self.out[r, :k].numpy() for r, k in enumerate(counts)]


def scatter_results(parts_scores: list, owner: np.ndarray) -> np.ndarray:
    """Rank-0 reassembly: per-rank result vectors (in each rank's local job order) back to global job order."""
    out = np.empty(owner.shape[0], np.float32)
    for r, s in enumerate(parts_scores):
        out[owner == r] = s.cpu().numpy() if hasattr(s, "cpu") else np.asarray(s)
    return out
