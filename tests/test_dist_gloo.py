"""The N>1 path on CPU: world_size-2 gloo.  Each rank scores its shard of reads (with the oracle here —
there is no GPU in this container; on the GPU box the same plumbing carries the CUDA scores over NCCL),
rank 0 gathers with ONE padded gather and must reproduce the single-process job order and values."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanopolish_b200 import synth
from nanopolish_b200.dist import gather_records_to_rank0, gather_to_rank0, job_owner, partition_reads, scatter_results


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_py import PortOracle
    model = synth.load_model("nucleotide")
    rs = synth.gen_reads(7, 900, model, seed=3)
    rs.reads["n_events"]  # same data on every rank (deterministic); each scores only its shard
    jobs = synth.scorereads_jobs(rs, 150, rc_every=3)
    parts = partition_reads(rs.reads["n_events"], world)
    owner = job_owner(jobs.jobs["read"], parts)
    mine = np.flatnonzero(owner == rank)
    port_o = PortOracle()
    local, _ = port_o.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, [model], jobs.kmer_ranks,
                                      np.ascontiguousarray(jobs.jobs[mine]))
    got = gather_to_rank0(torch.from_numpy(local))
    if rank == 0:
        full = scatter_results(got, owner)
        want, _ = port_o.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, [model], jobs.kmer_ranks, jobs.jobs)
        np.save(out_path, np.stack([full, want]))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    n = rng.integers(500, 9000, 1000)
    for world in (1, 2, 4, 8):
        parts = partition_reads(n, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(1000))
        loads = np.array([n[p].sum() for p in parts])
        assert loads.max() - loads.min() <= n.max()
        for p in parts:
            assert np.all(np.diff(p) > 0)


def test_two_rank_gather_matches_single_process(tmp_path):
    out = str(tmp_path / "res.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    full, want = np.load(out)
    assert np.array_equal(full.view(np.uint32), want.view(np.uint32))


def _records_worker(rank, world, port, out_path):
    """eventalign sharded by read: every rank aligns its reads (restatement + plain-C Viterbi here, the chain kernel on
    the GPU box) and holds a different number of 12-byte records; rank 0 gets them with one padded gather."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import eventalign_py as EP
    from oracle.oracle_py import PortOracle
    from tests import eventalign_cases as EC
    model, rs, cases = EC.build_cases(5, 700, seed=41)
    cases = cases[:5]
    parts = partition_reads(rs.reads["n_events"], world)
    po = PortOracle()

    def align(c):
        al = EP.align_read_to_ref(c["read"], c["contig_name"], c["fetched"], c["ref_pos"], c["flag"], c["cigar"], c["read_idx"],
                                  EC.port_align_fn(po, rs, model, EC.read_slot(c, rs.n_reads)))
        rec = np.zeros(len(al), synth.EA_RECORD_DT)
        for i, a in enumerate(al):
            rec[i] = (a.ref_position, a.event_idx, a.hmm_state.encode(), (c["read_idx"], 0, 0))      # read index rides in the padding
        return rec
    mine = np.concatenate([align(cases[i]) for i in parts[rank]]) if parts[rank].size else np.zeros(0, synth.EA_RECORD_DT)
    got = gather_records_to_rank0(mine)
    if rank == 0:
        assert len(got) == world and sum(g.shape[0] for g in got) > 2000 and len({g.shape[0] for g in got}) == world
        merged = np.concatenate(got)
        want = np.concatenate([align(c) for c in cases])
        order = np.argsort(merged["reserved"][:, 0], kind="stable")          # back to read order
        np.save(out_path, np.stack([merged[order].view(np.uint8).reshape(-1), want.view(np.uint8).reshape(-1)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_of_variable_length_records(tmp_path):
    out = str(tmp_path / "rec.npy")
    port = _free_port()
    mp.spawn(_records_worker, args=(2, port, out), nprocs=2, join=True)
    merged, want = np.load(out)
    assert np.array_equal(merged, want)


def _bytes_worker(rank, world, port, out_path):
    """ByteGather (the page-locked, allocate-once form bench.py uses for the TSV bytes): three rounds with different lengths per rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nanopolish_b200.dist import ByteGather
    g = ByteGather(5000)
    ok = True
    for rnd in range(3):
        n = 1000 * (rnd + 1) + 37 * rank + (0 if rnd < 2 else -1000 * 3 + 5)        # last round: 5 + 37 * rank bytes
        mine = torch.from_numpy(((np.arange(5000) * (rank + 3) + rnd) % 251).astype(np.uint8))
        got = g.gather(mine, n)
        if rank == 0:
            for r, part in enumerate(got):
                want_n = 1000 * (rnd + 1) + 37 * r + (0 if rnd < 2 else -1000 * 3 + 5)
                want = ((np.arange(5000) * (r + 3) + rnd) % 251).astype(np.uint8)[:want_n]
                ok = ok and part.shape[0] == want_n and np.array_equal(part, want)
    if rank == 0:
        np.save(out_path, np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_byte_gather_with_reused_staging(tmp_path):
    out = str(tmp_path / "bytes.npy")
    port = _free_port()
    mp.spawn(_bytes_worker, args=(2, port, out), nprocs=2, join=True)
    assert bool(np.load(out)[0])
