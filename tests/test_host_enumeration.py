"""Host-side job enumeration of call-methylation (nanopolish_b200/host/nph_methylation.*) without a
device: MethylationCaller::add_reads — reads enumerated by worker threads into private job lists and spliced in read
order — must queue exactly the jobs (and k-mer ranks) that add_read queues read by read."""
import ctypes as C
import os

import numpy as np
import pytest

from nanopolish_b200 import synth
from tests.test_host_mirror import HOST_SO, _register, _register_reads

K = 6


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(HOST_SO)
    lib.nphh_last_error.restype = C.c_char_p
    lib.nphh_methylation_enumerate_seconds.restype = C.c_double
    return lib


def _enumerate(host, rh, metas, parallel):
    n = len(metas)
    flat = np.concatenate([m["pairs"] for m in metas]).reshape(-1).astype(np.int32)
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum([m["pairs"].shape[0] for m in metas])
    names = (C.c_char_p * n)(*[f"read_{i}".encode() for i in range(n)])
    refs = (C.c_char_p * n)(*[m["ref"].encode() for m in metas])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = np.array([m["rc"] for m in metas], np.uint8)
    nj, nr = C.c_uint64(), C.c_uint64()
    jobs = np.zeros(40000, synth.HMM_JOB_DT); ranks = np.zeros(4_000_000, np.uint32)
    secs = host.nphh_methylation_enumerate_seconds(n, p(np.array(rh, np.int32)), names, p(rc), p(rc), p(np.full(n, 10_000, np.int32)), refs,
                                                   p(flat), p(off), b"chr1", C.byref(nj), int(parallel), p(jobs), C.c_size_t(jobs.shape[0]),
                                                   p(ranks), C.c_size_t(ranks.shape[0]), C.byref(nr))
    assert secs >= 0, host.nphh_last_error()
    return jobs[:nj.value].copy(), ranks[:nr.value].copy()


def test_parallel_enumeration_equals_read_by_read(host):
    nuc, cpg = synth.load_model("nucleotide"), synth.load_model("cpg")
    rs = synth.gen_reads(70, 1500, nuc, seed=88, cpg_keep=0.3)
    host.nphh_clear()
    mh, ch = _register(host, nuc), _register(host, cpg)
    rh = _register_reads(host, rs, mh)
    for r in rh:
        host.nphh_read_add_model(r, b"cpg", ch)
    metas = []
    for i in range(rs.n_reads):
        codes = rs.seq_codes[i]
        nk = codes.shape[0] - K + 1
        kfe = np.minimum(rs.kmer_first_event[i], int(rs.reads[i]["n_events"]) - 1)
        if i % 3 == 2:          # reverse-strand record: events fall as reference positions rise
            ref = synth._CODE2DNA[(3 - codes[::-1]).astype(np.uint8)].tobytes().decode()
            pairs = np.stack([10_000 + np.arange(K, nk - K), kfe[nk - 1 - np.arange(K, nk - K)]], 1)
            rc = 1
        else:
            ref = synth._CODE2DNA[codes].tobytes().decode()
            pairs = np.stack([10_000 + np.arange(K, nk - K), kfe[K:nk - K]], 1)
            rc = 0
        metas.append(dict(ref=ref, pairs=pairs.astype(np.int32), rc=rc))
    os.environ.pop("NPH_HOST_THREADS", None)
    j0, r0 = _enumerate(host, rh, metas, parallel=False)
    j1, r1 = _enumerate(host, rh, metas, parallel=True)
    assert j0.shape[0] > 1000 and j0.shape[0] % 2 == 0            # two jobs per scored group
    assert j0.tobytes() == j1.tobytes() and np.array_equal(r0, r1)
    assert (j0["n_kmers"] >= 16).all() and (j0["flags"] == 3).all()
    host.nphh_clear()


def test_rolling_ranks_equal_get_kmer_rank(host):
    """HmmBatch::add fills the ranks by one rolling pass; they must be HMMInputSequence::get_kmer_rank's, both strands,
    plain and methylation alphabets (the rc strand of a methylated sequence is not a per-base complement)."""
    rng = np.random.default_rng(12)
    for alphabet, sym in (("nucleotide", "ACGT"), ("cpg", "ACGT")):
        for rep in range(20):
            seq = "".join(sym[c] for c in rng.integers(0, 4, int(rng.integers(6, 260))))
            if alphabet == "cpg" and rep % 2:
                seq = seq.replace("CG", "MG")
            for rc in (0, 1):
                assert host.nphh_kmer_ranks_rolling_check(alphabet.encode(), seq.encode(), 6, rc) == 0, (alphabet, seq, rc)
    assert host.nphh_kmer_ranks_rolling_check(b"nucleotide", b"ACG", 6, 0) == 0


def test_enumeration_equals_python_restatement(host):
    """The job list itself (event bounds, strands, k-mer ranks of the unmethylated and methylated windows) against an
    independent restatement of calculate_methylation_for_read's enumeration (src/basemods/nanopolish_basemods.cpp:238-
    372) — the same one the GPU test scores — so the host-side fast paths are covered without a device."""
    from tests.test_host_methylation import _find_by_ref_bounds, _ranks
    nuc, cpg = synth.load_model("nucleotide"), synth.load_model("cpg")
    rs = synth.gen_reads(9, 2200, nuc, seed=314, cpg_keep=0.3)
    host.nphh_clear()
    mh, ch = _register(host, nuc), _register(host, cpg)
    rh = _register_reads(host, rs, mh)
    for r in rh:
        host.nphh_read_add_model(r, b"cpg", ch)
    metas, want = [], []
    for i in range(rs.n_reads):
        codes = rs.seq_codes[i]
        nk = codes.shape[0] - K + 1
        kfe = np.minimum(rs.kmer_first_event[i], int(rs.reads[i]["n_events"]) - 1)
        if i % 2:
            ref = synth._CODE2DNA[(3 - codes[::-1]).astype(np.uint8)].tobytes().decode()
            pairs = [(10_000 + p_, int(kfe[nk - 1 - p_])) for p_ in range(K, nk - K)]
            rc = 1
        else:
            ref = synth._CODE2DNA[codes].tobytes().decode()
            pairs = [(10_000 + p_, int(kfe[p_])) for p_ in range(K, nk - K)]
            rc = 0
        metas.append(dict(ref=ref, pairs=np.array(pairs, np.int32), rc=rc))
        sites = [j for j in range(len(ref) - 1) if ref[j:j + 2] == "CG"]
        cur = 0
        while cur < len(sites):
            end = cur + 1
            while end < len(sites) and sites[end] - sites[end - 1] <= 10:
                end += 1
            first, last = sites[cur], sites[end - 1]
            cur = end
            sub_start, sub_end = first - 10, last + 10
            if sub_start <= 10 or last - first > 200:
                continue
            b = _find_by_ref_bounds(pairs, sub_start + 10_000, sub_end + 10_000)
            if b is None or abs(b[1] - b[0]) <= 10:
                continue
            subseq = ref[sub_start:sub_end + 1]
            for seq in (subseq, subseq.replace("CG", "MG")):
                want.append((i, b[0], b[1], 1 if b[0] <= b[1] else -1, rc, _ranks(host, "cpg", seq, rc)))
    jobs, ranks = _enumerate(host, rh, metas, parallel=True)
    assert jobs.shape[0] == len(want) > 300
    for jb, (read, e1, e2, stride, rc, rk) in zip(jobs, want):
        assert (int(jb["read"]), int(jb["event_start"]), int(jb["event_stop"]), int(jb["stride"]), int(jb["rc"]), int(jb["flags"])) == \
            (read, e1, e2, stride, rc, 3)
        assert np.array_equal(ranks[int(jb["rank_off"]):int(jb["rank_off"]) + int(jb["n_kmers"])], rk)
    host.nphh_clear()
