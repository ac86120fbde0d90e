"""Plain-Python restatement of `nanopolish variants` candidate screening — test infrastructure (the checker of csrc/variants.cu).

Follows generate_candidate_single_base_edits (src/nanopolish_call_variants.cpp:288-361), AlignmentDB::get_event_subsequences
(src/alignment/nanopolish_alignment_db.cpp:172-221) and score_variant_thresholded (src/common/nanopolish_variant.cpp:765-799, the
single-thread order).  Pinned: tests/test_oracle_vs_ref.py runs its qualities against the compiled reference's score_variant_thresholded."""
import math

import numpy as np

from nanopolish_b200 import synth
from tests.meth_restatement import find_by_ref_bounds

BASES = "ACGT"


def candidates(ref: str, i0: int):
    """the Variants of position i (offset i0 into ref) in the order the reference generates them: (slot, ref_position offset, ref_seq, alt_seq)"""
    out = []
    b = ref[i0]
    for j in range(4):
        if BASES[j] != b:
            out.append((2 * j, i0, b, BASES[j]))                       # substitution
        if BASES[j] != b:
            out.append((2 * j + 1, i0, b, b + BASES[j]))               # insertion ("A" -> "AA" is redundant)
    if ref[i0 - 1] != ref[i0]:
        out.append((8, i0 - 1, ref[i0 - 1:i0 + 1], ref[i0 - 1]))       # deletion ("AA" -> "A" is redundant)
    return out


def event_sequences(records, pairs, cs, ce):
    """get_event_subsequences: [(record index, e1, e2)] in record order"""
    out = []
    for r, R in enumerate(records):
        pr = pairs[int(R["pair_off"]):int(R["pair_off"]) + int(R["n_pairs"])]
        if pr.shape[0] == 0:
            continue
        b = find_by_ref_bounds(pr["ref_pos"].tolist(), pr["read_pos"].tolist(), cs, ce)
        if b is None:
            continue
        if abs(b[0] - b[1]) / abs(ce - cs) < 20:
            out.append((r, b[0], b[1]))
    return out


def apply(window: str, off: int, ref_seq: str, alt_seq: str) -> str:
    assert window[off:off + len(ref_seq)] == ref_seq
    return window[:off] + alt_seq + window[off + len(ref_seq):]


def screen_position(port_oracle, rs, model, ref: str, region_start: int, i: int, records, pairs, flank=10, threshold=100, flags=0, indel_bias=1.0, k=6):
    """-> (qualities[9] with NaN for candidates the reference does not generate, number of event sequences, the windows' (record, e1, e2))"""
    n_ref = len(ref)
    cs, ce = i - flank, i + 1 + flank
    q = [math.nan] * 9
    if cs < region_start or ce > region_start + n_ref - 1:
        return q, 0, []
    window = ref[cs - region_start:ce - region_start + 1]
    seqs = event_sequences(records, pairs, cs, ce)
    cands = candidates(ref, i - region_start)
    hap = [window] + [apply(window, off - (cs - region_start), rseq, aseq) for (_, off, rseq, aseq) in cands]
    # one oracle batch: per read, the base and every candidate
    rows, ranks_list = [], []
    for (r, e1, e2) in seqs:
        rc = int(records[r]["rc"])
        for h in hap:
            codes = synth.encode(h, "nucleotide")
            ranks_list.append(synth.dna_rc_kmer_ranks(codes, k) if rc else synth.kmer_ranks_from_codes(codes, k, 4))
            rows.append((int(records[r]["read"]), 0, e1, e2, rc, flags))
    totals = [0.0] * len(cands)
    if rows:
        jobs = synth._finish_jobs(rows, ranks_list)
        jobs.jobs["stride"] = np.where(jobs.jobs["rc"] == 1, -1, 1)          # EventAlignmentRecord::stride
        sc, _ = port_oracle.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, [model], jobs.kmer_ranks, jobs.jobs, indel_bias=indel_bias)
        nh = len(hap)
        for ri in range(len(seqs)):
            base = float(sc[ri * nh])
            for c in range(len(cands)):
                if abs(totals[c]) < threshold:
                    totals[c] += float(sc[ri * nh + 1 + c]) - base
    for (slot, _, _, _), t in zip(cands, totals):
        q[slot] = t
    return q, len(seqs), seqs
