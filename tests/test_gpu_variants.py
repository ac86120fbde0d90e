"""SURVEY.md 8(f) row N2 on the device — nph_screen_edits_batch (csrc/variants.cu) through the C ABI: candidate generation, the
windows' event sequences, early-exit rounds and the accumulated Variant::quality of every candidate, against the restatement
(tests/var_restatement.py, pinned to the compiled reference in tests/test_oracle_vs_ref.py) and against the compiled reference itself."""
import math

import numpy as np
import pytest

from nanopolish_b200 import synth
from tests import var_restatement as vr

pytestmark = pytest.mark.gpu
K = 6
REGION = 5000


@pytest.fixture(scope="module")
def eng():
    from nanopolish_b200.engine import Engine
    e = Engine(0)
    e.model_upload(synth.load_model("nucleotide"))
    yield e
    e.close()


def _pileup(ref_len, depth, read_bases, seed, n_true=3):
    nuc = synth.load_model("nucleotide")
    ref, rs, recs, pairs = synth.gen_pileup(ref_len, depth, read_bases, nuc, seed=seed, region_start=REGION, n_true_variants=n_true)
    deltas, first = synth.compact_event_alignment(recs, pairs, int(recs["ref_len"].sum()))
    ref_chars = synth._CODE2DNA[ref]
    return nuc, ref, ref_chars, rs, recs, pairs, deltas, first


def _same(a, b):
    return (math.isnan(a) and math.isnan(b)) or a == b


@pytest.mark.parametrize("threshold,rpr,flags", [(30, 4, 0), (100, 8, 3), (10 ** 6, 5, 0)])
def test_qualities_equal_restatement(eng, port_oracle, threshold, rpr, flags):
    nuc, ref, ref_chars, rs, recs, pairs, deltas, first = _pileup(150, 14, 110, seed=7 + rpr)
    params = synth.screen_params(REGION, K, 10, threshold, flags, rpr)
    q, nr, scored = eng.screen_edits_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ref_chars, deltas, first, recs, params, indel_bias=0.9)
    ref_s = ref_chars.tobytes().decode()
    n_pos = ref.shape[0] - 1
    assert q.shape == (n_pos, 9)
    positive = exited = 0
    for pi in range(n_pos):
        want, n_seq, _ = vr.screen_position(port_oracle, rs, nuc, ref_s, REGION, REGION + pi, recs, pairs, 10, threshold, flags, 0.9, K)
        assert int(nr[pi]) == n_seq or math.isnan(want[0]) and math.isnan(want[2]) and math.isnan(want[4]) and math.isnan(want[6]), pi
        for c in range(9):
            assert _same(float(q[pi, c]), want[c]), (pi, c, q[pi], want)
        positive += sum(1 for v in want if v == v and v > 0)
        exited += sum(1 for v in want if v == v and abs(v) >= threshold)
    assert positive >= 2                                   # the planted substitutions win
    cnt = eng.screen_counts()
    if threshold < 10 ** 6:
        assert exited > 50 and cnt["jobs"] < cnt["jobs_without_exit"] and cnt["rounds"] >= 2
    else:
        assert exited == 0 and cnt["jobs"] == cnt["jobs_without_exit"]
    assert scored == cnt["scored_events"] > 0


def test_positions_outside_the_region_and_empty_pileup(eng):
    nuc, ref, ref_chars, rs, recs, pairs, deltas, first = _pileup(120, 6, 100, seed=3, n_true=0)
    params = synth.screen_params(REGION, K, 10, 100, 0, 8)
    q, nr, _ = eng.screen_edits_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ref_chars, deltas, first, recs, params)
    assert np.isnan(q[:10]).all() and np.isnan(q[-10:]).all()          # windows that leave the region: the reference skips the position
    assert not np.isnan(q[10:-10]).all(axis=1).any()
    # no records at all: every candidate keeps quality 0
    none = recs[:0]
    q0, nr0, ev0 = eng.screen_edits_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ref_chars, deltas[:0], first[:0], none, params)
    assert ev0 == 0 and (nr0 == 0).all() and np.nanmax(np.abs(q0)) == 0.0


def test_staged_form_and_compiled_reference(eng, ref_oracle):
    nuc, ref, ref_chars, rs, recs, pairs, deltas, first = _pileup(140, 12, 100, seed=21)
    params = synth.screen_params(REGION, K, 10, 40, 3, 4)
    q1, nr1, _ = eng.screen_edits_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ref_chars, deltas, first, recs, params, indel_bias=0.9)
    eng.reads_load(rs.reads, rs.ev_mean, rs.ev_start_time)
    eng.screen_load(ref_chars, deltas, first, recs, params, indel_bias=0.9)
    eng.screen_run()
    q2, nr2 = eng.screen_fetch()
    assert np.array_equal(nr1, nr2) and np.array_equal(np.nan_to_num(q1, nan=-1e300), np.nan_to_num(q2, nan=-1e300))
    # a few positions through the compiled reference's own score_variant_thresholded (one OpenMP thread)
    ref_s = ref_chars.tobytes().decode()
    ref_oracle.clear_reads()
    rh = ref_oracle.register_reads(rs.reads, rs.ev_mean, rs.ev_start_time, ref_oracle.builtin_model("nucleotide"))
    for pi in (15, 40, 77, 101):
        i = REGION + pi
        cs, ce = i - 10, i + 11
        seqs = vr.event_sequences(recs, pairs, cs, ce)
        cands = vr.candidates(ref_s, pi)
        got = ref_oracle.score_variants_thresholded([rh[r] for r, _, _ in seqs], [(e1, e2) for _, e1, e2 in seqs],
                                                    np.array([recs[r]["rc"] for r, _, _ in seqs], np.uint8), ref_s[cs - REGION:ce - REGION + 1], cs,
                                                    [(REGION + off, rseq, aseq) for _, off, rseq, aseq in cands], 3, 40, False, indel_bias=0.9)
        for (slot, _, _, _), v in zip(cands, got):
            assert float(q1[pi, slot]) == float(v), (pi, slot)
    ref_oracle.clear_reads()
