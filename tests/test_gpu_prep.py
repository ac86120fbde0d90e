"""Parity of the remaining load_from_raw steps on the device — raw trimming before event detection and calibration
after ABEA — with the oracle, through the C ABI.  Integer outputs (ranges, event maps, counts,
statuses) must be identical; the FP64 calibration is summed in the reference's order, so it is compared bit for bit."""
import numpy as np
import pytest

from nanopolish_b200 import synth
from nanopolish_b200._lib import NphError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nuc(engine):
    m = synth.load_model("nucleotide")
    return m, engine.model_upload(m)


def _stalled(model, seed, n, leader=0, tail=0):
    raw, reads = synth.gen_raw(1, n, model, seed=seed)
    x = raw[:int(reads[0]["n_samples"])]
    return np.concatenate([np.full(leader, 210.0, np.float32), x, np.full(tail, 95.5, np.float32)]).astype(np.float32)


def _pack(signals):
    reads = np.zeros(len(signals), synth.RAW_READ_DT)
    off = 0
    for i, x in enumerate(signals):
        reads[i] = (off, 0, x.shape[0], 0)
        off += x.shape[0]
    return np.concatenate(signals), reads


@pytest.mark.parametrize("perc,chunk", [(0.0, 100), (0.0, 64), (0.3, 100), (0.77, 37), (1.0, 100), (0.5, 128), (0.0, 2)])
def test_trim_ranges_identical(engine, nuc, port_oracle, perc, chunk):
    model, _ = nuc
    signals = [_stalled(model, 5, 12000), _stalled(model, 6, 9000, leader=730), _stalled(model, 7, 9037, leader=300, tail=1250),
               _stalled(model, 8, 150, leader=100), _stalled(model, 9, 260), _stalled(model, 10, 36000, leader=1100, tail=100),
               _stalled(model, 11, 57), np.full(1000, 80.0, np.float32)]        # shorter than a chunk; perfectly flat
    raw, reads = _pack(signals)
    got = engine.trim_raw_batch(raw, reads, 200, 10, chunk, perc)
    n_ok = 0
    for i, x in enumerate(signals):
        ok, s, e = port_oracle.trim_raw(x, 200, 10, chunk, perc)
        assert (int(got[i]["start"]), int(got[i]["end"])) == (s, e), f"signal {i}"
        n_ok += ok
    assert n_ok >= (0 if perc == 1.0 else 4)
    assert int(got[7]["end"]) == 0          # flat signal: every MAD equals the threshold, nothing survives


def test_trim_rejects_what_the_reference_asserts(engine, nuc):
    model, _ = nuc
    raw, reads = _pack([_stalled(model, 5, 3000)])
    for args in [(200, 10, 1, 0.0), (200, 10, 100, 1.5), (200, 10, 100, -0.1)]:
        with pytest.raises(NphError):
            engine.trim_raw_batch(raw, reads, *args)
    with pytest.raises(NphError):
        engine.trim_raw_batch(raw, reads, 200, 10, 500, 0.0)       # chunk above what a warp holds: unsupported, not wrong


def _check_cal(engine, port_oracle, model, mid, rs, jobs, ranks, pairs, res):
    b2e, cal = engine.recalibrate_batch(rs.reads, rs.ev_mean, ranks, jobs, mid, pairs, res)
    for j in range(jobs.shape[0]):
        wb, wc = port_oracle.recalibrate(rs.reads, rs.ev_mean, model, ranks, jobs[j], pairs, int(res[j]["n_pairs"]))
        o, nk = int(jobs[j]["rank_off"]), int(jobs[j]["n_kmers"])
        assert np.array_equal(b2e[o:o + nk], wb), f"job {j}: base_to_event_map differs"
        assert int(cal[j]["n_used"]) == int(wc["n_used"]) and int(cal[j]["status"]) == int(wc["status"]), f"job {j}"
        for f in ("shift", "scale", "drift", "var", "events_per_base"):
            assert np.float64(cal[j][f]).view(np.uint64) == np.float64(wc[f]).view(np.uint64), f"job {j}: {f} {cal[j][f]!r} != {wc[f]!r}"
    return b2e, cal


@pytest.mark.parametrize("n_events,n_reads", [(4000, 6), (900, 8), (250, 12), (8000, 2)])
def test_calibration_after_abea_identical(engine, nuc, port_oracle, n_events, n_reads):
    model, mid = nuc
    rs = synth.gen_reads(n_reads, n_events, model, seed=8100 + n_events, rng_scalings=True)
    jobs, ranks, total = synth.abea_jobs(rs)
    pairs, res = engine.abea_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ranks, jobs, mid, total)
    _, cal = _check_cal(engine, port_oracle, model, mid, rs, jobs, ranks, pairs, res)
    if n_events >= 900:
        assert (cal["status"] == 0).all() and (cal["n_used"] >= 200).all()
        assert np.abs(cal["shift"] - rs.reads["shift"]).max() < 2.5 and np.abs(cal["scale"] - rs.reads["scale"]).max() < 0.03
    else:
        assert (cal["status"] == 2).all()                 # fewer than 200 'M' events: left uncalibrated
        assert np.array_equal(cal["shift"], rs.reads["shift"]) and np.array_equal(cal["var"], rs.reads["var"])


def test_calibration_homopolymers_failed_reads_and_odd_pair_lists(engine, nuc, port_oracle):
    model, mid = nuc
    # long homopolymer runs: consecutive k-mers share a rank, so only the first of a run is an 'M' event
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, 1500, dtype=np.uint8)
    for s in range(40, 1400, 97):
        codes[s:s + 14] = codes[s]
    rs = synth.gen_reads_from_sequence(codes, 5, model, seed=12)
    jobs, ranks, total = synth.abea_jobs(rs)
    pairs, res = engine.abea_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ranks, jobs, mid, total)
    assert (res["n_pairs"] > 0).all()
    # read 1: pretend ABEA failed; read 2: a hand-made list with repeated events, gaps and an out-of-order tail
    res = res.copy(); pairs = pairs.copy()
    res[1]["n_pairs"] = 0
    o = int(jobs[2]["pairs_off"])
    hand = [(0, 0), (0, 1), (1, 1), (2, 1), (2, 2), (2, 2), (5, 9), (5, 7), (4, 8), (4, 8), (6, 3)]
    pairs[o:o + len(hand)] = np.array(hand, synth.PAIR_DT)
    res[2]["n_pairs"] = len(hand)
    # read 3: noise instead of signal, so the fit's residual variance is far above 2.5
    e0, ne = int(rs.reads[3]["event_off"]), int(rs.reads[3]["n_events"])
    ev = rs.ev_mean.copy()
    ev[e0:e0 + ne] = rng.uniform(60.0, 130.0, ne).astype(np.float32)
    rs2 = synth.ReadSet(rs.reads, ev, rs.ev_start_time, rs.seq_codes, rs.ev_kmer, rs.kmer_first_event, rs.k)
    _, cal = _check_cal(engine, port_oracle, model, mid, rs2, jobs, ranks, pairs, res)
    assert int(cal[1]["status"]) == 1 and cal[1]["events_per_base"] == 0.0
    assert int(cal[2]["status"]) == 2 and int(cal[2]["n_used"]) < 10
    assert int(cal[3]["status"]) == 4 and cal[3]["var"] > 2.5
    assert int(cal[0]["status"]) == 0 and int(cal[0]["n_used"]) < int(jobs[0]["n_kmers"]) - 100   # homopolymer runs collapsed
    # a pair outside the read is refused
    bad = pairs.copy(); bad[int(jobs[0]["pairs_off"]) + 3]["read_pos"] = 10 ** 6
    with pytest.raises(NphError):
        engine.recalibrate_batch(rs2.reads, rs2.ev_mean, ranks, jobs, mid, bad, res)


def test_raw_to_calibrated_read_chain(engine, nuc, port_oracle):
    """trim -> detect_events -> MoM -> ABEA -> calibration through the C ABI vs the same chain through the oracle: the
    order SquiggleRead::load_from_raw runs them in (src/nanopolish_squiggle_read.cpp:226-336)."""
    from oracle.prep_chain import oracle_chain, squiggle_events
    model, mid = nuc
    n_reads = 4
    raw, rreads, seqs = synth.gen_raw(n_reads, 30000, model, seed=501, return_seqs=True)
    signals = [raw[int(r["sample_off"]):int(r["sample_off"]) + int(r["n_samples"])] for r in rreads]
    want = oracle_chain(port_oracle, model, signals, seqs)
    rng = engine.trim_raw_batch(raw, rreads)
    trimmed = rreads.copy()
    trimmed["sample_off"] += rng["start"]
    trimmed["n_samples"] = rng["end"] - rng["start"]
    events = engine.detect_events_batch(raw, trimmed, synth.event_params(False))
    reads = np.zeros(n_reads, synth.READ_DT)
    means, times = [], []
    off = 0
    for i, ev in enumerate(events):
        assert np.array_equal(ev, want[i]["events"]) and ev.shape[0] > 1000
        _, t = squiggle_events(ev, 4000.0)
        reads[i] = (off, ev.shape[0], 0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
        means.append(ev["mean"]); times.append(t); off += ev.shape[0]
    rs = synth.ReadSet(reads, np.concatenate(means), np.concatenate(times), seqs, [None] * n_reads, [None] * n_reads, model.k)
    jobs, ranks, total = synth.abea_jobs(rs)
    ss = engine.mom_batch(rs.reads, rs.ev_mean, ranks, jobs, mid)
    for i in range(n_reads):
        assert tuple(ss[i]) == want[i]["mom"]
        reads[i]["shift"], reads[i]["scale"] = ss[i][0], ss[i][1]
    pairs, res = engine.abea_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ranks, jobs, mid, total)
    assert [int(v) for v in res["n_pairs"]] == [w["n_pairs"] for w in want] and (res["n_pairs"] > 0).all()
    b2e, cal = _check_cal(engine, port_oracle, model, mid, rs, jobs, ranks, pairs, res)
    for i in range(n_reads):
        assert cal[i].tobytes() == want[i]["cal"].tobytes()
        assert np.array_equal(b2e[int(jobs[i]["rank_off"]):][:int(jobs[i]["n_kmers"])], want[i]["b2e"])
    assert (cal["status"] == 0).all()
    assert np.abs(cal["scale"] - 1.0).max() < 0.06 and np.abs(cal["shift"]).max() < 6.0 and (cal["var"] < 2.0).all()


def _raw_jobs(signals, seqs, k, sample_rate=4000.0):
    jobs = np.zeros(len(signals), synth.RAW_JOB_DT)
    ranks = []
    soff = roff = 0
    for i, (x, c) in enumerate(zip(signals, seqs)):
        rk = synth.kmer_ranks_from_codes(c, k, 4)
        jobs[i] = (soff, roff, x.shape[0], rk.shape[0], sample_rate)
        ranks.append(rk)
        soff += x.shape[0]
        roff += rk.shape[0]
    return np.concatenate(signals), np.concatenate(ranks).astype(np.uint32), jobs


def test_load_from_raw_in_one_call(engine, nuc, port_oracle):
    """nph_load_from_raw_batch (raw samples + basecall ranks in; events, event map, scalings and QC out) vs the chain
    through the oracle, including reads that die at each stage."""
    from oracle.prep_chain import oracle_chain
    model, mid = nuc
    raw, rr, seqs = synth.gen_raw(5, 20000, model, seed=640, return_seqs=True)
    signals = [raw[int(r["sample_off"]):int(r["sample_off"]) + int(r["n_samples"])] for r in rr]
    signals.insert(2, np.full(3000, 77.0, np.float32)); seqs.insert(2, seqs[0][:300])          # nothing survives the trim
    short, _, sq = synth.gen_raw(1, 2400, model, seed=77, return_seqs=True)                      # alignment fails
    signals.append(short); seqs.append(sq[0])
    signals.append(_stalled(model, 12, 9000, leader=700, tail=400)); seqs.append(synth.gen_raw(1, 9000, model, seed=12, return_seqs=True)[2][0])
    signals.append(np.full(57, 90.0, np.float32)); seqs.append(seqs[0][:20])                    # shorter than one chunk
    tiny, _, tq = synth.gen_raw(1, 330, model, seed=5, return_seqs=True)                         # a handful of events
    signals.append(tiny); seqs.append(tq[0])
    want = oracle_chain(port_oracle, model, signals, seqs)
    flat, ranks, jobs = _raw_jobs(signals, seqs, model.k)
    off, mean, stdv, start, dur, b2e, cal = engine.load_from_raw_batch(flat, ranks, jobs, mid, synth.event_params(False))
    kinds = set()
    for i, w in enumerate(want):
        o, n = int(off[i]), int(off[i + 1] - off[i])
        if w["events"] is None:
            assert n == 0 and int(cal[i]["status"]) & 16
            kinds.add("trim")
            continue
        ev = w["events"]
        assert n == ev.shape[0]
        assert np.array_equal(mean[o:o + n], ev["mean"]) and np.array_equal(stdv[o:o + n], ev["stdv"])
        assert np.array_equal(dur[o:o + n], w["duration"]) and np.array_equal(start[o:o + n], w["start_time"])
        assert cal[i].tobytes() == w["cal"].tobytes(), f"read {i}: {cal[i]} != {w['cal']}"
        assert np.array_equal(b2e[int(jobs[i]["rank_off"]):][:int(jobs[i]["n_kmers"])], w["b2e"])
        kinds.add(int(w["cal"]["status"]))
    assert {"trim", 0, 1} <= kinds
    # the sample range each read kept (what SRF_LOAD_RAW_SAMPLES stores; eventalign --samples reads it back)
    kept = engine.last_trim_ranges(len(want))
    assert [(int(a), int(b)) for a, b in zip(kept["start"], kept["end"])] == [w["range"] for w in want]
    with pytest.raises(NphError):
        engine.last_trim_ranges(len(want) + 1)
    ms, launches = engine.last_kernel_ms()
    assert ms > 0 and launches >= 9
    # capacity too small for the events is reported, not overrun
    with pytest.raises(NphError):
        engine.load_from_raw_batch(flat, ranks, jobs, mid, synth.event_params(False), events_cap=1000)
    # a context that just ran the chain still scores (resident state was invalidated, not corrupted)
    rs = synth.gen_reads(3, 600, model, seed=2)
    hj = synth.scorereads_jobs(rs, 200, model_id=mid)
    got = engine.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, hj.kmer_ranks, hj.jobs)
    oj = hj.jobs.copy(); oj["model_id"] = 0
    wantS, _ = port_oracle.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, [model], hj.kmer_ranks, oj)
    assert np.array_equal(got.view(np.uint32), wantS.view(np.uint32))


def test_load_from_raw_direct_rna(engine, port_oracle):
    """The RNA branch of load_from_raw (src/nanopolish_squiggle_read.cpp:204-213,262-265): 5-mer model, scrappie's RNA
    detector parameters, MoM on the events in acquisition order, events turned around to 5'->3' (start times then
    descend) before ABEA and calibration — against the same chain through the oracle."""
    from oracle.prep_chain import oracle_chain
    # stands in for r9.4_70bps.u_to_t_rna.5mer: the 6-mer table averaged over its last base (neighbouring k-mers keep
    # correlated levels, which events straddling a boundary need to survive ABEA's emission QC)
    m6 = synth.load_model("nucleotide")
    sd5 = m6.level_stdv.reshape(1024, 4).mean(1)
    model = synth.PoreModel("derived.nucleotide.5mer", 5, "nucleotide", m6.level_mean.reshape(1024, 4).mean(1), sd5, np.log(sd5))
    mid = engine.model_upload(model)
    raw, rr, seqs = synth.gen_raw(4, 60000, model, seed=911, mean_dwell=40.0, return_seqs=True)
    # the pore reads RNA 3'->5': the trace is the basecall's levels back to front
    signals = [np.ascontiguousarray(raw[int(r["sample_off"]):int(r["sample_off"]) + int(r["n_samples"])][::-1]) for r in rr]
    want = oracle_chain(port_oracle, model, signals, seqs, sample_rate=3012.0, rna=True)
    flat, ranks, jobs = _raw_jobs(signals, seqs, model.k, sample_rate=3012.0)
    prm = synth.event_params(True)
    assert int(prm[0]["reverse_events"]) == 1
    off, mean, stdv, start, dur, b2e, cal = engine.load_from_raw_batch(flat, ranks, jobs, mid, prm)
    for i, w in enumerate(want):
        o, n = int(off[i]), int(off[i + 1] - off[i])
        ev = w["events"]
        assert n == ev.shape[0] and n > 500
        assert np.array_equal(mean[o:o + n], ev["mean"]) and np.array_equal(stdv[o:o + n], ev["stdv"])
        assert np.array_equal(dur[o:o + n], w["duration"]) and np.array_equal(start[o:o + n], w["start_time"])
        assert start[o] > start[o + n - 1]                                  # acquisition-order times on reversed events
        assert cal[i].tobytes() == w["cal"].tobytes(), f"read {i}: {cal[i]} != {w['cal']}"
        assert np.array_equal(b2e[int(jobs[i]["rank_off"]):][:int(jobs[i]["n_kmers"])], w["b2e"])
    assert (cal["status"] == 0).sum() >= 3                                  # the reversed events do align to the basecall
    # without the reversal the same signals cannot be aligned to the basecall
    prm_fwd = prm.copy(); prm_fwd[0]["reverse_events"] = 0
    cal_fwd = engine.load_from_raw_batch(flat, ranks, jobs, mid, prm_fwd)[6]
    assert (cal_fwd["status"] != 0).all()
