"""The plain-C oracle against the committed golden vectors (outputs of the compiled reference recorded
by scripts/make_golden.py) and against the known answers the reference's own unit tests hold
(src/test/nanopolish_test.cpp: "math" :267-275, "string functions" :243-245, "scalings" :277-325)."""
import os

import numpy as np
import pytest

from nanopolish_b200 import synth
from tests.golden_cases import make_abea_cases, make_hmm_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", ["segments", "short_bias08", "methylation"])
def test_hmm_golden(port_oracle, name):
    gold = np.load(os.path.join(GOLD, "hmm_golden.npz"))[name]
    case = make_hmm_cases()[name]
    rs, jobs = case["rs"], case["jobs"]
    models = [synth.load_model(a) for a in case["alphabets"]]
    got, _ = port_oracle.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, models, jobs.kmer_ranks,
                                         jobs.jobs, indel_bias=case["indel_bias"])
    assert gold.shape == got.shape
    assert np.array_equal(_bits(gold), _bits(got))


@pytest.mark.parametrize("name", ["reads_2k", "reads_short"])
def test_abea_golden(port_oracle, name):
    z = np.load(os.path.join(GOLD, "abea_golden.npz"))
    gp, gn = z[name + "_pairs"], z[name + "_npairs"]
    rs = make_abea_cases()[name]["rs"]
    jobs, ranks, total = synth.abea_jobs(rs)
    pairs, res, _ = port_oracle.abea_batch(rs.reads, rs.ev_mean, rs.ev_start_time, synth.load_model("nucleotide"),
                                           ranks, jobs, total)
    assert [int(x) for x in gn] == [int(x) for x in res["n_pairs"]]
    o = 0
    for i in range(rs.n_reads):
        n = int(gn[i])
        b = pairs[int(jobs[i]["pairs_off"]):int(jobs[i]["pairs_off"]) + n]
        assert np.array_equal(gp[o:o + n, 0], b["ref_pos"]) and np.array_equal(gp[o:o + n, 1], b["read_pos"])
        o += n


def test_known_answers(port_oracle):
    # "math": normal_pdf(2.25; mu 4, sigma 2) = 0.1360275 and log_normal_pdf == log(normal_pdf)
    lp = port_oracle.lib.npo_log_normal_pdf(2.25, 4.0, 2.0, float(np.float32(np.log(2.0))))
    assert abs(np.exp(lp) - 0.1360275) < 1e-6
    # "string functions": kmer_rank("GATGA", 5) == 568
    assert int(synth.kmer_ranks_from_codes(synth.encode("GATGA", "nucleotide"), 5, 4)[0]) == 568
    # p7_FLogsum: approximates log(e^a + e^b) to the table's 0.001-nat quantisation
    for a, b in [(-0.4, -0.5), (-3.0, -9.5), (-100.0, -100.0)]:
        got = port_oracle.lib.npo_logsum(a, b)
        assert abs(got - np.logaddexp(a, b)) < 6e-4
    assert port_oracle.lib.npo_logsum(-1.0, float("-inf")) == -1.0
    assert port_oracle.lib.npo_logsum(-1.0, -17.0) == -1.0


def test_scalings_known_answer(port_oracle):
    """The reference's "scalings" test: events drawn from k-mer rank 100 of the r9.4 6-mer model under
    set4(shift 10, scale 1.2, drift 0.5, var 1.3); log_probability_match_r9 must equal the closed form."""
    import ctypes as C
    model = synth.load_model("nucleotide")
    rank, n = 100, 100
    shift, scale, drift, var = 10.0, 1.2, 0.5, 1.3
    rng = np.random.default_rng(1)
    t = np.arange(n) * (10.0 / 4000.0)
    g_mean = (shift + scale * model.level_mean[rank] + t * drift).astype(np.float32)
    g_stdv = np.float32(var * model.level_stdv[rank])
    ev = (g_mean + g_stdv * rng.standard_normal(n)).astype(np.float32)
    reads = np.zeros(1, synth.READ_DT)
    reads[0] = (0, n, 0, scale, shift, drift, var, np.log(var), 1.7)
    marr = port_oracle.models([model])
    port_oracle.lib.npo_log_probability_match.restype = C.c_float
    for i in range(n):
        lp = port_oracle.lib.npo_log_probability_match(reads.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p),
                                                       t.ctypes.data_as(C.c_void_p), marr, C.c_uint32(rank), C.c_uint32(i))
        a = (ev[i] - g_mean[i]) / g_stdv
        want = np.log(0.3989422804014327) - np.log(g_stdv) - 0.5 * a * a
        assert abs(lp - want) < 1e-3 * max(1.0, abs(want))


def test_flank_table_is_both_pre_and_post(port_oracle):
    f = port_oracle.flank_table(50)
    assert abs(f[0] - np.log(0.5)) < 1e-7
    assert abs(f[1] - (np.log(0.5) - 3.0 + np.log(0.1))) < 1e-6
    assert np.all(np.diff(f[1:]) < 0)
