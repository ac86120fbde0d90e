"""Seeded inputs shared by scripts/make_golden.py (which records the compiled reference's outputs for
them in tests/golden/) and by the tests that replay them against the oracle and the CUDA path."""
from nanopolish_b200 import synth


def make_hmm_cases():
    nuc = synth.load_model("nucleotide")
    cases = {}
    # scorereads-style segments, both strands, with drift
    rs = synth.gen_reads(6, 1300, nuc, seed=101, drift=True)
    cases["segments"] = dict(rs=rs, jobs=synth.scorereads_jobs(rs, 250, model_id=0, rc_every=2, keep_seqs=True),
                             alphabets=["nucleotide"], indel_bias=1.0)
    # short segments, variants-style indel bias
    rs = synth.gen_reads(5, 700, nuc, seed=202)
    cases["short_bias08"] = dict(rs=rs, jobs=synth.scorereads_jobs(rs, 60, model_id=0, rc_every=3, keep_seqs=True),
                                 alphabets=["nucleotide"], indel_bias=0.8)
    # call-methylation windows over the cpg alphabet, PRE|POST clip
    rs = synth.gen_reads(6, 1500, nuc, seed=303, cpg_keep=0.3)
    cases["methylation"] = dict(rs=rs, jobs=synth.methylation_jobs(rs, model_id=1, keep_seqs=True),
                                alphabets=["nucleotide", "cpg"], indel_bias=1.0)
    return cases


def make_abea_cases():
    nuc = synth.load_model("nucleotide")
    cases = {}
    cases["reads_2k"] = dict(rs=synth.gen_reads(4, 2000, nuc, seed=404, rng_scalings=False))
    cases["reads_short"] = dict(rs=synth.gen_reads(6, 300, nuc, seed=505, rng_scalings=False))
    return cases
