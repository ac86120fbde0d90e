"""Ad-hoc device-resident timing of the forward kernel (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopolish_b200 import synth
from nanopolish_b200.engine import Engine

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
kind = sys.argv[3] if len(sys.argv) > 3 else "scorereads"
model = synth.load_model("nucleotide")
t = time.time(); rs = synth.gen_reads(n_reads, n_events, model, seed=42, cpg_keep=0.3 if kind == "methylation" else 1.0); print("gen", time.time() - t)
eng = Engine(0)
mid = eng.model_upload(model)
if kind == "methylation":
    cpg = synth.load_model("cpg"); cid = eng.model_upload(cpg)
    jobs = synth.methylation_jobs(rs, model_id=cid)
else:
    jobs = synth.scorereads_jobs(rs, 500)
print("jobs", jobs.jobs.shape[0], "scored events", jobs.scored_events, "cells", jobs.block_cells)
eng.reads_load(rs.reads, rs.ev_mean, rs.ev_start_time)
eng.hmm_jobs_load(jobs.kmer_ranks, jobs.jobs)
for it in range(4):
    eng.hmm_score(); eng.sync()
    ms, n = eng.last_kernel_ms()
    print(f"iter {it}: {ms:.3f} ms, {n} launches, {jobs.scored_events/ms*1e3:.3e} events/s, {jobs.block_cells/ms*1e3:.3e} cells/s")
t = time.time(); out = eng.hmm_score_batch(rs.reads, rs.ev_mean, rs.ev_start_time, jobs.kmer_ranks, jobs.jobs); dt = time.time() - t
print(f"e2e one-shot {dt*1e3:.1f} ms -> {jobs.scored_events/dt:.3e} events/s; mean score/event {out.sum()/jobs.scored_events:.3f}")
