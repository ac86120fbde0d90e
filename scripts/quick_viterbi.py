"""Ad-hoc timing of the Viterbi kernel on eventalign-shaped segments (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopolish_b200 import synth
from nanopolish_b200.engine import Engine
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nuc = synth.load_model("nucleotide")
rs = synth.gen_reads(n_reads, 4000, nuc, seed=42)
eng = Engine(0); mid = eng.model_upload(nuc)
jobs = synth.scorereads_jobs(rs, 170, model_id=mid)
print("jobs", jobs.jobs.shape[0], "events", jobs.scored_events, "cells", jobs.block_cells)
for it in range(3):
    t = time.time(); al, sc = eng.hmm_align_batch(rs.reads, rs.ev_mean, rs.ev_start_time, jobs.kmer_ranks, jobs.jobs); dt = time.time() - t
    ms, nl = eng.last_kernel_ms()
    print(f"viterbi: kernels {ms:.2f} ms ({nl} launches) -> {jobs.scored_events/ms*1e3:.3e} events/s, {jobs.block_cells/ms*1e3:.3e} cells/s; e2e {dt*1e3:.0f} ms")
