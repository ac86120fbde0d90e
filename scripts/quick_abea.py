"""Ad-hoc device-resident timing of the ABEA kernel (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopolish_b200 import synth
from nanopolish_b200.engine import Engine
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
model = synth.load_model("nucleotide")
rs = synth.gen_reads(n_reads, n_events, model, seed=42, rng_scalings=False)
jobs, ranks, total = synth.abea_jobs(rs)
eng = Engine(0); mid = eng.model_upload(model)
eng.reads_load(rs.reads, rs.ev_mean, rs.ev_start_time)
eng.abea_jobs_load(ranks, jobs, mid, total)
ev = int(rs.reads["n_events"].sum()); bands = int((rs.reads["n_events"] + jobs["n_kmers"] + 2).sum())
for it in range(3):
    eng.abea_run(); eng.sync(); ms, _ = eng.last_kernel_ms()
    print(f"iter {it}: {ms:.2f} ms  {ev/ms*1e3:.3e} events/s  {bands*100/ms*1e3:.3e} cells/s")
pairs, res = eng.abea_fetch()
print("ok reads", int((res['n_pairs'] > 0).sum()), "of", n_reads)
