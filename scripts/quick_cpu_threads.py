"""The reference arm's throughput against its OpenMP thread count (the GPU boxes expose 128 logical CPUs under a
16-CPU cgroup quota): which thread count is the fair 'all the host threads it can use'?  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from nanopolish_b200 import synth  # noqa: E402
from oracle.oracle_py import RefOracle  # noqa: E402

nuc = synth.load_model("nucleotide")
rs = synth.gen_reads(700, 4000, nuc, seed=42)
jobs = synth.scorereads_jobs(rs, 500, model_id=0, keep_seqs=True)
ro = RefOracle()
mh = [ro.builtin_model("nucleotide")]
rh = ro.register_reads(rs.reads, rs.ev_mean, rs.ev_start_time, mh[0])
sel = np.arange(min(4096, jobs.jobs.shape[0]))
sub = np.ascontiguousarray(jobs.jobs[sel])
seqs = [jobs.seqs[j] for j in sel]
E = np.abs(sub["event_stop"].astype(np.int64) - sub["event_start"].astype(np.int64)) + 1
out = {}
for t in (16, 32, 64, 128):
    best = None
    for rep in range(2):
        _, secs = ro.score_batch(rh, sub, seqs, mh, threads=t)
        best = secs if best is None else min(best, secs)
    out[str(t)] = float(E.sum() / best)
print(json.dumps({"workload": "scorereads windows, compiled reference, events/s by OpenMP thread count", "jobs": int(sel.shape[0]), "events_per_sec": out}))
