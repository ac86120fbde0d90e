"""Ad-hoc timing of the raw-trim and calibration kernels and of the whole load_from_raw chain (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopolish_b200 import synth
from nanopolish_b200.engine import Engine
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 36000
nuc = synth.load_model("nucleotide")
raw, rreads, seqs = synth.gen_raw(n_reads, n_samples, nuc, seed=5, return_seqs=True)
eng = Engine(0)
mid = eng.model_upload(nuc)
prm = synth.event_params(False)
for it in range(3):
    T = {}
    t0 = time.time()
    rng = eng.trim_raw_batch(raw, rreads); T["trim"] = eng.last_kernel_ms()[0]
    trimmed = rreads.copy(); trimmed["sample_off"] += rng["start"]; trimmed["n_samples"] = rng["end"] - rng["start"]
    events = eng.detect_events_batch(raw, trimmed, prm); T["events"] = eng.last_kernel_ms()[0]
    t1 = time.time()
    reads = np.zeros(n_reads, synth.READ_DT)
    off = 0
    means, times = [], []
    for i, ev in enumerate(events):
        dur = (ev["length"].astype(np.float64) / 4000.0).astype(np.float32)
        t = np.concatenate([[0.0], np.cumsum(dur.astype(np.float64))[:-1]])
        reads[i] = (off, ev.shape[0], 0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
        means.append(ev["mean"]); times.append(t); off += ev.shape[0]
    rs = synth.ReadSet(reads, np.concatenate(means), np.concatenate(times), seqs, [None] * n_reads, [None] * n_reads, nuc.k)
    jobs, ranks, total = synth.abea_jobs(rs)
    t2 = time.time()
    ss = eng.mom_batch(rs.reads, rs.ev_mean, ranks, jobs, mid)
    reads["shift"], reads["scale"] = ss[:, 0], ss[:, 1]
    pairs, res = eng.abea_batch(rs.reads, rs.ev_mean, rs.ev_start_time, ranks, jobs, mid, total); T["abea"] = eng.last_kernel_ms()[0]
    b2e, cal = eng.recalibrate_batch(rs.reads, rs.ev_mean, ranks, jobs, mid, pairs, res); T["recal"] = eng.last_kernel_ms()[0]
    t3 = time.time()
    print(f"{n_reads} reads x {n_samples} samples: kernels " + " ".join(f"{k} {v:.2f} ms" for k, v in T.items()) +
          f" | wall: trim+events {1e3*(t1-t0):.1f} ms, host glue {1e3*(t2-t1):.1f} ms, mom+abea+recal {1e3*(t3-t2):.1f} ms"
          f" | ok {int((cal['status']==0).sum())}/{n_reads}, events {off}")
