"""Ad-hoc timing of the event detector (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopolish_b200 import synth
from nanopolish_b200.engine import Engine
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 36000
nuc = synth.load_model("nucleotide")
raw, reads = synth.gen_raw(min(n_reads, 512), n_samples, nuc, seed=5)
if n_reads > 512:      # tile the same signals (generation in numpy is the slow part)
    reps = n_reads // 512
    per = raw.shape[0]
    raw = np.tile(raw, reps)
    reads = np.tile(reads, reps)
    for r in range(reps):
        reads["sample_off"][r * 512:(r + 1) * 512] += r * per
        reads["event_off"][r * 512:(r + 1) * 512] += r * int(reads["event_off"][511] + reads["event_cap"][511])
eng = Engine(0)
prm = synth.event_params(False)
for it in range(3):
    t = time.time(); ev = eng.detect_events_batch(raw, reads, prm); dt = time.time() - t
    ms, nl = eng.last_kernel_ms()
    print(f"detect_events {reads.shape[0]} reads: kernels {ms:.2f} ms ({nl} launches) -> {raw.shape[0]/ms*1e3:.3e} samples/s ({raw.nbytes/ms/1e6:.1f} GB/s in), e2e {dt*1e3:.1f} ms, events {sum(e.shape[0] for e in ev)}")
