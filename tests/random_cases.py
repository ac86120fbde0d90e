"""Random job shapes shared by the GPU parity tests (CUDA == port oracle) and the CPU pin (port oracle == compiled reference)."""
import numpy as np

from nanopolish_b200 import synth

HMM_SHAPES = [
    dict(kmin=1, kmax=40, emin=1, emax=60, n=400),       # tiny, includes K=1 and E=1
    dict(kmin=16, kmax=220, emin=11, emax=400, n=300),   # call-methylation window range
    dict(kmin=250, kmax=330, emin=450, emax=520, n=40),  # scorereads segments
    dict(kmin=600, kmax=1100, emin=20, emax=45, n=30),   # many strips, fewer rows than the chain period
    dict(kmin=900, kmax=1400, emin=700, emax=1200, n=12) # wide and tall
]


def random_hmm_jobs(rs, rng, n_jobs, kmin, kmax, emin, emax, flags_choices, model_id=0):
    rows, ranks, seqs, codes = [], [], [], []
    for _ in range(n_jobs):
        r = int(rng.integers(0, rs.n_reads))
        E = int(rs.reads[r]["n_events"])
        nk_all = rs.seq_codes[r].shape[0] - rs.k + 1
        K = int(rng.integers(kmin, min(kmax, nk_all) + 1))
        k0 = int(rng.integers(0, nk_all - K + 1))
        ne = int(rng.integers(emin, min(emax, E) + 1))
        e0 = int(rng.integers(0, E - ne + 1))
        e1 = e0 + ne - 1
        sub = rs.seq_codes[r][k0:k0 + K + rs.k - 1]
        rc = int(rng.integers(0, 2)) if ne > 1 else 0
        fl = int(rng.choice(flags_choices))
        if rc:
            rcsub = (3 - sub[::-1]).astype(np.uint8)
            ranks.append(synth.dna_rc_kmer_ranks(rcsub, rs.k)); rows.append((r, model_id, e1, e0, 1, fl)); seqs.append(synth._CODE2DNA[rcsub].tobytes()); codes.append(sub)
        else:
            ranks.append(synth.kmer_ranks_from_codes(sub, rs.k, 4)); rows.append((r, model_id, e0, e1, 0, fl)); seqs.append(synth._CODE2DNA[sub].tobytes()); codes.append(sub)
    return synth._finish_jobs(rows, ranks, seqs, codes)
