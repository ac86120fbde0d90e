"""eventalign's segment chaining and writers, restated — TEST INFRASTRUCTURE ONLY (see oracle_py.py).

  align_read_to_ref            src/alignment/nanopolish_eventalign.cpp:612-827
  get_aligned_segments         src/alignment/nanopolish_anchor.cpp:20-87
  trim_aligned_pairs_* / get_end_pair          src/alignment/nanopolish_eventalign.cpp:166-207
  SquiggleRead::get_closest_event_to           src/nanopolish_squiggle_read.cpp:160-186
  emit_event_alignment_tsv     src/alignment/nanopolish_eventalign.cpp:398-484
  event_alignment_to_cigar     :256-325;  emit_event_alignment_sam :327-396 (as the SAM text htslib prints)
  summarize_alignment          :486-537 and the summary row :600-607

Pinned: tests/test_oracle_vs_ref.py runs the compiled reference's align_read_to_ref + emit_event_alignment_tsv
(oracle/_ref, npref_eventalign) on the same reads and requires identical TSV bytes, event CIGAR and (ref_position,
event_idx, state) triples; tests/golden/eventalign_golden.json holds those reference outputs for boxes without
/root/reference.  The Viterbi inside is PortOracle.hmm_align (np_oracle.c, itself pinned to profile_hmm_align).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BAM_FUNMAP, BAM_FREVERSE = 4, 16
CIGAR_OPS = "MIDNSHP=XB"
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
# IUPAC::getPossibleSymbols(c)[0] (src/common/nanopolish_iupac.cpp): the lexicographically lowest base of each code
_IUPAC_LOWEST = {"A": "A", "C": "C", "G": "G", "T": "T", "M": "A", "R": "A", "W": "A", "S": "C", "Y": "C", "K": "G",
                 "V": "A", "H": "A", "D": "A", "B": "C", "N": "A"}
ALIGN_STRIDE, OUTPUT_STRIDE = 100, 50


def pack_cigar(ops):
    """[(len, 'M'), ...] -> packed BAM CIGAR (uint32: len << 4 | op)."""
    return np.array([(n << 4) | CIGAR_OPS.index(o) for n, o in ops], np.uint32)


def disambiguate(seq: str) -> str:
    return "".join(_IUPAC_LOWEST[c] for c in seq.upper())


def reverse_complement(seq: str) -> str:
    return "".join(_COMP[c] for c in reversed(seq))


def kmer_rank(kmer: str) -> int:
    r = 0
    for c in kmer:
        r = r * 4 + ("ACGT".index(c) if c in "ACGT" else 0)       # Alphabet::rank of a foreign symbol is 0
    return r


def get_aligned_segments(ref_pos, cigar, read_stride=1):
    out = [[]]
    read_pos = 0
    for c in cigar:
        n, op = int(c) >> 4, CIGAR_OPS[int(c) & 0xf]
        read_inc = ref_inc = 0
        aligned = False
        if op in "M=X":
            aligned, read_inc, ref_inc = True, read_stride, 1
        elif op == "D":
            ref_inc = 1
        elif op == "N":
            out.append([]); ref_inc = 1
        elif op == "I":
            read_inc = read_stride
        elif op == "S":
            read_inc = 1
        elif op == "H":
            pass
        else:
            raise ValueError("unhandled cigar operation " + op)
        for _ in range(n):
            if aligned:
                out[-1].append((ref_pos, read_pos))
            read_pos += read_inc
            ref_pos += ref_inc
    return out


def get_end_pair(pairs, ref_pos_max, pair_idx):
    while pair_idx < len(pairs):
        if pairs[pair_idx][0] > ref_pos_max:
            return pair_idx - 1
        pair_idx += 1
    return len(pairs) - 1


@dataclass
class EARead:
    """The SquiggleRead members eventalign touches (strand 0)."""
    name: str
    read_sequence: str
    b2e_start: np.ndarray            # base_to_event_map[k].indices[0].start, -1 where the k-mer has no event
    mean: np.ndarray                 # f32
    stdv: np.ndarray                 # f32
    duration: np.ndarray             # f32
    start_time: np.ndarray           # f64
    shift: float
    scale: float
    drift: float
    var: float
    model: object                    # synth.PoreModel (nucleotide)
    model_name: str = "r9.4_450bps.nucleotide.6mer.template.model"
    k: int = 6

    def get_next_event(self, start, stop, stride):
        while start != stop:
            ei = int(self.b2e_start[start])
            if ei != -1:
                return ei
            start += stride
        return -1

    def get_closest_event_to(self, k_idx):
        stop_before = max(0, k_idx - 1000)
        stop_after = min(k_idx + 1000, len(self.b2e_start) - 1)
        before = self.get_next_event(k_idx, stop_before, -1)
        after = self.get_next_event(k_idx, stop_after, 1)
        return after if before == -1 else before

    def flip_k_strand(self, k_idx):
        return len(self.read_sequence) - k_idx - self.k

    def drift_scaled_level(self, e):
        level = np.float32(self.mean[e])
        time = np.float32(self.start_time[e] - self.start_time[0])
        return np.float32(np.float64(level) - np.float64(time) * self.drift)

    def scaled_gaussian(self, rank):
        m = self.model
        return (np.float32(self.scale * m.level_mean[rank] + self.shift), np.float32(m.level_stdv[rank] * self.var))


@dataclass
class EA:
    ref_name: str
    ref_position: int
    ref_kmer: str
    read_idx: int
    event_idx: int
    rc: bool
    model_kmer: str
    hmm_state: str
    strand_idx: int = 0


def align_read_to_ref(read: EARead, ref_name, ref_seq_fetched, ref_offset, flag, cigar, read_idx, align_fn,
                      region_start=-1, region_end=-1, stats=None):
    """align_fn(fwd_subseq, rc_subseq, event_start, event_stop, stride, rc) -> [(event_idx, kmer_idx, state)] in
    ascending event order == profile_hmm_align(HMMInputSequence(fwd, rc), input)."""
    k = read.k
    out = []
    ref_seq = disambiguate(ref_seq_fetched)
    rc_ref_seq = reverse_complement(ref_seq)
    if flag & BAM_FUNMAP:
        return out
    for aligned_pairs in get_aligned_segments(ref_offset, cigar):
        if region_start != -1 and region_end != -1:
            aligned_pairs = [p for p in aligned_pairs if region_start <= p[0] <= region_end]
        max_kmer_idx = len(read.read_sequence) - k
        idx = len(aligned_pairs) - 1
        while idx >= 0 and aligned_pairs[idx][1] > max_kmer_idx:
            idx -= 1
        aligned_pairs = aligned_pairs[:idx + 1] if idx >= 0 else []
        if not aligned_pairs:
            return out
        do_base_rc = bool(flag & BAM_FREVERSE)
        rc_flag = do_base_rc                                   # rc_flags[strand 0]
        read_kidx_start, read_kidx_end = aligned_pairs[0][1], aligned_pairs[-1][1]
        if do_base_rc:
            read_kidx_start, read_kidx_end = read.flip_k_strand(read_kidx_start), read.flip_k_strand(read_kidx_end)
        first_event = read.get_closest_event_to(read_kidx_start)
        last_event = read.get_closest_event_to(read_kidx_end)
        forward = first_event < last_event
        curr_start_event, curr_start_ref, curr_pair_idx = first_event, aligned_pairs[0][0], 0
        while (forward and curr_start_event < last_event) or (not forward and curr_start_event > last_event):
            end_pair_idx = get_end_pair(aligned_pairs, curr_start_ref + ALIGN_STRIDE, curr_pair_idx)
            curr_end_ref, curr_end_read = aligned_pairs[end_pair_idx]
            if do_base_rc:
                curr_end_read = read.flip_k_strand(curr_end_read)
            s = curr_start_ref - ref_offset
            l = curr_end_ref - curr_start_ref + 1
            fwd_subseq = ref_seq[s:s + l]
            rc_subseq = rc_ref_seq[len(ref_seq) - s - l:len(ref_seq) - s]
            if len(fwd_subseq) < 2 * k:
                break
            event_start = curr_start_event
            event_stop = read.get_closest_event_to(curr_end_read)
            if abs(event_start - event_stop) < 2:
                break
            stride = 1 if event_start < event_stop else -1
            path = align_fn(fwd_subseq, rc_subseq, event_start, event_stop, stride, rc_flag)
            if stats is not None:
                stats["segments"] = stats.get("segments", 0) + 1
            num_output = 0
            last_section = end_pair_idx == len(aligned_pairs) - 1
            last_event_output = last_ref_kmer_output = 0
            for (event_idx, kmer_idx, state) in path:
                if not (num_output < OUTPUT_STRIDE or last_section):
                    break
                if state != "K" and event_idx != curr_start_event:
                    ref_position = curr_start_ref + kmer_idx
                    if state != "B":
                        model_kmer = fwd_subseq[kmer_idx:kmer_idx + k] if not rc_flag else \
                            rc_subseq[len(rc_subseq) - kmer_idx - k:len(rc_subseq) - kmer_idx]
                    else:
                        model_kmer = "N" * k
                    out.append(EA(ref_name, ref_position, ref_seq[ref_position - ref_offset:ref_position - ref_offset + k], read_idx,
                                  event_idx, rc_flag, model_kmer, state))
                    last_event_output, last_ref_kmer_output = event_idx, ref_position
                    num_output += 1
            curr_start_event, curr_start_ref = last_event_output, last_ref_kmer_output
            curr_pair_idx = get_end_pair(aligned_pairs, curr_start_ref, curr_pair_idx)
            if num_output == 0:
                break
    return out


def _f(x):
    return float(x)


def event_sample_idx(read: EARead, e, sample_rate):
    """SquiggleRead::get_event_sample_idx with sample_start_time = 0 (src/nanopolish_squiggle_read.cpp:393-428)"""
    start = np.float64(read.start_time[e])
    dur = np.float64(np.float32(read.duration[e]))
    return int(start * sample_rate), int((start + dur) * sample_rate)


def scaled_samples(read: EARead, e, samples, sample_rate):
    a, b = event_sample_idx(read, e, sample_rate)
    out = []
    for i in range(a, b):
        t = i / sample_rate
        s = np.float64(samples[i]) - read.shift
        s -= (t - 0.0 / sample_rate) * read.drift
        s /= read.scale
        out.append(np.float32(s))
    return out


def tsv(read: EARead, alignment, print_read_names=False, scale_events=False, samples=None, sample_rate=4000.0) -> str:
    """samples: the read's raw samples -> the --signal-index and --samples columns are appended"""
    rows = []
    sqrt_var = np.sqrt(np.float64(read.var))
    for ea in alignment:
        event_mean = np.float32(read.mean[ea.event_idx])
        event_stdv = np.float32(read.stdv[ea.event_idx])
        event_duration = np.float32(read.duration[ea.event_idx])
        rank = kmer_rank(ea.model_kmer)
        model_mean = model_stdv = np.float32(0.0)
        if scale_events:
            event_mean = np.float32((np.float64(read.drift_scaled_level(ea.event_idx)) - read.shift) / read.scale)
            if ea.hmm_state != "B":
                model_mean, model_stdv = np.float32(read.model.level_mean[rank]), np.float32(read.model.level_stdv[rank])
        elif ea.hmm_state != "B":
            model_mean, model_stdv = read.scaled_gaussian(rank)
        with np.errstate(divide="ignore", invalid="ignore"):
            standard_level = np.float32(np.float64(np.float32(event_mean - model_mean)) / (sqrt_var * np.float64(model_stdv)))
        who = read.name if print_read_names else "%d" % ea.read_idx
        row = "%s\t%d\t%s\t%s\t%s\t%d\t%.2f\t%.3f\t%.5f\t%s\t%.2f\t%.2f\t%.2f" % (
            ea.ref_name, ea.ref_position, ea.ref_kmer, who, "tc"[ea.strand_idx], ea.event_idx, _f(event_mean), _f(event_stdv),
            _f(event_duration), ea.model_kmer, _f(model_mean), _f(model_stdv), _f(standard_level))
        if samples is not None:
            a, b = event_sample_idx(read, ea.event_idx, sample_rate)
            row += "\t%d\t%d\t%s" % (a, b, ",".join("%g" % float(v) for v in scaled_samples(read, ea.event_idx, samples, sample_rate)))
        rows.append(row + "\n")
    return "".join(rows)


def event_cigar(alignment) -> str:
    if not alignment:
        return ""
    ops = []
    if alignment[0].event_idx > 0:
        ops.append([alignment[0].event_idx, "S"])
    ops.append([1, "M"])
    prev_r, prev_e = alignment[0].ref_position, alignment[0].event_idx
    for ea in alignment[1:]:
        r_step, e_step = abs(ea.ref_position - prev_r), abs(ea.event_idx - prev_e)
        if r_step == 1 and e_step == 1:
            inc = [1, "M"]
        elif r_step > 1:
            ops.append([r_step - 1, "D"])
            inc = [1, "M"]
        else:
            inc = [1, "I"]
        if ops[-1][1] == inc[1]:
            ops[-1][0] += inc[0]
        else:
            ops.append(inc)
        prev_r, prev_e = ea.ref_position, ea.event_idx
    return "".join("%d%s" % (n, o) for n, o in ops)


def sam(read: EARead, alignment, mapq) -> str:
    if not alignment:
        return ""
    stride = 1 if alignment[0].event_idx < alignment[-1].event_idx else -1
    return "%s.template\t%d\t%s\t%d\t%d\t%s\t*\t0\t0\t*\t*\tES:i:%d\n" % (
        read.name, 16 if alignment[0].rc else 0, alignment[0].ref_name, alignment[0].ref_position + 1, mapq, event_cigar(alignment), stride)


def summarize(read: EARead, alignment) -> dict:
    """summarize_alignment (:486-537): movement counters in size_t arithmetic from npos, duration and z-score sums"""
    num_events = num_steps = num_stays = num_skips = 0
    sum_duration = sum_z = 0.0
    prev = None
    for i, ea in enumerate(alignment):
        num_events += 1
        ref_move = (ea.ref_position - (prev if prev is not None else -1)) % (1 << 64)     # size_t arithmetic from npos
        if ref_move == 0:
            num_stays += 1
        elif i != 0 and ref_move > 1:
            num_skips += 1
        elif i != 0 and ref_move == 1:
            num_steps += 1
        sum_duration += float(np.float32(read.duration[ea.event_idx]))
        if ea.hmm_state == "M":
            mean, stdv = read.scaled_gaussian(kmer_rank(ea.model_kmer))
            sum_z += float(np.float32(np.float32(read.drift_scaled_level(ea.event_idx) - mean) / stdv))     # z_score: float arithmetic
        prev = ea.ref_position
    span = alignment[-1].ref_position - alignment[0].ref_position + 1 if alignment else 0
    return dict(num_events=num_events, num_steps=num_steps, num_stays=num_stays, num_skips=num_skips, reference_span=span,
                sum_duration=sum_duration, sum_z_score=sum_z)


def summary_row(read: EARead, alignment, read_idx, fast5_path="read.fast5") -> str:
    if not alignment:
        return ""
    sm = summarize(read, alignment)
    num_events, num_steps, num_stays, num_skips, sum_duration = (sm["num_events"], sm["num_steps"], sm["num_stays"], sm["num_skips"],
                                                                sm["sum_duration"])
    return "%d\t%s\t%s\t%s\t%s\t%d\t%d\t%d\t%d\t%.2f\t%.3f\t%.3f\t%.3f\t%.3f\n" % (
        read_idx, read.name, fast5_path, read.model_name, "template", num_events, num_steps, num_skips, num_stays,
        sum_duration, read.shift, read.scale, read.drift, read.var)
