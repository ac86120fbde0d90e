// nph_scorereads.hpp — the caller of profile_hmm_score in `nanopolish scorereads` (BASELINE configs[1]):
//
//   model_score                         ref: src/nanopolish_scorereads.cpp:116-203
//
// The reference walks a read's event alignment (the output of align_read_to_ref) in steps of events_per_segment
// entries, scores the reference stretch between the two boundary entries against the events between them with one
// profile_hmm_score call per segment, and returns sum(segment scores) / sum(events).  Here add_read() only enumerates the
// segments into one HmmBatch for a whole batch of reads; run() launches once and folds the per-read ratios.  The
// per-segment recalibration the reference prints on its SEGMENT lines does not feed the score and is not reproduced.
#pragma once
#include "nph_eventalign.hpp"

namespace nph {

struct ReadScore {
    double score = 1.0;          // curr_score / nevents, or +1 when no segment was scored (the reference's convention)
    size_t n_events = 0;
    size_t n_segments = 0;
};

class ScoreReads {
public:
    explicit ScoreReads(size_t events_per_segment = 500) : m_events_per_segment(events_per_segment) {}
    // alignment_output: the read's EventAlignments (EventAligner::alignment(i)); ref_seq: the reference over
    // [ref_offset, ...] as the record's faidx fetch returns it (any case, IUPAC codes allowed).  Returns the read's index.
    size_t add_read(SquiggleRead& sr, size_t strand_idx, const std::vector<EventAlignment>& alignment_output,
                    const std::string& ref_seq, int ref_offset);
    void run(Engine& engine, double indel_bias = hmm_indel_bias_factor);
    const ReadScore& score(size_t read_idx) const { return m_scores[read_idx]; }
    size_t num_reads() const { return m_scores.size(); }
    const HmmBatch& batch() const { return m_batch; }
    void clear();

private:
    struct Segment { size_t read; size_t job; int events; };
    size_t m_events_per_segment;
    HmmBatch m_batch;
    std::vector<Segment> m_segments;
    std::vector<ReadScore> m_scores;
};

} // namespace nph
