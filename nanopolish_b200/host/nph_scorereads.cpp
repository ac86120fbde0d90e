// nph_scorereads.cpp — see nph_scorereads.hpp.
#include "nph_scorereads.hpp"

#include <cstdlib>

namespace nph {

void ScoreReads::clear()
{
    m_batch.clear(); m_segments.clear(); m_scores.clear();
}

size_t ScoreReads::add_read(SquiggleRead& sr, size_t strand_idx, const std::vector<EventAlignment>& alignment_output,
                            const std::string& ref_seq, int ref_offset)
{
    const size_t read_idx = m_scores.size();
    m_scores.emplace_back();
    const PoreModel* pore_model = sr.get_model((uint32_t)strand_idx, "nucleotide");
    if (!pore_model) throw Error(NPH_ERR_INVALID, "read has no nucleotide model");
    const Alphabet* alphabet = pore_model->pmalphabet;
    const int eps = (int)m_events_per_segment;
    for (int align_start_idx = eps; align_start_idx < (int)alignment_output.size() - eps; align_start_idx += eps) {
        const EventAlignment& align_start = alignment_output[align_start_idx];
        const EventAlignment& align_end = alignment_output[align_start_idx + eps];
        HMMInputData data;
        data.read = &sr;
        data.pore_model = pore_model;
        data.strand = (uint8_t)strand_idx;
        data.rc = alignment_output.front().rc;
        data.event_start_idx = (uint32_t)align_start.event_idx;
        data.event_stop_idx = (uint32_t)align_end.event_idx;
        data.event_stride = data.event_start_idx <= data.event_stop_idx ? 1 : -1;
        const int ref_start_pos = align_start.ref_position, ref_end_pos = align_end.ref_position;
        if (ref_end_pos < ref_start_pos) throw Error(NPH_ERR_INVALID, "event alignment runs backwards on the reference");   // the reference asserts
        // get_reference_region_ts(fai, contig, start, end): end inclusive, clipped to what the index holds
        const long long first = (long long)ref_start_pos - ref_offset;
        if (first < 0 || first > (long long)ref_seq.size()) throw Error(NPH_ERR_INVALID, "segment outside the fetched reference");
        std::string seg = ref_seq.substr((size_t)first, (size_t)(ref_end_pos - ref_start_pos + 1));
        if ((int)seg.size() <= (int)sr.get_model_k((uint32_t)strand_idx)) continue;
        seg = alphabet->disambiguate(seg);
        HMMInputSequence sequence(seg, alphabet->reverse_complement(seg), alphabet);
        const int events_in_segment = std::abs((int)data.event_start_idx - (int)data.event_stop_idx) + 1;
        m_segments.push_back(Segment{read_idx, m_batch.add(sequence, data, 0), events_in_segment});
    }
    return read_idx;
}

void ScoreReads::run(Engine& engine, double indel_bias)
{
    const std::vector<float> s = m_batch.run(engine, indel_bias);
    std::vector<double> sum(m_scores.size(), 0.0);
    for (const Segment& g : m_segments) {
        sum[g.read] += s[g.job];                    // double accumulation of the float segment scores, in segment order
        m_scores[g.read].n_events += (size_t)g.events;
        m_scores[g.read].n_segments += 1;
    }
    for (size_t i = 0; i < m_scores.size(); ++i)
        m_scores[i].score = m_scores[i].n_events == 0 ? 1.0 : sum[i] / (double)m_scores[i].n_events;
    m_segments.clear();
    m_batch.clear();
}

} // namespace nph
