// nph_eventalign.cpp — see nph_eventalign.hpp.
#include "nph_eventalign.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace nph {

// ---------------------------------------------------------------------------------------------
// AlignBatch / profile_hmm_align
// ---------------------------------------------------------------------------------------------
size_t AlignBatch::add(const HMMInputSequence& sequence, const HMMInputData& data, uint32_t flags)
{
    if (!data.read || !data.pore_model) throw Error(NPH_ERR_INVALID, "HMMInputData without read or pore_model");
    if (data.read->pore_type != PORETYPE_R9) throw Error(NPH_ERR_UNSUPPORTED, "only R9 reads are supported (load_from_raw always makes R9)");
    const uint32_t k = data.pore_model->k;
    if (data.pore_model->states.size() != sequence.get_num_kmer_ranks(k))
        throw Error(NPH_ERR_INVALID, "sequence alphabet does not match the pore model's state space");
    if (!((data.rc && data.event_stride == -1) || (!data.rc && data.event_stride == 1)))
        throw Error(NPH_ERR_INVALID, "rc and event_stride disagree");                                       // ref asserts (profile_hmm_r9.inl:275)
    if (sequence.length() < k) throw Error(NPH_ERR_INVALID, "sequence shorter than k");
    ReadKey key{data.read, data.strand};
    auto it = m_read_index.find(key);
    uint32_t ridx;
    if (it == m_read_index.end()) {
        ridx = (uint32_t)m_reads.size();
        m_read_index[key] = ridx;
        m_reads.push_back(key);
    } else {
        ridx = it->second;
    }
    const uint32_t n_kmers = (uint32_t)(sequence.length() - k + 1);
    nph_hmm_job j;
    j.rank_off = m_ranks.size();
    j.read = ridx;
    j.model_id = 0;   // resolved against the engine in run()
    j.event_start = data.event_start_idx;
    j.event_stop = data.event_stop_idx;
    j.n_kmers = n_kmers;
    j.stride = data.event_stride;
    j.rc = data.rc;
    j.flags = (uint8_t)flags;
    j.reserved = 0;
    sequence.append_kmer_ranks(k, data.rc != 0, m_ranks);
    m_jobs.push_back(j);
    m_job_models.push_back(data.pore_model);
    return m_jobs.size() - 1;
}

void AlignBatch::clear_jobs()
{
    m_job_models.clear(); m_jobs.clear(); m_ranks.clear();
}

void AlignBatch::clear()
{
    clear_jobs();
    m_read_index.clear(); m_reads.clear(); m_reads_uploaded = 0;
}

std::vector<std::vector<HMMAlignmentState>> AlignBatch::run(Engine& engine, double indel_bias, bool reads_resident)
{
    std::vector<std::vector<HMMAlignmentState>> out(m_jobs.size());
    if (m_jobs.empty()) return out;
    for (size_t j = 0; j < m_jobs.size(); ++j) m_jobs[j].model_id = engine.model_id(m_job_models[j]);
    if (!reads_resident || m_reads_uploaded != m_reads.size()) {
        std::vector<std::pair<const SquiggleRead*, uint8_t>> rl;
        for (auto& r : m_reads) rl.push_back({r.read, r.strand});
        const detail::FlatReads fr = detail::flatten_reads(engine, rl);
        engine.check(nph_reads_load(engine.ctx(), fr.reads.data(), fr.reads.size(), fr.mean, fr.time, fr.n_events), "nph_reads_load");
        m_reads_uploaded = m_reads.size();
    }
    std::vector<uint64_t> off(m_jobs.size() + 1, 0);
    for (size_t j = 0; j < m_jobs.size(); ++j) {
        const nph_hmm_job& jb = m_jobs[j];
        const uint64_t E = (jb.event_stop > jb.event_start ? jb.event_stop - jb.event_start : jb.event_start - jb.event_stop) + 1;
        off[j + 1] = off[j] + E + jb.n_kmers + 2;
    }
    std::vector<nph_align_state> states(off.back());
    std::vector<uint32_t> counts(m_jobs.size());
    engine.check(nph_hmm_align(engine.ctx(), m_ranks.data(), m_ranks.size(), m_jobs.data(), m_jobs.size(), indel_bias,
                               states.data(), off.data(), counts.data(), nullptr), "nph_hmm_align");
    for (size_t j = 0; j < m_jobs.size(); ++j) {
        out[j].resize(counts[j]);
        for (uint32_t i = 0; i < counts[j]; ++i) {
            const nph_align_state& s = states[off[j] + i];
            HMMAlignmentState& as = out[j][i];
            as.event_idx = s.event_idx;
            as.kmer_idx = s.kmer_idx;
            as.l_posterior = -INFINITY;
            as.l_fm = s.l_fm;
            as.log_transition_probability = -INFINITY;
            as.state = s.state;
        }
    }
    return out;
}

std::vector<HMMAlignmentState> profile_hmm_align(const HMMInputSequence& sequence, const HMMInputData& data, const uint32_t flags)
{
    AlignBatch b;
    b.add(sequence, data, flags);
    return b.run(Engine::thread_default())[0];
}

// ---------------------------------------------------------------------------------------------
// aligned pairs from the CIGAR
// ---------------------------------------------------------------------------------------------
std::vector<AlignedSegment> get_aligned_segments(int ref_pos, const std::vector<uint32_t>& cigar, int read_stride)
{
    std::vector<AlignedSegment> out(1);
    int read_pos = 0;
    for (uint32_t c : cigar) {
        const int len = (int)(c >> 4), op = (int)(c & 0xf);
        int read_inc = 0, ref_inc = 0;
        bool is_aligned = false;
        switch (op) {
            case NPH_CIGAR_M: case NPH_CIGAR_EQ: case NPH_CIGAR_X: is_aligned = true; read_inc = read_stride; ref_inc = 1; break;
            case NPH_CIGAR_D: ref_inc = 1; break;
            case NPH_CIGAR_N: out.push_back(AlignedSegment()); ref_inc = 1; break;     // a reference skip starts a new segment
            case NPH_CIGAR_I: read_inc = read_stride; break;
            case NPH_CIGAR_S: read_inc = 1; break;                                     // soft clips ignore read_stride
            case NPH_CIGAR_H: break;
            default: throw Error(NPH_ERR_INVALID, "unhandled CIGAR operation");        // the reference asserts
        }
        for (int j = 0; j < len; ++j) {
            if (is_aligned) out.back().push_back({ref_pos, read_pos});
            read_pos += read_inc;
            ref_pos += ref_inc;
        }
    }
    return out;
}

void trim_aligned_pairs_to_kmer(std::vector<AlignedPair>& aligned_pairs, int max_kmer_idx)
{
    int idx = (int)aligned_pairs.size() - 1;
    while (idx >= 0 && aligned_pairs[idx].read_pos > max_kmer_idx) idx -= 1;
    if (idx < 0) aligned_pairs.clear();
    else aligned_pairs.resize(idx + 1);
}

void trim_aligned_pairs_to_ref_region(std::vector<AlignedPair>& aligned_pairs, int ref_start, int ref_end)
{
    std::vector<AlignedPair> trimmed;
    for (const AlignedPair& p : aligned_pairs)
        if (p.ref_pos >= ref_start && p.ref_pos <= ref_end) trimmed.push_back(p);
    aligned_pairs.swap(trimmed);
}

// index of the pair with the highest ref_pos not above ref_pos_max, searching from pair_idx
int get_end_pair(const std::vector<AlignedPair>& aligned_pairs, int ref_pos_max, int pair_idx)
{
    while (pair_idx < (int)aligned_pairs.size()) {
        if (aligned_pairs[pair_idx].ref_pos > ref_pos_max) return pair_idx - 1;
        pair_idx += 1;
    }
    return (int)aligned_pairs.size() - 1;
}

// ---------------------------------------------------------------------------------------------
// EventAligner
// ---------------------------------------------------------------------------------------------
static const int ALIGN_STRIDE = 100;     // approximately how many reference bases to align to at once (eventalign.cpp:666)
static const int OUTPUT_STRIDE = 50;     // approximately how many event alignments to output at once (:667)

void EventAligner::clear()
{
    m_reads.clear(); m_round.clear(); m_batch.clear();
}

size_t EventAligner::add_read(const EventAlignmentParameters& params)
{
    if (!params.sr) throw Error(NPH_ERR_INVALID, "EventAlignmentParameters without a read");
    if (params.strand_idx >= 2) throw Error(NPH_ERR_INVALID, "strand_idx out of range");
    if (!((params.region_start == -1 && params.region_end == -1) || params.region_start <= params.region_end))
        throw Error(NPH_ERR_INVALID, "region_start > region_end");
    m_reads.emplace_back();
    ReadState& rs = m_reads.back();
    rs.params = params;
    rs.pore_model = params.get_model();
    if (!rs.pore_model) throw Error(NPH_ERR_INVALID, "the read has no model for this strand / alphabet");
    rs.k = rs.pore_model->k;
    // upper case, ambiguity codes to their lexicographically lowest base (Alphabet::disambiguate upper-cases)
    rs.ref_seq = rs.pore_model->pmalphabet->disambiguate(params.ref_seq);
    rs.rc_ref_seq = rs.pore_model->pmalphabet->reverse_complement(rs.ref_seq);
    if ((params.flag & NPH_BAM_FUNMAP) != 0) { rs.done = true; return m_reads.size() - 1; }
    rs.segments = get_aligned_segments(params.ref_pos, params.cigar);
    rs.do_base_rc = (params.flag & NPH_BAM_FREVERSE) != 0;
    return m_reads.size() - 1;
}

// The head of the reference's per-segment loop body (eventalign.cpp:654-689): trims, then the events nearest to the
// segment's first and last aligned k-mer.  Depends only on the record, never on earlier paths.
bool EventAligner::setup_segment(ReadState& rs, size_t segment_idx, SegmentStart& out)
{
    AlignedSegment& aligned_pairs = rs.segments[segment_idx];
    const EventAlignmentParameters& p = rs.params;
    if (p.region_start != -1 && p.region_end != -1) trim_aligned_pairs_to_ref_region(aligned_pairs, p.region_start, p.region_end);
    const int max_kmer_idx = (int)p.sr->read_sequence.size() - (int)rs.k;
    trim_aligned_pairs_to_kmer(aligned_pairs, max_kmer_idx);
    if (aligned_pairs.empty()) return false;                 // the reference returns from the whole function here
    int read_kidx_start = aligned_pairs.front().read_pos;
    int read_kidx_end = aligned_pairs.back().read_pos;
    if (rs.do_base_rc) {
        read_kidx_start = p.sr->flip_k_strand(read_kidx_start, rs.k);
        read_kidx_end = p.sr->flip_k_strand(read_kidx_end, rs.k);
    }
    const int n_map = (int)p.sr->base_to_event_map.size();
    if (read_kidx_start < 0 || read_kidx_end < 0 || read_kidx_start >= n_map || read_kidx_end >= n_map)
        throw Error(NPH_ERR_INVALID, "aligned read position outside the base-to-event map");      // the reference asserts / reads out of bounds
    out.first_event = p.sr->get_closest_event_to(read_kidx_start, (uint32_t)p.strand_idx);
    out.last_event = p.sr->get_closest_event_to(read_kidx_end, (uint32_t)p.strand_idx);
    out.start_ref = aligned_pairs.front().ref_pos;
    return true;
}

bool EventAligner::enter_segment(ReadState& rs)
{
    if (rs.segment_idx >= rs.segments.size()) return false;
    SegmentStart st;
    if (!setup_segment(rs, rs.segment_idx, st)) return false;
    rs.last_event = st.last_event;
    rs.forward = st.first_event < st.last_event;
    rs.curr_start_event = st.first_event;
    rs.curr_start_ref = st.start_ref;
    rs.curr_pair_idx = 0;
    rs.in_segment = true;
    return true;
}

// The part of one iteration of the reference's while loop that comes before profile_hmm_align (eventalign.cpp:691-743).
bool EventAligner::prepare(ReadState& rs, AlignBatch& batch)
{
    while (!rs.done) {
        if (!rs.in_segment) {
            if (!enter_segment(rs)) { rs.done = true; break; }
        }
        const EventAlignmentParameters& p = rs.params;
        const AlignedSegment& aligned_pairs = rs.segments[rs.segment_idx];
        bool issued = false;
        if ((rs.forward && rs.curr_start_event < rs.last_event) || (!rs.forward && rs.curr_start_event > rs.last_event)) {
            const int end_pair_idx = get_end_pair(aligned_pairs, rs.curr_start_ref + ALIGN_STRIDE, rs.curr_pair_idx);
            if (end_pair_idx >= 0) {                              // (-1: the reference would index aligned_pairs[-1])
                const int curr_end_ref = aligned_pairs[end_pair_idx].ref_pos;
                int curr_end_read = aligned_pairs[end_pair_idx].read_pos;
                if (rs.do_base_rc) curr_end_read = p.sr->flip_k_strand(curr_end_read, rs.k);
                const int s = rs.curr_start_ref - p.ref_pos;
                const int l = curr_end_ref - rs.curr_start_ref + 1;
                const int n = (int)rs.ref_seq.length();
                const int n_map = (int)p.sr->base_to_event_map.size();
                if (curr_end_read >= 0 && curr_end_read < n_map && s >= 0 && l >= 0 && s + l <= n) {   // (outside: substr throws / the map is read out of bounds in the reference)
                    rs.fwd_subseq = rs.ref_seq.substr(s, l);
                    rs.rc_subseq = rs.rc_ref_seq.substr(n - s - l, l);
                    // require a minimum amount of sequence to align to
                    if (rs.fwd_subseq.length() >= 2 * rs.k) {
                        const int event_stop = p.sr->get_closest_event_to(curr_end_read, (uint32_t)p.strand_idx);
                        // segments with very few alignable events (large deletions) end the chain
                        if (event_stop >= 0 && rs.curr_start_event >= 0 && std::abs(rs.curr_start_event - event_stop) >= 2) {
                            HMMInputData input;
                            input.read = p.sr;
                            input.pore_model = rs.pore_model;
                            input.event_start_idx = (uint32_t)rs.curr_start_event;
                            input.event_stop_idx = (uint32_t)event_stop;
                            input.strand = (uint8_t)p.strand_idx;
                            input.event_stride = input.event_start_idx < input.event_stop_idx ? 1 : -1;
                            input.rc = p.strand_idx == 0 ? rs.do_base_rc : !rs.do_base_rc;      // rc_flags[strand]
                            HMMInputSequence hmm_sequence(rs.fwd_subseq, rs.rc_subseq, rs.pore_model->pmalphabet);
                            batch.add(hmm_sequence, input, 0);
                            rs.end_pair_idx = end_pair_idx;
                            rs.job_rc = input.rc;
                            rs.pending = true;
                            issued = true;
                        }
                    }
                }
            }
        }
        if (issued) return true;
        // the while loop of this BAM segment is over (condition false or a break): next segment
        rs.in_segment = false;
        rs.segment_idx += 1;
    }
    return false;
}

bool EventAligner::next_round(AlignBatch& batch)
{
    batch.clear_jobs();
    m_round.clear();
    for (size_t i = 0; i < m_reads.size(); ++i) {
        if (m_reads[i].done) continue;
        if (prepare(m_reads[i], batch)) m_round.push_back(i);
    }
    return !m_round.empty();
}

// The part after profile_hmm_align (eventalign.cpp:745-823).
void EventAligner::consume(const std::vector<std::vector<HMMAlignmentState>>& paths)
{
    if (paths.size() != m_round.size()) throw Error(NPH_ERR_STATE, "consume(): one path per job of the round");
    for (size_t j = 0; j < m_round.size(); ++j) {
        ReadState& rs = m_reads[m_round[j]];
        const std::vector<HMMAlignmentState>& event_alignment = paths[j];
        const AlignedSegment& aligned_pairs = rs.segments[rs.segment_idx];
        rs.pending = false;
        rs.segments_aligned += 1;

        size_t num_output = 0;
        // if we aligned to the last pair, output everything and stop
        const bool last_section = rs.end_pair_idx == (int)aligned_pairs.size() - 1;
        int last_event_output = 0;
        int last_ref_kmer_output = 0;
        for (size_t idx = 0; idx < event_alignment.size() && (num_output < (size_t)OUTPUT_STRIDE || last_section); idx++) {
            const HMMAlignmentState& as = event_alignment[idx];
            if (as.state != 'K' && (int)as.event_idx != rs.curr_start_event) {
                rs.output.push_back(Rec{rs.curr_start_ref + (int)as.kmer_idx, (int)as.event_idx, as.state});
                last_event_output = (int)as.event_idx;
                last_ref_kmer_output = rs.curr_start_ref + (int)as.kmer_idx;
                num_output += 1;
            }
        }
        // advance the cursor to where the output stopped
        rs.curr_start_event = last_event_output;
        rs.curr_start_ref = last_ref_kmer_output;
        rs.curr_pair_idx = get_end_pair(aligned_pairs, rs.curr_start_ref, rs.curr_pair_idx);
        if (num_output == 0) {        // break: on to the next BAM segment
            rs.in_segment = false;
            rs.segment_idx += 1;
        }
    }
    m_round.clear();
}

size_t EventAligner::run_rounds(Engine& engine, double indel_bias)
{
    size_t rounds = 0;
    m_batch.clear();
    while (next_round(m_batch)) {
        // AlignBatch uploads the read table again whenever a read shows up that was not resident
        consume(m_batch.run(engine, indel_bias, rounds > 0));
        rounds += 1;
    }
    m_batch.clear();
    return rounds;
}

// ref_seq's k-mer at offset pos and what HMMInputSequence::get_kmer hands the model for it: the k-mer itself, or for rc
// reads rc_subseq.substr(l - kmer_idx - k, k) == rc_ref_seq.substr(n - pos - k, k)
void EventAligner::kmers_at(const ReadState& rs, const Rec& r, char* ref_kmer, char* model_kmer)
{
    const size_t k = rs.k, n = rs.ref_seq.size();
    const size_t pos = (size_t)(r.ref_position - rs.params.ref_pos);
    size_t len = pos <= n ? std::min(k, n - pos) : 0;          // std::string::substr clips at the end
    std::memcpy(ref_kmer, rs.ref_seq.data() + (len ? pos : 0), len);
    ref_kmer[len] = 0;
    if (r.state == 'B') {
        std::memset(model_kmer, 'N', k);
    } else {
        const bool rc = rs.params.strand_idx == 0 ? rs.do_base_rc : !rs.do_base_rc;
        std::memcpy(model_kmer, rc ? rs.rc_ref_seq.data() + (n - pos - k) : rs.ref_seq.data() + pos, k);   // inside the window: always k long
    }
    model_kmer[k] = 0;
}

EventAlignment EventAligner::materialize(const ReadState& rs, const Rec& r) const
{
    char ref_kmer[64], model_kmer[64];
    kmers_at(rs, r, ref_kmer, model_kmer);
    EventAlignment ea;
    ea.ref_name = rs.params.ref_name;
    ea.ref_position = r.ref_position;
    ea.ref_kmer = ref_kmer;
    ea.read_idx = (size_t)rs.params.read_idx;
    ea.strand_idx = (int)rs.params.strand_idx;
    ea.event_idx = r.event_idx;
    ea.rc = rs.params.strand_idx == 0 ? rs.do_base_rc : !rs.do_base_rc;
    ea.model_kmer = model_kmer;
    ea.hmm_state = r.state;
    return ea;
}

std::vector<EventAlignment> EventAligner::alignment(size_t read_idx) const
{
    const ReadState& rs = m_reads[read_idx];
    std::vector<EventAlignment> out;
    out.reserve(rs.output.size());
    for (const Rec& r : rs.output) out.push_back(materialize(rs, r));
    return out;
}

// ranks of every k-mer of seq by one rolling pass: rank(pos + 1) = (rank(pos) mod A^(k-1)) * A + rank(seq[pos + k])
void rolling_kmer_ranks(const Alphabet* alphabet, const std::string& seq, uint32_t k, uint32_t* out)
{
    const size_t n = seq.size();
    if (n < k) return;
    uint64_t top = 1;
    for (uint32_t i = 1; i < k; ++i) top *= alphabet->size();
    uint64_t r = 0;
    for (uint32_t i = 0; i < k; ++i) r = r * alphabet->size() + alphabet->rank(seq[i]);
    out[0] = (uint32_t)r;
    for (size_t pos = 1; pos + k <= n; ++pos) {
        r = (r % top) * alphabet->size() + alphabet->rank(seq[pos + k - 1]);
        out[pos] = (uint32_t)r;
    }
}

size_t EventAligner::run(Engine& engine, double indel_bias)
{
    const size_t nr = m_reads.size();
    // ---- phase 1 (parallel over reads): trims and start/stop events of every BAM segment, up to the first segment
    //      that trims to nothing (where the reference returns from align_read_to_ref) ----
    std::vector<std::vector<SegmentStart>> starts(nr);
    std::vector<std::string> errors(nr);
#pragma omp parallel for schedule(dynamic, 8) num_threads(host_threads())
    for (long long i = 0; i < (long long)nr; ++i) {
        ReadState& rs = m_reads[i];
        if (rs.done) continue;
        try {
            if (rs.k >= 64) throw Error(NPH_ERR_UNSUPPORTED, "k-mer length");
            for (size_t sidx = 0; sidx < rs.segments.size(); ++sidx) {
                SegmentStart st;
                if (!setup_segment(rs, sidx, st)) break;
                starts[i].push_back(st);
            }
        } catch (const std::exception& e) { errors[i] = e.what(); }
    }
    for (size_t i = 0; i < nr; ++i) if (!errors[i].empty()) throw Error(NPH_ERR_INVALID, errors[i]);

    // ---- phase 2 (serial, O(reads)): read table, offsets ----
    struct Slot { uint32_t read_index; uint64_t map_off; };
    std::vector<std::pair<const SquiggleRead*, uint8_t>> read_table;
    std::map<std::pair<const SquiggleRead*, uint8_t>, Slot> read_slot;
    std::vector<Slot> slot_of(nr);
    std::vector<char> fills_map(nr, 0);
    std::vector<uint64_t> rank_off(nr, 0), chain_first(nr + 1, 0);
    std::vector<uint32_t> model_of(nr, 0);
    uint64_t n_map = 0, n_ranks = 0, n_pairs = 0, records_total = 0;
    for (size_t i = 0; i < nr; ++i) {
        chain_first[i + 1] = chain_first[i] + starts[i].size();
        if (starts[i].empty()) continue;
        const ReadState& rs = m_reads[i];
        const EventAlignmentParameters& p = rs.params;
        const auto key = std::make_pair((const SquiggleRead*)p.sr, (uint8_t)p.strand_idx);
        auto it = read_slot.find(key);
        if (it == read_slot.end()) {
            it = read_slot.insert({key, Slot{(uint32_t)read_table.size(), n_map}}).first;
            read_table.push_back(key);
            n_map += p.sr->base_to_event_map.size();
            fills_map[i] = 1;
        }
        slot_of[i] = it->second;
        rank_off[i] = n_ranks;
        n_ranks += rs.ref_seq.size() >= rs.k ? rs.ref_seq.size() - rs.k + 1 : 0;
        model_of[i] = engine.model_id(rs.pore_model);
    }
    const size_t n_chains = (size_t)chain_first[nr];
    for (ReadState& rs : m_reads) rs.done = true;             // nothing left for the round driver unless re-armed below
    if (n_chains == 0) return 0;
    std::vector<nph_ea_chain> chains(n_chains);
    std::vector<size_t> owner(n_chains);
    for (size_t i = 0; i < nr; ++i) {
        const ReadState& rs = m_reads[i];
        for (size_t sidx = 0; sidx < starts[i].size(); ++sidx) {
            const SegmentStart& st = starts[i][sidx];
            nph_ea_chain& c = chains[chain_first[i] + sidx];
            std::memset(&c, 0, sizeof(c));
            c.pair_off = n_pairs;
            c.n_pairs = (uint32_t)rs.segments[sidx].size();
            n_pairs += c.n_pairs;
            c.out_cap = (uint32_t)std::abs(st.last_event - st.first_event) + 2;
            c.out_off = records_total;
            records_total += c.out_cap;
            owner[chain_first[i] + sidx] = i;
        }
    }

    // ---- phase 3 (parallel over reads): pairs, event maps, rank tables, chain records ----
    std::vector<nph_aligned_pair> pairs(std::max<uint64_t>(n_pairs, 1));
    std::vector<int32_t> map_start(std::max<uint64_t>(n_map, 1));
    std::vector<uint32_t> ranks_fwd(std::max<uint64_t>(n_ranks, 1)), ranks_rc(std::max<uint64_t>(n_ranks, 1));
#pragma omp parallel for schedule(dynamic, 8) num_threads(host_threads())
    for (long long i = 0; i < (long long)nr; ++i) {
        if (starts[i].empty()) continue;
        const ReadState& rs = m_reads[i];
        const EventAlignmentParameters& p = rs.params;
        if (fills_map[i]) {
            int32_t* m = map_start.data() + slot_of[i].map_off;
            const std::vector<EventRangeForBase>& b2e = p.sr->base_to_event_map;
            for (size_t j = 0; j < b2e.size(); ++j) m[j] = b2e[j].indices[p.strand_idx].start;
        }
        const size_t n = rs.ref_seq.size();
        if (n >= rs.k) {
            const size_t nk = n - rs.k + 1;
            rolling_kmer_ranks(rs.pore_model->pmalphabet, rs.ref_seq, rs.k, ranks_fwd.data() + rank_off[i]);
            // entry pos of the rc table = rank of rc_ref_seq's k-mer at n - pos - k (what get_kmer_rank(ki, k, true) resolves to)
            std::vector<uint32_t> tmp(nk);
            rolling_kmer_ranks(rs.pore_model->pmalphabet, rs.rc_ref_seq, rs.k, tmp.data());
            uint32_t* rc = ranks_rc.data() + rank_off[i];
            for (size_t pos = 0; pos < nk; ++pos) rc[pos] = tmp[nk - 1 - pos];
        }
        for (size_t sidx = 0; sidx < starts[i].size(); ++sidx) {
            const SegmentStart& st = starts[i][sidx];
            nph_ea_chain& c = chains[chain_first[i] + sidx];
            nph_aligned_pair* dst = pairs.data() + c.pair_off;
            const AlignedSegment& seg = rs.segments[sidx];
            for (size_t j = 0; j < seg.size(); ++j) dst[j] = nph_aligned_pair{seg[j].ref_pos, seg[j].read_pos};
            c.map_off = slot_of[i].map_off;
            c.map_len = (uint32_t)p.sr->base_to_event_map.size();
            c.rank_off = rank_off[i];
            c.ref_len = (uint32_t)n;
            c.read = slot_of[i].read_index;
            c.model_id = model_of[i];
            c.read_seq_len = (uint32_t)p.sr->read_sequence.size();
            c.ref_offset = p.ref_pos;
            c.first_event = st.first_event;
            c.last_event = st.last_event;
            c.do_base_rc = rs.do_base_rc;
            c.rc = p.strand_idx == 0 ? rs.do_base_rc : !rs.do_base_rc;
            c.k = (uint8_t)rs.k;
        }
    }

    // ---- the device: reads up, one launch, records back ----
    const detail::FlatReads fr = detail::flatten_reads(engine, read_table);
    engine.check(nph_reads_load(engine.ctx(), fr.reads.data(), fr.reads.size(), fr.mean, fr.time, fr.n_events), "nph_reads_load");
    nph_ea_record* const records = static_cast<nph_ea_record*>(engine.pinned(0, sizeof(nph_ea_record) * std::max<uint64_t>(records_total, 1)));
    std::vector<nph_ea_result> results(n_chains);
    engine.check(nph_eventalign_chain(engine.ctx(), pairs.data(), pairs.size(), map_start.data(), map_start.size(), ranks_fwd.data(),
                                      ranks_rc.data(), ranks_fwd.size(), chains.data(), n_chains, indel_bias, records, records_total,
                                      results.data()),
                 "nph_eventalign_chain");

    // ---- scatter (parallel over reads); a read with a window the chain kernel could not hold goes through the round
    //      driver instead ----
    std::vector<char> redo(nr, 0);
    for (size_t c = 0; c < n_chains; ++c) {
        const int st = results[c].status;
        if (st & NPH_EA_RC_STRIDE) throw Error(NPH_ERR_INVALID, "rc and event_stride disagree");     // ref asserts (profile_hmm_r9.inl:275)
        if (st & NPH_EA_BAD_EVENT) throw Error(NPH_ERR_INVALID, "event index outside the read");
        if (st & NPH_EA_OUT_OVERFLOW) throw Error(NPH_ERR_STATE, "eventalign chain: record room exceeded");
        if (st & NPH_EA_WINDOW_TOO_LARGE) redo[owner[c]] = 1;
    }
#pragma omp parallel for schedule(dynamic, 8) num_threads(host_threads())
    for (long long i = 0; i < (long long)nr; ++i) {
        if (redo[i] || starts[i].empty()) continue;
        ReadState& rs = m_reads[i];
        size_t total = 0;
        for (size_t c = chain_first[i]; c < chain_first[i + 1]; ++c) total += results[c].n_records;
        rs.output.reserve(rs.output.size() + total);
        for (size_t c = chain_first[i]; c < chain_first[i + 1]; ++c) {
            const nph_ea_record* r = records + chains[c].out_off;
            for (uint32_t j = 0; j < results[c].n_records; ++j) rs.output.push_back(Rec{r[j].ref_position, r[j].event_idx, (char)r[j].hmm_state});
            rs.segments_aligned += results[c].n_windows;
        }
    }
    size_t n_redo = 0;
    for (size_t i = 0; i < nr; ++i) {
        if (!redo[i]) continue;
        ReadState& rs = m_reads[i];
        rs.done = false; rs.in_segment = false; rs.segment_idx = 0; rs.pending = false;
        rs.output.clear(); rs.segments_aligned = 0;
        ++n_redo;
    }
    return 1 + (n_redo ? run_rounds(engine, indel_bias) : 0);
}

// ---------------------------------------------------------------------------------------------
// writers
// ---------------------------------------------------------------------------------------------
static inline char* put_int(char* o, long long v)
{
    char tmp[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *o++ = '-';
    while (n) *o++ = tmp[--n];
    return o;
}
static inline char* put_str(char* o, const char* s, size_t n) { std::memcpy(o, s, n); return o + n; }

std::string EventAligner::tsv_header(const EventalignOptions& opt)
{
    std::string h = "contig\tposition\treference_kmer\t";
    h += opt.print_read_names ? "read_name" : "read_index";
    h += "\tstrand\tevent_index\tevent_level_mean\tevent_stdv\tevent_length\tmodel_kmer\tmodel_mean\tmodel_stdv\tstandardized_level";
    if (opt.write_signal_index) h += "\tstart_idx\tend_idx";
    if (opt.write_samples) h += "\tsamples";
    h += "\n";
    return h;
}

std::string EventAligner::tsv(size_t read_idx, const EventalignOptions& opt) const
{
    const ReadState& rs = m_reads[read_idx];
    const SquiggleRead& sr = *rs.params.sr;
    const PoreModel* pore_model = rs.pore_model;
    const uint32_t k = pore_model->k;
    const uint32_t strand = (uint32_t)rs.params.strand_idx;
    const std::string& ref_name = rs.params.ref_name;
    // the read column: read_idx as %zu, or the read name with -n
    const std::string who_s = opt.print_read_names ? sr.read_name : std::to_string((size_t)rs.params.read_idx);
    const double sqrt_var = std::sqrt(sr.scalings[strand].var);
    if ((opt.write_signal_index || opt.write_samples) && sr.samples.empty())
        throw Error(NPH_ERR_STATE, "--signal-index / --samples need the raw samples on the read (load_from_raw with SRF_LOAD_RAW_SAMPLES)");
    // room per row: six numbers of at most 47 characters; with the sample columns, 16 characters per raw sample of the
    // events written (%g, six significant digits) + two indices
    size_t extra = 0;
    if (opt.write_signal_index) extra += 48 * rs.output.size();
    if (opt.write_samples)
        for (const Rec& r : rs.output) {
            const std::pair<size_t, size_t> si = sr.get_event_sample_idx(strand, r.event_idx);
            extra += 2 + 16 * (si.second > si.first ? si.second - si.first : 0);
        }
    std::string out;
    out.resize(rs.output.size() * (ref_name.size() + who_s.size() + 2 * (size_t)k + 340) + extra);
    char* const base = &out[0];
    char* o = base;
    char ref_kmer[64], model_kmer[64];
    for (const Rec& r : rs.output) {
        kmers_at(rs, r, ref_kmer, model_kmer);
        // contig, position, reference_kmer, read, strand
        o = put_str(o, ref_name.data(), ref_name.size()); *o++ = '\t';
        o = put_int(o, r.ref_position); *o++ = '\t';
        o = put_str(o, ref_kmer, std::strlen(ref_kmer)); *o++ = '\t';
        o = put_str(o, who_s.data(), who_s.size()); *o++ = '\t';
        *o++ = "tc"[strand]; *o++ = '\t';

        float event_mean = sr.get_unscaled_level(r.event_idx, strand);
        const float event_stdv = sr.get_stdv(r.event_idx, strand);
        const float event_duration = sr.get_duration(r.event_idx, strand);
        float model_mean = 0.0, model_stdv = 0.0;
        if (opt.scale_events) {
            // scale reads to the model; unscaled model parameters
            event_mean = sr.get_fully_scaled_level(r.event_idx, strand);
            if (r.state != 'B') {
                const PoreModelStateParams model = pore_model->get_parameters(pore_model->pmalphabet->kmer_rank(model_kmer, k));
                model_mean = (float)model.level_mean;
                model_stdv = (float)model.level_stdv;
            }
        } else if (r.state != 'B') {
            // scale model to the reads
            const GaussianParameters model = sr.get_scaled_gaussian_from_pore_model_state(*pore_model, strand, pore_model->pmalphabet->kmer_rank(model_kmer, k));
            model_mean = model.mean;
            model_stdv = model.stdv;
        }
        // float difference over a double product, narrowed to float (a 'B' state divides by zero: inf, like the reference)
        const float standard_level = (float)((event_mean - model_mean) / (sqrt_var * model_stdv));
        // event_index %d, event_level_mean %.2lf, event_stdv %.3lf, event_length %.5lf
        o = put_int(o, r.event_idx); *o++ = '\t';
        o += format_fixed(o, event_mean, 2); *o++ = '\t';
        o += format_fixed(o, event_stdv, 3); *o++ = '\t';
        o += format_fixed(o, event_duration, 5); *o++ = '\t';
        // model_kmer, model_mean %.2lf, model_stdv %.2lf, standardized_level %.2lf
        o = put_str(o, model_kmer, k); *o++ = '\t';
        o += format_fixed(o, model_mean, 2); *o++ = '\t';
        o += format_fixed(o, model_stdv, 2); *o++ = '\t';
        o += format_fixed(o, standard_level, 2);
        if (opt.write_signal_index) {
            const std::pair<size_t, size_t> si = sr.get_event_sample_idx(strand, r.event_idx);
            *o++ = '\t'; o = put_int(o, (long long)si.first); *o++ = '\t'; o = put_int(o, (long long)si.second);
        }
        if (opt.write_samples) {
            // the reference streams the floats through an ostream (%g, 6 significant digits) with ',' after each and
            // drops the last comma (an event without samples makes it resize() to npos and throw; here: an empty column)
            const std::vector<float> samples = sr.get_scaled_samples_for_event(strand, r.event_idx);
            *o++ = '\t';
            for (size_t i = 0; i < samples.size(); ++i) {
                if (i) *o++ = ',';
                o += snprintf(o, 32, "%g", (double)samples[i]);
            }
        }
        *o++ = '\n';
    }
    out.resize((size_t)(o - base));
    return out;
}

std::vector<std::string> EventAligner::tsv_batch(const EventalignOptions& opt) const
{
    std::vector<std::string> out(m_reads.size());
    // rows of different reads are independent; the reference formats them one read at a time inside an omp critical
#pragma omp parallel for schedule(dynamic, 4) num_threads(host_threads())
    for (long long i = 0; i < (long long)m_reads.size(); ++i) out[i] = tsv((size_t)i, opt);
    return out;
}

std::vector<uint32_t> event_alignment_to_cigar(const std::vector<EventAlignment>& alignments)
{
    std::vector<uint32_t> out;
    if (alignments.empty()) return out;
    // a soft clip accounts for unaligned events at the beginning of the read
    if (alignments[0].event_idx > 0) out.push_back((uint32_t)alignments[0].event_idx << 4 | NPH_CIGAR_S);
    out.push_back(1u << 4 | NPH_CIGAR_M);       // always starts with a match
    int prev_r_idx = alignments[0].ref_position;
    int prev_e_idx = alignments[0].event_idx;
    for (size_t ai = 1; ai < alignments.size(); ++ai) {
        const int r_idx = alignments[ai].ref_position, e_idx = alignments[ai].event_idx;
        const int r_step = std::abs(r_idx - prev_r_idx), e_step = std::abs(e_idx - prev_e_idx);
        uint32_t incoming;
        if (r_step == 1 && e_step == 1) {
            incoming = 1u << 4 | NPH_CIGAR_M;
        } else if (r_step > 1) {
            // a reference jump of more than one is a deletion followed by a new match (the reference asserts e_step == 1)
            out.push_back((uint32_t)(r_step - 1) << 4 | NPH_CIGAR_D);
            incoming = 1u << 4 | NPH_CIGAR_M;
        } else {
            incoming = 1u << 4 | NPH_CIGAR_I;   // (the reference asserts e_step == 1 && r_step == 0)
        }
        if ((out.back() & 0xf) == (incoming & 0xf)) out.back() = ((out.back() >> 4) + (incoming >> 4)) << 4 | (incoming & 0xf);
        else out.push_back(incoming);
        prev_r_idx = r_idx;
        prev_e_idx = e_idx;
    }
    return out;
}

std::string cigar_ops_to_string(const std::vector<uint32_t>& ops)
{
    std::string s;
    for (uint32_t c : ops) { s += std::to_string(c >> 4); s += "MIDNSHP=XB"[c & 0xf]; }
    return s;
}

std::string EventAligner::event_cigar(size_t read_idx) const
{
    return cigar_ops_to_string(event_alignment_to_cigar(alignment(read_idx)));
}

std::string EventAligner::sam(size_t read_idx) const
{
    const ReadState& rs = m_reads[read_idx];
    const std::vector<Rec>& al = rs.output;
    if (al.empty()) return std::string();
    const bool rc = rs.params.strand_idx == 0 ? rs.do_base_rc : !rs.do_base_rc;
    const std::string qname = rs.params.sr->read_name + (rs.params.strand_idx == 0 ? ".template" : ".complement");
    const int stride = al.front().event_idx < al.back().event_idx ? 1 : -1;
    // QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL + the event-stride tag
    return qname + "\t" + std::to_string(rc ? 16 : 0) + "\t" + rs.params.ref_name + "\t" + std::to_string(al.front().ref_position + 1) + "\t" +
           std::to_string((int)rs.params.mapq) + "\t" + event_cigar(read_idx) + "\t*\t0\t0\t*\t*\tES:i:" + std::to_string(stride) + "\n";
}

EventalignSummary EventAligner::summarize(size_t read_idx) const
{
    const ReadState& rs = m_reads[read_idx];
    const SquiggleRead& sr = *rs.params.sr;
    EventalignSummary summary;
    uint64_t prev_ref_pos = ~(uint64_t)0;                  // std::string::npos
    const uint32_t strand = (uint32_t)rs.params.strand_idx;
    char ref_kmer[64], model_kmer[64];
    for (size_t i = 0; i < rs.output.size(); ++i) {
        const Rec& ea = rs.output[i];
        summary.num_events += 1;
        const uint64_t ref_move = (uint64_t)(int64_t)ea.ref_position - prev_ref_pos;    // size_t arithmetic, as the reference
        if (ref_move == 0) summary.num_stays += 1;
        else if (i != 0 && ref_move > 1) summary.num_skips += 1;
        else if (i != 0 && ref_move == 1) summary.num_steps += 1;
        summary.sum_duration += sr.get_duration(ea.event_idx, strand);
        if (ea.state == 'M') {
            kmers_at(rs, ea, ref_kmer, model_kmer);
            const uint32_t rank = rs.pore_model->pmalphabet->kmer_rank(model_kmer, rs.k);
            // z_score (src/hmm/nanopolish_emissions.h:32-41): float arithmetic
            const float level = sr.get_drift_scaled_level(ea.event_idx, strand);
            const GaussianParameters gp = sr.get_scaled_gaussian_from_pore_model_state(*rs.pore_model, strand, rank);
            const float z = (level - gp.mean) / gp.stdv;
            summary.sum_z_score += z;
        }
        prev_ref_pos = (uint64_t)(int64_t)ea.ref_position;
    }
    summary.alignment_edit_distance = rs.params.edit_distance;
    if (!rs.output.empty()) summary.reference_span = rs.output.back().ref_position - rs.output.front().ref_position + 1;
    return summary;
}

std::string EventAligner::summary_row(size_t read_idx, const std::string& fast5_path) const
{
    const ReadState& rs = m_reads[read_idx];
    const EventalignSummary summary = summarize(read_idx);
    if (summary.num_events <= 0) return std::string();
    const SquiggleScalings& scalings = rs.params.sr->scalings[rs.params.strand_idx];
    char buf[2048];
    std::string out;
    snprintf(buf, sizeof(buf), "%zu\t%s\t%s\t", (size_t)rs.params.read_idx, rs.params.sr->read_name.c_str(), fast5_path.c_str());
    out += buf;
    snprintf(buf, sizeof(buf), "%s\t%s\t", rs.pore_model->name.c_str(), rs.params.strand_idx == 0 ? "template" : "complement");
    out += buf;
    snprintf(buf, sizeof(buf), "%d\t%d\t%d\t%d\t", summary.num_events, summary.num_steps, summary.num_skips, summary.num_stays);
    out += buf;
    snprintf(buf, sizeof(buf), "%.2lf\t%.3lf\t%.3lf\t%.3lf\t%.3lf\n", summary.sum_duration, scalings.shift, scalings.scale, scalings.drift, scalings.var);
    out += buf;
    return out;
}

} // namespace nph
