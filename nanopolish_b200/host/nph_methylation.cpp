// nph_methylation.cpp — see nph_methylation.hpp.
#include "nph_methylation.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>

namespace nph {

static const int MAX_EVENT_TO_BP_RATIO = 20;       // ref: src/alignment/nanopolish_alignment_db.h:18

// lower_bound on ref_pos for both ends; bounded only if an aligned base sits at or outside each boundary.
// (The right-hand test compares the successor with ref_START, as the reference does.)
bool find_by_ref_bounds(const std::vector<AlignedPair>& pairs, int ref_start, int ref_stop, int& read_start, int& read_stop)
{
    auto lb = [](const AlignedPair& o, int v) { return o.ref_pos < v; };
    auto start_iter = std::lower_bound(pairs.begin(), pairs.end(), ref_start, lb);
    auto stop_iter = std::lower_bound(pairs.begin(), pairs.end(), ref_stop, lb);
    if (start_iter == pairs.end() || stop_iter == pairs.end()) return false;
    bool left_bounded = start_iter->ref_pos <= ref_start || (start_iter != pairs.begin() && (start_iter - 1)->ref_pos <= ref_start);
    bool right_bounded = stop_iter->ref_pos >= ref_stop || ((stop_iter + 1) != pairs.end() && (stop_iter + 1)->ref_pos >= ref_start);
    if (!(left_bounded && right_bounded)) return false;
    read_start = start_iter->read_pos;
    read_stop = stop_iter->read_pos;
    return true;
}

// ---------------------------------------------------------------------------------------------
// modBAM tags
// ---------------------------------------------------------------------------------------------
char unmodified_symbol_of(const Alphabet* alphabet)
{
    if (alphabet->num_recognition_sites() != 1) throw Error(NPH_ERR_UNSUPPORTED, "modBAM output needs an alphabet with one recognition site");
    char unmodified = 'N';
    const char* modified_motif = alphabet->get_recognition_site_methylated(0);
    for (size_t i = 0; i < alphabet->recognition_length(); ++i)
        if (modified_motif[i] == METHYLATED_SYMBOL) unmodified = alphabet->get_recognition_site(0)[i];
    if (unmodified == 'N') throw Error(NPH_ERR_INVALID, "alphabet without a methylated symbol");
    return unmodified;
}

void calculate_call_vectors(const std::map<int, ScoredSite>& calls, const Alphabet* alphabet,
                            std::vector<size_t>& call_reference_positions, std::vector<uint8_t>& call_probabilities)
{
    for (const auto& kv : calls) {
        const ScoredSite& call = kv.second;
        // the called positions are where methylate() puts the symbol; start_position is the first site of the group, so
        // the flank in front of it is subtracted
        const std::string m_seq = alphabet->methylate(call.sequence);
        const size_t flank_offset = m_seq.find_first_of(METHYLATED_SYMBOL);
        if (flank_offset == std::string::npos) throw Error(NPH_ERR_INVALID, "scored site without a motif");      // the reference asserts
        // shared by every motif of the group; strand 0 only, as in the reference
        const double methylation_probability = std::exp(call.ll_methylated[0]) / (std::exp(call.ll_methylated[0]) + std::exp(call.ll_unmethylated[0]));
        const uint8_t code = (uint8_t)std::min(255, (int)(methylation_probability * 255));
        for (size_t j = 0; j < m_seq.size(); j++) {
            if (m_seq[j] == METHYLATED_SYMBOL) {
                call_reference_positions.push_back((size_t)(call.start_position + (int)j - (int)flank_offset));
                call_probabilities.push_back(code);
            }
        }
    }
}

std::string generate_mm_tag(char unmodified_symbol, const std::string& sequence, const std::vector<size_t>& call_seq_indices)
{
    std::string delta_str;
    delta_str += unmodified_symbol;
    delta_str += "+m?";
    size_t count_start = 0;
    for (size_t call_index = 0; call_index < call_seq_indices.size(); ++call_index) {
        // unmodified bases skipped since the previous listed position
        int count = 0;
        for (size_t j = count_start; j < call_seq_indices[call_index]; ++j) count += sequence[j] == unmodified_symbol;
        delta_str += ',';
        delta_str += std::to_string(count);
        count_start = call_seq_indices[call_index] + 1;
    }
    delta_str += ';';
    return delta_str;
}

ModbamTags modbam_tags(const std::string& bam_seq, const std::vector<AlignedPair>& aligned_bases, bool is_reverse,
                       const std::map<int, ScoredSite>& calls, const MethylationCallingParameters& params)
{
    const Alphabet* alphabet = params.alphabet ? params.alphabet : get_alphabet_by_name(params.methylation_type);
    if (std::string(alphabet->get_name()) != "cpg") throw Error(NPH_ERR_UNSUPPORTED, "modBAM output supports the cpg alphabet only");   // the reference asserts
    const char unmodified_symbol = unmodified_symbol_of(alphabet);
    std::vector<size_t> call_reference_positions;
    std::vector<uint8_t> call_reference_probabilities;
    calculate_call_vectors(calls, alphabet, call_reference_positions, call_reference_probabilities);

    // reference position -> index into the read as sequenced
    const std::string original_sequence = !is_reverse ? bam_seq : gDNAAlphabet.reverse_complement(bam_seq);
    std::map<size_t, size_t> reference_to_read_map;
    for (const AlignedPair& ap : aligned_bases)
        reference_to_read_map[(size_t)ap.ref_pos] = !is_reverse ? (size_t)ap.read_pos : original_sequence.length() - (size_t)ap.read_pos - 1;
    // a CG read from the opposite strand has its C aligned to the reference G
    const size_t strand_offset = !is_reverse ? 0 : 1;
    ModbamTags out;
    std::vector<size_t> call_seq_indices;
    for (size_t i = 0; i < call_reference_positions.size(); ++i) {
        auto iter = reference_to_read_map.find(call_reference_positions[i] + strand_offset);
        if (iter == reference_to_read_map.end()) continue;
        const size_t read_index = iter->second;
        if (read_index < original_sequence.size() && original_sequence[read_index] == unmodified_symbol) {
            call_seq_indices.push_back(read_index);
            out.ml.push_back(call_reference_probabilities[i]);
        }
    }
    // SEQ is reverse complemented for reverse-strand records: list the calls in the original direction
    if (is_reverse) {
        std::reverse(call_seq_indices.begin(), call_seq_indices.end());
        std::reverse(out.ml.begin(), out.ml.end());
    }
    out.mm = generate_mm_tag(unmodified_symbol, original_sequence, call_seq_indices);
    return out;
}

ModbamTags reference_modbam_tags(const std::string& ref_seq, int ref_start_pos, const std::map<int, ScoredSite>& calls,
                                 const MethylationCallingParameters& params)
{
    const Alphabet* alphabet = params.alphabet ? params.alphabet : get_alphabet_by_name(params.methylation_type);
    const char unmodified_symbol = unmodified_symbol_of(alphabet);
    std::vector<size_t> call_reference_positions;
    ModbamTags out;
    calculate_call_vectors(calls, alphabet, call_reference_positions, out.ml);
    std::vector<size_t> indices;
    for (size_t pos : call_reference_positions) indices.push_back(pos - (size_t)ref_start_pos);
    out.mm = generate_mm_tag(unmodified_symbol, ref_seq, indices);
    return out;
}

MethylationCaller::MethylationCaller(const MethylationCallingParameters& params) : m_params(params)
{
    if (!m_params.alphabet) m_params.alphabet = get_alphabet_by_name(m_params.methylation_type);
}

void MethylationCaller::clear()
{
    m_batch.clear(); m_pending.clear(); m_reads.clear();
}

size_t MethylationCaller::add_read(const EventAlignedRead& r, int region_start, int region_end)
{
    const size_t read_idx = m_reads.size();
    m_reads.push_back(ReadEntry{r.read_name, r.is_reverse, {}});
    std::map<int, ScoredSite>& site_score_map = m_reads.back().sites;
    const std::string& ref_seq = r.ref_seq;
    if (ref_seq.empty()) return read_idx;
    const Alphabet* alphabet = m_params.alphabet;

    for (size_t strand_idx = 0; strand_idx < 2; ++strand_idx) {
        if (!r.read->has_events_for_strand(strand_idx)) continue;
        const size_t k = r.read->get_model_k(strand_idx);
        const PoreModel* motif_model = r.read->get_model(strand_idx, m_params.methylation_type);
        if (!motif_model) continue;                          // no model for this motif on this strand
        const std::vector<AlignedPair>& aligned_events = r.aligned_events[strand_idx];

        // scan for motifs, then batch them into groups separated by more than min_separation
        // (a site can only start where its first symbol stands — position 0 aside, where the reference's matcher also
        // accepts a string that lies wholly inside a site — so the full matcher runs on those positions only)
        bool first_symbol[256] = {false};
        for (size_t s = 0; s < alphabet->num_recognition_sites(); ++s) first_symbol[(unsigned char)alphabet->get_recognition_site(s)[0]] = true;
        std::vector<int> motif_sites;
        for (size_t i = 0; i + 1 < ref_seq.size(); ++i)
            if ((i == 0 || first_symbol[(unsigned char)ref_seq[i]]) && alphabet->is_motif_match(ref_seq, i)) motif_sites.push_back((int)i);
        std::vector<std::pair<size_t, size_t>> groups;
        size_t curr_idx = 0;
        while (curr_idx < motif_sites.size()) {
            size_t end_idx = curr_idx + 1;
            while (end_idx < motif_sites.size()) {
                if (motif_sites[end_idx] - motif_sites[end_idx - 1] > m_params.min_separation) break;
                end_idx += 1;
            }
            groups.push_back({curr_idx, end_idx});
            curr_idx = end_idx;
        }

        for (const auto& g : groups) {
            const size_t start_idx = g.first, end_idx = g.second;
            const int sub_start_pos = motif_sites[start_idx] - m_params.min_flank;
            const int sub_end_pos = motif_sites[end_idx - 1] + m_params.min_flank;
            const int span = motif_sites[end_idx - 1] - motif_sites[start_idx];
            if (sub_start_pos <= m_params.min_separation || span > 200) continue;

            std::string subseq = ref_seq.substr(sub_start_pos, sub_end_pos - sub_start_pos + 1);
            const int calling_start = sub_start_pos + r.ref_start_pos;
            const int calling_end = sub_end_pos + r.ref_start_pos;

            int e1 = 0, e2 = 0;
            const bool bounded = find_by_ref_bounds(aligned_events, calling_start, calling_end, e1, e2);
            // (the reference divides by calling_start - calling_end, a negative number, so this ratio never trips)
            const double ratio = std::fabs((double)(e2 - e1)) / (calling_start - calling_end);
            if (!bounded || std::abs(e2 - e1) <= 10 || ratio > MAX_EVENT_TO_BP_RATIO) continue;

            const uint32_t hmm_flags = HAF_ALLOW_PRE_CLIP | HAF_ALLOW_POST_CLIP;
            HMMInputData data;
            data.read = r.read;
            data.pore_model = motif_model;
            data.strand = (uint8_t)strand_idx;
            data.rc = r.rc[strand_idx];
            data.event_start_idx = (uint32_t)e1;
            data.event_stop_idx = (uint32_t)e2;
            data.event_stride = data.event_start_idx <= data.event_stop_idx ? 1 : -1;

            const int start_position = motif_sites[start_idx] + r.ref_start_pos;
            const int end_position = motif_sites[end_idx - 1] + r.ref_start_pos;
            // the reference scores first and filters by region afterwards; filtering first yields the same output
            if ((region_start != -1 && start_position < region_start) || (region_end != -1 && end_position >= region_end)) continue;

            std::string m_subseq = alphabet->methylate(subseq);
            // the jobs of this read only ever ask for the strand data.rc selects: the other strand's strings (the
            // reference builds all four) are left empty
            const bool need_rc = data.rc != 0;
            std::string rc_subseq = need_rc ? alphabet->reverse_complement(subseq) : std::string();
            std::string rc_m_subseq = need_rc ? alphabet->reverse_complement(m_subseq) : std::string();
            HMMInputSequence unmethylated(std::move(subseq), std::move(rc_subseq), alphabet);
            HMMInputSequence methylated(std::move(m_subseq), std::move(rc_m_subseq), alphabet);
            const size_t ju = m_batch.add(unmethylated, data, hmm_flags);
            const size_t jm = m_batch.add(methylated, data, hmm_flags);

            auto iter = site_score_map.find(start_position);
            if (iter == site_score_map.end()) {
                ScoredSite ss;
                ss.chromosome = r.contig;
                ss.start_position = start_position;
                ss.end_position = end_position;
                ss.n_motif = (int)(end_idx - start_idx);
                const size_t site_output_start = motif_sites[start_idx] - k + 1;
                const size_t site_output_end = motif_sites[end_idx - 1] + k;
                ss.sequence = ref_seq.substr(site_output_start, site_output_end - site_output_start);
                iter = site_score_map.insert({start_position, ss}).first;
            }
            m_pending.push_back(Pending{read_idx, start_position, strand_idx, ju, jm});
        }
    }
    return read_idx;
}

size_t MethylationCaller::add_reads(const std::vector<EventAlignedRead>& reads, int region_start, int region_end)
{
    const size_t first = m_reads.size();
    const size_t n = reads.size();
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads(), (n + 15) / 16));
    if (T == 1) {
        for (const EventAlignedRead& r : reads) add_read(r, region_start, region_end);
        return first;
    }
    // contiguous blocks of reads per worker: splicing the workers' lists in worker order reproduces the sequential order
    std::vector<std::unique_ptr<MethylationCaller>> locals((size_t)T);     // separately allocated: no false sharing of their cursors
    for (auto& l : locals) l.reset(new MethylationCaller(m_params));
    std::vector<std::string> errors((size_t)T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        const size_t b = n * (size_t)t / (size_t)T, e = n * ((size_t)t + 1) / (size_t)T;
        try {
            for (size_t i = b; i < e; ++i) locals[t]->add_read(reads[i], region_start, region_end);
        } catch (const std::exception& ex) { errors[t] = ex.what(); }
    }
    for (const std::string& e : errors) if (!e.empty()) throw Error(NPH_ERR_INVALID, e);
    for (int t = 0; t < T; ++t) {
        MethylationCaller& l = *locals[t];
        const size_t read_base = m_reads.size(), job_base = m_batch.size();
        for (Pending p : l.m_pending) { p.read += read_base; p.job_u += job_base; p.job_m += job_base; m_pending.push_back(p); }
        for (ReadEntry& re : l.m_reads) m_reads.push_back(std::move(re));
        m_batch.append(std::move(l.m_batch));
    }
    return first;
}

void MethylationCaller::run(Engine& engine, double indel_bias)
{
    const std::vector<float> ll = m_batch.run(engine, indel_bias);
    for (const Pending& p : m_pending) {
        ScoredSite& ss = m_reads[p.read].sites[p.site_key];
        ss.ll_unmethylated[p.strand] = ll[p.job_u];       // float -> double, like `double s = profile_hmm_score(...)`
        ss.ll_methylated[p.strand] = ll[p.job_m];
        ss.strands_scored += 1;
    }
    m_pending.clear();
    m_batch.clear();
}

std::string MethylationCaller::tsv(size_t read_idx) const
{
    const ReadEntry& re = m_reads[read_idx];
    std::string out;
    char buf[512];
    for (const auto& kv : re.sites) {
        const ScoredSite& ss = kv.second;
        const double sum_ll_m = ss.ll_methylated[0] + ss.ll_methylated[1];
        const double sum_ll_u = ss.ll_unmethylated[0] + ss.ll_unmethylated[1];
        const double diff = sum_ll_m - sum_ll_u;
        // chromosome, strand, start, end, read_name, log_lik_ratio %.2lf, log_lik_methylated %.2lf, log_lik_unmethylated %.2lf
        out += ss.chromosome; out += '\t'; out += re.is_reverse ? '-' : '+'; out += '\t';
        out += std::to_string(ss.start_position); out += '\t'; out += std::to_string(ss.end_position); out += '\t';
        out += re.name; out += '\t';
        out.append(buf, format_fixed(buf, diff, 2)); out += '\t';
        out.append(buf, format_fixed(buf, sum_ll_m, 2)); out += '\t';
        out.append(buf, format_fixed(buf, sum_ll_u, 2)); out += '\t';
        out += std::to_string(ss.strands_scored) + "\t" + std::to_string(ss.n_motif) + "\t" + ss.sequence + "\n";
    }
    return out;
}

std::vector<std::string> MethylationCaller::tsv_batch() const
{
    std::vector<std::string> out(m_reads.size());
#pragma omp parallel for schedule(dynamic, 16) num_threads(host_threads()) if (m_reads.size() > 64)
    for (long long i = 0; i < (long long)m_reads.size(); ++i) out[i] = tsv((size_t)i);
    return out;
}

void MethylationCaller::write_tsv(FILE* fp, size_t read_idx) const
{
    const std::string s = tsv(read_idx);
    fwrite(s.data(), 1, s.size(), fp);
}

} // namespace nph
