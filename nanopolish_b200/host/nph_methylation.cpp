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

            const std::string subseq = ref_seq.substr(sub_start_pos, sub_end_pos - sub_start_pos + 1);
            const std::string rc_subseq = alphabet->reverse_complement(subseq);
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

            HMMInputSequence unmethylated(subseq, rc_subseq, alphabet);
            const std::string m_subseq = alphabet->methylate(subseq);
            const std::string rc_m_subseq = alphabet->reverse_complement(m_subseq);
            HMMInputSequence methylated(m_subseq, rc_m_subseq, alphabet);
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
