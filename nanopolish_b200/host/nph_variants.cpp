// nph_variants.cpp — see nph_variants.hpp.
#include "nph_variants.hpp"

#include <cmath>

namespace nph {

const size_t Haplotype::INSERTED_POSITION = std::string::npos;

Haplotype::Haplotype(const std::string& ref_name, size_t ref_position, const std::string& ref_sequence)
    : m_ref_name(ref_name), m_ref_position(ref_position), m_reference(ref_sequence), m_sequence(ref_sequence)
{
    m_coordinate_map.resize(m_reference.size());
    for (size_t i = 0; i < m_coordinate_map.size(); ++i) m_coordinate_map[i] = m_ref_position + i;
}

size_t Haplotype::find_derived_index_by_ref_lower_bound(size_t ref_index) const
{
    for (size_t i = 0; i < m_coordinate_map.size(); ++i)
        if (m_coordinate_map[i] != INSERTED_POSITION && m_coordinate_map[i] >= ref_index) return i;
    return m_coordinate_map.size();
}

// Replace ref_seq by alt_seq at the haplotype position that still maps to v.ref_position; refuse (leaving the
// haplotype untouched) when that reference base was already replaced or the reference allele does not match.
bool Haplotype::apply_variant(const Variant& v)
{
    const size_t idx = find_derived_index_by_ref_lower_bound(v.ref_position);
    if (idx == m_coordinate_map.size() || m_coordinate_map[idx] != v.ref_position) return false;
    const size_t rl = v.ref_seq.length(), al = v.alt_seq.length();
    if (m_sequence.substr(idx, rl) != v.ref_seq) return false;
    m_sequence.replace(idx, rl, v.alt_seq);
    auto first = m_coordinate_map.begin() + idx;
    auto it = m_coordinate_map.erase(first, first + rl);
    m_coordinate_map.insert(it, al, INSERTED_POSITION);     // inserted bases have no reference coordinate
    m_variants.push_back(v);
    return true;
}

bool Haplotype::apply_variants(const std::vector<Variant>& variants)
{
    bool good = true;
    for (const Variant& v : variants) good = good && apply_variant(v);
    return good;
}

size_t Haplotype::get_reference_position_for_haplotype_base(size_t i) const
{
    return m_coordinate_map[i] == INSERTED_POSITION ? std::string::npos : m_coordinate_map[i];
}

std::vector<HMMInputSequence> generate_methylated_alternatives(const HMMInputSequence& sequence,
                                                               const std::vector<std::string>& methylation_types)
{
    std::vector<HMMInputSequence> out;
    out.push_back(sequence);
    for (const std::string& name : methylation_types) {
        const Alphabet* alphabet = get_alphabet_by_name(name);
        std::string methylated = alphabet->methylate(sequence.get_sequence());
        if (methylated != sequence.get_sequence()) out.emplace_back(methylated, alphabet);
    }
    return out;
}

namespace {

// Adds the jobs of profile_hmm_score_set(sequences, data) to the batch; returns (first job, count).
// `ranks` has one RankCache per sequence: a haplotype is scored against every read of the pile-up, its k-mer ranks are
// computed and shipped once per strand.
std::pair<size_t, size_t> add_score_set(HmmBatch& batch, const std::vector<HMMInputSequence>& sequences, std::vector<RankCache>& ranks,
                                        const HMMInputData& data, uint32_t flags)
{
    const size_t first = batch.size();
    ranks.resize(sequences.size());
    batch.add(sequences[0], data, flags, ranks[0]);
    for (size_t i = 1; i < sequences.size(); ++i) {
        HMMInputData alt = data;
        alt.pore_model = data.read->get_model(data.strand, sequences[i].get_alphabet()->get_name());
        if (!alt.pore_model) throw Error(NPH_ERR_INVALID, std::string("read has no pore model for alphabet ") + sequences[i].get_alphabet()->get_name());
        batch.add(sequences[i], alt, flags, ranks[i]);
    }
    return {first, sequences.size()};
}

double combine(const std::vector<float>& scores, std::pair<size_t, size_t> span)
{
    float out = 0.0f;
    int rc = nph_score_set_combine(scores.data() + span.first, 1, (uint32_t)span.second, &out);
    if (rc != NPH_OK) throw Error(rc, "nph_score_set_combine");
    return out;     // profile_hmm_score_set returns float; callers hold it in a double
}

} // namespace

std::vector<std::vector<double>> score_haplotypes(const std::vector<Haplotype>& haplotypes, const std::vector<HMMInputData>& input,
                                                  uint32_t alignment_flags, const std::vector<std::string>& methylation_types,
                                                  Engine& engine, double indel_bias)
{
    HmmBatch batch;
    std::vector<std::vector<HMMInputSequence>> alts;
    for (const Haplotype& h : haplotypes) alts.push_back(generate_methylated_alternatives(HMMInputSequence(h.get_sequence()), methylation_types));
    std::vector<std::vector<RankCache>> alt_ranks(haplotypes.size());
    std::vector<std::vector<std::pair<size_t, size_t>>> spans(input.size(), std::vector<std::pair<size_t, size_t>>(haplotypes.size()));
    for (size_t ri = 0; ri < input.size(); ++ri)
        for (size_t hi = 0; hi < haplotypes.size(); ++hi) spans[ri][hi] = add_score_set(batch, alts[hi], alt_ranks[hi], input[ri], alignment_flags);
    const std::vector<float> s = batch.run(engine, indel_bias);
    std::vector<std::vector<double>> out(input.size(), std::vector<double>(haplotypes.size()));
    for (size_t ri = 0; ri < input.size(); ++ri)
        for (size_t hi = 0; hi < haplotypes.size(); ++hi) out[ri][hi] = combine(s, spans[ri][hi]);
    return out;
}

size_t nChoosek(size_t n, size_t k)
{
    if (k > n) return 0;
    if (k * 2 > n) k = n - k;
    if (k == 0) return 1;
    int result = (int)n;                                       // the reference accumulates in an int
    for (int i = 2; i <= (int)k; ++i) {
        result *= (int)(n - i + 1);
        result /= i;
    }
    return (size_t)result;
}

VariantGroupScores score_variant_group(const std::vector<Variant>& variants, const Haplotype& base_haplotype,
                                       const std::vector<HMMInputData>& input, int max_haplotypes, uint32_t alignment_flags,
                                       const std::vector<std::string>& methylation_types, Engine& engine, double indel_bias)
{
    const size_t num_variants = variants.size();
    // the largest number of variants that can be tested jointly without exceeding max_haplotypes
    size_t sum_num_haplotypes = 0, max_r = 1;
    while (max_r <= num_variants) {
        const size_t num_haplotypes_r = nChoosek(num_variants, max_r);
        if (num_haplotypes_r + sum_num_haplotypes < (size_t)max_haplotypes) sum_num_haplotypes += num_haplotypes_r;
        else break;
        max_r += 1;
    }
    max_r -= 1;
    VariantGroupScores out;
    out.max_r = max_r;
    std::vector<Haplotype> haplotypes;
    for (size_t r = 0; r <= max_r; ++r) {
        std::vector<size_t> idx(r);
        for (size_t i = 0; i < r; ++i) idx[i] = i;
        for (;;) {
            Haplotype current = base_haplotype;
            std::vector<Variant> subset;
            for (size_t i : idx) subset.push_back(variants[i]);
            if (current.apply_variants(subset)) { out.combinations.push_back(idx); haplotypes.push_back(current); }
            // next r-subset of {0..n-1} in lexicographic order
            int i = (int)r - 1;
            while (i >= 0 && idx[i] == num_variants - r + (size_t)i) --i;
            if (i < 0) break;
            ++idx[i];
            for (size_t j = (size_t)i + 1; j < r; ++j) idx[j] = idx[j - 1] + 1;
        }
    }
    const std::vector<std::vector<double>> by_read = score_haplotypes(haplotypes, input, alignment_flags, methylation_types, engine, indel_bias);
    out.scores.assign(haplotypes.size(), std::vector<double>(input.size()));
    for (size_t ri = 0; ri < input.size(); ++ri)
        for (size_t hi = 0; hi < haplotypes.size(); ++hi) out.scores[hi][ri] = by_read[ri][hi];
    return out;
}

std::vector<Variant> score_variants_thresholded(const std::vector<Variant>& input_variants, const Haplotype& base_haplotype,
                                                const std::vector<HMMInputData>& input, uint32_t alignment_flags,
                                                uint32_t score_threshold, const std::vector<std::string>& methylation_types,
                                                Engine& engine, double indel_bias)
{
    std::vector<Variant> out = input_variants;
    HmmBatch batch;
    const std::vector<HMMInputSequence> base_seqs =
        generate_methylated_alternatives(HMMInputSequence(base_haplotype.get_sequence()), methylation_types);
    std::vector<std::pair<size_t, size_t>> base_span(input.size());
    std::vector<RankCache> base_ranks;
    for (size_t j = 0; j < input.size(); ++j) base_span[j] = add_score_set(batch, base_seqs, base_ranks, input[j], alignment_flags);   // once per read

    std::vector<char> applies(input_variants.size(), 0);
    std::vector<std::vector<std::pair<size_t, size_t>>> var_span(input_variants.size());
    for (size_t v = 0; v < input_variants.size(); ++v) {
        Haplotype hap = base_haplotype;
        // the reference applies the variant without checking the result: an incompatible variant leaves the
        // haplotype unchanged, so variant_score == base_score and the quality stays 0
        applies[v] = hap.apply_variant(input_variants[v]) ? 1 : 0;
        if (!applies[v]) continue;
        const std::vector<HMMInputSequence> seqs = generate_methylated_alternatives(HMMInputSequence(hap.get_sequence()), methylation_types);
        var_span[v].resize(input.size());
        std::vector<RankCache> var_ranks;
        for (size_t j = 0; j < input.size(); ++j) var_span[v][j] = add_score_set(batch, seqs, var_ranks, input[j], alignment_flags);
    }
    const std::vector<float> s = batch.run(engine, indel_bias);
    std::vector<double> base(input.size());
    for (size_t j = 0; j < input.size(); ++j) base[j] = combine(s, base_span[j]);
    for (size_t v = 0; v < input_variants.size(); ++v) {
        double total = 0.0f;
        if (applies[v]) {
            for (size_t j = 0; j < input.size(); ++j) {
                if (std::fabs(total) < score_threshold) total += (combine(s, var_span[v][j]) - base[j]);
            }
        }
        out[v].quality = total;
    }
    return out;
}

Variant score_variant_thresholded(const Variant& input_variant, Haplotype base_haplotype, const std::vector<HMMInputData>& input,
                                  uint32_t alignment_flags, uint32_t score_threshold, const std::vector<std::string>& methylation_types)
{
    return score_variants_thresholded({input_variant}, base_haplotype, input, alignment_flags, score_threshold, methylation_types,
                                      Engine::thread_default())[0];
}

} // namespace nph
