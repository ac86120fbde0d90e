// nph_variants.hpp — SURVEY.md section 8(f) row N2: the variant-scoring callers of profile_hmm_score_set,
// re-expressed as job generators over the batched forward kernel.
//
//   Variant (the fields scoring touches)        ref: src/common/nanopolish_variant.h:21-110
//   Haplotype::apply_variant(s)                  ref: src/nanopolish_haplotype.cpp:30-85
//   generate_methylated_alternatives             ref: src/common/nanopolish_variant.cpp:158-178
//   score_variant_thresholded                    ref: src/common/nanopolish_variant.cpp:765-799
//   the scoring loop of score_variant_group      ref: src/common/nanopolish_variant.cpp:242-258
//
// The reference scores one (read, sequence) pair per profile_hmm_score call inside an OpenMP loop.  Here every
// (read, haplotype, alphabet alternative) triple of a whole candidate set becomes one job of ONE launch; the base
// haplotype is scored once per read instead of once per candidate; the reference's host arithmetic on the returned
// floats (profile_hmm_score_set's table logsum in double, the early-exit accumulation of score_variant_thresholded
// in read order) is applied afterwards, unchanged.
#pragma once
#include "nph_host.hpp"

namespace nph {

struct Variant {
    std::string ref_name;
    size_t ref_position = 0;
    std::string ref_seq;
    std::string alt_seq;
    double quality = 0.0;
    std::string info, genotype, filter;
    std::string key() const { return ref_name + ":" + std::to_string(ref_position) + ":" + ref_seq + ":" + alt_seq; }
};

class Haplotype {
public:
    static const size_t INSERTED_POSITION;
    Haplotype(const std::string& ref_name, size_t ref_position, const std::string& ref_sequence);
    bool apply_variant(const Variant& v);                       // false (and no change) if incompatible
    bool apply_variants(const std::vector<Variant>& variants);
    const std::string& get_sequence() const { return m_sequence; }
    const std::string& get_reference() const { return m_reference; }
    size_t get_reference_position() const { return m_ref_position; }
    size_t get_reference_position_for_haplotype_base(size_t i) const;
    const std::vector<Variant>& get_variants() const { return m_variants; }

private:
    size_t find_derived_index_by_ref_lower_bound(size_t ref_index) const;
    std::string m_ref_name;
    size_t m_ref_position;
    std::string m_reference, m_sequence;
    std::vector<size_t> m_coordinate_map;
    std::vector<Variant> m_variants;
};

// {sequence} plus one methylated version per alphabet in methylation_types that changes the sequence
std::vector<HMMInputSequence> generate_methylated_alternatives(const HMMInputSequence& sequence,
                                                               const std::vector<std::string>& methylation_types);

// scores[read][haplotype] = profile_hmm_score_set(generate_methylated_alternatives(haplotype), input[read], flags)
// — the scoring loop of score_variant_group — as one batch.
std::vector<std::vector<double>> score_haplotypes(const std::vector<Haplotype>& haplotypes, const std::vector<HMMInputData>& input,
                                                  uint32_t alignment_flags, const std::vector<std::string>& methylation_types,
                                                  Engine& engine, double indel_bias = hmm_indel_bias_factor);

// score_variant_group (ref: src/common/nanopolish_variant.cpp:182-262): every combination of up to max_r of the group's variants
// (max_r = the largest r for which the haplotype count stays below max_haplotypes, :193-206) that applies cleanly to the base
// haplotype — r = 0, the base haplotype, included — scored against every read with profile_hmm_score_set, all as ONE batch.
// combinations[c] lists the variant ids of combination c (ascending ids, combinations of size r in lexicographic order; the
// reference's Combinations generator visits the same sets in a different order), scores[c][read] is what
// VariantGroup::set_combination_read_score receives.
struct VariantGroupScores {
    std::vector<std::vector<size_t>> combinations;
    std::vector<std::vector<double>> scores;
    size_t max_r = 0;
};
size_t nChoosek(size_t n, size_t k);                           // ref: src/common/nanopolish_common.cpp:93-105 (int arithmetic inside)
VariantGroupScores score_variant_group(const std::vector<Variant>& variants, const Haplotype& base_haplotype,
                                       const std::vector<HMMInputData>& input, int max_haplotypes, uint32_t alignment_flags,
                                       const std::vector<std::string>& methylation_types, Engine& engine,
                                       double indel_bias = hmm_indel_bias_factor);

// score_variant_thresholded for a whole candidate list: quality_i = sum over reads, in input order, of
// (variant_score - base_score), where a read is skipped once |running total| >= score_threshold — exactly the
// single-thread behaviour of the reference loop (its OpenMP version makes the skip set racy; the result here is
// the deterministic one).  Candidates that do not apply to the base haplotype keep quality 0.
std::vector<Variant> score_variants_thresholded(const std::vector<Variant>& input_variants, const Haplotype& base_haplotype,
                                                const std::vector<HMMInputData>& input, uint32_t alignment_flags,
                                                uint32_t score_threshold, const std::vector<std::string>& methylation_types,
                                                Engine& engine, double indel_bias = hmm_indel_bias_factor);

// the reference's single-candidate signature
Variant score_variant_thresholded(const Variant& input_variant, Haplotype base_haplotype, const std::vector<HMMInputData>& input,
                                  uint32_t alignment_flags, uint32_t score_threshold, const std::vector<std::string>& methylation_types);

} // namespace nph
