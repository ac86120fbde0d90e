// nph_methylation.hpp — SURVEY.md section 8(f) row N3: call-methylation's per-read logic split into
// enumerate / one batched launch / scatter, plus the TSV writer.
//
//   calculate_methylation_for_read      ref: src/basemods/nanopolish_basemods.cpp:238-457
//   ScoredSite, MethylationCallingParameters   ref: src/basemods/nanopolish_basemods.h:45-77
//   AlignmentDB::_find_by_ref_bounds    ref: src/alignment/nanopolish_alignment_db.cpp:688-731
//   write_methylation_results_as_tsv    ref: src/nanopolish_call_methylation.cpp:532-550
//   create_modbam_record / create_reference_modbam_record (the Mm / Ml tags of --modbam-output)
//                                       ref: src/basemods/nanopolish_basemods.cpp:35-177, 179-238
//
// The reference scores two sequences per CpG group with two profile_hmm_score calls inside the per-read OpenMP
// loop.  Here add_read() only enumerates (motif scan -> groups -> window -> event bounds) and appends two jobs per
// group to one HmmBatch for the whole BamProcessor batch; run() launches once and scatters log-likelihoods back into
// the ScoredSites; write_tsv() formats them exactly like the reference.  BAM/FASTA access stays with the caller, which
// hands over what the reference pulls out of them: the reference substring and the (ref_pos, event_idx) pairs of
// EventAlignmentRecord (src/alignment/nanopolish_alignment_db.cpp:50-91).
#pragma once
#include <cstdio>
#include "nph_host.hpp"

namespace nph {

struct ScoredSite {
    ScoredSite() { ll_unmethylated[0] = ll_unmethylated[1] = ll_methylated[0] = ll_methylated[1] = 0; strands_scored = 0; }
    std::string chromosome;
    int start_position = 0;
    int end_position = 0;
    int n_motif = 0;
    std::string sequence;
    double ll_unmethylated[2];
    double ll_methylated[2];
    int strands_scored;
};

struct MethylationCallingParameters {
    int min_separation = 10;
    int min_flank = 10;
    std::string methylation_type = "cpg";
    const Alphabet* alphabet = nullptr;      // get_alphabet_by_name(methylation_type)
};

// What the reference derives from the BAM record, the FASTA and the SquiggleRead before the group loop.
struct EventAlignedRead {
    SquiggleRead* read = nullptr;
    std::string read_name;                   // bam_get_qname
    bool is_reverse = false;                 // bam_is_rev: the "+"/"-" column
    std::string contig;
    int ref_start_pos = 0;                   // record->core.pos
    std::string ref_seq;                     // reference over [ref_start_pos, bam_endpos], already gDNAAlphabet.disambiguate()d
    // per strand: EventAlignmentRecord::aligned_events (ref_pos ascending, read_pos = event index) and ::rc
    std::vector<AlignedPair> aligned_events[2];
    bool rc[2] = {false, false};
};

// ---- modBAM tags (SAM specification: Mm = delta-encoded positions of the modified base, Ml = probabilities 0..255) ----
struct ModbamTags {
    std::string mm;                // e.g. "C+m?,3,0,12;"
    std::vector<uint8_t> ml;       // one entry per position listed in mm
};
// the unmodified base that the alphabet's methylated symbol replaces ('C' for cpg); the reference asserts one recognition site
char unmodified_symbol_of(const Alphabet* alphabet);
// reference position and probability code of every called site (strand-0 likelihoods, like the reference)
void calculate_call_vectors(const std::map<int, ScoredSite>& calls, const Alphabet* alphabet,
                            std::vector<size_t>& call_reference_positions, std::vector<uint8_t>& call_probabilities);
std::string generate_mm_tag(char unmodified_symbol, const std::string& sequence, const std::vector<size_t>& call_seq_indices);
// create_modbam_record: tags for the read's own BAM record.  bam_seq = SEQ as stored (reference orientation),
// aligned_bases = get_aligned_segments(record)[0] (nph_eventalign.hpp), is_reverse = bam_is_rev(record).
ModbamTags modbam_tags(const std::string& bam_seq, const std::vector<AlignedPair>& aligned_bases, bool is_reverse,
                       const std::map<int, ScoredSite>& calls, const MethylationCallingParameters& params);
// create_reference_modbam_record: tags against the (disambiguated) reference over the record's span
ModbamTags reference_modbam_tags(const std::string& ref_seq, int ref_start_pos, const std::map<int, ScoredSite>& calls,
                                 const MethylationCallingParameters& params);

bool find_by_ref_bounds(const std::vector<AlignedPair>& pairs, int ref_start, int ref_stop, int& read_start, int& read_stop);

class MethylationCaller {
public:
    explicit MethylationCaller(const MethylationCallingParameters& params);
    // enumerate the read's motif groups and queue their jobs; returns the read's index in this batch.
    // region_start/region_end = -1 for no window restriction (the reference's -w option).
    size_t add_read(const EventAlignedRead& r, int region_start = -1, int region_end = -1);
    // The same for a whole BamProcessor batch, enumerated by host_threads() workers into private job lists that are
    // spliced in read order (so job order, scores and output equal add_read called read by read).  Returns the index of
    // the first read.  The reference runs this enumeration inside its per-read OpenMP loop too.
    size_t add_reads(const std::vector<EventAlignedRead>& reads, int region_start = -1, int region_end = -1);
    void run(Engine& engine, double indel_bias = hmm_indel_bias_factor);       // one launch for every queued group
    const std::map<int, ScoredSite>& sites(size_t read_idx) const { return m_reads[read_idx].sites; }
    size_t num_reads() const { return m_reads.size(); }
    size_t num_jobs() const { return m_batch.size(); }
    const HmmBatch& batch() const { return m_batch; }      // the queued jobs (read-only; for inspection and tests)
    void write_tsv(FILE* fp, size_t read_idx) const;
    std::string tsv(size_t read_idx) const;
    std::vector<std::string> tsv_batch() const;            // every read's rows, formatted by host_threads() workers
    // the Mm / Ml tags of --modbam-output for one read of the batch (after run()): bam_seq and aligned_bases as in
    // nph::modbam_tags
    ModbamTags modbam(size_t read_idx, const std::string& bam_seq, const std::vector<AlignedPair>& aligned_bases) const
    {
        return modbam_tags(bam_seq, aligned_bases, m_reads[read_idx].is_reverse, m_reads[read_idx].sites, m_params);
    }
    void clear();

private:
    struct Pending { size_t read; int site_key; size_t strand; size_t job_u, job_m; };
    struct ReadEntry { std::string name; bool is_reverse; std::map<int, ScoredSite> sites; };
    MethylationCallingParameters m_params;
    HmmBatch m_batch;
    std::vector<Pending> m_pending;
    std::vector<ReadEntry> m_reads;
};

} // namespace nph
