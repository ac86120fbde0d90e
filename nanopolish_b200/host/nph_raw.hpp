// nph_raw.hpp — SURVEY.md section 8(f) row N4: the read prologue SquiggleRead::load_from_raw runs before any HMM
// call, as ONE batch over many reads with every numeric step on the device.
//
//   trim_and_segment_raw + detect_events            ref: src/nanopolish_squiggle_read.cpp:226-235
//   events -> SquiggleEvent (duration, start_time)  ref: src/nanopolish_squiggle_read.cpp:243-250
//   estimate_scalings_using_mom                     ref: src/nanopolish_squiggle_read.cpp:238-240
//   adaptive_banded_simple_event_align              ref: src/nanopolish_squiggle_read.cpp:270
//   base_to_event_map, events_per_base              ref: src/nanopolish_squiggle_read.cpp:273-302
//   recalibrate_model + QC                          ref: src/nanopolish_squiggle_read.cpp:304-336
//
// The reference does this per read inside the SquiggleRead constructor (one read per OpenMP thread); here the reads
// of a BamProcessor batch / an AlignmentDB region go through one device call (nph_load_from_raw_batch).  DNA and direct
// RNA (squiggle_read.cpp:192-213,262-265: U->T, 5-mer u_to_t_rna model, RNA detector parameters, events reversed to 5'->3').
#pragma once
#include "nph_host.hpp"

namespace nph {

// what Fast5Data / slow5 hands to load_from_raw (src/io/nanopolish_fast5_io.h:33-49), reduced to what is read
struct RawRead {
    std::string read_name;
    std::string read_sequence;        // the basecalled sequence
    std::vector<float> samples;       // picoamps, like raw_table::raw
    double sample_rate = 4000.0;
    SquiggleReadNucleotideType nucleotide_type = SRNT_DNA;   // SRNT_RNA: experiment_type "rna"/"internal_rna" (squiggle_read.cpp:192-195)
};

// the reference's global counters g_failed_alignment_reads / g_failed_calibration_reads / g_qc_fail_reads
// (src/nanopolish_squiggle_read.cpp:27-34), per batch
struct LoadFromRawStats {
    size_t total = 0, empty_after_trim = 0, failed_alignment = 0, failed_calibration = 0, qc_fail = 0;
};

// One SquiggleRead per input, in input order: strand-0 events, scalings, events_per_base and base_to_event_map set as
// load_from_raw leaves them; a read that fails a QC step has its events cleared, exactly like the reference.
// A signal that trims to nothing aborts the reference (assert et.n > 0); here it yields a read without events and is
// counted in empty_after_trim.
enum SquiggleReadFlags { SRF_NO_MODEL = 1, SRF_LOAD_RAW_SAMPLES = 2 };     // ref: src/nanopolish_squiggle_read.h:46-50
// flags & SRF_LOAD_RAW_SAMPLES: keep the trimmed samples on the read (eventalign --samples / --signal-index)
std::vector<std::unique_ptr<SquiggleRead>> load_from_raw(Engine& engine, const PoreModel& base_model, const std::vector<RawRead>& raw,
                                                         LoadFromRawStats* stats = nullptr, uint32_t flags = 0);

// scrappie's detect_events / trim_and_segment_raw for one signal (batch of one; the building blocks above)
std::vector<nph_event> detect_events(Engine& engine, const std::vector<float>& samples, const nph_event_params& params);
nph_raw_range trim_and_segment_raw(Engine& engine, const std::vector<float>& samples, int trim_start = 200, int trim_end = 10,
                                   int varseg_chunk = 100, float varseg_thresh = 0.0f);
extern const nph_event_params event_detection_defaults, event_detection_rna;   // ref: scrappie/event_detection.h:15-29

} // namespace nph
