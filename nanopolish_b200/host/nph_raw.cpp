// nph_raw.cpp — see nph_raw.hpp.
#include "nph_raw.hpp"

#include <cmath>

namespace nph {

const nph_event_params event_detection_defaults = {3, 6, 1.4f, 9.0f, 0.2f};
const nph_event_params event_detection_rna = {7, 14, 2.5f, 9.0f, 1.0f};

namespace {

// event detection over [start, end) of each signal; grows the per-read room once if the first guess was short
std::vector<std::vector<nph_event>> detect_batch(Engine& engine, const std::vector<float>& flat, const std::vector<uint64_t>& off,
                                                 const std::vector<nph_raw_range>& range, const nph_event_params& prm)
{
    const size_t n = range.size();
    std::vector<std::vector<nph_event>> out(n);
    std::vector<uint32_t> live;
    for (size_t i = 0; i < n; ++i) if (range[i].end > range[i].start) live.push_back((uint32_t)i);
    if (live.empty()) return out;
    for (int attempt = 0; attempt < 2; ++attempt) {
        std::vector<nph_raw_read> rr(live.size());
        uint64_t room = 0;
        for (size_t t = 0; t < live.size(); ++t) {
            const uint32_t i = live[t];
            const uint32_t ns = range[i].end - range[i].start;
            rr[t].sample_off = off[i] + range[i].start;
            rr[t].n_samples = ns;
            rr[t].event_cap = attempt == 0 ? ns / 2 + 8 : ns;      // n_samples always suffices
            rr[t].event_off = room;
            room += rr[t].event_cap;
        }
        std::vector<nph_event> ev(room);
        std::vector<uint32_t> cnt(live.size());
        const int rc = nph_detect_events_batch(engine.ctx(), flat.data(), flat.size(), rr.data(), rr.size(), &prm, ev.data(), ev.size(), cnt.data());
        if (rc == NPH_ERR_UNSUPPORTED && attempt == 0) continue;
        engine.check(rc, "nph_detect_events_batch");
        for (size_t t = 0; t < live.size(); ++t) out[live[t]].assign(ev.begin() + rr[t].event_off, ev.begin() + rr[t].event_off + cnt[t]);
        break;
    }
    return out;
}

} // namespace

std::vector<nph_event> detect_events(Engine& engine, const std::vector<float>& samples, const nph_event_params& params)
{
    if (samples.empty()) return {};
    return detect_batch(engine, samples, {0}, {nph_raw_range{0u, (uint32_t)samples.size()}}, params)[0];
}

nph_raw_range trim_and_segment_raw(Engine& engine, const std::vector<float>& samples, int trim_start, int trim_end, int varseg_chunk,
                                   float varseg_thresh)
{
    nph_raw_range out{0, 0};
    if (samples.empty()) return out;
    nph_raw_read rr{};
    rr.n_samples = (uint32_t)samples.size();
    engine.check(nph_trim_raw_batch(engine.ctx(), samples.data(), samples.size(), &rr, 1, trim_start, trim_end, varseg_chunk, varseg_thresh, &out),
                 "nph_trim_raw_batch");
    return out;
}

std::vector<std::unique_ptr<SquiggleRead>> load_from_raw(Engine& engine, const PoreModel& base_model, const std::vector<RawRead>& raw,
                                                         LoadFromRawStats* stats)
{
    const size_t n = raw.size();
    LoadFromRawStats st;
    st.total = n;
    std::vector<std::unique_ptr<SquiggleRead>> reads(n);
    const uint32_t k = base_model.k;
    for (size_t i = 0; i < n; ++i) {
        reads[i].reset(new SquiggleRead());
        SquiggleRead& sr = *reads[i];
        sr.read_name = raw[i].read_name;
        sr.read_sequence = raw[i].read_sequence;
        sr.pore_type = PORETYPE_R9;
        sr.base_model[0] = &base_model;
        sr.sample_rate = raw[i].sample_rate;
    }
    if (n == 0) { if (stats) *stats = st; return reads; }

    // 1. trim: scrappie's defaults, hard-coded at the call site
    std::vector<uint64_t> off(n);
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) { off[i] = total; total += raw[i].samples.size(); }
    std::vector<float> flat(total);
    std::vector<nph_raw_read> rr(n);
    for (size_t i = 0; i < n; ++i) {
        std::copy(raw[i].samples.begin(), raw[i].samples.end(), flat.begin() + off[i]);
        rr[i] = nph_raw_read{off[i], 0, (uint32_t)raw[i].samples.size(), 0};
    }
    std::vector<nph_raw_range> range(n, nph_raw_range{0, 0});
    if (total) engine.check(nph_trim_raw_batch(engine.ctx(), flat.data(), total, rr.data(), n, 200, 10, 100, 0.0f, range.data()), "nph_trim_raw_batch");

    // 2. events
    std::vector<std::vector<nph_event>> events = detect_batch(engine, flat, off, range, event_detection_defaults);

    // 3. SquiggleEvent conversion; reads that can go on to alignment
    std::vector<uint32_t> live;
    for (size_t i = 0; i < n; ++i) {
        SquiggleRead& sr = *reads[i];
        const std::vector<nph_event>& et = events[i];
        if (et.empty() || sr.read_sequence.size() < k) { ++st.empty_after_trim; continue; }
        sr.events[0].resize(et.size());
        double start_time = 0;
        for (size_t e = 0; e < et.size(); ++e) {
            const float length_in_seconds = (float)(et[e].length / sr.sample_rate);
            sr.events[0][e] = SquiggleEvent{et[e].mean, et[e].stdv, start_time, length_in_seconds, logf(et[e].stdv)};
            start_time += length_in_seconds;
        }
        live.push_back((uint32_t)i);
    }
    if (live.empty()) { if (stats) *stats = st; return reads; }

    // 4. flatten once: events, ranks, jobs shared by MoM, ABEA and the calibration
    std::vector<nph_read> nr(live.size());
    std::vector<nph_abea_job> jobs(live.size());
    std::vector<float> mean;
    std::vector<double> time;
    std::vector<uint32_t> ranks;
    uint64_t pairs_total = 0;
    for (size_t t = 0; t < live.size(); ++t) {
        const SquiggleRead& sr = *reads[live[t]];
        const std::vector<SquiggleEvent>& ev = sr.events[0];
        const uint32_t n_kmers = (uint32_t)(sr.read_sequence.size() - k + 1);
        nr[t] = nph_read{mean.size(), (uint32_t)ev.size(), 0, 1.0, 0.0, 0.0, 1.0, 0.0, 0.0};
        jobs[t] = nph_abea_job{ranks.size(), pairs_total, (uint32_t)t, n_kmers, (uint32_t)ev.size() + n_kmers, 0};
        pairs_total += jobs[t].pairs_cap;
        for (const SquiggleEvent& e : ev) { mean.push_back(e.mean); time.push_back(e.start_time); }
        for (uint32_t i = 0; i < n_kmers; ++i) ranks.push_back(base_model.pmalphabet->kmer_rank(sr.read_sequence.c_str() + i, k));
    }
    const uint32_t mid = engine.model_id(&base_model);

    // 5. method-of-moments scalings (drift 0, var 1)
    std::vector<double> ss(2 * live.size());
    engine.check(nph_mom_batch(engine.ctx(), nr.data(), nr.size(), mean.data(), mean.size(), ranks.data(), ranks.size(), jobs.data(), jobs.size(),
                               mid, ss.data()), "nph_mom_batch");
    for (size_t t = 0; t < live.size(); ++t) {
        reads[live[t]]->scalings[0].set4(ss[2 * t], ss[2 * t + 1], 0.0, 1.0);
        nr[t].shift = ss[2 * t]; nr[t].scale = ss[2 * t + 1];
    }

    // 6. event alignment
    std::vector<nph_aligned_pair> pairs(pairs_total);
    std::vector<nph_abea_result> res(live.size());
    engine.check(nph_abea_batch(engine.ctx(), nr.data(), nr.size(), mean.data(), time.data(), mean.size(), ranks.data(), ranks.size(),
                                jobs.data(), jobs.size(), mid, pairs.data(), pairs.size(), res.data()), "nph_abea_batch");

    // 7. base-to-event map, events per base, recalibration, QC
    std::vector<nph_event_range> b2e(ranks.size());
    std::vector<nph_calibration> cal(live.size());
    engine.check(nph_recalibrate_batch(engine.ctx(), nr.data(), nr.size(), mean.data(), mean.size(), ranks.data(), ranks.size(), jobs.data(),
                                       jobs.size(), mid, pairs.data(), pairs.size(), res.data(), b2e.data(), cal.data()), "nph_recalibrate_batch");
    for (size_t t = 0; t < live.size(); ++t) {
        SquiggleRead& sr = *reads[live[t]];
        const nph_calibration& c = cal[t];
        if (c.status & NPH_CAL_NOT_ALIGNED) {
            sr.events[0].clear();
            sr.events_per_base[0] = 0.0;
            ++st.failed_alignment;
            continue;
        }
        sr.base_to_event_map.resize(jobs[t].n_kmers);
        for (uint32_t i = 0; i < jobs[t].n_kmers; ++i) {
            const nph_event_range& r = b2e[jobs[t].rank_off + i];
            sr.base_to_event_map[i].indices[0] = IndexPair{r.start, r.stop};
        }
        sr.events_per_base[0] = c.events_per_base;
        if (!(c.status & NPH_CAL_TOO_FEW_EVENTS)) sr.scalings[0].set4(c.shift, c.scale, c.drift, c.var);
        if (c.status & (NPH_CAL_TOO_FEW_EVENTS | NPH_CAL_HIGH_VAR)) { sr.events[0].clear(); ++st.failed_calibration; }
        else if (c.status & NPH_CAL_TOO_MANY_STAYS) { sr.events[0].clear(); sr.events[1].clear(); ++st.qc_fail; }
    }
    if (stats) *stats = st;
    return reads;
}

} // namespace nph
