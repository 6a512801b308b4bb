// Read sources of the host pipeline and the step between an event table and the hot path: the reference's
// Fast5_Summary (Fast5_Summary.hpp:138-370, 505-745) without HDF5.
//
//   raw event tables ("NCRW0001", one or many reads per file): EventDetection events as the fast5 file holds them
//       (mean, stdv as double, start, length in samples) + sampling rate + read id.  Everything Fast5_Summary does
//       downstream of File::get_eventdetection_events is done here: --max-ed-events truncation (:505-525), abasic
//       level (:527-543), hairpin detection and trimming (:545-571, 653-731), the event filter (:734-745), the
//       time base (:348-364), the initial scaling (:223-278 in Pipeline::init_read_params).
//   segmented event tables (.ncev, .events.tsv): per-strand float events, already trimmed and on the strand's time
//       base (round-1 formats; kept for inputs that are produced by another segmenter).
//   synthetic source: a pool of R7.3-like 2D reads (template, hairpin island, complement) generated from the builtin
//       models and replayed with fresh read ids, so that sweeps of 10^6 reads never sit in memory or on disk.
#ifndef NC_READS_HPP
#define NC_READS_HPP

#include "pipeline.hpp"

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace nchost {

struct Ed_Event   // fast5::EventDetection_Event_Entry (fast5.hpp:55-68)
{
    double mean, stdv;
    long long start, length;
};

struct Raw_Read
{
    std::string file_name;   // path as given (base_file_name = last component without ".fast5")
    std::string read_id;     // EventDetection read_id attribute; empty = use base_file_name
    double sampling_rate = 0;
    std::vector< Ed_Event > ed;
};

// Fast5_Summary::summarize for one raw read: fills r (names, strand bounds, abasic level, per-strand events,
// scale_strands_together).  Returns false when the reference would leave num_ed_events == 0 (the read is skipped by
// training and basecalling but still gets a --stats row); `why` then holds the reference's log message.
bool summarize_raw_read(const Options& opt, Raw_Read&& raw, Read& r, std::string& why);
bool summarize_events(const Options& opt, const std::string& file_name, const std::string& read_id, double sampling_rate,
                      const Ed_Event* ed, size_t n_raw, Read& r, std::string& why);

// one record after the other; the callback returns false to stop
bool read_ncrw_file(const std::string& path, std::vector< Raw_Read >& out, std::string& err);
bool is_ncrw_file(const std::string& path);
void write_ncrw_file(const std::string& path, const std::vector< Raw_Read >& reads);

class Read_Source
{
public:
    virtual ~Read_Source() {}
    // next read in input order, summarised; false at the end.  Thread-safe: several loader threads may call it;
    // `index` is the read's position in the input order.
    virtual bool next(Read& r, size_t& index) = 0;
    virtual size_t size_hint() const { return 0; }   // number of reads when known, else 0
};

std::unique_ptr< Read_Source > make_file_source(const Options& opt, const std::vector< std::string >& files, int log_level);
// spec = "n_reads[:seed[:pool[:shape]]]"; shape = "2d" (default: 5000 + 5000 events), "2d:<nt>:<nc>", "1d:<n>",
// or "mix" (1D, the length mixture of BASELINE.json configs[4])
std::unique_ptr< Read_Source > make_synth_source(const Options& opt, const std::string& spec, const std::string& data_dir,
                                                 int log_level);

} // namespace nchost

#endif
