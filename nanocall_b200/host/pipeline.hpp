// Host side of nanocall_b200: the batched replacement of nanocall's per-read pfor loops.
//
//   reference                                   here
//   init_models            nanocall.cpp:97-178   Pipeline::init_models (builtin tables or TSV model files)
//   Fast5_Summary          Fast5_Summary.hpp     Read (event tables come from .events.tsv / .ncev, not fast5)
//     initial scaling      :223-278              Pipeline::init_read_params
//   train_reads            nanocall.cpp:275-582  Pipeline::train_reads   -> nc_train_round_batch per EM round
//   basecall_reads         nanocall.cpp:593-869  Pipeline::basecall_reads -> nc_viterbi_packed for all candidates
//   write_fasta            nanocall.cpp:584-591  write_fasta
//   --stats TSV            Fast5_Summary.hpp:460-502  Pipeline::write_stats
// Everything numeric happens behind the C ABI (include/nanocall_b200.h); this file is bookkeeping.
#ifndef NC_PIPELINE_HPP
#define NC_PIPELINE_HPP

#include "nanocall_b200.h"

#include <array>
#include <iosfwd>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace nchost {

typedef std::array< std::string, 2 > Model_Key;  // {template model, complement model}; "" = unused strand

struct Options
{
    // defaults = nanocall.cpp:56-94
    unsigned min_ed_events = 10;
    unsigned max_ed_events = 100000;
    std::array< unsigned, 4 > trim_margins{ { 50u, 50u, 50u, 50u } };  // after start, before end, before hairpin, after hairpin
    bool template_only = false;            // --1d
    double abasic_level_top_percent = 1.0; // Fast5_Summary.hpp:92-104, set by the pore preset (nanocall.cpp:943-969)
    double abasic_level_top_offset = -1.0; // -1: by pore preset (r73 -> 5, r9 -> 0)
    std::string ed_group;                  // accepted for compatibility: event tables hold one EventDetection group
    unsigned chunk_size = 1;               // accepted for compatibility: batching replaces pfor's chunks
    unsigned fasta_line_width = 80;
    float scaling_select_threshold = 20.0f;
    float scaling_min_progress = 1.0f;
    unsigned scaling_max_rounds = 10;
    unsigned scaling_num_events = 200;
    bool double_strand_scaling = true;
    bool train_transitions = true;
    bool train_scaling = true;
    bool train = true;
    bool basecall = true;
    float pr_skip = 0.3f;
    float pr_stay = 0.1f;
    std::string pore = "r9";
    int train_drift = -1;          // -1: by pore preset (r73 -> 1, r9 -> 0; nanocall.cpp:943-969)
    int log_level = 2;             // 0 error, 1 warning, 2 info, 3 debug
    std::vector< std::string > model_files;  // "strand:file"
    std::string model_fofn;        // file of "strand:file" lines (nanocall.cpp:118-127)
    std::string trans_fn;          // custom initial state transitions (nanocall.cpp:180-193)
    std::string data_dir;          // where builtin_models.{bin,txt} live
    unsigned host_threads = 4;     // helper threads of a dispatcher for its per-read host work (packing, base sequences)
};

struct Strand_Events
{
    std::vector< float > mean, stdv, start, length;
    size_t size() const { return mean.size(); }
};

struct Read
{
    std::string read_id;
    std::string base_file_name;
    // Fast5_Summary fields (Fast5_Summary.hpp:30-43).  num_ed_events == 0: the read is skipped by training and
    // basecalling (nanocall.cpp:293,623) but keeps its --stats row.
    unsigned num_ed_events = 0;
    float abasic_level = 0.f;
    float sampling_rate = 0.f;
    std::array< unsigned, 4 > strand_bounds{ { 0u, 0u, 0u, 0u } };
    Strand_Events events[2];
    bool scale_strands_together = false;
    std::array< Model_Key, 3 > preferred_model;                            // Fast5_Summary.hpp:32-34
    std::map< Model_Key, nc_pm_params > pm_params_m;                       // :35
    std::map< Model_Key, std::array< nc_st_params, 2 > > st_params_m;      // :36
    // outputs of basecall_reads
    std::array< std::string, 2 > base_seq;
    std::array< float, 2 > log_path_prob{ { 0.f, 0.f } };
    std::array< bool, 2 > called{ { false, false } };
};

struct Model
{
    std::string name;
    int strand = 2;
    int id = -1;          // nc_model_register id in the context
    float mean = 0.f, stdv = 0.f;
    std::vector< float > table;  // 4096 x 4
};

class Pipeline
{
public:
    Pipeline(const Options& o, int device, size_t viterbi_events_hint = 0);
    ~Pipeline();
    Pipeline(const Pipeline&) = delete;
    Pipeline& operator=(const Pipeline&) = delete;

    void init_models();
    void init_transitions();
    // allocate now what batches of up to `reads` reads / `events` events will need (device scratch, pinned staging), so
    // that the first batch does not pay for it
    void reserve(size_t reads, size_t events);
    void init_read_params(Read& r) const;
    void init_reads_params(std::vector< Read* >& reads) const;   // the same for a batch, on the helper threads
    void train_reads(std::vector< Read* >& reads);
    void basecall_reads(std::vector< Read* >& reads);
    static void write_fasta(std::ostream& os, const std::string& name, const std::string& seq, unsigned width);
    void write_output(std::ostream& os, const Read& r) const;
    static void write_stats_header(std::ostream& os);
    static void write_stats(std::ostream& os, const Read& r, const Options& opt);

    const std::map< std::string, Model >& models() const { return models_; }
    nc_ctx* ctx() { return ctx_; }
    // device time spent in the hot-path kernels (ms), for the run summary
    double train_kernel_ms = 0, viterbi_kernel_ms = 0;
    double train_call_s = 0, viterbi_call_s = 0;   // wall time inside the two C-ABI calls (copies included)
    size_t train_rounds = 0, fwbw_events = 0, viterbi_events = 0;

private:
    void check(int rc, const char* what) const;
    nc_st_params default_st() const { nc_st_params s; s.p_stay = opt_.pr_stay; s.p_skip = opt_.pr_skip; return s; }
    // grow-only page-locked staging buffers of the Viterbi calls (nc_host_alloc): events in, states / moves out
    struct Pinned
    {
        void* p = nullptr;
        size_t cap = 0;
        void* reserve(size_t bytes);
        ~Pinned();
    };
    Pinned pin_mean_, pin_stdv_, pin_start_, pin_states_, pin_moves_;
    Pinned pin_tr_mean_, pin_tr_stdv_, pin_tr_start_;   // the packed training sequences of an EM round (train_reads)
    Options opt_;
    nc_ctx* ctx_ = nullptr;
    // train_reads and basecall_reads may run on two threads (dispatch.cpp trains batch k+1 while batch k is basecalled):
    // the calls into the context are serialised here, the host passes around them are not
    std::mutex gpu_mu_;
    std::map< std::string, Model > models_;  // ordered by name, as Pore_Model_Dict (std::map)
};

// event-table readers (the always-available input path; fast5 needs libhdf5, absent from this build)
bool load_events_tsv(const std::string& path, Read& r, std::string& err);
bool load_events_ncev(const std::string& path, std::vector< Read >& reads, std::string& err);

void log_line(int level, int threshold, const std::string& msg);

} // namespace nchost

#endif
