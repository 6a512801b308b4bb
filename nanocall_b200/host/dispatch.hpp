// Streaming, dynamically dispatched execution of the host pipeline: the replacement of nanocall's two pfor loops
// (pfor.hpp:169-285 as used at nanocall.cpp:282-574 and :610-860).
//
//   loader threads  pull reads from a Read_Source (file list or synthetic), summarise them (Fast5_Summary) and put
//                   them into a bounded queue: the input never sits in memory as a whole
//   one dispatcher  per GPU (own thread, own nc_ctx): takes the lowest-numbered reads of the queue as a batch whenever
//                   it is free -- pfor's first-come scheduling with a batch instead of a chunk of reads (inside a batch
//                   the Viterbi call orders its jobs longest first) -- runs training, selection and basecalling on it;
//                   a second thread of the dispatcher basecalls batch k while batch k+1 is being trained, so the host
//                   passes of one batch (packing, base sequences, records) run under the kernels of the other
//   ordered sink    results are written in input order, like pfor's heap of finished chunks (pfor.hpp:216-235)
#ifndef NC_DISPATCH_HPP
#define NC_DISPATCH_HPP

#include "reads.hpp"

#include <iosfwd>
#include <string>
#include <vector>

namespace nchost {

struct Run_Config
{
    Options opt;
    std::vector< int > devices;           // CUDA device of every dispatcher
    size_t batch_reads = 4096;            // reads per batch (upper bound)
    size_t batch_events = (size_t)48 << 20;  // events per batch (upper bound; both strands)
    size_t queue_events = (size_t)192 << 20; // loaders pause above this many queued events
    unsigned loader_threads = 0;          // 0 = pick from the host's core count
    size_t pool_bytes = 0;                // Viterbi scratch per GPU; 0 = sized from the first batch
    bool overlap = true;                  // a second thread per GPU basecalls batch k while batch k+1 is trained
};

struct Device_Stats
{
    int device = -1;
    size_t reads = 0, batches = 0, train_rounds = 0, fwbw_events = 0, viterbi_events = 0, read_events = 0;
    double train_kernel_ms = 0, viterbi_kernel_ms = 0;
    double emission_ms = 0, fwbw_ms = 0, pm_stats_ms = 0, st_stats_ms = 0;
    double init_s = 0, train_s = 0, basecall_s = 0, wait_s = 0, hand_wait_s = 0, train_call_s = 0, viterbi_call_s = 0;
    double first_batch_at_s = 0, last_batch_done_s = 0;   // relative to the start of the run
};

struct Run_Stats
{
    std::vector< Device_Stats > dev;
    size_t reads = 0, read_events = 0;
    double wall_s = 0;          // whole run, including context creation
    double steady_wall_s = 0;   // from the first batch handed out to the last result written
    std::string error;
};

// Runs the whole pipeline over `src`.  fasta / stats may be null.  Returns false (and stats.error) on failure.
bool run_pipeline(const Run_Config& cfg, Read_Source& src, std::ostream* fasta, std::ostream* stats_tsv, Run_Stats& stats);

std::string stats_json(const Run_Config& cfg, const Run_Stats& s);

} // namespace nchost

#endif
