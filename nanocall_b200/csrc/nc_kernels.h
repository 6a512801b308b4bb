// Kernel argument blocks and launch-side declarations.
#ifndef NC_KERNELS_H
#define NC_KERNELS_H

#include "nc_internal.h"
#include <cuda_runtime.h>

namespace nc {

enum { VIT_THREADS = 512 };

// alpha-column kernel: a finished forward pass, queued for a traceback service warp
struct TbTicket
{
    unsigned job;          // index into jobs
    unsigned slab;         // the forward CTA that ran the job (its release counter)
    unsigned final_state;  // arg max of the last column
    unsigned ready;        // written last (release); polled with acquire
    unsigned col0;         // first column of the job's extent in the pool
    unsigned pad[3];
};

struct VitArgs
{
    const DevJob* jobs;
    const unsigned* order;      // job indices, longest first
    unsigned n_jobs;
    unsigned* next_job;         // device counter, zeroed before the launch
    const float* models;        // n_models x MODEL_FLOATS
    const float* mean;
    const float* stdv;
    const float* start;
    const float* log_stdv;
    const unsigned long long* landed;  // null, or the number of events already copied to the device (streamed input)
    unsigned long long ev_total;       // events in the call
    unsigned char* bp_pool;     // viterbi_kernel: gridDim.x slabs of slab_bytes of backpointers (4096 B/event);
                                // viterbi_alpha_kernel: slab_bytes of alpha columns (16384 B/event), one allocator
    size_t slab_bytes;
    float* path_logprob;        // n_jobs
    unsigned short* states;     // packed like the events, may be null
    unsigned char* moves;       // packed like the events, may be null
    float log_2pi;              // (float)log(2*pi)  (Pore_Model.hpp:28)
    float log_n_states;         // logf(4096.f)      (Viterbi.hpp:51)
    // alpha-column kernel only: CTAs [0, n_tb) are traceback service warps fed through `tickets`, the other n_fwd
    // CTAs run forward passes (columns from the device-wide allocator; one release counter per forward CTA)
    unsigned n_fwd, n_tb;
    TbTicket* tickets;          // n_jobs entries, zeroed before the launch
    unsigned* tb_tail;          // tickets published
    unsigned* tb_head;          // tickets claimed
    unsigned* slab_free;        // per forward CTA: number of its jobs the traceback service has released
    void* colalloc;             // device-wide allocator of alpha columns (ColAlloc in nc_viterbi_alpha.cu)
    const unsigned* job_col0;   // viterbi_cluster_kernel: first pool column of job order[q]
    unsigned* abort_word;       // zeroed before the launch; a CTA that waited longer than the deadline for columns or for a
                                // ticket (the persistent grid is not making progress), or that found the allocator's
                                // extent list full, writes a non-zero code here and every waiting loop gives up on seeing it:
                                // the call then fails with NC_ERR_STATE instead of hanging (1 = wait for a CTA's release slot,
                                // 2 = wait for columns, 3 = wait for a ticket, 4 = extent list full)
    long long wait_limit;       // nanoseconds of polling back-off a CTA may spend waiting for columns / tickets before it
                                // raises abort_word
    unsigned long long* stats;  // optional (may be null): [0] forward cycles, [1] forward cycles waiting for a slab,
                                // [2] traceback busy cycles, [3] traceback cycles waiting for a ticket, [4] passes,
                                // [5] lane steps, [6] jobs traced   (sums over CTAs / service warps)
};

__global__ void viterbi_kernel(const VitArgs a);        // backpointer form (long reads: 4 KiB/event of scratch)
size_t viterbi_smem_bytes();
__global__ void viterbi_alpha_kernel(const VitArgs a);  // alpha-column form (fast path: 16 KiB/event of scratch)
__global__ void viterbi_cluster_kernel(const VitArgs a);   // two CTAs (256 threads each, cluster of 2) per job, grid = 2 n_jobs
size_t viterbi_alpha_smem_bytes();
size_t viterbi_alpha_colalloc_bytes();
unsigned viterbi_alpha_max_forward_ctas();   // bound of the allocator's extent list
void viterbi_alpha_colalloc_init(void* host_image, unsigned pool_columns);

// ---- Forward/Backward + trainer statistics
enum { FB_EV_TILE = 16 };   // events per CTA of pm_stats_kernel
#ifndef NC_EM_TILE
#define NC_EM_TILE 50
#endif
enum { EM_TILE = NC_EM_TILE };   // events per CTA of emission_kernel

struct FbSeq
{
    unsigned long long ev_off;  // first event in the packed arrays
    unsigned long long slab;    // float offset of this sequence's E | alpha | beta slab in the scratch pool
    unsigned long long ev_out;  // first row of this sequence in pm_stats
    unsigned n_events;
    unsigned job;               // index into jobs: model + scaling + transition LUT of the sequence's strand
    unsigned strand;
    unsigned generic;           // 1: the strand's transition parameters are the defaults and a custom table is in force
                                // (fwbw_generic_kernel takes the sequence, fwbw_kernel skips it)
};

struct FbGroup
{
    unsigned seq_begin, seq_end;   // sequences of the group, in the caller's order
    float log_p_stay[2];           // logf(p_stay)                          (Parameter_Trainer.hpp:443)
    float log_p_step_4[2];         // (float)(log(1 - p_stay - p_skip) - log(4))   (:444)
};

struct FbArgs
{
    const DevJob* jobs;
    const FbSeq* seqs;
    const FbGroup* groups;
    unsigned n_seqs;
    unsigned n_groups;
    unsigned* next_item;
    const float* models;
    const float4* pm_consts;       // per model and state {mu, sigma^2, 1/sigma^2, eta}, {1/eta, lambda, 0, 0} (unscaled model)
    const float* mean;
    const float* stdv;
    const float* start;
    const float* log_stdv;
    const float* logsum_tbl;       // 16000 floats (logsum.hpp:113-127)
    const unsigned* train_kmers;   // Parameter_Trainer::st_train_kmers()
    unsigned n_train_kmers;
    float* scratch;
    float* log_pr_data;            // n_seqs
    float* pm_stats;               // total training events x 6
    float* st_stats;               // n_groups x 2 strands x {denom, stay, skip}
    float log_2pi;
    float log_n_states;
};

// An arbitrary transition table (nanocall --trans) as stored lists: predecessors of j in from_idx/from_lp
// [from_off[j], from_off[j+1]) in ascending source order (ties in file order), successors in to_* in file order
struct GenTrans
{
    const unsigned* from_off;   // 4097
    const unsigned* from_idx;
    const float* from_lp;
    const unsigned* to_off;     // 4097
    const unsigned* to_idx;
    const float* to_lp;
};
__global__ void viterbi_generic_kernel(const VitArgs a, const GenTrans g);
size_t viterbi_generic_smem_bytes();
__global__ void fwbw_generic_kernel(const FbArgs a, const GenTrans g);   // uses next_item[1] as its work counter
size_t fwbw_generic_smem_bytes();

__global__ void emission_kernel(const FbArgs a);
__global__ void fwbw_kernel(const FbArgs a);
__global__ void pm_stats_kernel(const FbArgs a);
__global__ void st_stats_kernel(const FbArgs a);
size_t fwbw_smem_bytes();
size_t st_stats_smem_bytes();
unsigned st_stats_threads();
unsigned st_stats_max_kmers();
size_t pm_stats_smem_bytes();

} // namespace nc

#endif
