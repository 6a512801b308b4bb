/* nanocall_b200 -- C ABI of the B200-native decoding hot path of nanocall.
 *
 * The reference (mateidavid/nanocall v0.7.4) has no FFI: its seam is the C++ class-template API
 *   Viterbi<float,6>::fill(pm, st, ev) + path_probability()        src/nanocall/Viterbi.hpp:35,44-46
 *   Forward_Backward<float,6>::fill(pm, st, ev) + log_pr_data()    src/nanocall/Forward_Backward.hpp:38-48
 *   Parameter_Trainer<float,6>::train_one_round(...)               src/nanocall/Parameter_Trainer.hpp:541-552
 * called from the two per-read lambdas of nanocall.cpp (:292-574 training, :621-857 basecalling).
 * This header is what a maintainer binds instead: the per-read calls become
 * BATCH calls, one job per (read-strand, candidate model), executed by hand-written sm_100a
 * kernels.  Plain pointers and sizes only; no exceptions cross the boundary; every function
 * returns NC_OK or a negative nc_status and nc_last_error() explains the last failure.
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device and fails with
 * NC_ERR_CUDA when there is none.
 *
 * Numerics: IEEE binary32, no FMA contraction, bit-identical to the reference for Viterbi
 * (path log-probability, states, moves) and for Forward/Backward alpha, beta, log Pr[data];
 * trainer outputs agree to ~1e-6 relative (device expf/logf and fixed-shape reductions).
 */
#ifndef NANOCALL_B200_H
#define NANOCALL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NC_N_STATES 4096u /* 4^6 k-mers; state index = 2-bit packed k-mer, A=0 C=1 G=2 T=3, first base in the
                             high bits (src/nanocall/Kmer.hpp:13-50) */
#define NC_KMER 6u

typedef enum {
    NC_OK = 0,
    NC_ERR_ARG = -1,    /* bad argument (NULL pointer, zero-length job, unknown model id, ...) */
    NC_ERR_CUDA = -2,   /* CUDA runtime error or no device */
    NC_ERR_NOMEM = -3,  /* host or device allocation failed / job exceeds the backpointer pool */
    NC_ERR_STATE = -4   /* call order (e.g. no model registered) */
} nc_status;

typedef enum { NC_MEM_HOST = 0, NC_MEM_DEVICE = 1 } nc_mem;

typedef struct nc_ctx nc_ctx;

/* One context per GPU (and per host dispatcher thread).  bp_pool_bytes = device bytes reserved for the Viterbi
 * scratch of the jobs in flight (alpha columns, 16 KiB per event, or backpointers, 4 KiB per event, for reads
 * too long for a forward CTA's share of the pool); 0 = 3/4 of the free memory, at most 140 GB.
 * Environment (read here, for tests): NC_STREAM_IN_MIN_EVENTS / NC_STREAM_IN_CHUNK = size from which
 * host-memory calls stream their events behind the kernel launch, and the chunk size, in events. */
int nc_ctx_create(int device, size_t bp_pool_bytes, nc_ctx** out);
void nc_ctx_destroy(nc_ctx* ctx);
const char* nc_last_error(const nc_ctx* ctx); /* ctx may be NULL: error of a failed nc_ctx_create */
/* cudaStream_t all work of this context is enqueued on (for timing with CUDA events). */
void* nc_ctx_stream(nc_ctx* ctx);
int nc_ctx_sync(nc_ctx* ctx);
/* device time (CUDA events on the context's stream) of the most recent hot-path kernel launch, ms */
float nc_ctx_last_kernel_ms(nc_ctx* ctx);
/* kernels launched by the most recent nc_viterbi_packed call (bench.py's gpu_launches) */
int nc_ctx_last_launches(nc_ctx* ctx);
/* Viterbi kernel choice: NC_VIT_AUTO = alpha-column kernel for every job whose columns fit a forward CTA's ring (16 KiB per
 * event), backpointer kernel (4 KiB per event) for longer jobs; NC_VIT_BACKPOINTER = backpointer kernel only.
 * Both produce the reference's bits; the switch exists for A/B measurements and tests. */
typedef enum { NC_VIT_AUTO = 0, NC_VIT_BACKPOINTER = 2 } nc_vit_mode;
int nc_ctx_set_viterbi_mode(nc_ctx* ctx, int mode);
/* Device-side counters of the alpha-column kernel, summed over CTAs / traceback service warps since the last
 * reset, in SM clock cycles: [0] forward passes, [1] forward CTAs waiting for ring space, [2] traceback busy,
 * [3] traceback waiting for work, then counts: [4] traceback passes, [5] traceback lane steps, [6] jobs traced. */
int nc_ctx_viterbi_stats(nc_ctx* ctx, uint64_t* out8, int reset);
/* number of SMs / name of the device, for reports */
int nc_ctx_device_info(nc_ctx* ctx, int* n_sms, size_t* total_mem, char* name, int name_cap);

/* Register an UNSCALED pore model: table = 4096 x {level_mean, level_stdv, sd_mean, sd_stdv}, the layout
 * of Builtin_Model::init_lists consumed by Pore_Model::load_from_vector (Pore_Model.hpp:220-239).
 * Derives sd_lambda and the logs on the host exactly as Pore_Model_State::update_sd_lambda /
 * update_logs do (Pore_Model.hpp:112,118-124).  strand: 0 template, 1 complement, 2 both. */
int nc_model_register(nc_ctx* ctx, const float* table, int strand, int* model_id);
/* mean / stdv of level_mean (Pore_Model::update_statistics, Pore_Model.hpp:308-314): initial scaling needs them */
int nc_model_stats(nc_ctx* ctx, int model_id, float* mean, float* stdv);

/* Pore_Model_Parameters (Pore_Model.hpp:42-52), in this order everywhere in the ABI. */
typedef struct { float scale, shift, drift, var, scale_sd, var_sd; } nc_pm_params;
/* State_Transition_Parameters (State_Transitions.hpp:14-51) */
typedef struct { float p_stay, p_skip; } nc_st_params;

/* ------------------------------------------------------------------------------------------------
 * Viterbi: replaces basecall_strand = Pore_Model::scale + compute_transitions_fast +
 * apply_drift_correction + Viterbi::fill (nanocall.cpp:645-690, Viterbi.hpp:44-150).
 *
 * Packed form.  Job k owns events [ev_off[k], ev_off[k+1]) of the concatenated event arrays:
 *   mean, stdv, start : Event::mean / stdv / start (Event.hpp:20-23; start in seconds from strand start)
 *   log_stdv          : Event::log_stdv = logf(stdv) after the stdv==0 -> 0.01 fix (Event.hpp:39-43); may be NULL,
 *                       the kernels then derive it on the device with a bit-compatible port of glibc's logf
 * ev_off, model_id, pm, st and path_logprob are always host arrays; the event arrays and
 * states/moves live where `mem` says.
 * Outputs: path_logprob[k] = Viterbi::path_probability(); states[i] = Event::model_state_idx;
 *          moves[i] = Event::move (0..6).  states/moves may be NULL.
 * Errors: NC_ERR_ARG for a job with zero events (the reference dereferences ev[0]). */
int nc_viterbi_packed(nc_ctx* ctx, uint32_t n_jobs, const uint64_t* ev_off,
                      const float* mean, const float* stdv, const float* start, const float* log_stdv,
                      const int32_t* model_id, const nc_pm_params* pm, const nc_st_params* st,
                      nc_mem mem,
                      float* path_logprob, uint16_t* states, uint8_t* moves);

/* Per-job-pointer form (host memory only); thin wrapper that packs and calls nc_viterbi_packed. */
typedef struct {
    const float* mean;
    const float* stdv;
    const float* start;
    uint32_t n_events;
    int32_t model_id;
    nc_pm_params pm;
    nc_st_params st;
} nc_vit_job;
typedef struct {
    float path_logprob;
    uint16_t* states;  /* n_events, caller-owned, may be NULL */
    uint8_t* moves;    /* n_events, caller-owned, may be NULL */
    char* bases;       /* Event_Sequence::get_base_seq (Event.hpp:85-99); caller-owned, may be NULL */
    uint32_t bases_cap;
    uint32_t n_bases;  /* out: bases needed (6 + sum of moves) */
} nc_vit_out;
int nc_viterbi_batch(nc_ctx* ctx, uint32_t n_jobs, const nc_vit_job* jobs, nc_vit_out* outs);

/* Event_Sequence::get_base_seq on the host from states/moves (Event.hpp:85-99). Returns bases needed. */
uint32_t nc_base_seq(uint32_t n_events, const uint16_t* states, const uint8_t* moves, char* out, uint32_t cap);

/* ------------------------------------------------------------------------------------------------
 * Forward/Backward: Forward_Backward::fill on a scaled model and drift-corrected events, as
 * Parameter_Trainer::fill_train_data runs it (Parameter_Trainer.hpp:141-155).  Debug/parity entry
 * point: alpha/beta are n_events x 4096 host arrays (either may be NULL). */
int nc_fwbw(nc_ctx* ctx, int32_t model_id, const nc_pm_params* pm, const nc_st_params* st,
            uint32_t n_events, const float* mean, const float* stdv, const float* start,
            float* alpha, float* beta, float* log_pr_data);

/* ------------------------------------------------------------------------------------------------
 * Training: Parameter_Trainer::train_one_round (Parameter_Trainer.hpp:541-579) for a batch of groups.
 * A group = the training sequences of one (read, candidate model pair): up to NC_MAX_TRAIN_SEQS
 * sequences, each tagged with its strand, two unscaled model ids (model_id[st]), the current
 * scaling parameters (common to both strands) and per-strand transition parameters.
 * Group g owns sequences [seq_off[g], seq_off[g+1]); sequence s owns events
 * [ev_off[s], ev_off[s+1]) of the concatenated host arrays. */
#define NC_MAX_TRAIN_SEQS 8u
typedef struct {
    int32_t model_id[2];
    nc_pm_params pm;
    nc_st_params st[2];
} nc_train_in;
typedef struct {
    nc_pm_params pm;     /* new scaling parameters (unchanged copy when done != 0 or !train_scaling) */
    nc_st_params st[2];  /* new transition parameters; NaN for a strand without sequences, as the reference */
    float fit;           /* sum over sequences of log Pr[data] under the CURRENT parameters */
    int32_t done;        /* 1 = singular scaling system, no further rounds possible */
} nc_train_out;
typedef struct {
    int train_scaling;      /* not --no-train-scaling */
    int train_transitions;  /* not --no-train-transitions */
    int train_drift;        /* Parameter_Trainer::pm_train_drift() (nanocall.cpp:970) */
} nc_train_opts;
int nc_train_round_batch(nc_ctx* ctx, uint32_t n_groups, const uint32_t* seq_off,
                         const uint64_t* ev_off, const uint8_t* seq_strand,
                         const float* mean, const float* stdv, const float* start,
                         const nc_train_in* in, const nc_train_opts* opts, nc_train_out* out);

/* Custom initial state transitions: nanocall's --trans (nanocall.cpp:180-193; file format of
 * State_Transitions::operator>>, State_Transitions.hpp:237-252): n_edges edges from[k] -> to[k] with log probability
 * logp[k], in FILE ORDER.  From then on every Viterbi job and every training strand whose transition parameters EQUAL
 * (p_stay_default, p_skip_default) -- State_Transition_Parameters::is_default(), State_Transitions.hpp:34-37 -- uses this
 * table (kernels that walk stored predecessor / successor lists) instead of the parametric compute_transitions_fast table;
 * all others keep the parametric table, exactly as basecall_strand (nanocall.cpp:651-661) and fill_train_data
 * (Parameter_Trainer.hpp:118-131) choose.  n_edges == 0 removes the custom table. */
int nc_ctx_set_default_transitions(nc_ctx* ctx, float p_stay_default, float p_skip_default, uint32_t n_edges,
                                   const uint16_t* from, const uint16_t* to, const float* logp);

/* Optional: allocate now what calls of this size will need (Forward/Backward scratch for train_events events per call, up
 * to the context's limit; device staging for NC_MEM_HOST Viterbi calls of viterbi_events events), so that the first
 * call does not pay for it.  No reference counterpart (the reference allocates per read, Forward_Backward.hpp:52-56). */
int nc_ctx_reserve(nc_ctx* ctx, uint64_t train_events, uint64_t viterbi_events);

/* Page-locked host memory for the event / state / move arrays of NC_MEM_HOST calls.  Optional: any host memory works,
 * but only pinned buffers are copied at full PCIe rate and let nc_viterbi_packed stream the events behind the kernel
 * launch.  nc_host_alloc returns NULL when the allocation fails (or there is no CUDA device). */
void* nc_host_alloc(size_t bytes);
void nc_host_free(void* p);

/* Device time of the training kernels since the last reset (CUDA events on the context's stream, summed over the waves
 * of every nc_train_round_batch / nc_fwbw call): out8 = { emission_kernel ms, fwbw_kernel ms, pm_stats_kernel ms,
 * st_stats_kernel ms, Forward/Backward events processed, kernels launched, waves, 0 }. */
int nc_ctx_train_stats(nc_ctx* ctx, double* out8, int reset);

/* ------------------------------------------------------------------------------------------------
 * Host-side helpers that restate small pieces of reference arithmetic needed around the calls. */
/* alg::mean_stdv_of<float> (hpptools alg.hpp:466-482) */
void nc_mean_stdv(uint32_t n, const float* x, float* mean, float* stdv);
/* log weight of an edge with 6-bit overlap mask m (bit0: i==j, bit l: suffix(i,6-l)==prefix(j,6-l)),
 * summed in State_Transitions::get_trans_prob's order (State_Transitions.hpp:125-144): lut[64] */
void nc_transition_lut(float p_stay, float p_skip, float* lut64);
/* Kmer::min_skip (Kmer.hpp:51-68) */
uint32_t nc_min_skip(uint32_t k1, uint32_t k2);
/* Dispatch order of the Viterbi jobs of one call (host helper, no device needed): lens[] descending, pool_columns =
 * alpha columns (16 KiB each) in the scratch pool, n_workers = forward CTAs.  perm[k] = index of the k-th job to
 * start: longest-first whenever the columns of the running jobs leave room, shorter jobs while memory is tight
 * (lens[] in any order is accepted: the first-wave check looks for the n_workers longest jobs itself).
 * Returns 1 when the order was changed, 0 when longest-first is kept (perm = identity). */
int nc_plan_dispatch_order(uint32_t n_jobs, const uint32_t* lens, uint64_t pool_columns, uint32_t n_workers, uint32_t* perm);
/* library version string */
const char* nc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NANOCALL_B200_H */
