// Internal declarations shared by the host translation units and the CUDA ones.
#ifndef NC_INTERNAL_H
#define NC_INTERNAL_H

#include "nanocall_b200.h"

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace nc {

// Unscaled model on the host, SoA (Pore_Model_State fields the emission reads, Pore_Model.hpp:79-96)
struct HostModel
{
    std::vector< float > level_mean, level_stdv, sd_mean, sd_lambda, log_level_stdv, log_sd_lambda;
    float mean, stdv;
    int strand;
};

void host_model_prepare(const float* table, HostModel& m);
void host_job_logs(const nc_pm_params& p, float& log_var, float& log_var_sd);
void host_event_logs(size_t n, const float* stdv, float* log_stdv, unsigned n_threads);

// Device-side view of one unscaled model: 6 arrays of 4096 floats, contiguous
// [level_mean | level_stdv | sd_mean | sd_lambda | log_level_stdv | log_sd_lambda]
enum { MODEL_FLOATS = 6 * NC_N_STATES };

// One decoding job as the kernels see it (built on the host, uploaded once per batch)
struct DevJob
{
    unsigned long long ev_off;  // first event in the packed arrays
    unsigned n_events;
    int model;                  // index into the context's model table
    float scale, shift, drift, var, scale_sd, var_sd;
    float log_var, log_var_sd;  // host logf of the scalars (Pore_Model.hpp:193-195)
    float lut[64];              // transition log-weights by overlap mask (nc_transition_lut)
};

} // namespace nc

#endif
