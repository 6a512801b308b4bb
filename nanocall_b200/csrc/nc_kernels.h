// Kernel argument blocks and launch-side declarations.
#ifndef NC_KERNELS_H
#define NC_KERNELS_H

#include "nc_internal.h"
#include <cuda_runtime.h>

namespace nc {

enum { VIT_THREADS = 512 };

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
    unsigned char* bp_pool;     // gridDim.x slabs of slab_bytes
    size_t slab_bytes;
    float* path_logprob;        // n_jobs
    unsigned short* states;     // packed like the events, may be null
    unsigned char* moves;       // packed like the events, may be null
    float log_2pi;              // (float)log(2*pi)  (Pore_Model.hpp:28)
    float log_n_states;         // logf(4096.f)      (Viterbi.hpp:51)
};

__global__ void viterbi_kernel(const VitArgs a);
size_t viterbi_smem_bytes();

} // namespace nc

#endif
