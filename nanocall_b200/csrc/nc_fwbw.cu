// K2 (Forward/Backward + trainer statistics): under construction; the entry points exist so the
// ABI is complete, and fail loudly rather than fall back to anything.
#include "nc_kernels.h"
#include <string>

extern "C" {
int nc_fwbw(nc_ctx*, int32_t, const nc_pm_params*, const nc_st_params*, uint32_t, const float*, const float*,
            const float*, float*, float*, float*)
{
    return NC_ERR_STATE;
}
int nc_train_round_batch(nc_ctx*, uint32_t, const uint32_t*, const uint64_t*, const uint8_t*, const float*,
                         const float*, const float*, const nc_train_in*, const nc_train_opts*, nc_train_out*)
{
    return NC_ERR_STATE;
}
}
