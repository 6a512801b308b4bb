// Context object and error plumbing shared by the ABI translation units.
#ifndef NC_CTX_H
#define NC_CTX_H

#include "nc_kernels.h"

#include <cstdio>
#include <string>
#include <vector>

extern thread_local std::string g_create_error;


struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
};

// grow-only page-locked host buffer
struct PinBuf
{
    void* p = nullptr;
    size_t cap = 0;
};

// One of the two training waves in flight (nc_train.cu): page-locked images of the wave's descriptors (uploaded
// asynchronously), its statistics on the device and page-locked on the host, and the events around its kernels.
struct TrainSlot
{
    PinBuf h_seqs, h_groups, h_jobs, h_lz, h_pm, h_st;
    DevBuf d_seqs, d_groups, d_jobs, d_counter, d_lz, d_pm, d_st;
    cudaEvent_t evk[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };   // around the four training kernels
    cudaEvent_t done = nullptr;                                              // the statistics are on the host
};

struct nc_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // backpointer-kernel launches that run next to the alpha kernel
    cudaEvent_t ev2 = nullptr;
    cudaStream_t stream3 = nullptr;   // streamed event upload of host-memory calls
    cudaEvent_t ev3 = nullptr;
    unsigned long long* d_landed = nullptr;
    unsigned long long* h_landed = nullptr;   // pinned: the values the copy stream writes into d_landed
    static constexpr int LANDED_SLOTS = 1024;
    std::vector< unsigned char > colalloc_image;   // host image of the alpha kernel's column allocator (one free extent)
    uint64_t stream_in_min_events = (uint64_t)8 << 20;   // calls with fewer events copy everything before the launch
    uint64_t stream_in_chunk = (uint64_t)4 << 20;        // events per chunk of the streamed upload
    cudaDeviceProp prop;
    std::vector< nc::HostModel > models;
    float* d_models = nullptr;
    int d_models_cap = 0;
    unsigned char* d_bp = nullptr;
    size_t bp_bytes = 0;
    float* d_logsum_tbl = nullptr;
    std::string err;
    // grow-only scratch
    DevBuf jobs, order, counter, path, mean, stdv, start, lstd, states, moves, tb, cl_col0;
    uint32_t cluster_min_events = 2000;   // calls with at most n_sms/2 jobs, each at least this long, give every job a CTA pair
    bool cluster_on = false;              // NC_VIT_CLUSTER=1 switches the cluster kernel on: measured slower than one CTA per
                                          // job (55 ms against 37 ms for a 60 k-event read), see nc_viterbi_alpha.cu
    DevBuf fb_scratch, fb_mean, fb_stdv, fb_start, fb_lstd;
    TrainSlot tslot[2];
    cudaStream_t stream4 = nullptr;   // statistics of a training wave back to the host while the next wave computes
    // custom default transition table (nc_ctx_set_default_transitions): in force for jobs / strands whose transition
    // parameters equal gen_default
    bool gen_on = false;
    nc_st_params gen_default{ 0.f, 0.f };
    DevBuf gen_from_off, gen_from_idx, gen_from_lp, gen_to_off, gen_to_idx, gen_to_lp, gen_bp, gen_order, gen_counter;
    float* d_pm_consts = nullptr;     // pm_stats_kernel's per-state constants of every registered model (FbArgs::pm_consts)
    size_t pm_consts_models = 0;      // how many models the table covers
    unsigned* d_train_kmers = nullptr;
    unsigned n_train_kmers = 0;
    size_t fb_scratch_limit = 0;  // bytes of E|alpha|beta slabs per wave (0 = pick from free memory)
    size_t fb_scratch_auto = 0;    // the limit derived from the free memory, once
    unsigned host_threads = 1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double train_ms[4] = { 0, 0, 0, 0 };   // emission, fwbw, pm_stats, st_stats: device time since the last reset
    double train_events = 0, train_launches = 0, train_waves = 0;
    float last_kernel_ms = 0.f;
    int last_launches = 0;        // kernels launched by the most recent nc_viterbi_packed
    unsigned long long* d_stats = nullptr;   // 8 counters of the alpha kernel (nc_ctx_viterbi_stats)
    unsigned* d_abort = nullptr;             // abort word of the alpha kernel's persistent grid (VitArgs::abort_word)
    unsigned* h_abort = nullptr;             // pinned copy read back after every launch
    double wait_limit_s = 120.0;             // NC_WAIT_LIMIT_S
    int vit_mode = 0;             // 0 = auto, 2 = backpointer kernel only (nc_ctx_set_viterbi_mode)
};

#define NC_FAIL(ctx, code, ...)                                        \
    do {                                                               \
        char _b[2048];                                                 \
        std::snprintf(_b, sizeof _b, __VA_ARGS__);                     \
        (ctx)->err = _b;                                               \
        return (code);                                                 \
    } while (0)

#define NC_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            NC_FAIL(ctx, NC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline int dev_reserve(nc_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return NC_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        NC_FAIL(ctx, NC_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return NC_OK;
}
inline int pin_reserve(nc_ctx* ctx, PinBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return NC_OK;
    if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        NC_FAIL(ctx, NC_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return NC_OK;
}
inline void pin_free(PinBuf& b)
{
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
}
inline void dev_free(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}


#endif
