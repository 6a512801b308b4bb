// Host side of K2: nc_fwbw (parity/debug entry) and nc_train_round_batch = one
// Parameter_Trainer::train_one_round (Parameter_Trainer.hpp:541-579) for a batch of groups.
// The device produces log Pr[data] per sequence, the six posterior sums per event and the three
// transition accumulators per (group, strand); this file finishes the round on the host exactly as the
// reference does: double accumulation over events in order, 3x3 Gaussian elimination with scaled partial
// pivoting (:339-390), var / scale_sd / var_sd (:406-426), exp of the accumulator differences and clamps
// (:516-530).
// A call's groups are cut into waves that fit the E|alpha|beta scratch; two waves are in flight (submit_wave /
// collect_wave): the host builds and queues wave k+1 while wave k's kernels run, statistics come back on a fourth stream.
#include "nc_ctx.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>
#include <thread>

namespace {

// Parameter_Trainer::init (Parameter_Trainer.hpp:30-57): k-mers without self-overlap whose one-step
// neighbours have self-overlap <= 1.  max_self_overlap as Kmer.hpp:81-110.
unsigned max_self_overlap(unsigned i)
{
    for (unsigned k = NC_KMER - 1; k >= 1; --k)
        if ((i & ((1u << (2 * k)) - 1)) == (i >> (2 * (NC_KMER - k)))) return k;
    return 0;
}

int ensure_train_tables(nc_ctx* ctx)
{
    // every entry point that touches the device selects the context's GPU first: the caller's thread may have
    // another one current (several contexts driven from one thread, or a host framework that switched devices)
    NC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->pm_consts_models != ctx->models.size())
    {
        // train_pm_params' per-state constants of the UNSCALED models (Parameter_Trainer.hpp:270-290): sigma^2 and the
        // correctly rounded reciprocals the kernel's three-instruction divisions start from (1.0f / x in IEEE float)
        const size_t nm = ctx->models.size();
        std::vector< float > pc(nm * NC_N_STATES * 8, 0.f);
        for (size_t m = 0; m < nm; ++m)
            for (unsigned j = 0; j < NC_N_STATES; ++j)
            {
                const nc::HostModel& M = ctx->models[m];
                float* q = pc.data() + (m * NC_N_STATES + j) * 8;
                const float sg2 = M.level_stdv[j] * M.level_stdv[j];
                q[0] = M.level_mean[j]; q[1] = sg2; q[2] = 1.0f / sg2; q[3] = M.sd_mean[j];
                q[4] = 1.0f / M.sd_mean[j]; q[5] = M.sd_lambda[j];
            }
        NC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->d_pm_consts) { cudaFree(ctx->d_pm_consts); ctx->d_pm_consts = nullptr; }
        ctx->pm_consts_models = 0;
        NC_CUDA(ctx, cudaMalloc(&ctx->d_pm_consts, std::max< size_t >(1, pc.size()) * sizeof(float)));
        if (!pc.empty()) NC_CUDA(ctx, cudaMemcpy(ctx->d_pm_consts, pc.data(), pc.size() * sizeof(float), cudaMemcpyHostToDevice));
        ctx->pm_consts_models = nm;
    }
    if (ctx->d_logsum_tbl && ctx->d_train_kmers) return NC_OK;
    if (!ctx->d_logsum_tbl)
    {
        // p7_FLogsumInit (logsum.hpp:113-127): double log/exp, stored as float.  Entry 15999 is set to 0: the reference
        // never reads it (its index is reached exactly when max - min >= 15.999f, where p7_FLogsum returns max), and the
        // kernels' p7_FLogsum relies on it to return max + 0 there without a compare (nc_fwbw_core.cuh)
        std::vector< float > tbl(16000);
        for (int i = 0; i < 16000; ++i) tbl[i] = (float)std::log(1. + std::exp((double)-i / 1000.f));
        tbl[15999] = 0.0f;
        NC_CUDA(ctx, cudaMalloc(&ctx->d_logsum_tbl, tbl.size() * sizeof(float)));
        NC_CUDA(ctx, cudaMemcpy(ctx->d_logsum_tbl, tbl.data(), tbl.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (!ctx->d_train_kmers)
    {
        std::vector< unsigned > km;
        for (unsigned i = 0; i < NC_N_STATES; ++i)
        {
            if (max_self_overlap(i) > 0) continue;
            bool all_good = true;
            for (unsigned b1 = 0; b1 < 4; ++b1)
            {
                unsigned j = ((i & 1023u) << 2) + b1;
                if (max_self_overlap(j) > 1) { all_good = false; break; }
            }
            if (all_good) km.push_back(i);
        }
        NC_CUDA(ctx, cudaMalloc(&ctx->d_train_kmers, km.size() * sizeof(unsigned)));
        NC_CUDA(ctx, cudaMemcpy(ctx->d_train_kmers, km.data(), km.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        ctx->n_train_kmers = (unsigned)km.size();
    }
    NC_CUDA(ctx, cudaFuncSetAttribute(nc::fwbw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nc::fwbw_smem_bytes()));
    NC_CUDA(ctx, cudaFuncSetAttribute(nc::pm_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nc::pm_stats_smem_bytes()));
    NC_CUDA(ctx, cudaFuncSetAttribute(nc::st_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nc::st_stats_smem_bytes()));
    return NC_OK;
}

// fn(i) for i in [0, n) on up to n_threads host threads (the per-group host work of a wave: job descriptors with their
// transition LUTs before the launch, the 3x3 solves after it)
template < typename Fn >
void parallel_for(size_t n, unsigned n_threads, Fn fn)
{
    n_threads = (unsigned)std::min< size_t >(std::min(n_threads, 8u), (n + 255) / 256);
    if (n_threads <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::vector< std::thread > th;
    for (unsigned t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            const size_t a = n * t / n_threads, b = n * (t + 1) / n_threads;
            for (size_t i = a; i < b; ++i) fn(i);
        });
    for (auto& x : th) x.join();
}

void fill_job(nc::DevJob& J, int model, const nc_pm_params& pm, const nc_st_params& st)
{
    J.ev_off = 0;
    J.n_events = 0;
    J.model = model;
    J.scale = pm.scale; J.shift = pm.shift; J.drift = pm.drift;
    J.var = pm.var; J.scale_sd = pm.scale_sd; J.var_sd = pm.var_sd;
    nc::host_job_logs(pm, J.log_var, J.log_var_sd);
    nc_transition_lut(st.p_stay, st.p_skip, J.lut);
}

size_t scratch_limit(nc_ctx* ctx)
{
    if (ctx->fb_scratch_limit) return ctx->fb_scratch_limit;
    // (asked once per context: cudaMemGetInfo takes ~13 ms on a 180 GB device with a large pool allocated, which was a
    // quarter of a second over the 20 EM rounds of a batch)
    if (ctx->fb_scratch_auto) return ctx->fb_scratch_auto;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (size_t)1 << 30;
    size_t have = ctx->fb_scratch.cap;
    ctx->fb_scratch_auto = std::min< size_t >((free_b + have) / 2, (size_t)24 << 30);
    return ctx->fb_scratch_auto;
}

// train_pm_params after the inner sums (Parameter_Trainer.hpp:297-427).  rows: per event {s0,s1,s2,l0,l1,l2};
// x/y/t: uncorrected mean, stdv (the 0 -> 0.01 fix of Event.hpp:39-43 is applied here) and start of the same events, in
// (sequence, event) order.
void finish_pm(size_t n_ev, const float* rows, const float* x, const float* y, const float* t, bool train_drift,
               const nc_pm_params& crt, nc_pm_params& out, int& done)
{
    done = 0;
    double A[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    double B[3] = { 0, 0, 0 };
    double D = 0.0, V_numer = 0.0, V_denom = 0.0, U_pos = 0.0;
    for (size_t i = 0; i < n_ev; ++i)
    {
        const float* s = rows + 6 * i;
        const float* l = s + 3;
        float x_i = x[i], y_i = (y[i] == 0.0f) ? 0.01f : y[i], t_i = t[i];
        A[0][0] += s[0];
        A[0][1] += s[1];
        A[1][1] += s[2];
        B[0] += s[0] * x_i;
        B[1] += s[1] * x_i;
        if (train_drift)
        {
            A[0][2] += s[0] * t_i;
            A[1][2] += s[1] * t_i;
            A[2][2] += s[0] * t_i * t_i;
            B[2] += s[0] * x_i * t_i;
        }
        D += s[0] * x_i * x_i;
        V_numer += l[2] * y_i;
        V_denom += l[1];
        U_pos += l[0] / y_i;
    }
    A[1][0] = A[0][1];
    A[2][0] = A[0][2];
    A[2][1] = A[1][2];
    if (!train_drift) A[2][2] = 1.0;
    double Ac[3][3], Bc[3], C[3];
    std::memcpy(Ac, A, sizeof A);
    std::memcpy(Bc, B, sizeof B);
    for (unsigned i = 0; i < 3; ++i)
    {
        C[i] = A[i][0];
        for (unsigned j = 1; j < 3; ++j) if (C[i] < A[i][j]) C[i] = A[i][j];
    }
    for (unsigned i = 0; i < 3; ++i)
    {
        unsigned p = i;
        double p_val = std::abs(A[i][i]) / C[p];
        for (unsigned i2 = i + 1; i2 < 3; ++i2)
        {
            double i2_val = std::abs(A[i2][i]) / C[i2];
            if (i2_val > p_val) { p = i2; p_val = i2_val; }
        }
        if (p_val < 1e-7)
        {
            done = 1;
            out = crt;
            return;
        }
        if (p > i)
        {
            for (unsigned j = 0; j < 3; ++j) std::swap(A[i][j], A[p][j]);
            std::swap(B[i], B[p]);
            std::swap(C[i], C[p]);
        }
        for (p = i + 1; p < 3; ++p)
        {
            double m = A[p][i] / A[i][i];
            A[p][i] = 0;
            for (unsigned j = i + 1; j < 3; ++j) A[p][j] -= m * A[i][j];
            B[p] -= m * B[i];
        }
    }
    // the solution is stored into float members as it is produced (:236-241,406-426)
    float c_hat = (float)(B[2] / A[2][2]);
    float b_hat = (float)((B[1] - A[1][2] * c_hat) / A[1][1]);
    float a_hat = (float)((B[0] - A[0][1] * b_hat - A[0][2] * c_hat) / A[0][0]);
    double d_numer = (D
                      + a_hat * a_hat * Ac[0][0]
                      + b_hat * b_hat * Ac[1][1]
                      + c_hat * c_hat * Ac[2][2]
                      + 2.0 * a_hat * b_hat * Ac[0][1]
                      + 2.0 * a_hat * c_hat * Ac[0][2]
                      + 2.0 * b_hat * c_hat * Ac[1][2]
                      - 2.0 * (a_hat * Bc[0] + b_hat * Bc[1] + c_hat * Bc[2]));
    float d_hat = (float)std::sqrt(d_numer / (double)n_ev);
    float v_hat = (float)(V_numer / V_denom);
    float u_hat = (float)((double)n_ev / (U_pos - V_denom / v_hat));
    out.shift = a_hat; out.scale = b_hat; out.drift = c_hat; out.var = d_hat; out.scale_sd = v_hat; out.var_sd = u_hat;
}

// Parameter_Trainer.hpp:516-530
nc_st_params finish_st(const float* acc /* denom, stay, skip */)
{
    nc_st_params r;
    r.p_stay = std::exp(acc[1] - acc[0]);
    r.p_skip = std::exp(acc[2] - acc[0]);
    if (r.p_stay < .05 || r.p_stay > .4 || r.p_skip < .05 || r.p_skip > .4)
    {
        float a_stay = std::max(r.p_stay, .05f);
        a_stay = std::min(a_stay, .4f);
        float a_skip = std::max(r.p_skip, .05f);
        a_skip = std::min(a_skip, .4f);
        r.p_stay = a_stay;
        r.p_skip = a_skip;
    }
    return r;
}

struct Wave
{
    std::vector< nc::FbSeq > seqs;
    std::vector< nc::FbGroup > groups;
    std::vector< nc::DevJob > jobs;
    size_t scratch_floats = 0;
    size_t n_events = 0;
    unsigned max_len = 0;
};

// Waves are pipelined two deep: while the kernels of wave k run, the host builds and queues wave k+1, and only then waits
// for wave k's statistics and finishes its groups (the 3x3 solves).  submit_wave queues a wave on ctx->stream --
// descriptors up from the slot's page-locked images, emission + fwbw (+ statistics kernels) -- and, on ctx->stream4
// behind the last kernel, the statistics back into the slot's page-locked buffers; nothing waits.  collect_wave waits for
// that copy.  The E|alpha|beta scratch is shared by both slots (stream order keeps wave k+1's kernels behind wave k's);
// everything a wave returns lives in its slot.  Event arrays are already on the device.
int submit_wave(nc_ctx* ctx, const Wave& w, TrainSlot& T, const float* d_mean, const float* d_stdv, const float* d_start,
                const float* d_lstd, bool pm_stats, bool st_stats, int& launches)
{
    int rc;
    const unsigned ns = (unsigned)w.seqs.size(), ng = (unsigned)w.groups.size();
    const size_t b_seqs = ns * sizeof(nc::FbSeq), b_groups = ng * sizeof(nc::FbGroup), b_jobs = w.jobs.size() * sizeof(nc::DevJob);
    const size_t b_lz = ns * sizeof(float), b_pm = w.n_events * 6 * sizeof(float), b_st = (size_t)ng * 6 * sizeof(float);
    if ((rc = dev_reserve(ctx, ctx->fb_scratch, w.scratch_floats * sizeof(float))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_seqs, b_seqs)) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_groups, std::max< size_t >(1, b_groups))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_jobs, b_jobs)) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_lz, b_lz)) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_pm, std::max< size_t >(1, b_pm))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_st, std::max< size_t >(1, b_st))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, T.d_counter, 2 * sizeof(unsigned))) != NC_OK) return rc;
    if ((rc = pin_reserve(ctx, T.h_seqs, b_seqs)) != NC_OK) return rc;
    if ((rc = pin_reserve(ctx, T.h_groups, std::max< size_t >(1, b_groups))) != NC_OK) return rc;
    if ((rc = pin_reserve(ctx, T.h_jobs, b_jobs)) != NC_OK) return rc;
    if ((rc = pin_reserve(ctx, T.h_lz, b_lz)) != NC_OK) return rc;
    if ((rc = pin_reserve(ctx, T.h_pm, std::max< size_t >(1, b_pm))) != NC_OK) return rc;
    if ((rc = pin_reserve(ctx, T.h_st, std::max< size_t >(1, b_st))) != NC_OK) return rc;
    for (int k = 0; k < 5; ++k)
        if (!T.evk[k]) NC_CUDA(ctx, cudaEventCreate(&T.evk[k]));
    if (!T.done) NC_CUDA(ctx, cudaEventCreateWithFlags(&T.done, cudaEventDisableTiming));
    std::memcpy(T.h_seqs.p, w.seqs.data(), b_seqs);
    if (ng) std::memcpy(T.h_groups.p, w.groups.data(), b_groups);
    std::memcpy(T.h_jobs.p, w.jobs.data(), b_jobs);
    cudaStream_t s = ctx->stream;
    NC_CUDA(ctx, cudaMemcpyAsync(T.d_seqs.p, T.h_seqs.p, b_seqs, cudaMemcpyHostToDevice, s));
    if (ng) NC_CUDA(ctx, cudaMemcpyAsync(T.d_groups.p, T.h_groups.p, b_groups, cudaMemcpyHostToDevice, s));
    NC_CUDA(ctx, cudaMemcpyAsync(T.d_jobs.p, T.h_jobs.p, b_jobs, cudaMemcpyHostToDevice, s));
    NC_CUDA(ctx, cudaMemsetAsync(T.d_counter.p, 0, 2 * sizeof(unsigned), s));

    nc::FbArgs a;
    a.jobs = (const nc::DevJob*)T.d_jobs.p;
    a.seqs = (const nc::FbSeq*)T.d_seqs.p;
    a.groups = (const nc::FbGroup*)T.d_groups.p;
    a.n_seqs = ns;
    a.n_groups = ng;
    a.next_item = (unsigned*)T.d_counter.p;
    a.models = ctx->d_models;
    a.pm_consts = reinterpret_cast< const float4* >(ctx->d_pm_consts);
    a.mean = d_mean; a.stdv = d_stdv; a.start = d_start; a.log_stdv = d_lstd;
    a.logsum_tbl = ctx->d_logsum_tbl;
    a.train_kmers = ctx->d_train_kmers;
    a.n_train_kmers = ctx->n_train_kmers;
    a.scratch = (float*)ctx->fb_scratch.p;
    a.log_pr_data = (float*)T.d_lz.p;
    a.pm_stats = (float*)T.d_pm.p;
    a.st_stats = (float*)T.d_st.p;
    a.log_2pi = (float)std::log(2.0 * M_PI);
    a.log_n_states = std::log((float)NC_N_STATES);

    const unsigned tiles = (w.max_len + nc::FB_EV_TILE - 1) / nc::FB_EV_TILE;
    NC_CUDA(ctx, cudaEventRecord(T.evk[0], s));
    nc::emission_kernel<<< dim3((w.max_len + nc::EM_TILE - 1) / nc::EM_TILE, ns), 512, 0, s >>>(a);
    NC_CUDA(ctx, cudaGetLastError());
    NC_CUDA(ctx, cudaEventRecord(T.evk[1], s));
    const unsigned grid = std::min< unsigned >(ns, 2u * (unsigned)ctx->prop.multiProcessorCount);
    nc::fwbw_kernel<<< grid, 512, nc::fwbw_smem_bytes(), s >>>(a);
    NC_CUDA(ctx, cudaGetLastError());
    launches = 2;
    bool any_generic = false;
    for (const auto& q : w.seqs) any_generic = any_generic || q.generic;
    if (any_generic)
    {
        // sequences under the custom default transition table (nc_ctx_set_default_transitions): stored-list kernel
        nc::GenTrans gt = { (const unsigned*)ctx->gen_from_off.p, (const unsigned*)ctx->gen_from_idx.p, (const float*)ctx->gen_from_lp.p,
                            (const unsigned*)ctx->gen_to_off.p, (const unsigned*)ctx->gen_to_idx.p, (const float*)ctx->gen_to_lp.p };
        nc::fwbw_generic_kernel<<< grid, 512, nc::fwbw_generic_smem_bytes(), s >>>(a, gt);
        NC_CUDA(ctx, cudaGetLastError());
        ++launches;
    }
    NC_CUDA(ctx, cudaEventRecord(T.evk[2], s));
    if (pm_stats)
    {
        nc::pm_stats_kernel<<< dim3(tiles, ns), 512, nc::pm_stats_smem_bytes(), s >>>(a);
        NC_CUDA(ctx, cudaGetLastError());
        ++launches;
    }
    NC_CUDA(ctx, cudaEventRecord(T.evk[3], s));
    if (st_stats && ng)
    {
        nc::st_stats_kernel<<< dim3(ng, 2), nc::st_stats_threads(), nc::st_stats_smem_bytes(), s >>>(a);
        NC_CUDA(ctx, cudaGetLastError());
        ++launches;
    }
    NC_CUDA(ctx, cudaEventRecord(T.evk[4], s));
    cudaStream_t c = ctx->stream4;
    NC_CUDA(ctx, cudaStreamWaitEvent(c, T.evk[4], 0));
    NC_CUDA(ctx, cudaMemcpyAsync(T.h_lz.p, T.d_lz.p, b_lz, cudaMemcpyDeviceToHost, c));
    if (pm_stats && b_pm) NC_CUDA(ctx, cudaMemcpyAsync(T.h_pm.p, T.d_pm.p, b_pm, cudaMemcpyDeviceToHost, c));
    if (st_stats && ng) NC_CUDA(ctx, cudaMemcpyAsync(T.h_st.p, T.d_st.p, b_st, cudaMemcpyDeviceToHost, c));
    NC_CUDA(ctx, cudaEventRecord(T.done, c));
    return NC_OK;
}

// Wait for a submitted wave; its statistics are then in T.h_lz / T.h_pm / T.h_st.  ctx->last_kernel_ms = its device time.
int collect_wave(nc_ctx* ctx, const Wave& w, TrainSlot& T, int launches)
{
    NC_CUDA(ctx, cudaEventSynchronize(T.done));
    float ms = 0.f;
    NC_CUDA(ctx, cudaEventElapsedTime(&ms, T.evk[0], T.evk[4]));
    ctx->last_kernel_ms = ms;
    for (int k = 0; k < 4; ++k)
    {
        float kms = 0.f;
        NC_CUDA(ctx, cudaEventElapsedTime(&kms, T.evk[k], T.evk[k + 1]));
        ctx->train_ms[k] += kms;
    }
    ctx->train_events += (double)w.n_events;
    ctx->train_launches += launches;
    ctx->train_waves += 1;
    return NC_OK;
}

// after a failure with work in flight: nothing may still be running against the context's buffers when the call returns
void drain_train(nc_ctx* ctx)
{
    cudaStreamSynchronize(ctx->stream);
    if (ctx->stream4) cudaStreamSynchronize(ctx->stream4);
}

int upload_events(nc_ctx* ctx, size_t total, const float* mean, const float* stdv, const float* start)
{
    int rc;
    if ((rc = dev_reserve(ctx, ctx->fb_mean, total * sizeof(float))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, ctx->fb_stdv, total * sizeof(float))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, ctx->fb_start, total * sizeof(float))) != NC_OK) return rc;
    NC_CUDA(ctx, cudaMemcpy(ctx->fb_mean.p, mean, total * sizeof(float), cudaMemcpyHostToDevice));
    NC_CUDA(ctx, cudaMemcpy(ctx->fb_stdv.p, stdv, total * sizeof(float), cudaMemcpyHostToDevice));
    NC_CUDA(ctx, cudaMemcpy(ctx->fb_start.p, start, total * sizeof(float), cudaMemcpyHostToDevice));
    return NC_OK;
}

} // namespace

extern "C" {

// NC_TRAIN_TIMING=1: where the host time of nc_train_round_batch goes (printed when the context is destroyed)
double g_train_t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
void nc_train_timing_report()
{
    if (!std::getenv("NC_TRAIN_TIMING")) return;
    std::fprintf(stderr, "nc_train_round_batch host time: upload %.3f s, scratch_limit %.3f, wave build %.3f, fill_job %.3f, waiting for a wave %.3f, finish %.3f, calls %.0f\n",
                 g_train_t[6], g_train_t[7], g_train_t[1], g_train_t[2], g_train_t[3], g_train_t[4], g_train_t[5]);
}


int nc_fwbw(nc_ctx* ctx, int32_t model_id, const nc_pm_params* pm, const nc_st_params* st,
            uint32_t n_events, const float* mean, const float* stdv, const float* start,
            float* alpha, float* beta, float* log_pr_data)
{
    if (!ctx) return NC_ERR_ARG;
    if (!pm || !st || !mean || !stdv || !start || n_events == 0) NC_FAIL(ctx, NC_ERR_ARG, "nc_fwbw: bad argument");
    if (model_id < 0 || model_id >= (int)ctx->models.size()) NC_FAIL(ctx, NC_ERR_ARG, "nc_fwbw: unknown model id %d", model_id);
    int rc;
    if ((rc = ensure_train_tables(ctx)) != NC_OK) return rc;
    if ((rc = upload_events(ctx, n_events, mean, stdv, start)) != NC_OK) return rc;
    Wave w;
    w.jobs.resize(1);
    fill_job(w.jobs[0], model_id, *pm, *st);
    nc::FbSeq q;
    std::memset(&q, 0, sizeof q);
    q.n_events = n_events;
    q.generic = ctx->gen_on && st->p_stay == ctx->gen_default.p_stay && st->p_skip == ctx->gen_default.p_skip;
    w.seqs.push_back(q);
    w.scratch_floats = (size_t)3 * n_events * NC_N_STATES;
    w.n_events = n_events;
    w.max_len = n_events;
    int launches = 0;
    TrainSlot& T = ctx->tslot[0];
    rc = submit_wave(ctx, w, T, (const float*)ctx->fb_mean.p, (const float*)ctx->fb_stdv.p, (const float*)ctx->fb_start.p,
                     nullptr, false, false, launches);
    if (rc == NC_OK) rc = collect_wave(ctx, w, T, launches);
    if (rc != NC_OK) { drain_train(ctx); return rc; }
    const size_t cells = (size_t)n_events * NC_N_STATES;
    const float* sc = (const float*)ctx->fb_scratch.p;
    if (alpha) NC_CUDA(ctx, cudaMemcpy(alpha, sc + cells, cells * sizeof(float), cudaMemcpyDeviceToHost));
    if (beta) NC_CUDA(ctx, cudaMemcpy(beta, sc + 2 * cells, cells * sizeof(float), cudaMemcpyDeviceToHost));
    if (log_pr_data) *log_pr_data = static_cast< const float* >(T.h_lz.p)[0];
    return NC_OK;
}

int nc_train_round_batch(nc_ctx* ctx, uint32_t n_groups, const uint32_t* seq_off,
                         const uint64_t* ev_off, const uint8_t* seq_strand,
                         const float* mean, const float* stdv, const float* start,
                         const nc_train_in* in, const nc_train_opts* opts, nc_train_out* out)
{
    if (!ctx) return NC_ERR_ARG;
    if (n_groups == 0) return NC_OK;
    if (!seq_off || !ev_off || !seq_strand || !mean || !stdv || !start || !in || !opts || !out)
        NC_FAIL(ctx, NC_ERR_ARG, "nc_train_round_batch: NULL argument");
    int rc;
    if ((rc = ensure_train_tables(ctx)) != NC_OK) return rc;
    const uint32_t n_seqs = seq_off[n_groups];
    for (uint32_t g = 0; g < n_groups; ++g)
    {
        if (seq_off[g + 1] <= seq_off[g] || seq_off[g + 1] - seq_off[g] > NC_MAX_TRAIN_SEQS)
            NC_FAIL(ctx, NC_ERR_ARG, "nc_train_round_batch: group %u has %u sequences (1..%u allowed)", g,
                    seq_off[g + 1] - seq_off[g], NC_MAX_TRAIN_SEQS);
        for (int st = 0; st < 2; ++st)
            if (in[g].model_id[st] < 0 || in[g].model_id[st] >= (int)ctx->models.size())
                NC_FAIL(ctx, NC_ERR_ARG, "nc_train_round_batch: group %u: unknown model id %d", g, in[g].model_id[st]);
    }
    for (uint32_t s = 0; s < n_seqs; ++s)
    {
        if (ev_off[s + 1] <= ev_off[s]) NC_FAIL(ctx, NC_ERR_ARG, "nc_train_round_batch: sequence %u has no events", s);
        if (seq_strand[s] > 1) NC_FAIL(ctx, NC_ERR_ARG, "nc_train_round_batch: sequence %u: strand must be 0 or 1", s);
    }
    if (ctx->n_train_kmers > nc::st_stats_max_kmers())   // st_stats_kernel: a fixed number of training k-mers per producer thread (2160 for 6-mers)
        NC_FAIL(ctx, NC_ERR_STATE, "nc_train_round_batch: %u training k-mers exceed the kernel's %u", ctx->n_train_kmers, nc::st_stats_max_kmers());
    const auto tt0 = std::chrono::steady_clock::now();
    auto lap = [&](int k, std::chrono::steady_clock::time_point& from) {
        const auto now = std::chrono::steady_clock::now();
        g_train_t[k] += std::chrono::duration< double >(now - from).count();
        from = now;
    };
    auto tl = tt0;
    g_train_t[5] += 1;
    const uint64_t base = ev_off[0];
    const size_t total = ev_off[n_seqs] - base;
    if ((rc = upload_events(ctx, total, mean + base, stdv + base, start + base)) != NC_OK) return rc;
    lap(6, tl);
    const size_t limit_floats = scratch_limit(ctx) / sizeof(float);
    lap(7, tl);
    if (!ctx->stream4) NC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream4, cudaStreamNonBlocking));

    // ---- a wave: consecutive groups whose slabs fit the scratch pool
    auto build_wave = [&](Wave& w, uint32_t g0) -> uint32_t {
        w = Wave();
        uint32_t g1 = g0;
        while (g1 < n_groups)
        {
            size_t need = 0;
            for (uint32_t s = seq_off[g1]; s < seq_off[g1 + 1]; ++s) need += (size_t)3 * (ev_off[s + 1] - ev_off[s]) * NC_N_STATES;
            if (g1 > g0 && (w.scratch_floats + need > limit_floats || w.seqs.size() + NC_MAX_TRAIN_SEQS > 60000)) break;
            nc::FbGroup G;
            G.seq_begin = (unsigned)w.seqs.size();
            for (int st = 0; st < 2; ++st)
            {
                // Parameter_Trainer.hpp:443-444
                G.log_p_stay[st] = std::log(in[g1].st[st].p_stay);
                G.log_p_step_4[st] = (float)(std::log(1.0 - in[g1].st[st].p_stay - in[g1].st[st].p_skip) - std::log(4.0));
                w.jobs.push_back(nc::DevJob());   // filled below, in parallel
            }
            for (uint32_t s = seq_off[g1]; s < seq_off[g1 + 1]; ++s)
            {
                nc::FbSeq q;
                std::memset(&q, 0, sizeof q);
                q.ev_off = ev_off[s] - base;
                q.n_events = (unsigned)(ev_off[s + 1] - ev_off[s]);
                q.slab = w.scratch_floats;
                q.ev_out = w.n_events;
                q.job = 2 * (g1 - g0) + seq_strand[s];
                q.strand = seq_strand[s];
                // State_Transition_Parameters::is_default() (Parameter_Trainer.hpp:118-131): the custom table, if any
                q.generic = ctx->gen_on && in[g1].st[seq_strand[s]].p_stay == ctx->gen_default.p_stay
                    && in[g1].st[seq_strand[s]].p_skip == ctx->gen_default.p_skip;
                w.scratch_floats += (size_t)3 * q.n_events * NC_N_STATES;
                w.n_events += q.n_events;
                w.max_len = std::max(w.max_len, q.n_events);
                w.seqs.push_back(q);
            }
            G.seq_end = (unsigned)w.seqs.size();
            w.groups.push_back(G);
            ++g1;
        }
        lap(1, tl);
        parallel_for(g1 - g0, ctx->host_threads, [&](size_t k) {
            for (int st = 0; st < 2; ++st) fill_job(w.jobs[2 * k + st], in[g0 + k].model_id[st], in[g0 + k].pm, in[g0 + k].st[st]);
        });
        lap(2, tl);
        return g1;
    };
    // ---- finish every group of a collected wave on the host (train_one_round, :541-579)
    auto finish_wave = [&](const Wave& w, const TrainSlot& T, uint32_t g0, uint32_t g1) {
        const float* lz = static_cast< const float* >(T.h_lz.p);
        const float* pm_rows = static_cast< const float* >(T.h_pm.p);
        const float* st_acc = static_cast< const float* >(T.h_st.p);
        parallel_for(g1 - g0, ctx->host_threads, [&](size_t gk)
        {
            const uint32_t g = g0 + (uint32_t)gk;
            const nc::FbGroup& G = w.groups[g - g0];
            nc_train_out& o = out[g];
            float fit = 0.0f;
            for (unsigned q = G.seq_begin; q < G.seq_end; ++q) fit += lz[q];
            o.fit = fit;
            o.pm = in[g].pm;
            o.st[0] = in[g].st[0];
            o.st[1] = in[g].st[1];
            o.done = 0;
            if (opts->train_scaling)
            {
                const nc::FbSeq& first = w.seqs[G.seq_begin];
                size_t n_ev = 0;
                for (unsigned q = G.seq_begin; q < G.seq_end; ++q) n_ev += w.seqs[q].n_events;
                // the group's events are contiguous in (sequence, event) order both in pm_rows and in the inputs
                int done = 0;
                finish_pm(n_ev, pm_rows + 6 * first.ev_out, mean + base + first.ev_off, stdv + base + first.ev_off,
                          start + base + first.ev_off, opts->train_drift != 0, in[g].pm, o.pm, done);
                o.done = done;
                if (done) return;  // new_st_params = crt_st_params (:566-570)
            }
            if (opts->train_transitions)
            {
                o.st[0] = finish_st(st_acc + (size_t)(g - g0) * 6);
                o.st[1] = finish_st(st_acc + (size_t)(g - g0) * 6 + 3);
            }
        });
        lap(4, tl);
    };

    // ---- two waves in flight: wave k+1 is built and queued while wave k's kernels run, then wave k is collected and finished
    Wave wv[2];
    uint32_t wg0[2] = { 0, 0 }, wg1[2] = { 0, 0 };
    int wl[2] = { 0, 0 };
    float kernel_ms = 0.f;   // device time of the call = sum over its waves
    auto submit = [&](int k) -> int {
        return submit_wave(ctx, wv[k], ctx->tslot[k], (const float*)ctx->fb_mean.p, (const float*)ctx->fb_stdv.p,
                           (const float*)ctx->fb_start.p, nullptr, opts->train_scaling != 0, opts->train_transitions != 0, wl[k]);
    };
    int cur = 0;
    wg0[0] = 0;
    wg1[0] = build_wave(wv[0], 0);
    if ((rc = submit(0)) != NC_OK) { drain_train(ctx); return rc; }
    for (;;)
    {
        const int nxt = cur ^ 1;
        const bool more = wg1[cur] < n_groups;
        if (more)
        {
            wg0[nxt] = wg1[cur];
            wg1[nxt] = build_wave(wv[nxt], wg0[nxt]);
            if ((rc = submit(nxt)) != NC_OK) { drain_train(ctx); return rc; }
        }
        if ((rc = collect_wave(ctx, wv[cur], ctx->tslot[cur], wl[cur])) != NC_OK) { drain_train(ctx); return rc; }
        kernel_ms += ctx->last_kernel_ms;
        lap(3, tl);
        finish_wave(wv[cur], ctx->tslot[cur], wg0[cur], wg1[cur]);
        if (!more) break;
        cur = nxt;
    }
    ctx->last_kernel_ms = kernel_ms;
    return NC_OK;
}

int nc_ctx_reserve(nc_ctx* ctx, uint64_t train_events, uint64_t viterbi_events)
{
    if (!ctx) return NC_ERR_ARG;
    NC_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (train_events)
    {
        if ((rc = ensure_train_tables(ctx)) != NC_OK) return rc;
        const size_t want = std::min< size_t >(scratch_limit(ctx), (size_t)train_events * 3 * NC_N_STATES * sizeof(float));
        if ((rc = dev_reserve(ctx, ctx->fb_scratch, want)) != NC_OK) return rc;
        // (the statistics buffers of the two wave slots are sized by the waves themselves: a wave is at most
        // scratch_limit / 48 KB events)
        for (DevBuf* b : { &ctx->fb_mean, &ctx->fb_stdv, &ctx->fb_start })
            if ((rc = dev_reserve(ctx, *b, (size_t)train_events * sizeof(float))) != NC_OK) return rc;
    }
    if (viterbi_events)
    {
        for (DevBuf* b : { &ctx->mean, &ctx->stdv, &ctx->start })
            if ((rc = dev_reserve(ctx, *b, (size_t)viterbi_events * sizeof(float))) != NC_OK) return rc;
        if ((rc = dev_reserve(ctx, ctx->states, (size_t)viterbi_events * sizeof(uint16_t))) != NC_OK) return rc;
        if ((rc = dev_reserve(ctx, ctx->moves, (size_t)viterbi_events)) != NC_OK) return rc;
    }
    return NC_OK;
}

int nc_ctx_train_stats(nc_ctx* ctx, double* out8, int reset)
{
    if (!ctx || !out8) return NC_ERR_ARG;
    for (int k = 0; k < 4; ++k) out8[k] = ctx->train_ms[k];
    out8[4] = ctx->train_events;
    out8[5] = ctx->train_launches;
    out8[6] = ctx->train_waves;
    out8[7] = 0;
    if (reset)
    {
        for (int k = 0; k < 4; ++k) ctx->train_ms[k] = 0;
        ctx->train_events = ctx->train_launches = ctx->train_waves = 0;
    }
    return NC_OK;
}

} // extern "C"
