// C ABI of nanocall_b200 (include/nanocall_b200.h): context, model registry, batch dispatch.
// No CPU fallback lives here: without a CUDA device nc_ctx_create fails and nothing else runs.
#include "nc_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <thread>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "nc_ctx.h"

thread_local std::string g_create_error;

extern "C" {

int nc_ctx_create(int device, size_t bp_pool_bytes, nc_ctx** out)
{
    if (!out) { g_create_error = "nc_ctx_create: out is NULL"; return NC_ERR_ARG; }
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
    {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (nanocall_b200 has no CPU fallback)";
        cudaGetLastError();
        return NC_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) { g_create_error = "nc_ctx_create: bad device index"; return NC_ERR_ARG; }
    nc_ctx* ctx = new nc_ctx();
    ctx->device = device;
    auto fail = [&](const char* what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        nc_ctx_destroy(ctx);
        return NC_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if (bp_pool_bytes == 0)
    {
        size_t free_b = 0, total_b = 0;
        if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) return fail("cudaMemGetInfo", e);
        // 3/4 of the free memory, at most 140 GB: the Forward/Backward scratch (<= 24 GB) and the event arrays of a
        // call come on top.  Long reads hold their alpha columns for their whole forward pass (a 150 k-event read:
        // 2.4 GB for ~85 ms), so a large pool is what keeps the other forward CTAs supplied next to them.
        bp_pool_bytes = std::min< size_t >(free_b / 4 * 3, (size_t)140 << 30);
    }
    bp_pool_bytes &= ~(size_t)4095;
    if ((e = cudaMalloc(&ctx->d_bp, bp_pool_bytes)) != cudaSuccess) return fail("cudaMalloc(backpointer pool)", e);
    ctx->bp_bytes = bp_pool_bytes;
    if ((e = cudaFuncSetAttribute(nc::viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)nc::viterbi_smem_bytes())) != cudaSuccess)
        return fail("cudaFuncSetAttribute(viterbi_kernel)", e);
    if ((e = cudaFuncSetAttribute(nc::viterbi_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)nc::viterbi_alpha_smem_bytes())) != cudaSuccess)
        return fail("cudaFuncSetAttribute(viterbi_alpha_kernel)", e);
    if ((e = cudaFuncSetAttribute(nc::viterbi_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)nc::viterbi_alpha_smem_bytes())) != cudaSuccess)
        return fail("cudaFuncSetAttribute(viterbi_cluster_kernel)", e);
    if ((e = cudaMalloc(&ctx->d_stats, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc(stats)", e);
    if ((e = cudaMemset(ctx->d_stats, 0, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMemset(stats)", e);
    if ((e = cudaMalloc(&ctx->d_abort, sizeof(unsigned))) != cudaSuccess) return fail("cudaMalloc(abort)", e);
    if ((e = cudaHostAlloc(&ctx->h_abort, sizeof(unsigned), cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc(abort)", e);
    if (const char* v = std::getenv("NC_WAIT_LIMIT_S")) ctx->wait_limit_s = std::max(1e-3, std::atof(v));
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev2, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream4, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev3, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_landed, sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc(landed)", e);
    if ((e = cudaHostAlloc(&ctx->h_landed, nc_ctx::LANDED_SLOTS * sizeof(unsigned long long), cudaHostAllocDefault)) != cudaSuccess)
        return fail("cudaHostAlloc(landed)", e);
    // streamed event upload (nc_viterbi_packed, host memory): thresholds, overridable for tests
    if (const char* v = std::getenv("NC_VIT_CLUSTER")) ctx->cluster_on = std::atoi(v) != 0;
    if (const char* v = std::getenv("NC_VIT_CLUSTER_MIN_EVENTS")) ctx->cluster_min_events = (uint32_t)std::strtoul(v, nullptr, 10);
    if (const char* v = std::getenv("NC_STREAM_IN_MIN_EVENTS")) ctx->stream_in_min_events = std::strtoull(v, nullptr, 10);
    if (const char* v = std::getenv("NC_STREAM_IN_CHUNK")) ctx->stream_in_chunk = std::max< uint64_t >(32, std::strtoull(v, nullptr, 10));
    unsigned hc = std::thread::hardware_concurrency();
    ctx->host_threads = hc ? std::min(hc, 32u) : 4u;
    *out = ctx;
    return NC_OK;
}

void nc_train_timing_report();
void nc_ctx_destroy(nc_ctx* ctx)
{
    nc_train_timing_report();
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : { &ctx->jobs, &ctx->order, &ctx->counter, &ctx->path, &ctx->mean, &ctx->stdv, &ctx->start,
                       &ctx->lstd, &ctx->states, &ctx->moves, &ctx->tb, &ctx->cl_col0, &ctx->fb_scratch, &ctx->fb_mean, &ctx->fb_stdv, &ctx->fb_start, &ctx->fb_lstd,
                       &ctx->gen_from_off, &ctx->gen_from_idx, &ctx->gen_from_lp, &ctx->gen_to_off, &ctx->gen_to_idx, &ctx->gen_to_lp,
                       &ctx->gen_bp, &ctx->gen_order, &ctx->gen_counter })
        dev_free(*b);
    if (ctx->d_models) cudaFree(ctx->d_models);
    if (ctx->d_bp) cudaFree(ctx->d_bp);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->d_abort) cudaFree(ctx->d_abort);
    if (ctx->h_abort) cudaFreeHost(ctx->h_abort);
    if (ctx->d_logsum_tbl) cudaFree(ctx->d_logsum_tbl);
    if (ctx->d_train_kmers) cudaFree(ctx->d_train_kmers);
    if (ctx->d_pm_consts) cudaFree(ctx->d_pm_consts);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->stream4) { cudaStreamSynchronize(ctx->stream4); cudaStreamDestroy(ctx->stream4); }
    for (TrainSlot& t : ctx->tslot)
    {
        for (DevBuf* b : { &t.d_seqs, &t.d_groups, &t.d_jobs, &t.d_counter, &t.d_lz, &t.d_pm, &t.d_st }) dev_free(*b);
        for (PinBuf* b : { &t.h_seqs, &t.h_groups, &t.h_jobs, &t.h_lz, &t.h_pm, &t.h_st }) pin_free(*b);
        for (cudaEvent_t e : t.evk) if (e) cudaEventDestroy(e);
        if (t.done) cudaEventDestroy(t.done);
    }
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream3) { cudaStreamSynchronize(ctx->stream3); cudaStreamDestroy(ctx->stream3); }
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->d_landed) cudaFree(ctx->d_landed);
    if (ctx->h_landed) cudaFreeHost(ctx->h_landed);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* nc_last_error(const nc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* nc_ctx_stream(nc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

float nc_ctx_last_kernel_ms(nc_ctx* ctx) { return ctx ? ctx->last_kernel_ms : -1.f; }

int nc_ctx_last_launches(nc_ctx* ctx) { return ctx ? ctx->last_launches : -1; }

int nc_ctx_set_viterbi_mode(nc_ctx* ctx, int mode)
{
    if (!ctx) return NC_ERR_ARG;
    if (mode != NC_VIT_AUTO && mode != NC_VIT_BACKPOINTER) NC_FAIL(ctx, NC_ERR_ARG, "nc_ctx_set_viterbi_mode: bad mode %d", mode);
    ctx->vit_mode = mode;
    return NC_OK;
}

int nc_ctx_viterbi_stats(nc_ctx* ctx, uint64_t* out8, int reset)
{
    if (!ctx || !out8) return NC_ERR_ARG;
    NC_CUDA(ctx, cudaSetDevice(ctx->device));
    NC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NC_CUDA(ctx, cudaMemcpy(out8, ctx->d_stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) NC_CUDA(ctx, cudaMemset(ctx->d_stats, 0, 8 * sizeof(uint64_t)));
    return NC_OK;
}

int nc_ctx_sync(nc_ctx* ctx)
{
    if (!ctx) return NC_ERR_ARG;
    NC_CUDA(ctx, cudaSetDevice(ctx->device));
    NC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NC_OK;
}

int nc_ctx_device_info(nc_ctx* ctx, int* n_sms, size_t* total_mem, char* name, int name_cap)
{
    if (!ctx) return NC_ERR_ARG;
    if (n_sms) *n_sms = ctx->prop.multiProcessorCount;
    if (total_mem) *total_mem = ctx->prop.totalGlobalMem;
    if (name && name_cap > 0) { std::strncpy(name, ctx->prop.name, name_cap - 1); name[name_cap - 1] = 0; }
    return NC_OK;
}

int nc_model_register(nc_ctx* ctx, const float* table, int strand, int* model_id)
{
    if (!ctx) return NC_ERR_ARG;
    if (!table || !model_id || strand < 0 || strand > 2) NC_FAIL(ctx, NC_ERR_ARG, "nc_model_register: bad argument");
    NC_CUDA(ctx, cudaSetDevice(ctx->device));
    nc::HostModel m;
    nc::host_model_prepare(table, m);
    m.strand = strand;
    int id = (int)ctx->models.size();
    if (id + 1 > ctx->d_models_cap)
    {
        int cap = std::max(8, ctx->d_models_cap * 2);
        float* nd = nullptr;
        NC_CUDA(ctx, cudaMalloc(&nd, (size_t)cap * nc::MODEL_FLOATS * sizeof(float)));
        if (ctx->d_models)
        {
            NC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            NC_CUDA(ctx, cudaMemcpy(nd, ctx->d_models, (size_t)id * nc::MODEL_FLOATS * sizeof(float), cudaMemcpyDeviceToDevice));
            cudaFree(ctx->d_models);
        }
        ctx->d_models = nd;
        ctx->d_models_cap = cap;
    }
    std::vector< float > blob(nc::MODEL_FLOATS);
    const std::vector< float >* parts[6] = { &m.level_mean, &m.level_stdv, &m.sd_mean, &m.sd_lambda, &m.log_level_stdv, &m.log_sd_lambda };
    for (int a = 0; a < 6; ++a) std::memcpy(blob.data() + (size_t)a * NC_N_STATES, parts[a]->data(), NC_N_STATES * sizeof(float));
    NC_CUDA(ctx, cudaMemcpy(ctx->d_models + (size_t)id * nc::MODEL_FLOATS, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->models.push_back(std::move(m));
    *model_id = id;
    return NC_OK;
}

int nc_model_stats(nc_ctx* ctx, int model_id, float* mean, float* stdv)
{
    if (!ctx) return NC_ERR_ARG;
    if (model_id < 0 || model_id >= (int)ctx->models.size()) NC_FAIL(ctx, NC_ERR_ARG, "nc_model_stats: unknown model id %d", model_id);
    if (mean) *mean = ctx->models[model_id].mean;
    if (stdv) *stdv = ctx->models[model_id].stdv;
    return NC_OK;
}

int nc_ctx_set_default_transitions(nc_ctx* ctx, float p_stay_default, float p_skip_default, uint32_t n_edges,
                                   const uint16_t* from, const uint16_t* to, const float* logp)
{
    if (!ctx) return NC_ERR_ARG;
    NC_CUDA(ctx, cudaSetDevice(ctx->device));
    NC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_edges == 0) { ctx->gen_on = false; return NC_OK; }
    if (!from || !to || !logp) NC_FAIL(ctx, NC_ERR_ARG, "nc_ctx_set_default_transitions: NULL argument");
    for (uint32_t k = 0; k < n_edges; ++k)
        if (from[k] >= NC_N_STATES || to[k] >= NC_N_STATES) NC_FAIL(ctx, NC_ERR_ARG, "nc_ctx_set_default_transitions: edge %u: state out of range", k);
    // to lists: the edges of a source in file order (to_v as State_Transitions::operator>> fills it, :237-252);
    // from lists: built source by source, ascending (update_fields, :79-99)
    std::vector< unsigned > to_off(NC_N_STATES + 1, 0), from_off(NC_N_STATES + 1, 0), to_idx(n_edges), from_idx(n_edges);
    std::vector< float > to_lp(n_edges), from_lp(n_edges);
    for (uint32_t k = 0; k < n_edges; ++k) { ++to_off[from[k] + 1]; ++from_off[to[k] + 1]; }
    for (unsigned i = 0; i < NC_N_STATES; ++i) { to_off[i + 1] += to_off[i]; from_off[i + 1] += from_off[i]; }
    {
        std::vector< unsigned > fill(to_off.begin(), to_off.end() - 1);
        for (uint32_t k = 0; k < n_edges; ++k) { const unsigned at = fill[from[k]]++; to_idx[at] = to[k]; to_lp[at] = logp[k]; }
        std::vector< unsigned > ffill(from_off.begin(), from_off.end() - 1);
        for (unsigned i = 0; i < NC_N_STATES; ++i)
            for (unsigned e = to_off[i]; e < to_off[i + 1]; ++e) { const unsigned at = ffill[to_idx[e]]++; from_idx[at] = i; from_lp[at] = to_lp[e]; }
    }
    int rc;
    struct Up { DevBuf* b; const void* src; size_t bytes; };
    const Up ups[] = { { &ctx->gen_from_off, from_off.data(), from_off.size() * 4 }, { &ctx->gen_from_idx, from_idx.data(), from_idx.size() * 4 },
                       { &ctx->gen_from_lp, from_lp.data(), from_lp.size() * 4 }, { &ctx->gen_to_off, to_off.data(), to_off.size() * 4 },
                       { &ctx->gen_to_idx, to_idx.data(), to_idx.size() * 4 }, { &ctx->gen_to_lp, to_lp.data(), to_lp.size() * 4 } };
    for (const Up& u : ups)
    {
        if ((rc = dev_reserve(ctx, *u.b, u.bytes)) != NC_OK) return rc;
        NC_CUDA(ctx, cudaMemcpy(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice));
    }
    NC_CUDA(ctx, cudaFuncSetAttribute(nc::viterbi_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nc::viterbi_generic_smem_bytes()));
    NC_CUDA(ctx, cudaFuncSetAttribute(nc::fwbw_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nc::fwbw_generic_smem_bytes()));
    ctx->gen_default.p_stay = p_stay_default;
    ctx->gen_default.p_skip = p_skip_default;
    ctx->gen_on = true;
    return NC_OK;
}

int nc_viterbi_packed(nc_ctx* ctx, uint32_t n_jobs, const uint64_t* ev_off,
                      const float* mean, const float* stdv, const float* start, const float* log_stdv,
                      const int32_t* model_id, const nc_pm_params* pm, const nc_st_params* st,
                      nc_mem mem, float* path_logprob, uint16_t* states, uint8_t* moves)
{
    if (!ctx) return NC_ERR_ARG;
    if (n_jobs == 0) return NC_OK;
    if (!ev_off || !mean || !stdv || !start || !model_id || !pm || !st || !path_logprob)
        NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_packed: NULL argument");
    if (moves && !states && mem == NC_MEM_DEVICE)
        NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_packed: moves need states for device-resident outputs");
    if (ctx->models.empty()) NC_FAIL(ctx, NC_ERR_STATE, "nc_viterbi_packed: no model registered");
    NC_CUDA(ctx, cudaSetDevice(ctx->device));

    // ---- job descriptors (host): scalars, logs, transition LUTs; longest-first order
    const uint64_t total = ev_off[n_jobs] - ev_off[0];
    std::vector< nc::DevJob > jobs(n_jobs);
    uint32_t max_len = 0;
    {
        float lut_key[2] = { -1.f, -1.f };
        float lut_val[64];
        for (uint32_t k = 0; k < n_jobs; ++k)
        {
            if (ev_off[k + 1] <= ev_off[k]) NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_packed: job %u has no events", k);
            uint64_t len = ev_off[k + 1] - ev_off[k];
            if (len > 0xffffffffull) NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_packed: job %u too long", k);
            if (model_id[k] < 0 || model_id[k] >= (int)ctx->models.size())
                NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_packed: job %u: unknown model id %d", k, model_id[k]);
            nc::DevJob& J = jobs[k];
            J.ev_off = ev_off[k] - ev_off[0];
            J.n_events = (unsigned)len;
            J.model = model_id[k];
            J.scale = pm[k].scale; J.shift = pm[k].shift; J.drift = pm[k].drift;
            J.var = pm[k].var; J.scale_sd = pm[k].scale_sd; J.var_sd = pm[k].var_sd;
            nc::host_job_logs(pm[k], J.log_var, J.log_var_sd);
            if (st[k].p_stay != lut_key[0] || st[k].p_skip != lut_key[1])
            {
                nc_transition_lut(st[k].p_stay, st[k].p_skip, lut_val);
                lut_key[0] = st[k].p_stay;
                lut_key[1] = st[k].p_skip;
            }
            std::memcpy(J.lut, lut_val, sizeof lut_val);
            max_len = std::max(max_len, J.n_events);
        }
    }
    // ---- which kernel decodes which job.  The alpha-column kernel is the fast path; it needs 16 KiB of scratch per
    // event of every job between the start of its forward pass and the end of its traceback, taken from a device-wide
    // allocator over the pool (nc_viterbi_alpha.cu), so a read of any length up to the pool takes it.  The
    // backpointer kernel needs 4 KiB per event and one slab per CTA; it serves the reads that exceed the alpha
    // pool (and NC_VIT_BACKPOINTER).  When both classes exist the two kernels run concurrently on disjoint sets of
    // SMs (every CTA of either kernel fills an SM) and disjoint parts of the pool.  A call that wants no states
    // (candidate ranking by path probability) needs no scratch at all.
    const bool want_path = states != nullptr || moves != nullptr;
    const size_t n_sms = (size_t)ctx->prop.multiProcessorCount;
    const size_t a_col = (size_t)NC_N_STATES * sizeof(float), b_col = (size_t)NC_N_STATES;
    // traceback service CTAs of the alpha kernel: one per ~48 forward CTAs (16 jobs in flight each; measured busy
    // 22 % of the time at one per 36)
    auto tb_ctas_for = [&](size_t fwd) { return want_path ? std::max< size_t >(1, std::min< size_t >(4, (fwd + 24) / 48)) : (size_t)0; };
    auto fwd_ctas_in = [&](size_t ctas, size_t n_alpha_jobs) {   // forward CTAs when `ctas` SMs serve n_alpha_jobs jobs
        size_t fwd = std::min< size_t >(n_alpha_jobs, ctas);
        fwd = std::min< size_t >(fwd, nc::viterbi_alpha_max_forward_ctas());   // the allocator's extent list holds them all
        while (fwd > 1 && fwd + tb_ctas_for(fwd) > ctas) --fwd;
        return fwd;
    };
    // jobs whose transition parameters are the defaults while a custom table is in force (--trans) go to the
    // list-walking kernel; the rest of this function deals with the others (n_jobs_p of them)
    std::vector< unsigned > order, order_g;
    order.reserve(n_jobs);
    for (uint32_t k = 0; k < n_jobs; ++k)
    {
        const bool gen = ctx->gen_on && st[k].p_stay == ctx->gen_default.p_stay && st[k].p_skip == ctx->gen_default.p_skip;
        (gen ? order_g : order).push_back(k);
    }
    const uint32_t n_jobs_all = n_jobs;
    n_jobs = (uint32_t)order.size();
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return jobs[a].n_events > jobs[b].n_events; });
    std::stable_sort(order_g.begin(), order_g.end(), [&](unsigned a, unsigned b) { return jobs[a].n_events > jobs[b].n_events; });
    // order = [long jobs (backpointer form) | the rest (alpha form)], each longest first
    uint32_t alpha_max_len = 0xffffffffu;
    if (want_path) alpha_max_len = (uint32_t)std::min< size_t >(ctx->bp_bytes / a_col, 0x7fffffffu);
    if (ctx->vit_mode == NC_VIT_BACKPOINTER) alpha_max_len = 0;   // forced backpointer form (tests, A/B measurements)
    uint32_t n_long = 0;
    while (n_long < n_jobs && jobs[order[n_long]].n_events > alpha_max_len) ++n_long;
    if (n_jobs) max_len = jobs[order[0]].n_events;   // (of the jobs this part handles)
    // ---- few jobs: a call with at most one job per two SMs would leave SMs idle, and a long read is 0.6 us per event
    // on one CTA: every job then gets a cluster of two CTAs (viterbi_cluster_kernel) and a fixed extent of the pool
    bool use_cluster = false;
    std::vector< unsigned > cl_col0;
    if (want_path && ctx->cluster_on && ctx->vit_mode != NC_VIT_BACKPOINTER && order_g.empty() && n_jobs > 0
        && 2 * (size_t)n_jobs <= n_sms && jobs[order[n_jobs - 1]].n_events >= ctx->cluster_min_events)
    {
        size_t cols = 0;
        for (uint32_t k = 0; k < n_jobs; ++k)
        {
            cl_col0.push_back((unsigned)cols);
            cols += jobs[order[k]].n_events;
        }
        use_cluster = cols * a_col <= ctx->bp_bytes && cols < 0xffffffffull;
    }
    unsigned grid_b = 0, fwd_a = 0, tb_a = 0;
    size_t slab_b = 0, slab_a = 0, pool_b = 0;
    if (use_cluster) n_long = 0;
    for (; !use_cluster;)
    {
        const uint32_t n_short = n_jobs - n_long;
        grid_b = fwd_a = tb_a = 0;
        slab_b = slab_a = pool_b = 0;
        size_t sms_b = 0;
        if (n_long)
        {
            slab_b = (size_t)max_len * b_col;
            const size_t fit = ctx->bp_bytes / slab_b;
            if (fit == 0)
                NC_FAIL(ctx, NC_ERR_NOMEM, "nc_viterbi_packed: a %u-event job needs %zu backpointer bytes, pool has %zu",
                        max_len, slab_b, ctx->bp_bytes);
            sms_b = std::min< size_t >(std::min< size_t >(n_long, n_sms), fit);
            if (n_short)
            {
                // share the SMs by estimated time: the backpointer form costs ~1.85x per event
                double ev_long = 0, ev_short = 0;
                for (uint32_t k = 0; k < n_jobs; ++k) (k < n_long ? ev_long : ev_short) += jobs[order[k]].n_events;
                const double share = 1.85 * ev_long / (1.85 * ev_long + ev_short);
                size_t want = (size_t)(share * (double)n_sms + 0.5);
                want = std::max< size_t >(1, std::min< size_t >(want, n_sms > 2 ? n_sms - 2 : 1));
                sms_b = std::min(sms_b, want);
            }
            grid_b = (unsigned)sms_b;
            pool_b = (size_t)grid_b * slab_b;
        }
        if (n_short)
        {
            const size_t fwd = fwd_ctas_in(n_sms - sms_b, n_short);
            fwd_a = (unsigned)fwd;
            tb_a = (unsigned)tb_ctas_for(fwd);
            if (want_path)
            {
                slab_a = (ctx->bp_bytes - pool_b) & ~(a_col - 1);   // the alpha pool: what the backpointer slabs leave
                // the longest alpha job must fit it: otherwise it joins the backpointer class
                if ((size_t)jobs[order[n_long]].n_events * a_col > slab_a) { ++n_long; continue; }
            }
        }
        break;
    }
    const uint32_t n_short = use_cluster ? 0u : n_jobs - n_long;

    // ---- dispatch order of the alpha class (nc_plan_dispatch_order, nc_host.cpp): longest-first whenever the pool
    // allows, shorter jobs while the long reads pin most of it
    if (want_path && n_short > fwd_a && fwd_a > 0)
    {
        std::vector< uint32_t > lens(n_short), perm(n_short);
        for (uint32_t k = 0; k < n_short; ++k) lens[k] = jobs[order[n_long + k]].n_events;
        if (nc_plan_dispatch_order(n_short, lens.data(), (uint64_t)(slab_a / a_col), fwd_a, perm.data()))
        {
            std::vector< unsigned > seq(n_short);
            for (uint32_t k = 0; k < n_short; ++k) seq[k] = order[n_long + perm[k]];
            std::copy(seq.begin(), seq.end(), order.begin() + n_long);
        }
    }

    int rc;
    if ((rc = dev_reserve(ctx, ctx->jobs, n_jobs_all * sizeof(nc::DevJob))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, ctx->order, std::max< size_t >(1, n_jobs) * sizeof(unsigned))) != NC_OK) return rc;
    // the list-walking kernel's share: its job order, a work counter, and one slab of 16-bit backpointers per CTA
    unsigned grid_g = 0;
    size_t slab_g = 0;
    if (!order_g.empty())
    {
        slab_g = (size_t)jobs[order_g[0]].n_events * NC_N_STATES * sizeof(unsigned short);
        grid_g = (unsigned)std::min< size_t >(order_g.size(), 2 * n_sms);
        while (grid_g > 1 && (size_t)grid_g * slab_g > ((size_t)16 << 30)) --grid_g;
        if ((rc = dev_reserve(ctx, ctx->gen_order, order_g.size() * sizeof(unsigned))) != NC_OK) return rc;
        if ((rc = dev_reserve(ctx, ctx->gen_counter, sizeof(unsigned))) != NC_OK) return rc;
        if (want_path && (rc = dev_reserve(ctx, ctx->gen_bp, (size_t)grid_g * slab_g)) != NC_OK) return rc;
    }
    if ((rc = dev_reserve(ctx, ctx->counter, 2 * sizeof(unsigned))) != NC_OK) return rc;
    if (use_cluster && (rc = dev_reserve(ctx, ctx->cl_col0, n_jobs * sizeof(unsigned))) != NC_OK) return rc;
    if ((rc = dev_reserve(ctx, ctx->path, n_jobs_all * sizeof(float))) != NC_OK) return rc;
    // alpha kernel control block: tickets | tail, head | release counter of every forward CTA | column allocator.
    // Every allocation of the call happens before its first asynchronous operation (a cudaFree inside dev_reserve is a
    // device-wide synchronisation, and an allocation failure must not leave copies or kernels in flight).
    const size_t tk_bytes = (size_t)n_short * sizeof(nc::TbTicket);
    const size_t ctl_bytes = ((2 + 2 * (size_t)fwd_a + 80) * sizeof(unsigned) + 15) & ~(size_t)15;   // + phase words (diagnostics)
    const size_t ca_bytes = nc::viterbi_alpha_colalloc_bytes();
    if (n_short && want_path && (rc = dev_reserve(ctx, ctx->tb, tk_bytes + ctl_bytes + ca_bytes)) != NC_OK) return rc;
    const uint64_t base = ev_off[0];
    bool stream_in = false;
    if (mem == NC_MEM_HOST)
    {
        if ((rc = dev_reserve(ctx, ctx->mean, total * sizeof(float))) != NC_OK) return rc;
        if ((rc = dev_reserve(ctx, ctx->stdv, total * sizeof(float))) != NC_OK) return rc;
        if ((rc = dev_reserve(ctx, ctx->start, total * sizeof(float))) != NC_OK) return rc;
        if (log_stdv && (rc = dev_reserve(ctx, ctx->lstd, total * sizeof(float))) != NC_OK) return rc;
        if ((states || moves) && (rc = dev_reserve(ctx, ctx->states, total * sizeof(uint16_t))) != NC_OK) return rc;
        if (moves && (rc = dev_reserve(ctx, ctx->moves, total)) != NC_OK) return rc;
        // Big batches in PINNED memory: the kernels start at once and the event arrays follow on a third stream in
        // chunks.  From pageable memory cudaMemcpyAsync is staged and blocks the host, so nothing would overlap
        // (and the chunking would only add small copies): such calls take the plain copy path.
        stream_in = total >= ctx->stream_in_min_events && order_g.empty() && !use_cluster;   // (the list-walking and the cluster kernel do not poll the landed counter)
        if (stream_in)
        {
            cudaPointerAttributes at;
            for (const float* p : { mean + base, stdv + base, start + base })
                if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeHost) stream_in = false;
            cudaGetLastError();
        }
    }
    // from here on work is in flight: a failure synchronises the context's streams before it returns, so that the
    // caller's buffers and the context's scratch are quiescent when the error is reported
#define NC_CUDA_INFLIGHT(call)                                                                                        \
    do {                                                                                                              \
        cudaError_t _e = (call);                                                                                      \
        if (_e != cudaSuccess)                                                                                        \
        {                                                                                                             \
            cudaStreamSynchronize(ctx->stream3); cudaStreamSynchronize(ctx->stream2); cudaStreamSynchronize(ctx->stream); \
            cudaGetLastError();                                                                                       \
            NC_FAIL(ctx, NC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);      \
        }                                                                                                             \
    } while (0)
    cudaStream_t s = ctx->stream;
    NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->jobs.p, jobs.data(), n_jobs_all * sizeof(nc::DevJob), cudaMemcpyHostToDevice, s));
    if (n_jobs) NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->order.p, order.data(), n_jobs * sizeof(unsigned), cudaMemcpyHostToDevice, s));
    NC_CUDA_INFLIGHT(cudaMemsetAsync(ctx->counter.p, 0, 2 * sizeof(unsigned), s));

    nc::VitArgs a;
    a.jobs = (const nc::DevJob*)ctx->jobs.p;
    a.order = (const unsigned*)ctx->order.p;
    a.n_jobs = n_long;
    a.next_job = (unsigned*)ctx->counter.p;
    a.models = ctx->d_models;
    a.bp_pool = ctx->d_bp;
    a.slab_bytes = slab_b;
    a.path_logprob = (float*)ctx->path.p;
    a.log_2pi = (float)std::log(2.0 * M_PI);
    a.log_n_states = std::log((float)NC_N_STATES);
    a.n_fwd = 0;
    a.n_tb = 0;
    a.tickets = nullptr;
    a.tb_tail = a.tb_head = a.slab_free = nullptr;
    a.colalloc = nullptr;
    a.job_col0 = nullptr;
    a.stats = ctx->d_stats;
    a.abort_word = ctx->d_abort;
    a.wait_limit = (long long)(ctx->wait_limit_s * 1e9);   // nanoseconds of back-off
    NC_CUDA_INFLIGHT(cudaMemsetAsync(ctx->d_abort, 0, sizeof(unsigned), s));

    a.landed = nullptr;
    a.ev_total = total;
    if (mem == NC_MEM_HOST)
    {
        const float* lsp = log_stdv ? log_stdv + base : nullptr;  // NULL: the kernel derives it (nc_logf)
        // streamed input: a job waits (wait_events_landed) until the copy engine has delivered its events
        if (stream_in)
        {
            NC_CUDA_INFLIGHT(cudaMemsetAsync(ctx->d_landed, 0, sizeof(unsigned long long), s));
            NC_CUDA_INFLIGHT(cudaEventRecord(ctx->ev3, s));
        }
        else
        {
            NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->mean.p, mean + base, total * sizeof(float), cudaMemcpyHostToDevice, s));
            NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->stdv.p, stdv + base, total * sizeof(float), cudaMemcpyHostToDevice, s));
            NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->start.p, start + base, total * sizeof(float), cudaMemcpyHostToDevice, s));
            if (lsp) NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->lstd.p, lsp, total * sizeof(float), cudaMemcpyHostToDevice, s));
        }
        if (stream_in) a.landed = ctx->d_landed;
        a.mean = (const float*)ctx->mean.p;
        a.stdv = (const float*)ctx->stdv.p;
        a.start = (const float*)ctx->start.p;
        a.log_stdv = lsp ? (const float*)ctx->lstd.p : nullptr;
        a.states = nullptr;
        a.moves = nullptr;
        if (states || moves) a.states = (unsigned short*)ctx->states.p;
        if (moves) a.moves = (unsigned char*)ctx->moves.p;
    }
    else
    {
        a.mean = mean + base;
        a.stdv = stdv + base;
        a.start = start + base;
        a.log_stdv = log_stdv ? log_stdv + base : nullptr;
        a.states = states ? states + base : nullptr;
        a.moves = moves ? moves + base : nullptr;
    }

    if (stream_in)
    {
        // Queue the chunked upload on the copy stream BEFORE the kernels are launched: the copies are asynchronous
        // (pinned memory) so the launch follows at once and the kernels overlap them, and nothing can deadlock when
        // launches are serialised (profilers, CUDA_LAUNCH_BLOCKING, pageable buffers): the data is then simply
        // there first.  Chunks are multiples of 32 events (whole 128-byte lines).
        cudaStream_t sc = ctx->stream3;
        NC_CUDA_INFLIGHT(cudaStreamWaitEvent(sc, ctx->ev3, 0));
        const float* lsp = log_stdv ? log_stdv + base : nullptr;
        uint64_t chunk = std::max< uint64_t >(ctx->stream_in_chunk, (total + nc_ctx::LANDED_SLOTS - 1) / nc_ctx::LANDED_SLOTS);
        chunk = (chunk + 31) & ~(uint64_t)31;
        int slot = 0;
        cudaError_t ce = cudaSuccess;
        for (uint64_t c0 = 0; c0 < total && ce == cudaSuccess; c0 += chunk, ++slot)
        {
            const uint64_t c1 = std::min(total, c0 + chunk), nb = (c1 - c0) * sizeof(float);
            ce = cudaMemcpyAsync((float*)ctx->mean.p + c0, mean + base + c0, nb, cudaMemcpyHostToDevice, sc);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync((float*)ctx->stdv.p + c0, stdv + base + c0, nb, cudaMemcpyHostToDevice, sc);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync((float*)ctx->start.p + c0, start + base + c0, nb, cudaMemcpyHostToDevice, sc);
            if (ce == cudaSuccess && lsp) ce = cudaMemcpyAsync((float*)ctx->lstd.p + c0, lsp + c0, nb, cudaMemcpyHostToDevice, sc);
            ctx->h_landed[slot] = c1;
            if (ce == cudaSuccess)
                ce = cudaMemcpyAsync(ctx->d_landed, ctx->h_landed + slot, sizeof(unsigned long long), cudaMemcpyHostToDevice, sc);
        }
        if (ce != cudaSuccess)
        {
            cudaStreamSynchronize(sc);
            cudaStreamSynchronize(s);
            NC_FAIL(ctx, NC_ERR_CUDA, "nc_viterbi_packed: streamed event upload failed: %s", cudaGetErrorString(ce));
        }
    }
    NC_CUDA_INFLIGHT(cudaEventRecord(ctx->ev0, s));
    ctx->last_launches = 0;
    if (!order_g.empty())
    {
        nc::VitArgs gkn = a;
        gkn.order = (const unsigned*)ctx->gen_order.p;
        gkn.n_jobs = (unsigned)order_g.size();
        gkn.next_job = (unsigned*)ctx->gen_counter.p;
        gkn.bp_pool = (unsigned char*)ctx->gen_bp.p;
        gkn.slab_bytes = slab_g;
        gkn.landed = nullptr;
        nc::GenTrans gt = { (const unsigned*)ctx->gen_from_off.p, (const unsigned*)ctx->gen_from_idx.p, (const float*)ctx->gen_from_lp.p,
                            (const unsigned*)ctx->gen_to_off.p, (const unsigned*)ctx->gen_to_idx.p, (const float*)ctx->gen_to_lp.p };
        NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->gen_order.p, order_g.data(), order_g.size() * sizeof(unsigned), cudaMemcpyHostToDevice, s));
        NC_CUDA_INFLIGHT(cudaMemsetAsync(ctx->gen_counter.p, 0, sizeof(unsigned), s));
        nc::viterbi_generic_kernel<<< grid_g, 512, nc::viterbi_generic_smem_bytes(), s >>>(gkn, gt);
        NC_CUDA_INFLIGHT(cudaGetLastError());
        ++ctx->last_launches;
    }
    if (use_cluster)
    {
        nc::VitArgs c = a;
        c.n_jobs = n_jobs;
        c.job_col0 = (const unsigned*)ctx->cl_col0.p;
        NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->cl_col0.p, cl_col0.data(), n_jobs * sizeof(unsigned), cudaMemcpyHostToDevice, s));
        nc::viterbi_cluster_kernel<<< 2 * n_jobs, nc::VIT_THREADS / 2, nc::viterbi_alpha_smem_bytes(), s >>>(c);
        NC_CUDA_INFLIGHT(cudaGetLastError());
        ++ctx->last_launches;
    }
    if (n_long)
    {
        // on the second stream, so it runs next to the alpha kernel (grid_b + fwd_a + tb_a <= number of SMs)
        cudaStream_t sb = n_short ? ctx->stream2 : s;
        if (n_short) NC_CUDA_INFLIGHT(cudaStreamWaitEvent(sb, ctx->ev0, 0));
        nc::viterbi_kernel<<< grid_b, nc::VIT_THREADS, nc::viterbi_smem_bytes(), sb >>>(a);
        NC_CUDA_INFLIGHT(cudaGetLastError());
        if (n_short) NC_CUDA_INFLIGHT(cudaEventRecord(ctx->ev2, sb));
        ++ctx->last_launches;
    }
    if (n_short)
    {
        nc::VitArgs b = a;
        b.order = a.order + n_long;
        b.n_jobs = n_short;
        b.next_job = a.next_job + 1;
        b.bp_pool = a.bp_pool + pool_b;
        b.slab_bytes = slab_a;
        b.n_fwd = fwd_a;
        b.n_tb = tb_a;
        if (want_path)
        {
            // tickets, counters (zeroed) and the column allocator (one free extent)
            NC_CUDA_INFLIGHT(cudaMemsetAsync(ctx->tb.p, 0, tk_bytes + ctl_bytes, s));
            ctx->colalloc_image.resize(ca_bytes);
            nc::viterbi_alpha_colalloc_init(ctx->colalloc_image.data(), (unsigned)(slab_a / a_col));
            NC_CUDA_INFLIGHT(cudaMemcpyAsync((char*)ctx->tb.p + tk_bytes + ctl_bytes, ctx->colalloc_image.data(), ca_bytes, cudaMemcpyHostToDevice, s));
            b.tickets = (nc::TbTicket*)ctx->tb.p;
            b.tb_tail = (unsigned*)((char*)ctx->tb.p + tk_bytes);
            b.tb_head = b.tb_tail + 1;
            b.slab_free = b.tb_tail + 2;
            b.colalloc = (char*)ctx->tb.p + tk_bytes + ctl_bytes;
        }
        // Cooperative launch: forward CTAs and traceback service CTAs wait for each other, so the whole grid must be
        // resident at once.  The driver guarantees that for a cooperative grid and REFUSES the launch (an error, not a
        // hang) when it cannot: too many CTAs for the SMs this context can use.  SMs busy with other work (the
        // backpointer kernel of this call, another context's kernels) only delay the start.
        {
            void* kargs[] = { (void*)&b };
            NC_CUDA_INFLIGHT(cudaLaunchCooperativeKernel((const void*)nc::viterbi_alpha_kernel, dim3(fwd_a + tb_a), dim3(nc::VIT_THREADS),
                                                         kargs, nc::viterbi_alpha_smem_bytes(), s));
        }
        ++ctx->last_launches;
        if (n_long) NC_CUDA_INFLIGHT(cudaStreamWaitEvent(s, ctx->ev2, 0));
    }
    NC_CUDA_INFLIGHT(cudaEventRecord(ctx->ev1, s));

    // what the alpha grid looks like (diagnostics of a stalled or aborted grid): counters, forward CTAs and service warps by
    // what they are doing (note_phase), tickets, the column allocator's free list.  Copies go through the copy stream so
    // the snapshot can be taken while the kernel is still running.
    auto alpha_state = [&]() -> std::string {
        if (!(n_short && want_path)) return std::string();
        cudaStream_t sc = ctx->stream3;
        const size_t n_ctl = 2 + 2 * (size_t)fwd_a + 80;
        std::vector< unsigned char > img(ca_bytes);
        std::vector< unsigned > ctl(n_ctl, 0u);
        std::vector< nc::TbTicket > tks(n_short);
        unsigned started[2] = { 0, 0 };
        cudaMemcpyAsync(img.data(), (char*)ctx->tb.p + tk_bytes + ctl_bytes, ca_bytes, cudaMemcpyDeviceToHost, sc);
        cudaMemcpyAsync(ctl.data(), (char*)ctx->tb.p + tk_bytes, n_ctl * sizeof(unsigned), cudaMemcpyDeviceToHost, sc);
        cudaMemcpyAsync(tks.data(), ctx->tb.p, tk_bytes, cudaMemcpyDeviceToHost, sc);
        cudaMemcpyAsync(started, ctx->counter.p, sizeof started, cudaMemcpyDeviceToHost, sc);
        cudaStreamSynchronize(sc);
        cudaGetLastError();
        const unsigned* w = reinterpret_cast< const unsigned* >(img.data());   // next_ticket, now_serving, n_free, max_free, start[], len[]
        const unsigned nf = std::min(w[2], 1024u);
        unsigned long long free_cols = 0;
        unsigned largest = 0;
        for (unsigned k = 0; k < nf; ++k) { free_cols += w[4 + 1024 + k]; largest = std::max(largest, w[4 + 1024 + k]); }
        const unsigned* rel = ctl.data() + 2;
        const unsigned* fph = rel + fwd_a;
        const unsigned* sph = fph + fwd_a;
        unsigned by_phase[16] = { 0 }, max_behind = 0, sum_done = 0, sum_rel = 0;
        for (unsigned k = 0; k < fwd_a; ++k)
        {
            const unsigned ph = fph[k] & 15u, done = fph[k] >> 4;
            ++by_phase[ph];
            sum_done += done;
            sum_rel += rel[k];
            if (done > rel[k]) max_behind = std::max(max_behind, done - rel[k]);
        }
        unsigned ready_below = 0, ready_above = 0, first_unready = n_short;
        for (unsigned k = 0; k < n_short; ++k)
        {
            if (tks[k].ready) ++(k < ctl[0] ? ready_below : ready_above);
            else if (k < first_unready) first_unready = k;
        }
        unsigned sv_phase[8] = { 0 }, sv_min = 0xffffffffu, sv_max = 0;
        for (unsigned k = 0; k < 16 * tb_a; ++k) { const unsigned v = sph[k]; ++sv_phase[v & 7u]; sv_min = std::min(sv_min, v >> 4); sv_max = std::max(sv_max, v >> 4); }
        char d[1800];
        std::snprintf(d, sizeof d, "alpha jobs %u (longest %u events), %u forward + %u service CTAs, handed out %u, tickets published %u, "
                      "claimed by the traceback service %u, released %u; forward CTAs: %u not started, %u waiting for a release, %u waiting for columns, "
                      "%u in the forward pass, %u publishing, %u waiting for input, %u done, %u gave up waiting for a release, %u gave up waiting for columns; "
                      "their finished jobs %u, most unreleased jobs of one CTA %u; pool %zu columns, free %llu in %u extents (largest %u, bound %u), lock waiters %u; "
                      "tickets ready below the published count %u, above it %u, first not ready %u; service warps: %u idle, %u waiting for a ticket, "
                      "%u tracing, %u releasing columns, %u released, %u out of tickets, %u gave up (tickets %u..%u), %u releases after the abort",
                      n_short, jobs[order[n_long]].n_events, fwd_a, tb_a, started[1], ctl[0], ctl[1], sum_rel,
                      by_phase[0], by_phase[1], by_phase[2], by_phase[3], by_phase[4], by_phase[5], by_phase[6], by_phase[7], by_phase[8],
                      sum_done, max_behind, slab_a / a_col, free_cols, nf, largest, w[3], w[0] - w[1], ready_below, ready_above, first_unready,
                      sv_phase[0], sv_phase[1], sv_phase[2], sv_phase[3], sv_phase[4], sv_phase[5], sv_phase[6], sv_min, sv_max, ctl[2 + 2 * (size_t)fwd_a + 64]);
        std::string out(d);
        if (std::getenv("NC_DEBUG_STALL_S"))
        {
            // raw state: per forward CTA (finished jobs, releases, phase), per service warp (ticket, phase), unready tickets
            // below the published count, the free list
            char e[96];
            out += "\n  forward CTAs (done/released/phase):";
            for (unsigned k = 0; k < fwd_a; ++k) { std::snprintf(e, sizeof e, " %u/%u/%u", fph[k] >> 4, rel[k], fph[k] & 15u); out += e; }
            out += "\n  service warps (ticket/phase):";
            for (unsigned k = 0; k < 16 * tb_a; ++k) { std::snprintf(e, sizeof e, " %u/%u", sph[k] >> 4, sph[k] & 7u); out += e; }
            out += "\n  tickets below the published count that are not ready:";
            for (unsigned k = 0; k < std::min(ctl[0], n_short); ++k)
                if (!tks[k].ready) { std::snprintf(e, sizeof e, " %u(job %u cta %u col0 %u)", k, tks[k].job, tks[k].slab, tks[k].col0); out += e; }
            out += "\n  free extents (start+len):";
            for (unsigned k = 0; k < std::min(nf, 64u); ++k) { std::snprintf(e, sizeof e, " %u+%u", w[4 + k], w[4 + 1024 + k]); out += e; }
        }
        return out;
    };
    // NC_DEBUG_STALL_S=<seconds>: if the kernels are still running after that long, print the grid's state and go on waiting
    // (before the copies back: a copy into pageable memory blocks the host until the kernels are done)
    if (const char* v = std::getenv("NC_DEBUG_STALL_S"))
    {
        const double lim = std::atof(v);
        const auto t0 = std::chrono::steady_clock::now();
        bool dumped = false;
        while (cudaStreamQuery(s) == cudaErrorNotReady)
        {
            const double el = std::chrono::duration< double >(std::chrono::steady_clock::now() - t0).count();
            if (!dumped && el > lim)
            {
                std::fprintf(stderr, "nc_viterbi_packed: still running after %.1f s: %s\n", el, alpha_state().c_str());
                std::this_thread::sleep_for(std::chrono::milliseconds(500));
                std::fprintf(stderr, "nc_viterbi_packed: 0.5 s later: %s\n", alpha_state().c_str());
                dumped = true;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        cudaGetLastError();
    }
    NC_CUDA_INFLIGHT(cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    NC_CUDA_INFLIGHT(cudaMemcpyAsync(path_logprob, ctx->path.p, n_jobs_all * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (mem == NC_MEM_HOST)
    {
        if (states) NC_CUDA_INFLIGHT(cudaMemcpyAsync(states + base, ctx->states.p, total * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
        if (moves) NC_CUDA_INFLIGHT(cudaMemcpyAsync(moves + base, ctx->moves.p, total, cudaMemcpyDeviceToHost, s));
    }
    NC_CUDA_INFLIGHT(cudaStreamSynchronize(s));
    if (stream_in) NC_CUDA_INFLIGHT(cudaStreamSynchronize(ctx->stream3));
    NC_CUDA_INFLIGHT(cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev0, ctx->ev1));
    if (*ctx->h_abort != 0u)
    {
        static const char* const why[] = { "", "a forward CTA waited for its release slot", "a forward CTA waited for alpha columns",
                                           "a traceback warp waited for a ticket", "the column allocator's extent list overflowed" };
        const unsigned code = *ctx->h_abort;
        const std::string diag = alpha_state();
        if (std::getenv("NC_DEBUG_STALL_S")) std::fprintf(stderr, "nc_viterbi_packed: after the abort: %s\n", diag.c_str());
        NC_FAIL(ctx, NC_ERR_STATE, "nc_viterbi_packed: the alpha-column kernel stopped without finishing (%s for more than %.0f s): "
                "results of this call are invalid%s%s", code < 5 ? why[code] : "unknown reason", ctx->wait_limit_s, diag.empty() ? "" : "; ", diag.c_str());
    }
    return NC_OK;
#undef NC_CUDA_INFLIGHT
}

void* nc_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void nc_host_free(void* p) { if (p) cudaFreeHost(p); }

int nc_viterbi_batch(nc_ctx* ctx, uint32_t n_jobs, const nc_vit_job* jobs, nc_vit_out* outs)
{
    if (!ctx) return NC_ERR_ARG;
    if (n_jobs == 0) return NC_OK;
    if (!jobs || !outs) NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_batch: NULL argument");
    std::vector< uint64_t > off(n_jobs + 1, 0);
    for (uint32_t k = 0; k < n_jobs; ++k)
    {
        if (jobs[k].n_events == 0 || !jobs[k].mean || !jobs[k].stdv || !jobs[k].start)
            NC_FAIL(ctx, NC_ERR_ARG, "nc_viterbi_batch: job %u is empty or has NULL events", k);
        off[k + 1] = off[k] + jobs[k].n_events;
    }
    const uint64_t total = off[n_jobs];
    std::vector< float > mean(total), stdv(total), start(total), path(n_jobs);
    std::vector< int32_t > mid(n_jobs);
    std::vector< nc_pm_params > pm(n_jobs);
    std::vector< nc_st_params > st(n_jobs);
    std::vector< uint16_t > states(total);
    std::vector< uint8_t > moves(total);
    for (uint32_t k = 0; k < n_jobs; ++k)
    {
        std::memcpy(mean.data() + off[k], jobs[k].mean, jobs[k].n_events * sizeof(float));
        std::memcpy(stdv.data() + off[k], jobs[k].stdv, jobs[k].n_events * sizeof(float));
        std::memcpy(start.data() + off[k], jobs[k].start, jobs[k].n_events * sizeof(float));
        mid[k] = jobs[k].model_id;
        pm[k] = jobs[k].pm;
        st[k] = jobs[k].st;
    }
    int rc = nc_viterbi_packed(ctx, n_jobs, off.data(), mean.data(), stdv.data(), start.data(), nullptr,
                               mid.data(), pm.data(), st.data(), NC_MEM_HOST, path.data(), states.data(), moves.data());
    if (rc != NC_OK) return rc;
    for (uint32_t k = 0; k < n_jobs; ++k)
    {
        outs[k].path_logprob = path[k];
        if (outs[k].states) std::memcpy(outs[k].states, states.data() + off[k], jobs[k].n_events * sizeof(uint16_t));
        if (outs[k].moves) std::memcpy(outs[k].moves, moves.data() + off[k], jobs[k].n_events);
        outs[k].n_bases = nc_base_seq(jobs[k].n_events, states.data() + off[k], moves.data() + off[k],
                                      outs[k].bases, outs[k].bases ? outs[k].bases_cap : 0);
    }
    return NC_OK;
}

} // extern "C"
