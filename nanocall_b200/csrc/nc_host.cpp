// Host-side arithmetic of the boundary: everything here runs once per model / job / event on the
// CPU with libm so that the transcendental inputs of the device recursion are the same bits the
// reference computes.  Compiled WITHOUT -march and with
// -ffp-contract=off: the reference's Release build is baseline x86-64 (no FMA).
#include "nc_internal.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <set>
#include <queue>
#include <cstring>
#include <thread>
#include <vector>

namespace nc {

static inline unsigned prefix(unsigned i, unsigned k) { return i >> (2 * (NC_KMER - k)); }
static inline unsigned suffix(unsigned i, unsigned k) { return i & ((1u << (2 * k)) - 1); }

// Pore_Model::load_from_vector + update_sd_lambda + update_logs (Pore_Model.hpp:112,118-124,220-239)
void host_model_prepare(const float* table, HostModel& m)
{
    m.level_mean.resize(NC_N_STATES); m.level_stdv.resize(NC_N_STATES); m.sd_mean.resize(NC_N_STATES);
    m.sd_lambda.resize(NC_N_STATES); m.log_level_stdv.resize(NC_N_STATES); m.log_sd_lambda.resize(NC_N_STATES);
    for (unsigned i = 0; i < NC_N_STATES; ++i)
    {
        float level_mean = table[4 * i + 0], level_stdv = table[4 * i + 1];
        float sd_mean = table[4 * i + 2], sd_stdv = table[4 * i + 3];
        // pow(float, double) promotes: evaluated in double, narrowed on store
        float sd_lambda = static_cast< float >(std::pow(static_cast< double >(sd_mean), 3.0)
                                               / std::pow(static_cast< double >(sd_stdv), 2.0));
        m.level_mean[i] = level_mean;
        m.level_stdv[i] = level_stdv;
        m.sd_mean[i] = sd_mean;
        m.sd_lambda[i] = sd_lambda;
        m.log_level_stdv[i] = std::log(level_stdv);
        m.log_sd_lambda[i] = std::log(sd_lambda);
    }
    nc_mean_stdv(NC_N_STATES, m.level_mean.data(), &m.mean, &m.stdv);
}

// Per-job scalars of Pore_Model::scale (Pore_Model.hpp:190-195): the three logs are float logs
void host_job_logs(const nc_pm_params& p, float& log_var, float& log_var_sd)
{
    log_var = std::log(p.var);
    log_var_sd = std::log(p.var_sd);
}

// Event::update_logs (Event.hpp:39-43) over a whole event array, threaded
void host_event_logs(size_t n, const float* stdv, float* log_stdv, unsigned n_threads)
{
    auto work = [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i)
        {
            float s = stdv[i];
            if (s == 0.0) s = 0.01;
            log_stdv[i] = std::log(s);
        }
    };
    if (n < (1u << 16) || n_threads <= 1) { work(0, n); return; }
    std::vector< std::thread > th;
    size_t chunk = (n + n_threads - 1) / n_threads;
    for (unsigned t = 0; t < n_threads; ++t)
    {
        size_t a = t * chunk, b = a + chunk < n ? a + chunk : n;
        if (a < b) th.emplace_back(work, a, b);
    }
    for (auto& t : th) t.join();
}

} // namespace nc

extern "C" {

// alg::mean_stdv_of<float> (alg.hpp:466-482): single pass in float, final expression in double
void nc_mean_stdv(uint32_t n, const float* x, float* mean_out, float* stdv_out)
{
    float s = 0.0f, s2 = 0.0f;
    unsigned long cnt = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        s += x[i];
        s2 += x[i] * x[i];
        ++cnt;
    }
    float mean = cnt > 0 ? s / cnt : 0.0f;
    float stdv = 0.0f;
    if (cnt > 1)
    {
        float sm = s * mean;
        float mmn = mean * mean * static_cast< float >(cnt);
        stdv = static_cast< float >(std::sqrt((static_cast< double >(s2) - static_cast< double >(sm) * 2.0
                                              + static_cast< double >(mmn)) / static_cast< double >(cnt - 1)));
    }
    *mean_out = mean;
    *stdv_out = stdv;
}

// State_Transitions::get_trans_prob (State_Transitions.hpp:125-144) as a function of the overlap
// mask alone; p_step / p_skip_1 as compute_transitions_fast derives them (:199-201).
void nc_transition_lut(float p_stay, float p_skip, float* lut64)
{
    float p_step = static_cast< float >(1.0 - p_stay - p_skip);
    float p_skip_1 = static_cast< float >(p_skip / (p_skip + 1.0));
    // the double-valued terms do not depend on the mask: evaluate each pow once
    double term[NC_KMER];
    for (unsigned l = 2; l < NC_KMER; ++l) term[l] = std::pow(p_skip_1, l - 1) / (1u << (2 * l));
    const double tail = (std::pow(p_skip_1, 5) / (1.0f - p_skip_1)) / NC_N_STATES;
    for (unsigned mask = 0; mask < 64; ++mask)
    {
        float p = 0;
        if (mask & 1u) p += p_stay;
        if (mask & 2u) p += p_step / 4;
        for (unsigned l = 2; l < NC_KMER; ++l)
            if (mask & (1u << l)) p += term[l];  // float += double: added in double, narrowed (as :139)
        p += tail;
        lut64[mask] = std::log(p);
    }
}

// Dispatch order of the alpha-column kernel's jobs.  lens[] is sorted descending (longest-first balances the forward
// CTAs), but that order starts all the long reads at once and they pin their alpha columns for their whole forward
// pass: measured on the configs[4] mixture, forward CTAs then wait for columns 14 % of the time.  The order returned
// is the one a simulated run of the launch produces with the rule "the longest job whose columns are free now"
// (n_workers forward CTAs, time proportional to the events of a job, columns held until the job ends plus a margin
// for its traceback, 90 % of the pool to allow for fragmentation): longest-first whenever memory allows, shorter
// jobs while it is tight.  perm[k] = index into lens[] of the k-th job to dispatch.  Returns 0 (perm = identity)
// when the first wave simply fits.
int nc_plan_dispatch_order(uint32_t n_jobs, const uint32_t* lens, uint64_t pool_columns, uint32_t n_workers, uint32_t* perm)
{
    for (uint32_t k = 0; k < n_jobs; ++k) perm[k] = k;
    if (n_jobs == 0 || n_workers == 0 || n_jobs <= n_workers) return 0;
    const uint64_t pool_cols = (uint64_t)((double)pool_columns * 0.9);
    // the first wave = the n_workers longest jobs, wherever they are in lens[] (the caller need not pass them sorted)
    uint64_t need_first = 0;
    {
        std::vector< uint32_t > top(lens, lens + n_jobs);
        std::nth_element(top.begin(), top.begin() + n_workers, top.end(), std::greater< uint32_t >());
        for (uint32_t k = 0; k < n_workers; ++k) need_first += top[k];
    }
    if (need_first * 11 / 10 <= pool_cols) return 0;
    std::multiset< std::pair< uint32_t, uint32_t > > remaining;   // (length, index): ascending
    for (uint32_t k = 0; k < n_jobs; ++k) remaining.insert({ lens[k], k });
    typedef std::pair< double, uint32_t > Fin;                    // (finish time, columns)
    std::priority_queue< Fin, std::vector< Fin >, std::greater< Fin > > running;
    uint64_t free_cols = pool_cols;
    uint32_t idle = n_workers, out = 0;
    double now = 0.0;
    while (!remaining.empty())
    {
        bool placed = false;
        if (idle > 0)
        {
            auto it = remaining.upper_bound({ (uint32_t)std::min< uint64_t >(free_cols, 0xffffffffu), 0xffffffffu });
            if (it != remaining.begin())
            {
                --it;   // the longest job that fits
                const uint32_t len = it->first;
                perm[out++] = it->second;
                remaining.erase(it);
                free_cols -= len;
                --idle;
                running.push({ now + 1.1 * (double)len, len });
                placed = true;
            }
        }
        if (!placed)
        {
            if (running.empty())   // a job larger than 90 % of the pool: it runs alone
            {
                auto it = std::prev(remaining.end());
                perm[out++] = it->second;
                remaining.erase(it);
                continue;
            }
            now = running.top().first;
            free_cols += running.top().second;
            ++idle;
            running.pop();
        }
    }
    return 1;
}

// Kmer::min_skip (Kmer.hpp:51-68)
uint32_t nc_min_skip(uint32_t k1, uint32_t k2)
{
    if (k1 == k2) return 0;
    for (unsigned k = NC_KMER - 1; k > 0; --k)
        if (nc::suffix(k1, k) == nc::prefix(k2, k)) return NC_KMER - k;
    return NC_KMER;
}

// Event_Sequence::get_base_seq (Event.hpp:85-99)
uint32_t nc_base_seq(uint32_t n_events, const uint16_t* states, const uint8_t* moves, char* out, uint32_t cap)
{
    static const char b2c[4] = { 'A', 'C', 'G', 'T' };
    if (n_events == 0) return 0;
    uint32_t len = 0;
    for (unsigned c = 0; c < NC_KMER; ++c)
    {
        if (out && len < cap) out[len] = b2c[(states[0] >> (2 * (NC_KMER - 1 - c))) & 3];
        ++len;
    }
    for (uint32_t i = 1; i < n_events; ++i)
    {
        unsigned a = moves[i] < NC_KMER ? moves[i] : NC_KMER;
        for (unsigned c = NC_KMER - a; c < NC_KMER; ++c)
        {
            if (out && len < cap) out[len] = b2c[(states[i] >> (2 * (NC_KMER - 1 - c))) & 3];
            ++len;
        }
    }
    return len;
}

const char* nc_version(void) { return "nanocall_b200 0.1 (sm_100a)"; }

} // extern "C"
