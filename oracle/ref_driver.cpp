// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
//
// C-ABI driver over the UNMODIFIED reference headers under /root/reference/src (compiled in
// place by oracle/Makefile into oracle/_ref/libncref.so; no reference source is copied here).
// It instantiates the reference's own Viterbi<float,6>, Forward_Backward<float,6>,
// Parameter_Trainer<float,6>, State_Transitions<float,6> and Pore_Model<float,6> and exposes
// them with plain pointers so that tests/ and bench.py's cpu_baseline leg can
//   (1) pin the C restatement in oracle/nc_oracle.c bit-for-bit, and
//   (2) generate the golden vectors committed under tests/golden/.
// Per-read orchestration mirrors the basecall_strand lambda (nanocall.cpp:645-690) and
// Parameter_Trainer::train_one_round (Parameter_Trainer.hpp:541-579).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <array>
#include <set>
#include <iomanip>
#include <cassert>

#include "Pore_Model.hpp"
#include "Builtin_Model.hpp"
#include "State_Transitions.hpp"
#include "Event.hpp"
#include "Viterbi.hpp"
#include "Forward_Backward.hpp"
#include "Parameter_Trainer.hpp"
#include "logger.hpp"

typedef State_Transitions< float, 6 > ST;
typedef State_Transition_Parameters< float > STP;
typedef Pore_Model< float, 6 > PM;
typedef Pore_Model_Parameters< float > PMP;
typedef Event< float, 6 > EV;
typedef Event_Sequence< float, 6 > EVS;
typedef Viterbi< float, 6 > VIT;
typedef Forward_Backward< float, 6 > FB;
typedef Parameter_Trainer< float, 6 > PT;

namespace
{
const unsigned S = 4096;

PM make_model(const float* table)
{
    std::vector< float > v(table, table + 4 * S);
    PM pm;
    pm.load_from_vector(v);
    return pm;
}
PMP make_params(const float* p)
{
    PMP r;
    r.scale = p[0]; r.shift = p[1]; r.drift = p[2]; r.var = p[3]; r.scale_sd = p[4]; r.var_sd = p[5];
    return r;
}
EVS make_events(unsigned n, const float* mean, const float* stdv, const float* start)
{
    // as Event::operator>> (Event.hpp:59-68) and Fast5_Summary::load_events (Fast5_Summary.hpp:352-360)
    EVS ev;
    ev.reserve(n);
    for (unsigned i = 0; i < n; ++i)
    {
        EV e;
        e.mean = mean[i];
        e.corrected_mean = e.mean;
        e.stdv = stdv[i];
        e.start = start[i];
        e.length = 0;
        e.update_logs();
        ev.push_back(e);
    }
    return ev;
}
} // namespace

extern "C"
{

// mirrors main()'s copies of CLI options into statics (nanocall.cpp:923-924,970) and
// train_reads' call to Parameter_Trainer::init (nanocall.cpp:280)
int ncref_init(float default_p_stay, float default_p_skip, int train_drift)
{
    logger::Logger::set_default_level(logger::level::warning);
    STP::default_p_stay() = default_p_stay;
    STP::default_p_skip() = default_p_skip;
    PT::pm_train_drift() = train_drift;
    PT::init();
    return (int)PT::st_train_kmers().size();
}

int ncref_st_train_kmers(uint32_t* out)
{
    unsigned n = PT::st_train_kmers().size();
    if (out) for (unsigned i = 0; i < n; ++i) out[i] = PT::st_train_kmers()[i];
    return (int)n;
}

int ncref_n_builtin() { return (int)Builtin_Model::num; }

int ncref_builtin(int idx, float* table, int* strand, char* name, int name_cap)
{
    if (idx < 0 or idx >= (int)Builtin_Model::num) return -1;
    const auto& v = Builtin_Model::init_lists[idx];
    if (v.size() != 4 * S) return -2;
    if (table) std::memcpy(table, v.data(), 4 * S * sizeof(float));
    if (strand) *strand = (int)Builtin_Model::strands[idx];
    if (name) { std::strncpy(name, Builtin_Model::names[idx].c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    return 0;
}

// out: S x 8 floats {level_mean, level_stdv, log_level_stdv, sd_mean, sd_lambda, log_sd_lambda, sd_stdv, log_sd_mean}
// stats: {mean, stdv} of level_mean (Pore_Model::update_statistics)
int ncref_scaled_model(const float* table, const float* pm_params, float* out, float* stats)
{
    PM pm = make_model(table);
    if (pm_params) pm.scale(make_params(pm_params));
    for (unsigned j = 0; j < S; ++j)
    {
        const auto& s = pm.state(j);
        float* o = out + 8 * j;
        o[0] = s.level_mean; o[1] = s.level_stdv; o[2] = s.log_level_stdv; o[3] = s.sd_mean;
        o[4] = s.sd_lambda; o[5] = s.log_sd_lambda; o[6] = s.sd_stdv; o[7] = s.log_sd_mean;
    }
    if (stats) { stats[0] = pm.mean(); stats[1] = pm.stdv(); }
    return 0;
}

// per-(state,event) emission table, emit[i*S + j] (Pore_Model.hpp:145-149)
int ncref_emissions(const float* table, const float* pm_params, unsigned n,
                    const float* mean, const float* stdv, const float* start, float* emit)
{
    PM pm = make_model(table);
    PMP p = make_params(pm_params);
    pm.scale(p);
    EVS ev = make_events(n, mean, stdv, start);
    ev.apply_drift_correction(p.drift);
    for (unsigned i = 0; i < n; ++i)
        for (unsigned j = 0; j < S; ++j)
            emit[(size_t)i * S + j] = pm.log_pr_corrected_emission(j, ev[i]);
    return 0;
}

// from/to neighbour lists of compute_transitions_fast (State_Transitions.hpp:181-224)
// cnt: S entries; idx/lp: S x 21 (row-padded)
int ncref_transitions(float p_skip, float p_stay,
                      uint32_t* from_cnt, uint32_t* from_idx, float* from_lp,
                      uint32_t* to_cnt, uint32_t* to_idx, float* to_lp)
{
    ST st;
    st.compute_transitions_fast(p_skip, p_stay);
    for (unsigned j = 0; j < S; ++j)
    {
        const auto& fv = st.neighbours(j).from_v;
        const auto& tv = st.neighbours(j).to_v;
        if (fv.size() > 21 or tv.size() > 21) return -1;
        from_cnt[j] = fv.size();
        to_cnt[j] = tv.size();
        for (unsigned k = 0; k < fv.size(); ++k) { from_idx[21 * j + k] = fv[k].first; from_lp[21 * j + k] = fv[k].second; }
        for (unsigned k = 0; k < tv.size(); ++k) { to_idx[21 * j + k] = tv[k].first; to_lp[21 * j + k] = tv[k].second; }
    }
    return 0;
}

// one basecall_strand (nanocall.cpp:645-690): scale, transitions, drift-correct, Viterbi::fill
// alpha_dump/beta_dump: optional n x S dumps of the DP matrix
int ncref_viterbi(const float* table, const float* pm_params, float p_stay, float p_skip,
                  unsigned n, const float* mean, const float* stdv, const float* start,
                  float* path_prob, uint32_t* states, int32_t* moves,
                  char* bases, unsigned bases_cap, unsigned* n_bases,
                  float* alpha_dump, uint32_t* beta_dump)
{
    PM pm = make_model(table);
    PMP p = make_params(pm_params);
    pm.scale(p);
    ST st;
    st.compute_transitions_fast(p_skip, p_stay);
    EVS ev = make_events(n, mean, stdv, start);
    ev.apply_drift_correction(p.drift);
    VIT vit;
    vit.fill(pm, st, ev);
    if (path_prob) *path_prob = vit.path_probability();
    for (unsigned i = 0; i < n; ++i)
    {
        if (states) states[i] = ev[i].model_state_idx;
        if (moves) moves[i] = ev[i].move;
    }
    if (bases or n_bases)
    {
        std::string s = ev.get_base_seq();
        if (n_bases) *n_bases = s.size();
        if (bases) { std::strncpy(bases, s.c_str(), bases_cap); }
    }
    if (alpha_dump or beta_dump)
    {
        for (unsigned i = 0; i < n; ++i)
            for (unsigned j = 0; j < S; ++j)
            {
                if (alpha_dump) alpha_dump[(size_t)i * S + j] = vit.cell(i, j).alpha;
                if (beta_dump) beta_dump[(size_t)i * S + j] = vit.cell(i, j).beta;
            }
    }
    return 0;
}

// Batch of Viterbi jobs sharing one model table, threaded like pfor with chunk size 1
// (pfor.hpp:183-205): workers pull the next job index under dynamic scheduling.
// Transitions are built once when every job uses the same (p_stay,p_skip) -- as the reference
// does with default_transitions (nanocall.cpp:180-193).
int ncref_viterbi_batch(const float* table, unsigned n_jobs, const uint64_t* ev_off,
                        const float* mean, const float* stdv, const float* start,
                        const float* pm_params /* n_jobs x 6 */, const float* st_params /* n_jobs x 2: p_stay,p_skip */,
                        unsigned n_threads,
                        float* path_prob, uint32_t* states, int32_t* moves)
{
    PM pm0 = make_model(table);
    ST st0;
    bool all_same = true;
    for (unsigned k = 1; k < n_jobs; ++k)
        if (st_params[2 * k] != st_params[0] or st_params[2 * k + 1] != st_params[1]) all_same = false;
    if (all_same and n_jobs > 0) st0.compute_transitions_fast(st_params[1], st_params[0]);
    std::atomic< unsigned > next(0);
    auto worker = [&] () {
        while (true)
        {
            unsigned k = next++;
            if (k >= n_jobs) break;
            PM pm(pm0);
            PMP p = make_params(pm_params + 6 * k);
            pm.scale(p);
            ST custom;
            const ST* stp = &st0;
            if (not all_same) { custom.compute_transitions_fast(st_params[2 * k + 1], st_params[2 * k]); stp = &custom; }
            unsigned n = ev_off[k + 1] - ev_off[k];
            EVS ev = make_events(n, mean + ev_off[k], stdv + ev_off[k], start + ev_off[k]);
            ev.apply_drift_correction(p.drift);
            VIT vit;
            vit.fill(pm, *stp, ev);
            if (path_prob) path_prob[k] = vit.path_probability();
            for (unsigned i = 0; i < n; ++i)
            {
                if (states) states[ev_off[k] + i] = ev[i].model_state_idx;
                if (moves) moves[ev_off[k] + i] = ev[i].move;
            }
        }
    };
    if (n_threads <= 1) worker();
    else
    {
        std::vector< std::thread > th;
        for (unsigned t = 0; t < n_threads; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    return 0;
}

// Forward_Backward::fill on a scaled model + drift-corrected events (as fill_train_data does,
// Parameter_Trainer.hpp:141-155). alpha/beta: n x S.
int ncref_fwbw(const float* table, const float* pm_params, float p_stay, float p_skip,
               unsigned n, const float* mean, const float* stdv, const float* start,
               float* alpha, float* beta, float* log_pr_data)
{
    PM pm = make_model(table);
    PMP p = make_params(pm_params);
    pm.scale(p);
    ST st;
    st.compute_transitions_fast(p_skip, p_stay);
    EVS ev = make_events(n, mean, stdv, start);
    ev.apply_drift_correction(p.drift);
    FB fb;
    fb.fill(pm, st, ev);
    if (log_pr_data) *log_pr_data = fb.log_pr_data();
    for (unsigned i = 0; i < n; ++i)
        for (unsigned j = 0; j < S; ++j)
        {
            if (alpha) alpha[(size_t)i * S + j] = fb.cell(i, j).alpha;
            if (beta) beta[(size_t)i * S + j] = fb.cell(i, j).beta;
        }
    return 0;
}

// Parameter_Trainer::train_one_round (Parameter_Trainer.hpp:541-579).
// seqs are concatenated; seq_strand[k] in {0,1}; table0/table1 = unscaled models per strand;
// st_params = {p_stay0, p_skip0, p_stay1, p_skip1}; default transitions use the statics set by
// ncref_init.
int ncref_train_one_round(unsigned n_seqs, const uint32_t* seq_len, const uint32_t* seq_strand,
                          const float* mean, const float* stdv, const float* start,
                          const float* table0, const float* table1,
                          const float* pm_params, const float* st_params,
                          int train_scaling, int train_transitions,
                          float* new_pm_params, float* new_st_params, float* fit, int* done)
{
    static ST default_transitions;
    static float def_key[2] = { -1, -1 };
    static std::mutex def_mutex;   // bench.py times this entry point from several threads
    {
        std::lock_guard< std::mutex > def_lock(def_mutex);
        if (def_key[0] != STP::default_p_stay() or def_key[1] != STP::default_p_skip())
        {
            default_transitions.compute_transitions_fast(STP::default_p_skip(), STP::default_p_stay());
            def_key[0] = STP::default_p_stay();
            def_key[1] = STP::default_p_skip();
        }
    }
    PM m0 = make_model(table0);
    PM m1 = make_model(table1);
    std::vector< EVS > seqs;
    size_t off = 0;
    for (unsigned k = 0; k < n_seqs; ++k)
    {
        seqs.push_back(make_events(seq_len[k], mean + off, stdv + off, start + off));
        off += seq_len[k];
    }
    std::vector< std::pair< const EVS*, unsigned > > ptrs;
    for (unsigned k = 0; k < n_seqs; ++k) ptrs.push_back(std::make_pair(&seqs[k], (unsigned)seq_strand[k]));
    PMP crt_pm = make_params(pm_params);
    std::array< STP, 2 > crt_st;
    crt_st[0].p_stay = st_params[0]; crt_st[0].p_skip = st_params[1];
    crt_st[1].p_stay = st_params[2]; crt_st[1].p_skip = st_params[3];
    // the reference passes crt_* as the destinations too (nanocall.cpp:374-380): start them equal
    PMP new_pm(crt_pm);
    std::array< STP, 2 > new_st(crt_st);
    float f = 0;
    bool d = false;
    PT::train_one_round(ptrs, {{ &m0, &m1 }}, default_transitions, crt_pm, crt_st,
                        new_pm, new_st, f, d, train_scaling, train_transitions);
    new_pm_params[0] = new_pm.scale; new_pm_params[1] = new_pm.shift; new_pm_params[2] = new_pm.drift;
    new_pm_params[3] = new_pm.var; new_pm_params[4] = new_pm.scale_sd; new_pm_params[5] = new_pm.var_sd;
    new_st_params[0] = new_st[0].p_stay; new_st_params[1] = new_st[0].p_skip;
    new_st_params[2] = new_st[1].p_stay; new_st_params[3] = new_st[1].p_skip;
    *fit = f;
    *done = d;
    return 0;
}

// alg::mean_stdv_of<float> over event means (alg.hpp:466-482), used for initial scaling
// (Fast5_Summary.hpp:225-267) and the means_apart check (nanocall.cpp:633-635)
int ncref_mean_stdv(unsigned n, const float* x, float* mean, float* stdv)
{
    std::vector< float > v(x, x + n);
    auto r = alg::mean_stdv_of< float >(v);
    *mean = r.first;
    *stdv = r.second;
    return 0;
}

// p7_FLogsum (logsum.hpp:141-154) and its table (logsum.hpp:113-127)
float ncref_flogsum(float a, float b) { return logsum::p7_FLogsum(a, b); }
int ncref_flogsum_table(float* out)
{
    (void)logsum::p7_FLogsum(0.f, 0.f);
    for (int i = 0; i < 16000; ++i) out[i] = logsum::p7_FLogsum_Helper::flogsum_lookup()[i];
    return 16000;
}

} // extern "C"
