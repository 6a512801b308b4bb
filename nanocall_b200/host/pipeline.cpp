// See pipeline.hpp for the map onto the reference.  Every decision below (candidate lists, training
// slices, EM stop rules, model selection, candidate ranking, FASTA naming) follows nanocall.cpp line
// by line, but each "call Parameter_Trainer / Viterbi for this read" becomes "append a job to the batch".
#include "pipeline.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace nchost {

void log_line(int level, int threshold, const std::string& msg)
{
    if (level <= threshold) std::clog << (msg + "\n") << std::flush;   // one insertion per line: several threads log
}

#define NLOG(lvl, expr)                                   \
    do {                                                  \
        if ((lvl) <= opt_.log_level) {                    \
            std::ostringstream _o;                        \
            _o << expr;                                   \
            _o << '\n';                                   \
            std::clog << _o.str() << std::flush;          \
        }                                                 \
    } while (0)

static std::ostream& operator<<(std::ostream& os, const nc_pm_params& p)  // Pore_Model.hpp:66-71
{
    os << "[scale=" << p.scale << " shift=" << p.shift << " drift=" << p.drift
       << " var=" << p.var << " scale_sd=" << p.scale_sd << " var_sd=" << p.var_sd << "]";
    return os;
}
static std::ostream& operator<<(std::ostream& os, const nc_st_params& p)  // State_Transitions.hpp:39-44
{
    os << "[p_stay=" << p.p_stay << " p_skip=" << p.p_skip << "]";
    return os;
}

static nc_pm_params default_pm()
{
    nc_pm_params p;
    p.scale = 1.0f; p.shift = 0.0f; p.drift = 0.0f; p.var = 1.0f; p.scale_sd = 1.0f; p.var_sd = 1.0f;
    return p;
}

Pipeline::Pipeline(const Options& o, int device, size_t viterbi_events_hint) : opt_(o)
{
    if (opt_.train_drift < 0) opt_.train_drift = (opt_.pore == "r73") ? 1 : 0;  // nanocall.cpp:943-969
    // Viterbi scratch: 16 KiB per event of the jobs in flight.  A shard that cannot use the default pool (3/4 of the
    // free memory, up to seconds of cudaMalloc) asks for what it can use.
    size_t pool = 0;
    if (viterbi_events_hint) pool = std::max< size_t >((size_t)1 << 30, viterbi_events_hint * 16384u + ((size_t)256 << 20));
    int rc = nc_ctx_create(device, pool, &ctx_);
    if (rc == NC_ERR_CUDA && pool) rc = nc_ctx_create(device, 0, &ctx_);   // larger than the device: take the default
    if (rc != NC_OK) throw std::runtime_error(std::string("nc_ctx_create: ") + nc_last_error(nullptr));
}

Pipeline::~Pipeline() { nc_ctx_destroy(ctx_); }

void Pipeline::check(int rc, const char* what) const
{
    if (rc != NC_OK) throw std::runtime_error(std::string(what) + ": " + nc_last_error(ctx_));
}

// ---------------------------------------------------------------- models (nanocall.cpp:97-178)
static bool read_model_tsv(const std::string& path, std::vector< float >& table, std::string& err)
{
    // Pore_Model::operator>> (Pore_Model.hpp:251-287): "kmer level_mean level_stdv sd_mean sd_stdv", '#' and
    // header lines skipped, rows sorted by k-mer
    std::ifstream is(path);
    if (!is) { err = "cannot open " + path; return false; }
    table.assign(4 * NC_N_STATES, 0.f);
    std::vector< bool > seen(NC_N_STATES, false);
    std::string line;
    unsigned n = 0;
    while (std::getline(is, line))
    {
        std::istringstream iss(line);
        std::string s;
        iss >> s;
        if (s.empty() || s[0] == '#') continue;
        if (line.find("kmer") != std::string::npos) continue;
        if (s.size() != NC_KMER) { err = "bad k-mer in " + path; return false; }
        unsigned idx = 0;
        for (char c : s)
        {
            int b = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
            if (b < 0) { err = "bad base in " + path; return false; }
            idx = (idx << 2) | (unsigned)b;
        }
        float* row = table.data() + 4 * idx;
        iss >> row[0] >> row[1] >> row[2] >> row[3];
        if (!seen[idx]) { seen[idx] = true; ++n; }
    }
    if (n != NC_N_STATES) { err = "unexpected number of states in " + path; return false; }
    return true;
}

void Pipeline::init_models()
{
    auto add = [&](const std::string& name, int strand, std::vector< float >&& table) {
        Model m;
        m.name = name;
        m.strand = strand;
        m.table = std::move(table);
        check(nc_model_register(ctx_, m.table.data(), strand, &m.id), "nc_model_register");
        check(nc_model_stats(ctx_, m.id, &m.mean, &m.stdv), "nc_model_stats");
        NLOG(2, "loaded module [" << name << "] for strand [" << strand << "] statistics [mean=" << m.mean
                                  << ", stdv=" << m.stdv << "]");
        models_[name] = std::move(m);
    };
    // -m strand:file (multi) and --model-fofn (one "strand:file" per line), nanocall.cpp:99-153
    std::vector< std::string > specs = opt_.model_files;
    if (!opt_.model_fofn.empty())
    {
        std::ifstream is(opt_.model_fofn);
        if (!is) throw std::runtime_error("cannot open " + opt_.model_fofn);
        std::string line;
        while (std::getline(is, line)) specs.push_back(line);
    }
    if (!specs.empty())
    {
        std::vector< std::string > by_strand[3];
        for (const auto& s : specs)
        {
            if (s.size() < 3 || (s[0] != '0' && s[0] != '1' && s[0] != '2') || s[1] != ':')
                throw std::runtime_error("could not parse model name: \"" + s + "\"; format should be \"[0|1|2]:<file>\"");
            by_strand[s[0] - '0'].push_back(s.substr(2));
        }
        if (by_strand[2].empty() && (by_strand[0].empty() != by_strand[1].empty()))
            throw std::runtime_error(std::string("models were specified only for strand ") + (by_strand[0].empty() ? "1" : "0")
                                     + "! give models for both strands, or for neither.");
        for (int st = 0; st < 3; ++st)
            for (const auto& f : by_strand[st])
            {
                std::vector< float > table;
                std::string err;
                if (!read_model_tsv(f, table, err)) throw std::runtime_error(err);
                add(f, st, std::move(table));
            }
        return;
    }
    // builtin models: names filtered by "<pore>." prefix (nanocall.cpp:157-170)
    std::string dir = opt_.data_dir;
    std::ifstream names(dir + "/builtin_models.txt");
    std::ifstream blob(dir + "/builtin_models.bin", std::ios::binary);
    if (!names || !blob) throw std::runtime_error("builtin model data not found under " + dir);
    std::string name;
    int strand;
    unsigned idx = 0;
    while (names >> name >> strand)
    {
        std::vector< float > table(4 * NC_N_STATES);
        blob.seekg((std::streamoff)idx * 4 * NC_N_STATES * sizeof(float));
        blob.read(reinterpret_cast< char* >(table.data()), table.size() * sizeof(float));
        if (!blob) throw std::runtime_error("builtin_models.bin is truncated");
        ++idx;
        if (name.compare(0, opt_.pore.size() + 1, opt_.pore + ".") != 0) continue;
        add(name, strand, std::move(table));
    }
    if (models_.empty()) throw std::runtime_error("no builtin models found for pore [" + opt_.pore + "]");
}

// init_transitions (nanocall.cpp:180-193): a custom initial table (-s/--trans, State_Transitions::operator>>,
// State_Transitions.hpp:237-252: lines "kmer_i kmer_j log_prob", kept in file order), or the parametric one
void Pipeline::reserve(size_t reads, size_t events)
{
    // training: two candidate pairs per read x two strands x scaling_num_events events, at most
    const size_t train_events = opt_.train ? reads * 2 * 2 * (size_t)opt_.scaling_num_events : 0;
    // basecalling: every strand once per candidate still in the race (two at most with the builtin presets)
    const size_t vit_events = opt_.basecall ? 2 * events : 0;
    check(nc_ctx_reserve(ctx_, train_events, vit_events), "nc_ctx_reserve");
    if (train_events)
    {
        pin_tr_mean_.reserve(train_events * sizeof(float));
        pin_tr_stdv_.reserve(train_events * sizeof(float));
        pin_tr_start_.reserve(train_events * sizeof(float));
    }
    if (vit_events)
    {
        pin_mean_.reserve(vit_events * sizeof(float));
        pin_stdv_.reserve(vit_events * sizeof(float));
        pin_start_.reserve(vit_events * sizeof(float));
        pin_states_.reserve(vit_events * sizeof(uint16_t));
        pin_moves_.reserve(vit_events);
    }
}

void Pipeline::init_transitions()
{
    if (opt_.trans_fn.empty())
    {
        NLOG(2, "init_state_transitions pr_skip=[" << opt_.pr_skip << "], pr_stay=[" << opt_.pr_stay << "]");
        return;
    }
    std::ifstream is(opt_.trans_fn);
    if (!is) throw std::runtime_error("cannot open " + opt_.trans_fn);
    auto to_int = [&](const std::string& s) {
        if (s.size() != NC_KMER) throw std::runtime_error("bad k-mer [" + s + "] in " + opt_.trans_fn);
        unsigned idx = 0;
        for (char c : s)
        {
            int b = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
            if (b < 0) throw std::runtime_error("bad base in " + opt_.trans_fn);
            idx = (idx << 2) | (unsigned)b;
        }
        return (uint16_t)idx;
    };
    std::vector< uint16_t > from, to;
    std::vector< float > lp;
    std::string ki, kj;
    float p;
    while (is >> ki >> kj >> p)
    {
        from.push_back(to_int(ki));
        to.push_back(to_int(kj));
        lp.push_back(p);
    }
    if (from.empty()) throw std::runtime_error("no transitions in " + opt_.trans_fn);
    check(nc_ctx_set_default_transitions(ctx_, opt_.pr_stay, opt_.pr_skip, (uint32_t)from.size(), from.data(), to.data(), lp.data()),
          "nc_ctx_set_default_transitions");
    NLOG(2, "loaded state transitions from [" << opt_.trans_fn << "]");
}

// ---------------------------------------------------------------- initial scaling (Fast5_Summary.hpp:210-278)
void Pipeline::init_read_params(Read& r) const
{
    r.pm_params_m.clear();
    r.st_params_m.clear();
    for (auto& k : r.preferred_model) k = Model_Key();
    if (r.num_ed_events == 0) return;
    const nc_st_params dst = default_st();
    // r.scale_strands_together was decided by the loader from the RAW strand bounds (Fast5_Summary.hpp:210-212)
    if (r.scale_strands_together)
    {
        float m0, s0, m1, s1;
        nc_mean_stdv((uint32_t)r.events[0].size(), r.events[0].mean.data(), &m0, &s0);
        nc_mean_stdv((uint32_t)r.events[1].size(), r.events[1].mean.data(), &m1, &s1);
        for (const auto& p0 : models_)
            if (p0.second.strand == 0 || p0.second.strand == 2)
                for (const auto& p1 : models_)
                    if (p1.second.strand == 1 || p1.second.strand == 2)
                    {
                        Model_Key key = { { p0.first, p1.first } };
                        nc_pm_params pm = default_pm();
                        pm.scale = (s0 / p0.second.stdv + s1 / p1.second.stdv) / 2;
                        pm.shift = (m0 - pm.scale * p0.second.mean + m1 - pm.scale * p1.second.mean) / 2;
                        r.pm_params_m[key] = pm;
                        r.st_params_m[key] = { { dst, dst } };
                    }
    }
    else
    {
        for (unsigned st = 0; st < 2; ++st)
        {
            if (r.events[st].size() < opt_.min_ed_events) continue;
            float m, s;
            nc_mean_stdv((uint32_t)r.events[st].size(), r.events[st].mean.data(), &m, &s);
            for (const auto& p : models_)
                if (p.second.strand == (int)st || p.second.strand == 2)
                {
                    Model_Key key;
                    key[st] = p.first;
                    nc_pm_params pm = default_pm();
                    pm.scale = s / p.second.stdv;
                    pm.shift = m - pm.scale * p.second.mean;
                    r.pm_params_m[key] = pm;
                    r.st_params_m[key] = { { dst, dst } };
                }
        }
    }
}

// ---------------------------------------------------------------- training (nanocall.cpp:275-582)
namespace {
// fn(i) for i in [0, n) on up to n_threads threads (contiguous ranges; fn must not throw)
template < typename Fn >
void parallel_for(size_t n, unsigned n_threads, Fn fn)
{
    n_threads = (unsigned)std::min< size_t >(n_threads, (n + 63) / 64);
    if (n_threads <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::vector< std::thread > th;
    for (unsigned t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] {
            const size_t a = n * t / n_threads, b = n * (t + 1) / n_threads;
            for (size_t i = a; i < b; ++i) fn(i);
        });
    for (auto& x : th) x.join();
}
} // namespace

void Pipeline::init_reads_params(std::vector< Read* >& reads) const
{
    // (a pass over every event of every read for the initial scaling: 40 ms per 4096-read batch on one thread)
    parallel_for(reads.size(), opt_.host_threads, [&](size_t k) { init_read_params(*reads[k]); });
}

namespace {
struct Candidate
{
    size_t read;
    Model_Key key;
    unsigned strand;           // 0/1 single-strand training, 2 = both strands scaled together
    int model_id[2];
    nc_pm_params crt_pm;
    std::array< nc_st_params, 2 > crt_st;
    float crt_fit;
    unsigned round;
    unsigned max_rounds;
    bool active;
    // packed training sequences of this candidate
    std::vector< uint8_t > seq_strand;
    std::vector< uint32_t > seq_len;
    std::vector< size_t > seq_from;   // first event of the sequence in its strand
    std::vector< float > mean, stdv, start;
};
} // namespace

void Pipeline::train_reads(std::vector< Read* >& reads)
{
    std::vector< Candidate > cands;
    for (size_t ri = 0; ri < reads.size(); ++ri)
    {
        Read& rd = *reads[ri];
        // per-strand list of models to try (:300-323)
        std::array< std::vector< std::string >, 2 > model_list;
        for (unsigned st = 0; st < 2; ++st)
        {
            if (rd.events[st].size() < opt_.min_ed_events) continue;
            if (!rd.preferred_model[st][st].empty()) model_list[st].push_back(rd.preferred_model[st][st]);
            else
                for (const auto& p : models_)
                    if (p.second.strand == (int)st || p.second.strand == 2) model_list[st].push_back(p.first);
        }
        // two training sequences per strand: first and last n/2 events (:327-338)
        auto add_seqs = [&](Candidate& c, unsigned st) {
            const Strand_Events& ev = rd.events[st];
            unsigned n = (unsigned)std::min< size_t >(opt_.scaling_num_events, ev.size());
            unsigned h = n / 2;
            size_t from[2] = { 0, ev.size() - h };
            for (int part = 0; part < 2; ++part)
            {
                c.seq_strand.push_back((uint8_t)st);
                c.seq_len.push_back(h);
                c.seq_from.push_back(from[part]);   // (the events are copied below, on the helper threads)
            }
        };
        auto make = [&](const Model_Key& key, unsigned strand) {
            Candidate c;
            c.read = ri;
            c.key = key;
            c.strand = strand;
            c.crt_pm = rd.pm_params_m.at(key);
            c.crt_st = rd.st_params_m.at(key);
            c.crt_fit = -std::numeric_limits< float >::infinity();
            c.round = 0;
            c.active = true;
            return c;
        };
        if (rd.scale_strands_together)
        {
            for (const auto& m0 : model_list[0])
                for (const auto& m1 : model_list[1])
                {
                    Candidate c = make(Model_Key{ { m0, m1 } }, 2);
                    c.model_id[0] = models_.at(m0).id;
                    c.model_id[1] = models_.at(m1).id;
                    c.max_rounds = 2u * opt_.scaling_max_rounds;  // :420
                    add_seqs(c, 0);
                    add_seqs(c, 1);
                    cands.push_back(std::move(c));
                }
        }
        else
        {
            for (unsigned st = 0; st < 2; ++st)
            {
                if (rd.events[st].size() < opt_.min_ed_events) continue;
                for (const auto& m : model_list[st])
                {
                    Model_Key key;
                    key[st] = m;
                    Candidate c = make(key, st);
                    c.model_id[0] = c.model_id[1] = models_.at(m).id;  // :491
                    c.max_rounds = opt_.scaling_max_rounds;            // :536
                    add_seqs(c, st);
                    cands.push_back(std::move(c));
                }
            }
        }
    }
    parallel_for(cands.size(), opt_.host_threads, [&](size_t k) {
        Candidate& c = cands[k];
        const Read& rd = *reads[c.read];
        size_t n = 0;
        for (auto l : c.seq_len) n += l;
        c.mean.resize(n); c.stdv.resize(n); c.start.resize(n);
        size_t at = 0;
        for (size_t q = 0; q < c.seq_len.size(); ++q)
        {
            const Strand_Events& ev = rd.events[c.seq_strand[q]];
            std::memcpy(c.mean.data() + at, ev.mean.data() + c.seq_from[q], c.seq_len[q] * sizeof(float));
            std::memcpy(c.stdv.data() + at, ev.stdv.data() + c.seq_from[q], c.seq_len[q] * sizeof(float));
            std::memcpy(c.start.data() + at, ev.start.data() + c.seq_from[q], c.seq_len[q] * sizeof(float));
            at += c.seq_len[q];
        }
    });
    // drop candidates whose sequences are empty (n/2 == 0 cannot happen with min_ed_events >= 2, but be safe)
    for (auto& c : cands)
        for (auto l : c.seq_len)
            if (l == 0) c.active = false;

    nc_train_opts topts;
    topts.train_scaling = opt_.train_scaling;
    topts.train_transitions = opt_.train_transitions;
    topts.train_drift = opt_.train_drift;

    // ---- EM rounds: every active candidate advances by one train_one_round per batch call (:367-426, :483-542)
    std::vector< size_t > act;
    std::vector< uint32_t > seq_off;
    std::vector< uint64_t > ev_off, cand_ev;
    std::vector< uint8_t > strands;
    std::vector< nc_train_in > tin;
    std::vector< nc_train_out > tout;
    for (;;)
    {
        act.clear();
        for (size_t k = 0; k < cands.size(); ++k)
            if (cands[k].active) act.push_back(k);
        if (act.empty()) break;
        // offsets first (serial, light), then the candidates' events into the packed arrays on the helper threads: the
        // copy is the bulk of the host time between two calls, and the GPU waits for it
        seq_off.assign(1, 0);
        ev_off.assign(1, 0);
        strands.clear();
        tin.resize(act.size());
        tout.resize(act.size());
        cand_ev.resize(act.size());
        for (size_t a = 0; a < act.size(); ++a)
        {
            const Candidate& c = cands[act[a]];
            cand_ev[a] = ev_off.back();
            for (size_t s = 0; s < c.seq_len.size(); ++s)
            {
                strands.push_back(c.seq_strand[s]);
                ev_off.push_back(ev_off.back() + c.seq_len[s]);
            }
            seq_off.push_back(seq_off.back() + (uint32_t)c.seq_len.size());
            tin[a].model_id[0] = c.model_id[0];
            tin[a].model_id[1] = c.model_id[1];
            tin[a].pm = c.crt_pm;
            tin[a].st[0] = c.crt_st[0];
            tin[a].st[1] = c.crt_st[1];
        }
        // (page-locked: the call's three uploads are plain DMA instead of staged copies)
        float* mean = static_cast< float* >(pin_tr_mean_.reserve(ev_off.back() * sizeof(float)));
        float* stdv = static_cast< float* >(pin_tr_stdv_.reserve(ev_off.back() * sizeof(float)));
        float* start = static_cast< float* >(pin_tr_start_.reserve(ev_off.back() * sizeof(float)));
        parallel_for(act.size(), opt_.host_threads, [&](size_t a) {
            const Candidate& c = cands[act[a]];
            std::memcpy(mean + cand_ev[a], c.mean.data(), c.mean.size() * sizeof(float));
            std::memcpy(stdv + cand_ev[a], c.stdv.data(), c.stdv.size() * sizeof(float));
            std::memcpy(start + cand_ev[a], c.start.data(), c.start.size() * sizeof(float));
        });
        {
            std::lock_guard< std::mutex > gpu(gpu_mu_);
            const auto c0 = std::chrono::steady_clock::now();
            check(nc_train_round_batch(ctx_, (uint32_t)act.size(), seq_off.data(), ev_off.data(), strands.data(),
                                       mean, stdv, start, tin.data(), &topts, tout.data()),
                  "nc_train_round_batch");
            train_call_s += std::chrono::duration< double >(std::chrono::steady_clock::now() - c0).count();
            train_kernel_ms += nc_ctx_last_kernel_ms(ctx_);
        }
        ++train_rounds;
        fwbw_events += ev_off.back();
        for (size_t a = 0; a < act.size(); ++a)
        {
            Candidate& c = cands[act[a]];
            const nc_pm_params old_pm = c.crt_pm;
            const std::array< nc_st_params, 2 > old_st = c.crt_st;
            const float old_fit = c.crt_fit;
            c.crt_pm = tout[a].pm;
            c.crt_st = { { tout[a].st[0], tout[a].st[1] } };
            c.crt_fit = tout[a].fit;
            const Read& rd = *reads[c.read];
            NLOG(3, "scaling_round read [" << rd.read_id << "] strand [" << c.strand << "] model [" << c.key[0]
                        << (c.strand == 2 ? "+" : "") << c.key[1] << "] old_pm_params [" << old_pm << "] old_fit ["
                        << old_fit << "] crt_pm_params [" << c.crt_pm << "] crt_fit [" << c.crt_fit << "] round ["
                        << c.round << "]");
            if (tout[a].done) { c.active = false; continue; }  // singularity detected; stop
            if (c.crt_fit < old_fit)
            {
                NLOG(2, "scaling_regression read [" << rd.read_id << "] strand [" << c.strand << "] model [" << c.key[0]
                            << (c.strand == 2 ? "+" : "") << c.key[1] << "] old_params [" << old_pm << "] old_fit ["
                            << old_fit << "] crt_pm_params [" << c.crt_pm << "] crt_fit [" << c.crt_fit
                            << "] round [" << c.round << "]");
                c.crt_pm = old_pm;
                c.crt_st = old_st;
                c.crt_fit = old_fit;
                c.active = false;
                continue;
            }
            ++c.round;
            if (c.round >= c.max_rounds || (c.round > 1 && c.crt_fit < old_fit + opt_.scaling_min_progress)) c.active = false;
        }
    }
    // ---- results back into the reads, model selection (:427-459, :543-570)
    std::map< std::pair< size_t, unsigned >, std::vector< const Candidate* > > by_read;  // (read, strand tag) -> candidates in map order
    for (auto& c : cands)
    {
        Read& rd = *reads[c.read];
        rd.pm_params_m[c.key] = c.crt_pm;
        rd.st_params_m[c.key] = c.crt_st;
        if (c.strand == 2)
            NLOG(2, "scaling_result read [" << rd.read_id << "] strand [2] model [" << c.key[0] << "+" << c.key[1]
                        << "] pm_params [" << c.crt_pm << "] st_params [" << c.crt_st[0] << "," << c.crt_st[1]
                        << "] fit [" << c.crt_fit << "] rounds [" << c.round << "]");
        else
            NLOG(2, "scaling_result read [" << rd.read_id << "] strand [" << c.strand << "] model [" << c.key[c.strand]
                        << "] pm_params [" << c.crt_pm << "] st_params [" << c.crt_st[c.strand] << "] fit ["
                        << c.crt_fit << "] rounds [" << c.round << "]");
        by_read[std::make_pair(c.read, c.strand)].push_back(&c);
    }
    if (opt_.scaling_select_threshold < std::numeric_limits< float >::infinity())
    {
        for (auto& e : by_read)
        {
            // candidates were generated in the key order of the reference's std::map; alg::max_of keeps the first maximum
            std::vector< const Candidate* > v = e.second;
            std::sort(v.begin(), v.end(), [](const Candidate* a, const Candidate* b) { return a->key < b->key; });
            const Candidate* best = v[0];
            for (const Candidate* c : v)
                if (best->crt_fit < c->crt_fit) best = c;
            bool unique = true;
            for (const Candidate* c : v)
                if (c != best && !(c->crt_fit + opt_.scaling_select_threshold < best->crt_fit)) unique = false;
            if (!unique) continue;
            Read& rd = *reads[e.first.first];
            if (e.first.second == 2)
            {
                rd.preferred_model[2] = best->key;
                NLOG(2, "selected_model read [" << rd.read_id << "] strand [2] model [" << best->key[0] << "+" << best->key[1] << "]");
            }
            else
            {
                unsigned st = e.first.second;
                rd.preferred_model[st][st] = best->key[st];
                NLOG(2, "selected_model read [" << rd.read_id << "] strand [" << st << "] model [" << best->key[st] << "]");
            }
        }
    }
}

// ---------------------------------------------------------------- basecalling (nanocall.cpp:593-869)
namespace {
struct VitJob
{
    size_t read;
    unsigned strand;
    Model_Key key;
    size_t cand;  // candidate index within the read's list
};
} // namespace

void* Pipeline::Pinned::reserve(size_t bytes)
{
    if (bytes <= cap) return p;
    if (p) nc_host_free(p);
    cap = bytes + bytes / 4 + 4096;
    p = nc_host_alloc(cap);
    if (!p) { cap = 0; throw std::runtime_error("nc_host_alloc failed"); }
    return p;
}
Pipeline::Pinned::~Pinned() { if (p) nc_host_free(p); }


void Pipeline::basecall_reads(std::vector< Read* >& reads)
{
    const uint64_t max_events_per_call = 48ull << 20;
    size_t r0 = 0;
    while (r0 < reads.size())
    {
        // ---- pass 1 (serial, light): the jobs of the call = the (read, strand, candidate model) triples the reference's
        // lambda iterates (:692-733, :787-818), in its order
        std::vector< VitJob > jobs;
        std::vector< uint64_t > off(1, 0);
        std::vector< int32_t > mid;
        std::vector< nc_pm_params > pm;
        std::vector< nc_st_params > st;
        size_t r1 = r0;
        while (r1 < reads.size() && (r1 == r0 || off.back() < max_events_per_call))
        {
            Read& rd = *reads[r1];
            auto add_job = [&](unsigned s, const Model_Key& key, size_t cand) {
                const Strand_Events& ev = rd.events[s];
                if (ev.size() == 0) return;   // (the reference would index an empty sequence: Viterbi.hpp:60)
                const Model& m = models_.at(key[s]);
                const nc_pm_params& p = rd.pm_params_m.at(key);
                const nc_st_params& t = rd.st_params_m.at(key)[s];
                NLOG(2, "basecalling read [" << rd.read_id << "] strand [" << s << "] model [" << key[s]
                            << "] pm_params [" << p << "] st_params [" << t << "]");
                jobs.push_back(VitJob{ r1, s, key, cand });
                off.push_back(off.back() + ev.size());
                mid.push_back(m.id);
                pm.push_back(p);
                st.push_back(t);
            };
            if (rd.scale_strands_together && (rd.events[0].size() == 0 || rd.events[1].size() == 0))
            {
                // every event of a strand was filtered out after the read was put on the joint path: the reference
                // would run Viterbi on an empty sequence (undefined); the read is not called
                NLOG(1, "empty_strand read [" << rd.read_id << "]: not basecalled");
            }
            else if (rd.scale_strands_together)
            {
                std::vector< Model_Key > sub;  // :697-709
                if (!rd.preferred_model[2][0].empty()) sub.push_back(rd.preferred_model[2]);
                else
                    for (const auto& p : rd.pm_params_m)
                        if (!p.first[0].empty() && !p.first[1].empty()) sub.push_back(p.first);
                for (size_t c = 0; c < sub.size(); ++c)
                    for (unsigned s = 0; s < 2; ++s) add_job(s, sub[c], c);
            }
            else
            {
                for (unsigned s = 0; s < 2; ++s)
                {
                    if (rd.events[s].size() < opt_.min_ed_events) continue;
                    std::vector< Model_Key > sub;  // :791-806
                    if (!rd.preferred_model[s][s].empty()) sub.push_back(rd.preferred_model[s]);
                    else
                        for (const auto& p : rd.pm_params_m)
                            if (!p.first[s].empty() && p.first[1 - s].empty()) sub.push_back(p.first);
                    for (size_t c = 0; c < sub.size(); ++c) add_job(s, sub[c], c);
                }
            }
            ++r1;
        }
        const uint32_t nj = (uint32_t)jobs.size();
        const size_t total = off.back();
        std::vector< float > path(nj);
        float* mean = static_cast< float* >(pin_mean_.reserve(total * sizeof(float)));
        float* stdv = static_cast< float* >(pin_stdv_.reserve(total * sizeof(float)));
        float* start = static_cast< float* >(pin_start_.reserve(total * sizeof(float)));
        uint16_t* states = static_cast< uint16_t* >(pin_states_.reserve(total * sizeof(uint16_t)));
        uint8_t* moves = static_cast< uint8_t* >(pin_moves_.reserve(total));
        // ---- pass 2 (parallel over jobs): events into the pinned staging arrays; the means_apart check (:633-683):
        // mean of the scaled model's level means against the mean of the events
        std::vector< std::string > warn(nj);
        parallel_for(nj, opt_.host_threads, [&](size_t j) {
            const VitJob& vj = jobs[j];
            const Read& rd = *reads[vj.read];
            const Strand_Events& ev = rd.events[vj.strand];
            std::memcpy(mean + off[j], ev.mean.data(), ev.size() * sizeof(float));
            std::memcpy(stdv + off[j], ev.stdv.data(), ev.size() * sizeof(float));
            std::memcpy(start + off[j], ev.start.data(), ev.size() * sizeof(float));
            const Model& m = models_.at(vj.key[vj.strand]);
            const nc_pm_params& p = pm[j];
            float lv[NC_N_STATES];
            for (unsigned k = 0; k < NC_N_STATES; ++k) lv[k] = m.table[4 * k] * p.scale + p.shift;
            float mm, ms, em, es;
            nc_mean_stdv(NC_N_STATES, lv, &mm, &ms);
            nc_mean_stdv((uint32_t)ev.size(), ev.mean.data(), &em, &es);
            if (std::abs(em - mm) > 5.0 && opt_.log_level >= 1)
            {
                std::ostringstream o;
                o << "means_apart read [" << rd.read_id << "] strand [" << vj.strand << "] model [" << vj.key[vj.strand]
                  << "] parameters [" << p << "] model_mean=[" << mm << "] events_mean=[" << em << "]";
                warn[j] = o.str();
            }
        });
        for (const auto& w : warn)
            if (!w.empty()) std::clog << (w + "\n") << std::flush;
        if (nj)
        {
            std::lock_guard< std::mutex > gpu(gpu_mu_);
            const auto c0 = std::chrono::steady_clock::now();
            check(nc_viterbi_packed(ctx_, nj, off.data(), mean, stdv, start, nullptr, mid.data(),
                                    pm.data(), st.data(), NC_MEM_HOST, path.data(), states, moves),
                  "nc_viterbi_packed");
            viterbi_call_s += std::chrono::duration< double >(std::chrono::steady_clock::now() - c0).count();
            viterbi_kernel_ms += nc_ctx_last_kernel_ms(ctx_);
            viterbi_events += off.back();
        }
        // ---- pass 3 (parallel over reads): rank the candidates of a read (:711-750, :808-836), assemble the sequences
        auto seq_of = [&](size_t j) {
            uint32_t n = (uint32_t)(off[j + 1] - off[j]);
            uint32_t need = nc_base_seq(n, states + off[j], moves + off[j], nullptr, 0);
            std::string s(need, 'N');
            nc_base_seq(n, states + off[j], moves + off[j], &s[0], need);
            return s;
        };
        std::vector< std::pair< size_t, size_t > > spans;   // jobs [first, last) of each read that has any
        for (size_t j = 0; j < jobs.size();)
        {
            size_t je = j;
            while (je < jobs.size() && jobs[je].read == jobs[j].read) ++je;
            spans.push_back(std::make_pair(j, je));
            j = je;
        }
        std::vector< std::string > info(spans.size());
        parallel_for(spans.size(), opt_.host_threads, [&](size_t k) {
            const size_t j = spans[k].first, je = spans[k].second;
            Read& rd = *reads[jobs[j].read];
            std::ostringstream lg;
            if (rd.scale_strands_together)
            {
                // jobs come in (strand 0, strand 1) pairs per candidate; best = last of a stable ascending sort by the sum
                size_t best = j;
                float best_sum = path[j] + path[j + 1];
                for (size_t q = j + 2; q + 1 < je; q += 2)
                {
                    float s = path[q] + path[q + 1];
                    if (!(s < best_sum)) { best = q; best_sum = s; }
                }
                const Model_Key key = jobs[best].key;
                const nc_pm_params best_pm = rd.pm_params_m.at(key);
                const std::array< nc_st_params, 2 > best_st = rd.st_params_m.at(key);
                for (unsigned s = 0; s < 2; ++s)
                {
                    if (opt_.log_level >= 2)
                        lg << "best_model read [" << rd.read_id << "] strand [" << s << "] model [" << key[s] << "] pm_params ["
                           << best_pm << "] st_params [" << best_st[s] << "] log_path_prob [" << path[best + s] << "]\n";
                    rd.preferred_model[s][s] = key[s];
                    rd.pm_params_m[rd.preferred_model[s]] = best_pm;
                    rd.st_params_m[rd.preferred_model[s]][s] = best_st[s];
                    rd.base_seq[s] = seq_of(best + s);
                    rd.log_path_prob[s] = path[best + s];
                    rd.called[s] = true;
                }
            }
            else
            {
                for (unsigned s = 0; s < 2; ++s)
                {
                    size_t best = SIZE_MAX;
                    for (size_t q = j; q < je; ++q)
                        if (jobs[q].strand == s && (best == SIZE_MAX || !(path[q] < path[best]))) best = q;
                    if (best == SIZE_MAX) continue;
                    const Model_Key key = jobs[best].key;
                    if (opt_.log_level >= 2)
                        lg << "best_model read [" << rd.read_id << "] strand [" << s << "] model [" << key[s] << "] pm_params ["
                           << rd.pm_params_m.at(key) << "] st_params [" << rd.st_params_m.at(key)[s]
                           << "] log_path_prob [" << path[best] << "]\n";
                    rd.preferred_model[s][s] = key[s];
                    rd.base_seq[s] = seq_of(best);
                    rd.log_path_prob[s] = path[best];
                    rd.called[s] = true;
                }
            }
            info[k] = lg.str();
        });
        for (const auto& l : info)
            if (!l.empty()) std::clog << l << std::flush;
        r0 = r1;
    }
}

void Pipeline::write_fasta(std::ostream& os, const std::string& name, const std::string& seq, unsigned width)
{
    os << ">" << name << "\n";  // nanocall.cpp:584-591
    for (size_t pos = 0; pos < seq.size(); pos += width) os << seq.substr(pos, width) << "\n";
}

void Pipeline::write_output(std::ostream& os, const Read& r) const
{
    for (unsigned st = 0; st < 2; ++st)
    {
        if (!r.called[st]) continue;
        std::ostringstream name;
        name << r.read_id << ":" << r.base_file_name << ":" << st;  // :764-769
        write_fasta(os, name.str(), r.base_seq[st], opt_.fasta_line_width);
    }
}

void Pipeline::write_stats_header(std::ostream& os)
{
    // Fast5_Summary::write_tsv_header (Fast5_Summary.hpp:460-476) + the endl of nanocall.cpp:896
    os << "file_name\tread_name\tnum_ed_events\tabasic_level\ttemplate_start_idx\ttemplate_end_idx"
       << "\tcomplement_start_idx\tcomplement_end_idx";
    for (unsigned st = 0; st < 2; ++st)
        os << "\tn" << st << "_model_name\tn" << st << "_scale\tn" << st << "_shift\tn" << st << "_drift\tn" << st
           << "_var\tn" << st << "_scale_sd\tn" << st << "_var_sd\tn" << st << "_p_stay\tn" << st << "_p_skip";
    os << std::endl;
}

void Pipeline::write_stats(std::ostream& os, const Read& r, const Options& opt)
{
    // Fast5_Summary::write_tsv (:478-502).  The parameter columns switch the stream to fixed notation with five
    // decimals (Pore_Model.hpp:72-76) and the reference never switches back, so abasic_level is printed in the default
    // notation in the first row only: the stream state is deliberately left as the reference leaves it.
    os << r.base_file_name << '\t' << r.read_id << '\t' << r.num_ed_events << '\t' << r.abasic_level << '\t' << r.strand_bounds[0]
       << '\t' << r.strand_bounds[1] << '\t' << r.strand_bounds[2] << '\t' << r.strand_bounds[3];
    auto pm_tsv = [&](const nc_pm_params& p) {
        os << std::fixed << std::setprecision(5) << p.scale << '\t' << p.shift << '\t' << p.drift << '\t' << p.var << '\t'
           << p.scale_sd << '\t' << p.var_sd;
    };
    auto st_tsv = [&](const nc_st_params& p) { os << std::fixed << std::setprecision(5) << p.p_stay << '\t' << p.p_skip; };
    for (unsigned st = 0; st < 2; ++st)
    {
        os << '\t';
        if (!r.preferred_model[st][st].empty())
        {
            os << r.preferred_model[st][st] << '\t';
            pm_tsv(r.pm_params_m.at(r.preferred_model[st]));
            os << '\t';
            st_tsv(r.st_params_m.at(r.preferred_model[st])[st]);
        }
        else
        {
            os << ".\t";
            pm_tsv(default_pm());
            os << '\t';
            nc_st_params d;
            d.p_stay = opt.pr_stay;
            d.p_skip = opt.pr_skip;
            st_tsv(d);
        }
    }
    os << std::endl;
}

// ---------------------------------------------------------------- event tables
// Fast5_Summary::filter_ed_event (Fast5_Summary.hpp:734-745): the reference drops an eventdetection event whose mean
// reaches the abasic level or whose stdv exceeds 4 before it builds the strands' event sequences (:352-363).  Event
// tables are the input here, so the same rule is applied on load; the abasic level is known only when the table
// carries it ("#abasic_level <pA>"), otherwise that half of the rule is off.
// segmented tables carry no raw-event indices: the strand bounds are the offsets of the two strands in the
// concatenated table, the abasic level is unknown (0)
static void finish_segmented_read(Read& r)
{
    const unsigned n0 = (unsigned)r.events[0].size(), n1 = (unsigned)r.events[1].size();
    r.num_ed_events = n0 + n1;
    r.strand_bounds = { { 0u, n0, n0, n0 + n1 } };
}

static bool keep_event(float mean, float stdv, float abasic_level)
{
    if (mean >= abasic_level) return false;
    if (stdv > 4.0f) return false;
    return true;
}

bool load_events_tsv(const std::string& path, Read& r, std::string& err)
{
    // "#read_id <id>" header (optional), then rows "strand mean stdv start length"; the last four columns are what
    // Event::operator>> reads (Event.hpp:59-68).  start is in seconds from the strand start the reference would use
    // (Fast5_Summary.hpp:359).
    std::ifstream is(path);
    if (!is) { err = "cannot open " + path; return false; }
    auto pos = path.find_last_of('/');
    r.base_file_name = pos != std::string::npos ? path.substr(pos + 1) : path;
    for (const char* ext : { ".events.tsv", ".tsv", ".txt" })
    {
        std::string e(ext);
        if (r.base_file_name.size() > e.size() && r.base_file_name.compare(r.base_file_name.size() - e.size(), e.size(), e) == 0)
        {
            r.base_file_name.resize(r.base_file_name.size() - e.size());
            break;
        }
    }
    r.read_id = r.base_file_name;
    float abasic_level = std::numeric_limits< float >::infinity();
    std::string line;
    while (std::getline(is, line))
    {
        if (line.empty()) continue;
        if (line[0] == '#')
        {
            std::istringstream iss(line.substr(1));
            std::string k, v;
            iss >> k >> v;
            if (k == "read_id" && !v.empty()) r.read_id = v;
            if (k == "abasic_level" && !v.empty()) abasic_level = std::strtof(v.c_str(), nullptr);
            continue;
        }
        std::istringstream iss(line);
        unsigned st;
        float mean, stdv, start, length;
        if (!(iss >> st >> mean >> stdv >> start >> length) || st > 1) { err = "bad event row in " + path + ": " + line; return false; }
        if (!keep_event(mean, stdv, abasic_level)) continue;
        Strand_Events& ev = r.events[st];
        ev.mean.push_back(mean); ev.stdv.push_back(stdv); ev.start.push_back(start); ev.length.push_back(length);
    }
    finish_segmented_read(r);
    return true;
}

bool load_events_ncev(const std::string& path, std::vector< Read >& reads, std::string& err)
{
    // binary container: "NCEV0001", u32 n_reads, then per read: u32 id_len, id, u32 n0, u32 n1,
    // and for each strand mean[n] stdv[n] start[n] length[n] as float32
    std::ifstream is(path, std::ios::binary);
    if (!is) { err = "cannot open " + path; return false; }
    char magic[8];
    uint32_t n_reads = 0;
    is.read(magic, 8);
    is.read(reinterpret_cast< char* >(&n_reads), 4);
    if (!is || std::memcmp(magic, "NCEV0001", 8) != 0) { err = path + " is not an NCEV0001 file"; return false; }
    auto pos = path.find_last_of('/');
    std::string base = pos != std::string::npos ? path.substr(pos + 1) : path;
    if (base.size() > 5 && base.compare(base.size() - 5, 5, ".ncev") == 0) base.resize(base.size() - 5);
    for (uint32_t k = 0; k < n_reads; ++k)
    {
        Read r;
        uint32_t id_len = 0, n[2] = { 0, 0 };
        is.read(reinterpret_cast< char* >(&id_len), 4);
        if (!is || id_len > 4096) { err = "corrupt read header in " + path; return false; }
        r.read_id.resize(id_len);
        is.read(&r.read_id[0], id_len);
        is.read(reinterpret_cast< char* >(n), 8);
        r.base_file_name = base;
        for (int st = 0; st < 2; ++st)
        {
            Strand_Events& ev = r.events[st];
            for (std::vector< float >* v : { &ev.mean, &ev.stdv, &ev.start, &ev.length })
            {
                v->resize(n[st]);
                is.read(reinterpret_cast< char* >(v->data()), (std::streamsize)n[st] * sizeof(float));
            }
        }
        if (!is) { err = "truncated " + path; return false; }
        for (int st = 0; st < 2; ++st)   // the stdv half of filter_ed_event (the container carries no abasic level)
        {
            Strand_Events& ev = r.events[st];
            size_t w = 0;
            for (size_t i = 0; i < ev.mean.size(); ++i)
                if (keep_event(ev.mean[i], ev.stdv[i], std::numeric_limits< float >::infinity()))
                {
                    ev.mean[w] = ev.mean[i]; ev.stdv[w] = ev.stdv[i]; ev.start[w] = ev.start[i]; ev.length[w] = ev.length[i];
                    ++w;
                }
            for (std::vector< float >* v : { &ev.mean, &ev.stdv, &ev.start, &ev.length }) v->resize(w);
        }
        finish_segmented_read(r);
        reads.push_back(std::move(r));
    }
    return true;
}

} // namespace nchost
