// See reads.hpp.  summarize_raw_read restates Fast5_Summary::summarize (Fast5_Summary.hpp:138-319) and its helpers
// line by line; the arithmetic types (float members, double event entries, unsigned bounds) are the reference's.
#include "reads.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <mutex>
#include <random>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace nchost {

namespace {

typedef std::vector< std::pair< unsigned, unsigned > > Islands;

std::string islands_str(const Islands& v)
{
    std::ostringstream os;
    for (size_t i = 0; i < v.size(); ++i) os << (i ? " " : "") << "[" << v[i].first << "," << v[i].second << "]";
    return os.str();
}

// Fast5_Summary::find_islands_5_consec (:545-571): runs of >= 5 consecutive events at or above the abasic level
Islands find_islands_5_consec(const Ed_Event* ed, size_t n_ed, float abasic_level)
{
    Islands islands;
    unsigned i = 0;
    while (i < n_ed)
    {
        if (ed[i].mean >= abasic_level)
        {
            unsigned j = i + 1;
            while (j < n_ed && ed[j].mean >= abasic_level) ++j;
            if (j - i >= 5) islands.push_back(std::make_pair(i, j));
            i = j + 1;
        }
        else ++i;
    }
    return islands;
}

// Fast5_Summary::detect_strands (:653-731)
void detect_strands(const Options& opt, const Ed_Event* ed, size_t n_ed, float abasic_level, const std::string& read_id,
                    std::array< unsigned, 4 >& sb)
{
    const auto& tm = opt.trim_margins;
    Islands islands = find_islands_5_consec(ed, n_ed, abasic_level);
    for (unsigned i = 1; i < islands.size(); ++i)
    {
        if (islands[i - 1].second + std::max(tm[2], tm[3]) >= islands[i].first)
        {
            islands[i - 1].second = islands[i].second;
            islands.erase(islands.begin() + i);
            i = 0;
        }
    }
    if (opt.log_level >= 3) log_line(3, opt.log_level, "final_islands: " + islands_str(islands));
    if (islands.empty())
    {
        log_line(2, opt.log_level, "template_only read_id=[" + read_id + "]");
        return;
    }
    auto dist_to_middle = [&](const std::pair< unsigned, unsigned >& p) {
        return std::min((unsigned)std::abs((long)p.first - (long)n_ed / 2),
                        (unsigned)std::abs((long)p.second - (long)n_ed / 2));
    };
    auto it = islands.begin();   // alg::min_of: the first minimum
    for (auto jt = islands.begin() + 1; jt != islands.end(); ++jt)
        if (dist_to_middle(*jt) < dist_to_middle(*it)) it = jt;
    if (dist_to_middle(*it) > n_ed / 6)
    {
        log_line(2, opt.log_level, "drop_read read_id=[" + read_id + "] islands=[" + islands_str(islands) + "]");
        return;
    }
    sb[0] = tm[0];
    if (islands[0].first < tm[0] + tm[2]) sb[0] = std::max(sb[0], islands[0].second);
    sb[1] = it->first - tm[2];
    sb[2] = it->first + tm[3];
    sb[3] = (unsigned)n_ed - tm[1];
    if (islands[islands.size() - 1].second > n_ed - (tm[3] + tm[1])) sb[3] = std::min(sb[3], islands[islands.size() - 1].first);
}

std::string base_name_of(const std::string& file_name)
{
    auto pos = file_name.find_last_of('/');
    std::string b = pos != std::string::npos ? file_name.substr(pos + 1) : file_name;
    if (b.size() >= 6 && b.compare(b.size() - 6, 6, ".fast5") == 0) b.resize(b.size() - 6);
    return b;
}

} // namespace

bool summarize_raw_read(const Options& opt, Raw_Read&& raw, Read& r, std::string& why)
{
    return summarize_events(opt, raw.file_name, raw.read_id, raw.sampling_rate, raw.ed.data(), raw.ed.size(), r, why);
}

// (the events are only read: the synthetic source summarises its pool reads in place, without a copy per replay)
bool summarize_events(const Options& opt, const std::string& fn, const std::string& raw_read_id, double raw_sampling_rate,
                      const Ed_Event* ed, size_t n_raw, Read& r, std::string& why)
{
    r = Read();
    r.base_file_name = base_name_of(fn);
    r.read_id = r.base_file_name;
    if (!(raw_sampling_rate > 0)) { why = fn + ": missing sampling rate"; return false; }
    r.sampling_rate = (float)raw_sampling_rate;
    if (r.sampling_rate < 1000.0 || r.sampling_rate > 10000.0)
    {
        std::ostringstream os;
        os << fn << ": unexpected sampling rate: " << r.sampling_rate;
        why = os.str();
        return false;
    }
    if (!raw_read_id.empty()) r.read_id = raw_read_id;
    // load_ed_events (:505-525)
    if (n_raw > opt.max_ed_events)
    {
        std::ostringstream os;
        os << fn << ": using only " << opt.max_ed_events << " of " << n_raw << " events";
        log_line(2, opt.log_level, os.str());
        r.num_ed_events = opt.max_ed_events;
    }
    else r.num_ed_events = (unsigned)n_raw;
    const size_t n_ed = r.num_ed_events;
    const auto& tm = opt.trim_margins;
    if (r.num_ed_events < tm[0] + tm[1] + opt.min_ed_events)
    {
        std::ostringstream os;
        os << fn << ": not enough eventdetection events: " << r.num_ed_events;
        why = os.str();
        r.num_ed_events = 0;
        return false;
    }
    // detect_abasic_level (:527-543): the level below the top percent of the event means, plus the preset's offset.
    // (the reference sorts; the order statistic is the same)
    {
        std::vector< float > s(n_ed);
        float lo = std::numeric_limits< float >::infinity(), hi = -lo;
        for (size_t i = 0; i < n_ed; ++i)
        {
            s[i] = (float)ed[i].mean;
            lo = std::min(lo, s[i]);
            hi = std::max(hi, s[i]);
        }
        size_t k = (size_t)((double)s.size() * (1.0 - opt.abasic_level_top_percent / 100.0));
        if (k >= s.size()) k = s.size() - 1;   // (top percent 0 indexes past the end in the reference)
        // The k-th smallest mean, exactly, in two light passes instead of a selection over the whole read: a histogram over
        // [lo, hi] finds the bin that holds it, the selection then runs inside that bin (a hundredth of the events).  A
        // value's bin is a monotone function of the value, so order statistics carry over; NaN or a flat read take the
        // plain selection.
        float kth;
        constexpr int NB = 1024;
        if (hi > lo && n_ed >= 4 * NB)
        {
            const float scale = (float)(NB - 1) / (hi - lo);
            auto bin = [&](float v) { int b = (int)((v - lo) * scale); return b < 0 ? 0 : (b >= NB ? NB - 1 : b); };
            unsigned cnt[NB] = { 0 };
            for (size_t i = 0; i < n_ed; ++i) ++cnt[bin(s[i])];
            size_t below = 0;
            int b = 0;
            while (below + cnt[b] <= k) below += cnt[b++];
            std::vector< float > in;
            in.reserve(cnt[b]);
            for (size_t i = 0; i < n_ed; ++i)
                if (bin(s[i]) == b) in.push_back(s[i]);
            std::nth_element(in.begin(), in.begin() + (k - below), in.end());
            kth = in[k - below];
        }
        else
        {
            std::nth_element(s.begin(), s.begin() + k, s.end());
            kth = s[k];
        }
        r.abasic_level = (float)(kth + opt.abasic_level_top_offset);
    }
    if (r.abasic_level <= 1.0)
    {
        std::ostringstream os;
        os << fn << ": abasic level too low: " << r.abasic_level;
        why = os.str();
        r.num_ed_events = 0;
        return false;
    }
    r.strand_bounds = { { tm[0], r.num_ed_events - tm[1], 0u, 0u } };
    if (!opt.template_only) detect_strands(opt, ed, n_ed, r.abasic_level, r.read_id, r.strand_bounds);
    const auto& sb = r.strand_bounds;
    if (sb[1] <= sb[0])
    {
        why = fn + ": no template strand detected";
        r.num_ed_events = 0;
        return false;
    }
    // decided on the raw bounds, before the event filter runs (:210-212)
    r.scale_strands_together = opt.double_strand_scaling && sb[1] - sb[0] >= opt.min_ed_events && sb[3] - sb[2] >= opt.min_ed_events;
    // load_events (:348-364) with filter_ed_event (:734-745)
    for (unsigned st = 0; st < 2; ++st)
    {
        Strand_Events& ev = r.events[st];
        const unsigned b0 = sb[2 * st], b1 = sb[2 * st + 1];
        if (b1 <= b0) continue;
        ev.mean.resize(b1 - b0); ev.stdv.resize(b1 - b0); ev.start.resize(b1 - b0); ev.length.resize(b1 - b0);
        const long long t0 = ed[sb[r.scale_strands_together ? 0 : 2 * st]].start;
        size_t w = 0;   // (written unconditionally, kept by advancing w: no push_back per field and event)
        for (unsigned j = b0; j < b1; ++j)
        {
            const Ed_Event& e = ed[j];
            ev.mean[w] = (float)e.mean;
            ev.stdv[w] = (float)e.stdv;
            ev.start[w] = (float)(e.start - t0) / r.sampling_rate;
            ev.length[w] = (float)e.length / r.sampling_rate;
            w += (e.mean >= r.abasic_level || e.stdv > 4.0) ? 0 : 1;
        }
        ev.mean.resize(w); ev.stdv.resize(w); ev.start.resize(w); ev.length.resize(w);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ NCRW0001
// "NCRW0001", u32 n_reads, then per read: u32 id_len, id, f64 sampling_rate, u32 n_events, n_events x
// { f64 mean, f64 stdv, i64 start, i64 length }.  A file with one record stands in for one fast5 file (the oracle's
// fast5::File reads the first record of the same format, oracle/stub_full/fast5.hpp).
bool is_ncrw_file(const std::string& path)
{
    std::ifstream is(path, std::ios::binary);
    char magic[8];
    return is.read(magic, 8) && std::memcmp(magic, "NCRW0001", 8) == 0;
}

static bool read_ncrw_record(std::istream& is, Raw_Read& r)
{
    uint32_t id_len = 0, n = 0;
    is.read(reinterpret_cast< char* >(&id_len), 4);
    if (!is || id_len > 4096) return false;
    r.read_id.resize(id_len);
    is.read(&r.read_id[0], id_len);
    is.read(reinterpret_cast< char* >(&r.sampling_rate), 8);
    is.read(reinterpret_cast< char* >(&n), 4);
    if (!is) return false;
    static_assert(sizeof(Ed_Event) == 32, "NCRW0001 record layout");
    r.ed.resize(n);
    is.read(reinterpret_cast< char* >(r.ed.data()), (std::streamsize)n * sizeof(Ed_Event));
    return (bool)is;
}

bool read_ncrw_file(const std::string& path, std::vector< Raw_Read >& out, std::string& err)
{
    std::ifstream is(path, std::ios::binary);
    char magic[8];
    uint32_t n_reads = 0;
    if (!is.read(magic, 8) || std::memcmp(magic, "NCRW0001", 8) != 0) { err = path + " is not an NCRW0001 file"; return false; }
    is.read(reinterpret_cast< char* >(&n_reads), 4);
    for (uint32_t k = 0; k < n_reads; ++k)
    {
        Raw_Read r;
        if (!read_ncrw_record(is, r)) { err = "truncated " + path; return false; }
        r.file_name = path;
        out.push_back(std::move(r));
    }
    return true;
}

void write_ncrw_file(const std::string& path, const std::vector< Raw_Read >& reads)
{
    std::ofstream os(path, std::ios::binary);
    os.write("NCRW0001", 8);
    uint32_t n = (uint32_t)reads.size();
    os.write(reinterpret_cast< const char* >(&n), 4);
    for (const auto& r : reads)
    {
        uint32_t id_len = (uint32_t)r.read_id.size(), ne = (uint32_t)r.ed.size();
        os.write(reinterpret_cast< const char* >(&id_len), 4);
        os.write(r.read_id.data(), id_len);
        os.write(reinterpret_cast< const char* >(&r.sampling_rate), 8);
        os.write(reinterpret_cast< const char* >(&ne), 4);
        os.write(reinterpret_cast< const char* >(r.ed.data()), (std::streamsize)ne * sizeof(Ed_Event));
    }
}

// ------------------------------------------------------------------------------------------------ file source
namespace {

bool has_ext(const std::string& p, const char* e)
{
    size_t n = std::strlen(e);
    return p.size() > n && p.compare(p.size() - n, n, e) == 0;
}

void finish_summary(const Options& opt, bool ok, const std::string& why, const Read& r)
{
    if (!ok && !why.empty()) log_line(why.find("unexpected sampling rate") != std::string::npos ? 1 : 2, opt.log_level, why);
    if (opt.log_level < 2) return;
    // "summary: " << Fast5_Summary (nanocall.cpp:270, Fast5_Summary.hpp:439-458)
    std::ostringstream os;
    os << "summary: [base_file_name=" << r.base_file_name << " valid=1 num_ed_events=" << r.num_ed_events;
    if (r.num_ed_events > 0)
    {
        float tl[2] = { 0.f, 0.f };
        for (unsigned st = 0; st < 2; ++st)
            if (r.events[st].size() >= opt.min_ed_events) tl[st] = r.events[st].start.back() + r.events[st].length.back();
        os << " read_id=" << r.read_id << " abasic_level=" << r.abasic_level << " strand_bounds=[" << r.strand_bounds[0] << ","
           << r.strand_bounds[1] << "," << r.strand_bounds[2] << "," << r.strand_bounds[3] << "] time_length=[" << tl[0] << ","
           << tl[1] << "]";
    }
    os << "]";
    log_line(2, opt.log_level, os.str());
}

class File_Source : public Read_Source
{
public:
    File_Source(const Options& opt, const std::vector< std::string >& files, int log_level)
        : opt_(opt), files_(files), log_level_(log_level) {}

    bool next(Read& r, size_t& index) override
    {
        Raw_Read raw;
        bool have_raw = false;
        {
            std::lock_guard< std::mutex > lock(mu_);
            for (;;)
            {
                if (!pending_.empty())
                {
                    r = std::move(pending_.back());
                    pending_.pop_back();
                    index = next_index_++;
                    // segmented tables: the strand bounds ARE the event counts (Fast5_Summary.hpp:210-212)
                    r.scale_strands_together = opt_.double_strand_scaling && r.events[0].size() >= opt_.min_ed_events
                        && r.events[1].size() >= opt_.min_ed_events;
                    return true;
                }
                if (ncrw_left_ > 0)
                {
                    if (!read_ncrw_record(ncrw_, raw)) throw std::runtime_error("truncated " + files_[file_ - 1]);
                    raw.file_name = files_[file_ - 1];
                    if (ncrw_total_ > 1)
                    {
                        // containers of many reads: the read id names the read, the container the file
                        if (raw.read_id.empty()) raw.read_id = "read" + std::to_string(ncrw_total_ - ncrw_left_);
                    }
                    --ncrw_left_;
                    index = next_index_++;
                    have_raw = true;
                    break;
                }
                if (file_ >= files_.size()) return false;
                const std::string& f = files_[file_++];
                if (has_ext(f, ".ncev"))
                {
                    std::vector< Read > v;
                    std::string err;
                    if (!load_events_ncev(f, v, err)) throw std::runtime_error(err);
                    for (auto it = v.rbegin(); it != v.rend(); ++it) pending_.push_back(std::move(*it));
                }
                else if (has_ext(f, ".events.tsv"))
                {
                    Read one;
                    std::string err;
                    if (!load_events_tsv(f, one, err)) throw std::runtime_error(err);
                    pending_.push_back(std::move(one));
                }
                else
                {
                    ncrw_.close();
                    ncrw_.clear();
                    ncrw_.open(f, std::ios::binary);
                    char magic[8];
                    uint32_t n = 0;
                    if (!ncrw_.read(magic, 8) || std::memcmp(magic, "NCRW0001", 8) != 0) throw std::runtime_error(f + " is not an event table");
                    ncrw_.read(reinterpret_cast< char* >(&n), 4);
                    ncrw_left_ = ncrw_total_ = n;
                }
            }
        }
        if (have_raw)
        {
            std::string why;
            const bool ok = summarize_raw_read(opt_, std::move(raw), r, why);
            finish_summary(opt_, ok, why, r);
        }
        return true;
    }

private:
    Options opt_;
    std::vector< std::string > files_;
    int log_level_;
    std::mutex mu_;
    size_t file_ = 0, next_index_ = 0;
    std::ifstream ncrw_;
    uint32_t ncrw_left_ = 0, ncrw_total_ = 0;
    std::vector< Read > pending_;   // segmented reads of the current container, last first
};

// ------------------------------------------------------------------------------------------------ synthetic source
struct Synth_Model
{
    std::vector< float > table;   // 4096 x {level_mean, level_stdv, sd_mean, sd_stdv}
};

class Synth_Source : public Read_Source
{
public:
    Synth_Source(const Options& opt, const std::string& spec, const std::string& data_dir, int log_level) : opt_(opt)
    {
        // spec = n_reads[:seed[:pool[:shape...]]]
        std::vector< std::string > f;
        {
            std::istringstream is(spec);
            std::string tok;
            while (std::getline(is, tok, ':')) f.push_back(tok);
        }
        if (f.empty() || f[0].empty()) throw std::runtime_error("--synth needs n_reads[:seed[:pool[:shape]]]");
        n_reads_ = std::stoull(f[0]);
        seed_ = f.size() > 1 && !f[1].empty() ? std::stoull(f[1]) : 1;
        size_t pool = f.size() > 2 && !f[2].empty() ? std::stoull(f[2]) : 2048;
        shape_ = f.size() > 3 ? f[3] : "2d";
        nt_ = f.size() > 4 ? std::stoul(f[4]) : 5000;
        nc_ = f.size() > 5 ? std::stoul(f[5]) : (shape_ == "2d" ? 5000 : 0);
        if (shape_ != "2d" && shape_ != "1d" && shape_ != "mix") throw std::runtime_error("--synth shape must be 2d, 1d or mix");
        pool = std::max< size_t >(1, std::min(pool, n_reads_));
        load_models(data_dir);
        pool_.resize(pool);
        unsigned nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector< std::thread > th;
        std::atomic< size_t > nxt(0);
        for (unsigned t = 0; t < nth; ++t)
            th.emplace_back([&] {
                for (size_t k; (k = nxt++) < pool_.size();) pool_[k] = make_raw(k);
            });
        for (auto& t : th) t.join();
        size_t ev = 0;
        for (const auto& p : pool_) ev += p.ed.size();
        std::ostringstream os;
        os << "synthetic source: " << n_reads_ << " reads replayed from a pool of " << pool_.size() << " distinct reads (" << ev
           << " raw events), shape " << shape_;
        log_line(2, log_level, os.str());
    }

    bool next(Read& r, size_t& index) override
    {
        const size_t k = counter_++;
        if (k >= n_reads_) return false;
        index = k;
        const Raw_Read& raw = pool_[k % pool_.size()];
        char id[32];
        std::snprintf(id, sizeof id, "synth%08zu", k);
        std::string why;
        const bool ok = summarize_events(opt_, "synth.fast5", id, raw.sampling_rate, raw.ed.data(), raw.ed.size(), r, why);
        if (!ok && !why.empty()) log_line(2, opt_.log_level, why);
        return true;
    }
    size_t size_hint() const override { return n_reads_; }

private:
    void load_models(const std::string& dir)
    {
        std::ifstream names(dir + "/builtin_models.txt");
        std::ifstream blob(dir + "/builtin_models.bin", std::ios::binary);
        if (!names || !blob) throw std::runtime_error("builtin model data not found under " + dir);
        std::string name;
        int strand;
        while (names >> name >> strand)
        {
            std::vector< float > t(4 * NC_N_STATES);
            blob.read(reinterpret_cast< char* >(t.data()), t.size() * sizeof(float));
            if (!blob) throw std::runtime_error("builtin_models.bin is truncated");
            if (name.compare(0, 4, "r73.") != 0) continue;
            if (strand == 0) tmpl_.table = t;
            else comp_.push_back(Synth_Model{ t });
        }
        if (tmpl_.table.empty() || comp_.empty()) throw std::runtime_error("r73 builtin models not found under " + dir);
    }

    // one strand's events appended to ed: stay/step/skip walk over a random base stream, Gaussian level, inverse-Gaussian
    // stdv, exponential durations (the generator of nanocall_b200/synth.py)
    void add_strand(std::mt19937_64& g, const Synth_Model& M, unsigned n, const float pm[6], long long t0, long long& clock,
                    std::vector< Ed_Event >& ed) const
    {
        std::uniform_real_distribution< double > U(0.0, 1.0);
        std::normal_distribution< double > N01(0.0, 1.0);
        std::exponential_distribution< double > EXPD(1.0 / 0.02);
        unsigned state = (unsigned)(g() & 4095u);
        for (unsigned i = 0; i < n; ++i)
        {
            if (i)
            {
                const double u = U(g);
                if (u >= 0.1)
                {
                    state = ((state << 2) | (unsigned)(g() & 3u)) & 4095u;
                    if (u >= 0.7) state = ((state << 2) | (unsigned)(g() & 3u)) & 4095u;
                }
            }
            const float* row = M.table.data() + 4 * state;
            const double mu = row[0], sigma = row[1], eta = row[2], sd = row[3];
            const double lam = eta * eta * eta / (sd * sd);
            const double t = (double)(clock - t0) / 5000.0;
            Ed_Event e;
            e.mean = (double)(float)(pm[0] * mu + pm[1] + pm[2] * t + pm[3] * sigma * N01(g));
            // inverse Gaussian(mean m, shape l) by Michael, Schucany and Haas
            const double m = pm[4] * eta, l = pm[5] * lam;
            const double y = N01(g), y2 = y * y;
            double x = m + m * m * y2 / (2 * l) - m / (2 * l) * std::sqrt(4 * m * l * y2 + m * m * y2 * y2);
            if (U(g) > m / (m + x)) x = m * m / x;
            e.stdv = (double)(float)std::min(4.0, std::max(1e-3, x));
            e.start = clock;
            e.length = std::max< long long >(10, (long long)std::llround(EXPD(g) * 5000.0));
            clock += e.length;
            ed.push_back(e);
        }
    }

    Raw_Read make_raw(size_t k) const
    {
        std::mt19937_64 g(seed_ * 0x9E3779B97F4A7C15ull + k * 0xD1B54A32D192ED03ull + 12345u);
        std::uniform_real_distribution< double > U(0.0, 1.0);
        float pm[6] = { (float)(0.9 + 0.2 * U(g)), (float)(-5 + 10 * U(g)), (float)(-0.005 + 0.01 * U(g)),
                        (float)(0.9 + 0.4 * U(g)), (float)(0.8 + 0.4 * U(g)), (float)(0.8 + 0.7 * U(g)) };
        unsigned nt = nt_, nc = nc_;
        if (shape_ == "mix")
        {
            // BASELINE.json configs[4]: 90 % LogNormal(median 5000, sigma 0.5) in [500, 20000), 9 % 20k-50k, 1 % 100k-150k
            const double u = U(g);
            std::normal_distribution< double > N01(0.0, 1.0);
            if (u < 0.90) nt = (unsigned)std::min(19999.0, std::max(500.0, std::exp(std::log(5000.0) + 0.5 * N01(g))));
            else if (u < 0.99) nt = (unsigned)(20000 + 30000 * U(g));
            else nt = (unsigned)(100000 + 50000 * U(g));
            nc = 0;
        }
        if (shape_ == "1d") nc = 0;
        Raw_Read r;
        r.sampling_rate = 5000.0;
        const unsigned lead = 60, tail = 60;
        r.ed.reserve(lead + nt + nc + tail + 16);
        long long clock = 1000;
        const float ident[6] = { 1.f, 0.f, 0.f, 1.f, 1.f, 1.f };
        add_strand(g, tmpl_, lead, ident, clock, clock, r.ed);
        const long long t0 = clock;
        add_strand(g, tmpl_, nt, pm, t0, clock, r.ed);
        if (nc)
        {
            std::normal_distribution< double > N01(0.0, 1.0);
            for (unsigned i = 0; i < 8; ++i)   // the hairpin: an island of abasic-level events
            {
                Ed_Event e;
                e.mean = (double)(float)(115.0 + 2.0 * N01(g));
                e.stdv = (double)(float)(1.0 + 0.2 * U(g));
                e.start = clock;
                e.length = 100;
                clock += e.length;
                r.ed.push_back(e);
            }
            add_strand(g, comp_[k % comp_.size()], nc, pm, t0, clock, r.ed);
        }
        add_strand(g, nc ? comp_[k % comp_.size()] : tmpl_, tail, ident, clock, clock, r.ed);
        return r;
    }

    Options opt_;
    size_t n_reads_ = 0;
    uint64_t seed_ = 1;
    std::string shape_;
    unsigned nt_ = 5000, nc_ = 5000;
    Synth_Model tmpl_;
    std::vector< Synth_Model > comp_;
    std::vector< Raw_Read > pool_;
    std::atomic< size_t > counter_{ 0 };
};

} // namespace

std::unique_ptr< Read_Source > make_file_source(const Options& opt, const std::vector< std::string >& files, int log_level)
{
    return std::unique_ptr< Read_Source >(new File_Source(opt, files, log_level));
}

std::unique_ptr< Read_Source > make_synth_source(const Options& opt, const std::string& spec, const std::string& data_dir, int log_level)
{
    return std::unique_ptr< Read_Source >(new Synth_Source(opt, spec, data_dir, log_level));
}

} // namespace nchost
