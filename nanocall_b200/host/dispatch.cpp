// See dispatch.hpp.
#include "dispatch.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace nchost {

namespace {

typedef std::chrono::steady_clock Clock;
double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration< double >(b - a).count(); }

struct Item
{
    size_t index;
    Read read;
    size_t events() const { return read.events[0].size() + read.events[1].size(); }
};

// bounded, index-ordered queue between the loaders and the dispatchers
class Read_Queue
{
public:
    Read_Queue(size_t cap_events, size_t n_workers) : cap_events_(cap_events), n_workers_(n_workers) {}

    void push(Item&& it)
    {
        std::unique_lock< std::mutex > lk(mu_);
        // the lowest outstanding read is always admitted: a full queue of later reads must not block it
        space_.wait(lk, [&] { return events_ < cap_events_ || q_.empty() || it.index < q_.begin()->first || failed_; });
        events_ += it.events();
        const size_t idx = it.index;
        q_.emplace(idx, std::move(it));
        ready_.notify_all();
    }
    void loader_done()
    {
        std::lock_guard< std::mutex > lk(mu_);
        if (--loaders_ == 0) ready_.notify_all();
    }
    void set_loaders(unsigned n) { loaders_ = n; }
    void fail()
    {
        std::lock_guard< std::mutex > lk(mu_);
        failed_ = true;
        ready_.notify_all();
        space_.notify_all();
    }
    // the lowest-numbered reads, up to max_reads / max_events; empty = nothing left.  While loaders are still running
    // a dispatcher waits for a full batch; once the input is exhausted the rest is shared evenly between dispatchers.
    std::vector< Item > pop_batch(size_t max_reads, size_t max_events)
    {
        std::unique_lock< std::mutex > lk(mu_);
        ready_.wait(lk, [&] {
            return failed_ || loaders_ == 0 || q_.size() >= max_reads || events_ >= max_events;
        });
        std::vector< Item > out;
        if (failed_ || q_.empty()) return out;
        size_t want = max_reads;
        if (loaders_ == 0) want = std::min(want, std::max< size_t >(1, (q_.size() + n_workers_ - 1) / n_workers_));
        size_t ev = 0;
        while (!q_.empty() && out.size() < want && (out.empty() || ev + q_.begin()->second.events() <= max_events))
        {
            ev += q_.begin()->second.events();
            out.push_back(std::move(q_.begin()->second));
            q_.erase(q_.begin());
        }
        events_ -= ev;
        space_.notify_all();
        return out;
    }

private:
    std::mutex mu_;
    std::condition_variable ready_, space_;
    std::map< size_t, Item > q_;
    size_t events_ = 0, cap_events_, n_workers_;
    unsigned loaders_ = 0;
    bool failed_ = false;
};

// results are written in input order (pfor.hpp:216-235)
class Ordered_Sink
{
public:
    Ordered_Sink(const Options& opt, std::ostream* fasta, std::ostream* stats) : opt_(opt), fasta_(fasta), stats_(stats)
    {
        if (stats_) Pipeline::write_stats_header(*stats_);
    }
    void put(std::vector< Item >&& batch)
    {
        std::lock_guard< std::mutex > lk(mu_);
        for (auto& it : batch)
        {
            const size_t idx = it.index;
            done_.emplace(idx, std::move(it));
        }
        while (!done_.empty() && done_.begin()->first == next_)
        {
            const Read& r = done_.begin()->second.read;
            if (fasta_)
                for (unsigned st = 0; st < 2; ++st)
                    if (r.called[st])
                        Pipeline::write_fasta(*fasta_, r.read_id + ":" + r.base_file_name + ":" + std::to_string(st), r.base_seq[st],
                                              opt_.fasta_line_width);
            if (stats_) Pipeline::write_stats(*stats_, r, opt_);
            done_.erase(done_.begin());
            ++next_;
        }
        last_write_ = Clock::now();
    }
    size_t written() const { return next_; }
    Clock::time_point last_write() const { return last_write_; }

private:
    Options opt_;
    std::ostream* fasta_;
    std::ostream* stats_;
    std::mutex mu_;
    std::map< size_t, Item > done_;
    size_t next_ = 0;
    Clock::time_point last_write_ = Clock::now();
};

// one trained batch on its way from a dispatcher's training thread to its basecalling thread
class Hand_Over
{
public:
    bool give(std::vector< Item >&& b)   // false: the other side failed
    {
        std::unique_lock< std::mutex > lk(mu_);
        cv_.wait(lk, [&] { return !full_ || failed_; });
        if (failed_) return false;
        slot_ = std::move(b);
        full_ = true;
        cv_.notify_all();
        return true;
    }
    bool take(std::vector< Item >& b)    // false: nothing will come any more
    {
        std::unique_lock< std::mutex > lk(mu_);
        cv_.wait(lk, [&] { return full_ || closed_ || failed_; });
        if (!full_) return false;
        b = std::move(slot_);
        slot_.clear();
        full_ = false;
        cv_.notify_all();
        return true;
    }
    void close() { std::lock_guard< std::mutex > lk(mu_); closed_ = true; cv_.notify_all(); }
    void fail() { std::lock_guard< std::mutex > lk(mu_); failed_ = true; cv_.notify_all(); }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector< Item > slot_;
    bool full_ = false, closed_ = false, failed_ = false;
};

} // namespace

bool run_pipeline(const Run_Config& cfg, Read_Source& src, std::ostream* fasta, std::ostream* stats_tsv, Run_Stats& stats)
{
    const auto t_start = Clock::now();
    const size_t n_dev = cfg.devices.size();
    if (n_dev == 0) { stats.error = "no device given"; return false; }
    stats.dev.assign(n_dev, Device_Stats());
    Read_Queue queue(cfg.queue_events, n_dev);
    Ordered_Sink sink(cfg.opt, fasta, stats_tsv);
    std::mutex err_mu;
    std::string error;
    auto fail = [&](const std::string& what) {
        {
            std::lock_guard< std::mutex > lk(err_mu);
            if (error.empty()) error = what;
        }
        queue.fail();
    };

    unsigned n_load = cfg.loader_threads;
    if (n_load == 0)
    {
        unsigned hc = std::thread::hardware_concurrency();
        n_load = std::max(1u, std::min(32u, hc > n_dev + 1 ? hc - (unsigned)n_dev - 1 : 1u));
    }
    queue.set_loaders(n_load);
    std::atomic< size_t > n_reads(0), n_events(0);
    std::vector< std::thread > loaders;
    for (unsigned l = 0; l < n_load; ++l)
        loaders.emplace_back([&] {
            try
            {
                for (;;)
                {
                    Item it;
                    if (!src.next(it.read, it.index)) break;
                    ++n_reads;
                    n_events += it.events();
                    queue.push(std::move(it));
                }
            }
            catch (const std::exception& e) { fail(e.what()); }
            queue.loader_done();
        });

    std::atomic< bool > first_batch_seen(false);
    Clock::time_point t_first_batch = t_start;
    std::mutex fb_mu;
    std::vector< std::thread > workers;
    for (size_t g = 0; g < n_dev; ++g)
        workers.emplace_back([&, g] {
            Device_Stats& ds = stats.dev[g];
            ds.device = cfg.devices[g];
            std::unique_ptr< Pipeline > p;
            // Two threads per GPU share the context: this one trains batch k+1 while the caller thread basecalls batch k
            // (Pipeline serialises the calls into the context; what overlaps is one thread's host passes -- packing, base
            // sequences, FASTA -- with the other's kernels).  At most one trained batch waits between them.
            Hand_Over hand;
            std::thread caller;
            auto finish = [&](std::vector< Item >& batch) {
                std::vector< Read* > rp;
                size_t ev = 0;
                for (auto& it : batch)
                {
                    rp.push_back(&it.read);
                    ev += it.events();
                }
                const auto a1 = Clock::now();
                if (cfg.opt.basecall) p->basecall_reads(rp);
                ds.basecall_s += secs(a1, Clock::now());
                ds.reads += batch.size();
                ds.read_events += ev;
                ++ds.batches;
                for (auto& it : batch)   // the events are not needed any more
                    for (unsigned st = 0; st < 2; ++st) it.read.events[st] = Strand_Events();
                sink.put(std::move(batch));
                ds.last_batch_done_s = secs(t_start, Clock::now());
            };
            try
            {
                size_t n_batches = 0;
                for (;;)
                {
                    const auto w0 = Clock::now();
                    std::vector< Item > batch = queue.pop_batch(cfg.batch_reads, cfg.batch_events);
                    const auto w1 = Clock::now();
                    ds.wait_s += secs(w0, w1);
                    if (batch.empty()) break;
                    if (n_batches++ == 0) ds.first_batch_at_s = secs(t_start, w1);
                    size_t longest = 0, ev = 0;
                    for (const auto& it : batch)
                        for (unsigned st = 0; st < 2; ++st)
                        {
                            ev += it.read.events[st].size();
                            longest = std::max(longest, it.read.events[st].size());
                        }
                    if (!p)
                    {
                        // Viterbi scratch: what the jobs in flight of a batch like this one can use (two candidate
                        // models per strand at most; every forward CTA holds up to five jobs: ~760 jobs of the longest
                        // strand), unless the caller fixed it
                        size_t hint = cfg.pool_bytes ? cfg.pool_bytes / 16384u : std::min(2 * ev, 760 * std::max< size_t >(longest, 1));
                        const auto i0 = Clock::now();
                        p.reset(new Pipeline(cfg.opt, cfg.devices[g], hint));
                        p->init_models();
                        p->init_transitions();
                        // one-time allocations for batches like this one (later, larger batches grow them as needed)
                        p->reserve(batch.size(), ev);
                        ds.init_s = secs(i0, Clock::now());
                        if (cfg.overlap && cfg.opt.train && cfg.opt.basecall)
                            caller = std::thread([&] {
                                try
                                {
                                    std::vector< Item > b;
                                    while (hand.take(b)) finish(b);
                                }
                                catch (const std::exception& e)
                                {
                                    fail(e.what());
                                    hand.fail();
                                }
                            });
                    }
                    if (!first_batch_seen.exchange(true))
                    {
                        // the steady-state clock starts when the first dispatcher has its context (a one-time cost per
                        // process: CUDA context + scratch pool) and begins to work on reads
                        std::lock_guard< std::mutex > lk(fb_mu);
                        t_first_batch = Clock::now();
                    }
                    std::vector< Read* > rp;
                    for (auto& it : batch) rp.push_back(&it.read);
                    p->init_reads_params(rp);
                    const auto a0 = Clock::now();
                    if (cfg.opt.train) p->train_reads(rp);
                    ds.train_s += secs(a0, Clock::now());
                    if (caller.joinable())
                    {
                        const auto h0 = Clock::now();
                        const bool ok = hand.give(std::move(batch));
                        ds.hand_wait_s += secs(h0, Clock::now());
                        if (!ok) break;
                    }
                    else finish(batch);
                }
            }
            catch (const std::exception& e) { fail(e.what()); }
            hand.close();
            if (caller.joinable()) caller.join();
            if (p)
            {
                ds.train_rounds = p->train_rounds;
                ds.fwbw_events = p->fwbw_events;
                ds.viterbi_events = p->viterbi_events;
                ds.train_kernel_ms = p->train_kernel_ms;
                ds.viterbi_kernel_ms = p->viterbi_kernel_ms;
                ds.train_call_s = p->train_call_s;
                ds.viterbi_call_s = p->viterbi_call_s;
                double ts[8];
                if (nc_ctx_train_stats(p->ctx(), ts, 0) == NC_OK)
                {
                    ds.emission_ms = ts[0]; ds.fwbw_ms = ts[1]; ds.pm_stats_ms = ts[2]; ds.st_stats_ms = ts[3];
                }
            }
        });
    for (auto& t : loaders) t.join();
    for (auto& t : workers) t.join();
    stats.reads = n_reads;
    stats.read_events = n_events;
    stats.wall_s = secs(t_start, Clock::now());
    stats.steady_wall_s = secs(t_first_batch, sink.last_write());
    if (!error.empty()) { stats.error = error; return false; }
    if (sink.written() != stats.reads) { stats.error = "internal error: not every read was written"; return false; }
    return true;
}

std::string stats_json(const Run_Config& cfg, const Run_Stats& s)
{
    std::ostringstream os;
    os.precision(10);
    double fb_ev = 0, fb_ms = 0, v_ev = 0, v_ms = 0;
    os << "{\"n_gpus\": " << cfg.devices.size() << ", \"reads\": " << s.reads << ", \"read_events\": " << s.read_events
       << ", \"wall_s\": " << s.wall_s << ", \"steady_wall_s\": " << s.steady_wall_s << ", \"devices\": [";
    for (size_t g = 0; g < s.dev.size(); ++g)
    {
        const Device_Stats& d = s.dev[g];
        fb_ev += (double)d.fwbw_events; v_ev += (double)d.viterbi_events;
        fb_ms = std::max(fb_ms, d.train_kernel_ms); v_ms = std::max(v_ms, d.viterbi_kernel_ms);
        os << (g ? ", " : "") << "{\"device\": " << d.device << ", \"reads\": " << d.reads << ", \"batches\": " << d.batches
           << ", \"read_events\": " << d.read_events << ", \"train_rounds\": " << d.train_rounds << ", \"fwbw_events\": " << d.fwbw_events
           << ", \"train_kernel_ms\": " << d.train_kernel_ms << ", \"emission_ms\": " << d.emission_ms << ", \"fwbw_ms\": " << d.fwbw_ms
           << ", \"pm_stats_ms\": " << d.pm_stats_ms << ", \"st_stats_ms\": " << d.st_stats_ms
           << ", \"viterbi_events\": " << d.viterbi_events << ", \"viterbi_kernel_ms\": " << d.viterbi_kernel_ms
           << ", \"init_s\": " << d.init_s << ", \"train_s\": " << d.train_s << ", \"basecall_s\": " << d.basecall_s
           << ", \"train_call_s\": " << d.train_call_s << ", \"viterbi_call_s\": " << d.viterbi_call_s
           << ", \"wait_s\": " << d.wait_s << ", \"hand_wait_s\": " << d.hand_wait_s << ", \"first_batch_at_s\": " << d.first_batch_at_s
           << ", \"last_batch_done_s\": " << d.last_batch_done_s << "}";
    }
    double last = 0, first_done = 1e300;
    for (const auto& d : s.dev) if (d.batches) { last = std::max(last, d.last_batch_done_s); first_done = std::min(first_done, d.last_batch_done_s); }
    os << "], \"fwbw_events\": " << fb_ev << ", \"viterbi_events\": " << v_ev
       << ", \"read_events_per_s\": " << (s.steady_wall_s > 0 ? (double)s.read_events / s.steady_wall_s : 0.0)
       << ", \"viterbi_events_per_s_kernel\": " << (v_ms > 0 ? v_ev / v_ms * 1e3 : 0.0)
       << ", \"fwbw_events_per_s_kernel\": " << (fb_ms > 0 ? fb_ev / fb_ms * 1e3 : 0.0)
       << ", \"tail_s\": " << (first_done < 1e299 ? last - first_done : 0.0) << "}";
    return os.str();
}

} // namespace nchost
