// nanocall-b200: command-line front end with nanocall's option surface (nanocall.cpp:50-95,908-1080) over the
// batched GPU hot path.  Inputs are EVENT TABLES (.events.tsv per read, or .ncev containers of many reads),
// directories of them, or files of file names ("-" = stdin); fast5 input needs libhdf5 and is not built here.
// Reads are sharded across --gpus devices by host threads (one context per GPU, no collectives) and written
// in input order, like pfor's ordered output (pfor.hpp:216-235).
#include "pipeline.hpp"

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

using namespace nchost;

namespace {

struct Cli
{
    Options opt;
    std::string output_fn, stats_fn;
    std::vector< std::string > inputs;
    int gpus = 1;
    int first_device = 0;
    bool single_strand_scaling = false, double_strand_flag = false, train_flag = false, no_train = false;
    bool basecall_flag = false, no_basecall = false;
};

void usage()
{
    std::cerr <<
        "USAGE: nanocall-b200 [options] <inputs...>\n"
        "  inputs: .events.tsv / .ncev event tables, directories of them, or files of file names (\"-\" = stdin)\n"
        "  --pore r73|r9 (r9)        --pr-stay F (.1)   --pr-skip F (.3)   -m/--model strand:file (multi)\n"
        "  --train / --no-train      --no-train-scaling --no-train-transitions   --train-drift 0|1\n"
        "  --single-strand-scaling | --double-strand-scaling (default)\n"
        "  --scaling-num-events N (200)  --scaling-max-rounds N (10)  --scaling-min-progress F (1.0)\n"
        "  --scaling-select-threshold F (20.0)   --min-ed-events N (10)\n"
        "  --basecall / --no-basecall   --fasta-line-width N (80)   -o/--output file   --stats file\n"
        "  --log error|warning|info|debug (info)   -t/--threads N (accepted, unused)   --gpus N (1)   --device K (0)\n"
        "  --version   --help\n";
}

bool is_dir(const std::string& p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool has_ext(const std::string& p, const char* e)
{
    size_t n = std::strlen(e);
    return p.size() > n && p.compare(p.size() - n, n, e) == 0;
}
bool is_event_file(const std::string& p) { return has_ext(p, ".events.tsv") || has_ext(p, ".ncev"); }

std::string self_dir()
{
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n <= 0) return ".";
    buf[n] = 0;
    std::string s(buf);
    auto pos = s.find_last_of('/');
    return pos == std::string::npos ? "." : s.substr(0, pos);
}

int parse(int argc, char** argv, Cli& c)
{
    auto need = [&](int& i) -> const char* {
        if (i + 1 >= argc) { std::cerr << "missing value for " << argv[i] << "\n"; std::exit(EXIT_FAILURE); }
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        if (a == "--help" || a == "-h") { usage(); std::exit(EXIT_SUCCESS); }
        else if (a == "--version") { std::cout << nc_version() << std::endl; std::exit(EXIT_SUCCESS); }
        else if (a == "--pore") c.opt.pore = need(i);
        else if (a == "--pr-stay") c.opt.pr_stay = std::strtof(need(i), nullptr);
        else if (a == "--pr-skip") c.opt.pr_skip = std::strtof(need(i), nullptr);
        else if (a == "-m" || a == "--model") c.opt.model_files.push_back(need(i));
        else if (a == "--train") c.train_flag = true;
        else if (a == "--no-train") c.no_train = true;
        else if (a == "--no-train-scaling") c.opt.train_scaling = false;
        else if (a == "--no-train-transitions") c.opt.train_transitions = false;
        else if (a == "--train-drift") c.opt.train_drift = std::atoi(need(i));
        else if (a == "--single-strand-scaling") c.single_strand_scaling = true;
        else if (a == "--double-strand-scaling") c.double_strand_flag = true;
        else if (a == "--scaling-num-events") c.opt.scaling_num_events = (unsigned)std::atoi(need(i));
        else if (a == "--scaling-max-rounds") c.opt.scaling_max_rounds = (unsigned)std::atoi(need(i));
        else if (a == "--scaling-min-progress") c.opt.scaling_min_progress = std::strtof(need(i), nullptr);
        else if (a == "--scaling-select-threshold") c.opt.scaling_select_threshold = std::strtof(need(i), nullptr);
        else if (a == "--min-ed-events") c.opt.min_ed_events = (unsigned)std::atoi(need(i));
        else if (a == "--basecall") c.basecall_flag = true;
        else if (a == "--no-basecall") c.no_basecall = true;
        else if (a == "--fasta-line-width") c.opt.fasta_line_width = (unsigned)std::atoi(need(i));
        else if (a == "-o" || a == "--output") c.output_fn = need(i);
        else if (a == "--stats") c.stats_fn = need(i);
        else if (a == "-t" || a == "--threads") (void)need(i);
        else if (a == "--gpus") c.gpus = std::atoi(need(i));
        else if (a == "--device") c.first_device = std::atoi(need(i));
        else if (a == "--data-dir") c.opt.data_dir = need(i);
        else if (a == "--log")
        {
            std::string l = need(i);
            auto pos = l.find(':');
            if (pos != std::string::npos) l = l.substr(pos + 1);
            c.opt.log_level = l == "error" ? 0 : l == "warning" ? 1 : l == "info" ? 2 : 3;
        }
        else if (a == "--") { for (++i; i < argc; ++i) c.inputs.push_back(argv[i]); }
        else if (a.size() > 1 && a[0] == '-' && a != "-") { std::cerr << "unknown option " << a << "\n"; usage(); return 1; }
        else c.inputs.push_back(a);
    }
    // validation mirrors nanocall.cpp:995-1059
    if (c.inputs.empty()) { usage(); return 1; }
    if (c.opt.pore != "r9" && c.opt.pore != "r73") { std::cerr << "unknown pore type: " << c.opt.pore << "\n"; return 1; }
    if (c.train_flag && c.no_train) { std::cerr << "either --train or --no-train may be used, but not both\n"; return 1; }
    if (c.basecall_flag && c.no_basecall) { std::cerr << "either --basecall or --no-basecall may be used, but not both\n"; return 1; }
    if (c.single_strand_scaling && c.double_strand_flag)
    {
        std::cerr << "either --single-strand-scaling or --double-strand-scaling may be used, but not both\n";
        return 1;
    }
    if (c.opt.train_drift > 1) { std::cerr << "train-drift not understood\n"; return 1; }
    if (c.opt.pr_stay <= 0.f || c.opt.pr_skip <= 0.f || c.opt.pr_stay + c.opt.pr_skip >= 1.f)
    {
        std::cerr << "invalid pr-stay / pr-skip\n";
        return 1;
    }
    c.opt.train = !c.no_train;
    c.opt.basecall = !c.no_basecall;
    c.opt.double_strand_scaling = !c.single_strand_scaling;
    if (!c.opt.train) { c.opt.train_scaling = false; c.opt.train_transitions = false; }
    if (c.opt.data_dir.empty())
    {
        const char* env = std::getenv("NANOCALL_B200_DATA");
        c.opt.data_dir = env ? env : self_dir() + "/../data";
    }
    if (c.gpus < 1) c.gpus = 1;
    return 0;
}

void collect_files(const Cli& c, std::vector< std::string >& files)
{
    for (const auto& f : c.inputs)
    {
        if (f != "-" && is_dir(f))
        {
            DIR* d = opendir(f.c_str());  // raw readdir order, as fs_support.hpp:22-36
            if (!d) continue;
            while (struct dirent* e = readdir(d))
            {
                std::string g = e->d_name;
                if (g == "." || g == "..") continue;
                std::string f2 = f + (f.back() != '/' ? "/" : "") + g;
                if (!is_dir(f2) && is_event_file(f2)) files.push_back(f2);
            }
            closedir(d);
        }
        else if (f != "-" && is_event_file(f)) files.push_back(f);
        else
        {
            std::ifstream ifs;
            std::istream* is = &std::cin;
            if (f != "-") { ifs.open(f); is = &ifs; }
            std::string g;
            while (std::getline(*is, g))
                if (is_event_file(g)) files.push_back(g);
        }
    }
}

} // namespace

int main(int argc, char** argv)
{
    Cli cli;
    if (int rc = parse(argc, argv, cli)) return rc;
    const int lvl = cli.opt.log_level;
    std::vector< std::string > files;
    collect_files(cli, files);
    if (files.empty()) { std::cerr << "no event files to process\n"; return EXIT_FAILURE; }

    std::vector< Read > reads;
    for (const auto& f : files)
    {
        std::string err;
        if (has_ext(f, ".ncev"))
        {
            if (!load_events_ncev(f, reads, err)) { std::cerr << err << "\n"; return EXIT_FAILURE; }
        }
        else
        {
            Read r;
            if (!load_events_tsv(f, r, err)) { std::cerr << err << "\n"; return EXIT_FAILURE; }
            reads.push_back(std::move(r));
        }
    }
    log_line(2, lvl, "loaded " + std::to_string(reads.size()) + " reads from " + std::to_string(files.size()) + " files");

    // shard reads across GPUs: contiguous ranges balanced by event count
    const int n_gpus = std::min< int >(cli.gpus, (int)reads.size());
    std::vector< size_t > bounds(n_gpus + 1, 0);
    {
        size_t total = 0;
        for (const auto& r : reads) total += r.events[0].size() + r.events[1].size();
        size_t acc = 0;
        int g = 1;
        for (size_t i = 0; i < reads.size() && g < n_gpus; ++i)
        {
            acc += reads[i].events[0].size() + reads[i].events[1].size();
            if (acc * n_gpus >= total * (size_t)g) bounds[g++] = i + 1;
        }
        for (; g <= n_gpus; ++g) bounds[g] = reads.size();
    }
    std::vector< std::string > errors(n_gpus);
    std::vector< std::string > summaries(n_gpus);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int g) {
        try
        {
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto secs_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration< double >(now() - t).count(); };
            auto w0 = now();
            // events that can be in flight in one Viterbi call: every job of the shard (two candidate models per
            // strand at most), but no more than ~300 jobs (2 per forward CTA) of the longest strand
            size_t shard_events = 0, longest = 0;
            for (size_t i = bounds[g]; i < bounds[g + 1]; ++i)
                for (unsigned st = 0; st < 2; ++st)
                {
                    shard_events += reads[i].events[st].size();
                    longest = std::max(longest, reads[i].events[st].size());
                }
            Pipeline p(cli.opt, cli.first_device + g, std::min(2 * shard_events, 300 * longest));
            p.init_models();
            std::vector< Read* > mine;
            for (size_t i = bounds[g]; i < bounds[g + 1]; ++i)
            {
                p.init_read_params(reads[i]);
                mine.push_back(&reads[i]);
            }
            const double init_s = secs_since(w0);
            auto w1 = now();
            if (cli.opt.train) p.train_reads(mine);
            const double train_s = secs_since(w1);
            auto w2 = now();
            if (cli.opt.basecall) p.basecall_reads(mine);
            const double basecall_s = secs_since(w2);
            std::ostringstream s;
            s << "gpu " << (cli.first_device + g) << ": reads=" << mine.size() << " train_rounds=" << p.train_rounds
              << " fwbw_events=" << p.fwbw_events << " train_kernel_ms=" << p.train_kernel_ms
              << " viterbi_events=" << p.viterbi_events << " viterbi_kernel_ms=" << p.viterbi_kernel_ms
              << " init_s=" << init_s << " train_s=" << train_s << " basecall_s=" << basecall_s;
            summaries[g] = s.str();
        }
        catch (const std::exception& e) { errors[g] = e.what(); }
    };
    std::vector< std::thread > th;
    for (int g = 0; g < n_gpus; ++g) th.emplace_back(worker, g);
    for (auto& t : th) t.join();
    for (const auto& e : errors)
        if (!e.empty()) { std::cerr << "error: " << e << "\n"; return EXIT_FAILURE; }
    double secs = std::chrono::duration< double >(std::chrono::steady_clock::now() - t0).count();
    for (const auto& s : summaries) log_line(2, lvl, s);
    log_line(2, lvl, "processed " + std::to_string(reads.size()) + " reads in " + std::to_string(secs) + " seconds");

    // ordered output
    if (cli.opt.basecall)
    {
        std::ofstream ofs;
        std::ostream* os = &std::cout;
        if (!cli.output_fn.empty()) { ofs.open(cli.output_fn); os = &ofs; }
        Options o = cli.opt;
        for (const auto& r : reads)
            for (unsigned st = 0; st < 2; ++st)
                if (r.called[st])
                    Pipeline::write_fasta(*os, r.read_id + ":" + r.base_file_name + ":" + std::to_string(st), r.base_seq[st],
                                          o.fasta_line_width);
    }
    if (!cli.stats_fn.empty())
    {
        std::ofstream ofs(cli.stats_fn);
        Pipeline::write_stats_header(ofs);
        // write_stats only needs the options: build the rows without a device context
        for (const auto& r : reads)
        {
            const size_t n0 = r.events[0].size(), n1 = r.events[1].size();
            ofs << r.base_file_name << '\t' << r.read_id << '\t' << (n0 + n1) << "\t0\t0\t" << n0 << '\t' << n0 << '\t' << (n0 + n1);
            for (unsigned st = 0; st < 2; ++st)
            {
                char buf[512];
                if (!r.preferred_model[st][st].empty() && r.pm_params_m.count(r.preferred_model[st]))
                {
                    const nc_pm_params& p = r.pm_params_m.at(r.preferred_model[st]);
                    const nc_st_params& s = r.st_params_m.at(r.preferred_model[st])[st];
                    std::snprintf(buf, sizeof buf, "\t%s\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f",
                                  r.preferred_model[st][st].c_str(), p.scale, p.shift, p.drift, p.var, p.scale_sd, p.var_sd,
                                  s.p_stay, s.p_skip);
                }
                else
                    std::snprintf(buf, sizeof buf, "\t.\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f", 1.0, 0.0, 0.0, 1.0, 1.0,
                                  1.0, cli.opt.pr_stay, cli.opt.pr_skip);
                ofs << buf;
            }
            ofs << "\n";
        }
    }
    return EXIT_SUCCESS;
}
