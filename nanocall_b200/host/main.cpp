// nanocall-b200: command-line front end with nanocall's option surface (nanocall.cpp:50-95,908-1080) over the
// batched GPU hot path.  Inputs are EVENT TABLES (.events.tsv per read, or .ncev containers of many reads),
// directories of them, or files of file names ("-" = stdin); fast5 input needs libhdf5 and is not built here.
// Reads are sharded across --gpus devices by host threads (one context per GPU, no collectives) and written
// in input order, like pfor's ordered output (pfor.hpp:216-235).
#include "dispatch.hpp"

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <memory>
#include <thread>

using namespace nchost;

namespace {

struct Cli
{
    Options opt;
    std::string output_fn, stats_fn, synth, summary_json;
    std::vector< std::string > inputs;
    int gpus = 1;
    int first_device = 0;
    size_t batch_reads = 4096, batch_mevents = 48, pool_gb = 0;
    bool overlap = true;
    unsigned loader_threads = 0;
    bool single_strand_scaling = false, double_strand_flag = false, train_flag = false, no_train = false;
    bool basecall_flag = false, no_basecall = false, write_fast5 = false, summarize_only = false;
};

void usage()
{
    std::cerr <<
        "USAGE: nanocall-b200 [options] <inputs...>\n"
        "  inputs: event tables (raw NCRW0001 files, one read per file like a fast5 or many; segmented .ncev /\n"
        "          .events.tsv), directories of them, or files of file names (\"-\" = stdin); --synth instead of inputs\n"
        "  nanocall's options (nanocall.cpp:56-94):\n"
        "  --pore r73|r9 (r9)        --pr-stay F (.1)   --pr-skip F (.3)   -m/--model strand:file (multi)   --model-fofn file\n"
        "  -s/--trans file           --train / --no-train      --no-train-scaling --no-train-transitions   --train-drift 0|1\n"
        "  --single-strand-scaling | --double-strand-scaling (default when training)\n"
        "  --scaling-num-events N (200)  --scaling-max-rounds N (10)  --scaling-min-progress F (1.0)\n"
        "  --scaling-select-threshold F (20.0)   --min-ed-events N (10)   --max-ed-events N (100000)\n"
        "  --trim-ed-sq-start N  --trim-ed-sq-end N  --trim-ed-hp-start N  --trim-ed-hp-end N (50 each)   --1d\n"
        "  --ed-group G   --chunk-size N   --basecall / --no-basecall   --fasta-line-width N (80)\n"
        "  -o/--output file   --stats file   --log [facility:]level (multi)   -t/--threads N (host loader threads)\n"
        "  --write-fast5 (needs HDF5: rejected in this build)   --version   --help\n"
        "  device options: --gpus N (1)   --device K (0)   --batch-reads N (4096)   --batch-mevents N (48)   --pool-gb N   --no-overlap\n"
        "  --synth n[:seed[:pool[:2d|1d|mix[:nt[:nc]]]]]   synthetic R7.3 reads instead of inputs\n"
        "  --summary-json file      run statistics (per-GPU device times, events/s, tail)\n"
        "  --summarize-only         segmentation and --stats without a GPU (no training, no basecalling)\n";
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
// fast5::File::is_valid_file's role (nanocall.cpp:212,225,247): is this an input the loaders understand?
bool is_event_file(const std::string& p) { return has_ext(p, ".events.tsv") || has_ext(p, ".ncev") || is_ncrw_file(p); }

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
    std::string train_drift;
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        if (a == "--help" || a == "-h") { usage(); std::exit(EXIT_SUCCESS); }
        else if (a == "--version") { std::cout << nc_version() << std::endl; std::exit(EXIT_SUCCESS); }
        else if (a == "--pore") c.opt.pore = need(i);
        else if (a == "--pr-stay") c.opt.pr_stay = std::strtof(need(i), nullptr);
        else if (a == "--pr-skip") c.opt.pr_skip = std::strtof(need(i), nullptr);
        else if (a == "-m" || a == "--model") c.opt.model_files.push_back(need(i));
        else if (a == "--model-fofn") c.opt.model_fofn = need(i);
        else if (a == "-s" || a == "--trans") c.opt.trans_fn = need(i);
        else if (a == "--train") c.train_flag = true;
        else if (a == "--no-train") c.no_train = true;
        else if (a == "--no-train-scaling") c.opt.train_scaling = false;
        else if (a == "--no-train-transitions") c.opt.train_transitions = false;
        else if (a == "--train-drift") train_drift = need(i);
        else if (a == "--single-strand-scaling") c.single_strand_scaling = true;
        else if (a == "--double-strand-scaling") c.double_strand_flag = true;
        else if (a == "--scaling-num-events") c.opt.scaling_num_events = (unsigned)std::atoi(need(i));
        else if (a == "--scaling-max-rounds") c.opt.scaling_max_rounds = (unsigned)std::atoi(need(i));
        else if (a == "--scaling-min-progress") c.opt.scaling_min_progress = std::strtof(need(i), nullptr);
        else if (a == "--scaling-select-threshold") c.opt.scaling_select_threshold = std::strtof(need(i), nullptr);
        else if (a == "--min-ed-events") c.opt.min_ed_events = (unsigned)std::atoi(need(i));
        else if (a == "--max-ed-events") c.opt.max_ed_events = (unsigned)std::atoi(need(i));
        else if (a == "--trim-ed-sq-start") c.opt.trim_margins[0] = (unsigned)std::atoi(need(i));
        else if (a == "--trim-ed-sq-end") c.opt.trim_margins[1] = (unsigned)std::atoi(need(i));
        else if (a == "--trim-ed-hp-start") c.opt.trim_margins[2] = (unsigned)std::atoi(need(i));
        else if (a == "--trim-ed-hp-end") c.opt.trim_margins[3] = (unsigned)std::atoi(need(i));
        else if (a == "--1d") c.opt.template_only = true;
        else if (a == "--ed-group") c.opt.ed_group = need(i);
        else if (a == "--chunk-size") c.opt.chunk_size = (unsigned)std::atoi(need(i));
        else if (a == "--basecall") c.basecall_flag = true;
        else if (a == "--no-basecall") c.no_basecall = true;
        else if (a == "--write-fast5") c.write_fast5 = true;
        else if (a == "--fasta-line-width") c.opt.fasta_line_width = (unsigned)std::atoi(need(i));
        else if (a == "-o" || a == "--output") c.output_fn = need(i);
        else if (a == "--stats") c.stats_fn = need(i);
        else if (a == "-t" || a == "--threads") c.loader_threads = (unsigned)std::atoi(need(i));
        else if (a == "--gpus") c.gpus = std::atoi(need(i));
        else if (a == "--device") c.first_device = std::atoi(need(i));
        else if (a == "--batch-reads") c.batch_reads = (size_t)std::atoll(need(i));
        else if (a == "--batch-mevents") c.batch_mevents = (size_t)std::atoll(need(i));
        else if (a == "--no-overlap") c.overlap = false;
        else if (a == "--pool-gb") c.pool_gb = (size_t)std::atoll(need(i));
        else if (a == "--synth") c.synth = need(i);
        else if (a == "--summary-json") c.summary_json = need(i);
        else if (a == "--summarize-only") c.summarize_only = true;
        else if (a == "--data-dir") c.opt.data_dir = need(i);
        else if (a == "--log")
        {
            // [facility:]level (logger.hpp): facilities are not distinguished here, the most verbose level wins
            std::string l = need(i);
            auto pos = l.find(':');
            if (pos != std::string::npos) l = l.substr(pos + 1);
            int lv = l == "error" ? 0 : l == "warning" ? 1 : l == "info" ? 2 : 3;
            static bool first = true;
            c.opt.log_level = first ? lv : std::max(c.opt.log_level, lv);
            first = false;
        }
        else if (a == "--") { for (++i; i < argc; ++i) c.inputs.push_back(argv[i]); }
        else if (a.size() > 1 && a[0] == '-' && a != "-") { std::cerr << "unknown option " << a << "\n"; usage(); return 1; }
        else c.inputs.push_back(a);
    }
    // validation and defaults mirror nanocall.cpp:920-1059
    if (c.inputs.empty() && c.synth.empty()) { usage(); return 1; }
    if (!train_drift.empty() && train_drift != "0" && train_drift != "1") { std::cerr << "train-drift not understdood: " << train_drift << "\n"; return 1; }
    if (c.opt.pore == "r9") { c.opt.abasic_level_top_percent = 1.0; c.opt.abasic_level_top_offset = 0.0; c.opt.train_drift = train_drift.empty() ? 0 : train_drift == "1"; }
    else if (c.opt.pore == "r73") { c.opt.abasic_level_top_percent = 1.0; c.opt.abasic_level_top_offset = 5.0; c.opt.train_drift = train_drift.empty() ? 1 : train_drift == "1"; }
    else { std::cerr << "unknown pore type: " << c.opt.pore << "\n"; return 1; }
    if (c.train_flag && c.no_train) { std::cerr << "either --train or --no-train may be used, but not both\n"; return 1; }
    if (c.basecall_flag && c.no_basecall) { std::cerr << "either --basecall or --no-basecall may be used, but not both\n"; return 1; }
    c.opt.train = !c.no_train;
    c.opt.basecall = !c.no_basecall;
    // --double-strand-scaling becomes the default only when scaling is trained (nanocall.cpp:1012-1026): with
    // --no-train or --no-train-scaling and neither flag given, strands are scaled (and ranked) separately
    c.opt.double_strand_scaling = c.double_strand_flag;
    if (c.opt.train && c.opt.train_scaling)
    {
        if (c.single_strand_scaling && c.double_strand_flag)
        {
            std::cerr << "either --single-strand-scaling or --double-strand-scaling may be used, but not both\n";
            return 1;
        }
        if (!c.single_strand_scaling && !c.double_strand_flag) c.opt.double_strand_scaling = true;
    }
    if (c.opt.scaling_select_threshold < 0.0f) { std::cerr << "invalid scaling_select_threshold: " << c.opt.scaling_select_threshold << "\n"; return 1; }
    if (c.opt.scaling_min_progress < 0.0f) { std::cerr << "invalid scaling_min_progress: " << c.opt.scaling_min_progress << "\n"; return 1; }
    if (c.write_fast5)
    {
        std::cerr << (c.output_fn.empty() ? "--write-fast5 needs libhdf5, which this build does not have: write FASTA with -o instead\n"
                                          : "output may be written to fast5 files or to a single output file, but not both\n");
        return 1;
    }
    if (c.opt.pr_stay <= 0.f || c.opt.pr_skip <= 0.f || c.opt.pr_stay + c.opt.pr_skip >= 1.f)
    {
        std::cerr << "invalid pr-stay / pr-skip\n";
        return 1;
    }
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
    const int lvl = c.opt.log_level;
    for (const auto& f : c.inputs)
    {
        if (f != "-" && is_dir(f))
        {
            DIR* d = opendir(f.c_str());  // raw readdir order, as fs_support.hpp:22-36
            if (!d) continue;
            while (struct dirent* e = readdir(d))
            {
                std::string g = e->d_name;
                std::string f2 = f + (f.back() != '/' ? "/" : "") + g;
                if (is_dir(f2)) { log_line(2, lvl, "ignoring subdirectory [" + f2 + "]"); continue; }
                if (is_event_file(f2)) { files.push_back(f2); log_line(2, lvl, "adding input file [" + f2 + "]"); }
                else log_line(2, lvl, "ignoring file [" + f2 + "]");
            }
            closedir(d);
        }
        else if (f != "-" && is_event_file(f)) { files.push_back(f); log_line(2, lvl, "adding input file [" + f + "]"); }
        else
        {
            log_line(2, lvl, "interpreting [" + f + "] as fofn");
            std::ifstream ifs;
            std::istream* is = &std::cin;
            if (f != "-") { ifs.open(f); is = &ifs; }
            std::string g;
            while (std::getline(*is, g))
                if (is_event_file(g)) { files.push_back(g); log_line(2, lvl, "adding input file [" + g + "]"); }
        }
    }
}

} // namespace

int main(int argc, char** argv)
{
    Cli cli;
    if (int rc = parse(argc, argv, cli)) return rc;
    const int lvl = cli.opt.log_level;
    std::unique_ptr< Read_Source > src;
    try
    {
        if (!cli.synth.empty()) src = make_synth_source(cli.opt, cli.synth, cli.opt.data_dir, lvl);
        else
        {
            std::vector< std::string > files;
            collect_files(cli, files);
            if (files.empty()) { std::cerr << "no fast5 files to process\n"; return EXIT_FAILURE; }
            src = make_file_source(cli.opt, files, lvl);
        }
    }
    catch (const std::exception& e) { std::cerr << "error: " << e.what() << "\n"; return EXIT_FAILURE; }

    std::ofstream fasta_fs, stats_fs;
    std::ostream* fasta = nullptr;
    if (cli.opt.basecall && !cli.summarize_only)
    {
        fasta = &std::cout;
        if (!cli.output_fn.empty()) { fasta_fs.open(cli.output_fn); fasta = &fasta_fs; }
    }
    std::ostream* stats_os = nullptr;
    if (!cli.stats_fn.empty()) { stats_fs.open(cli.stats_fn); stats_os = &stats_fs; }

    if (cli.summarize_only)
    {
        // Fast5_Summary alone: what init_reads + the --stats rows of untrained reads need, no device
        if (stats_os) Pipeline::write_stats_header(*stats_os);
        Read r;
        size_t idx;
        while (src->next(r, idx))
            if (stats_os) Pipeline::write_stats(*stats_os, r, cli.opt);
        return EXIT_SUCCESS;
    }

    Run_Config cfg;
    cfg.opt = cli.opt;
    for (int g = 0; g < cli.gpus; ++g) cfg.devices.push_back(cli.first_device + g);
    cfg.batch_reads = std::max< size_t >(1, cli.batch_reads);
    cfg.batch_events = std::max< size_t >(1, cli.batch_mevents) << 20;
    cfg.overlap = cli.overlap;
    cfg.queue_events = std::max(cfg.batch_events * (cfg.devices.size() + 1), (size_t)192 << 20);
    cfg.loader_threads = cli.loader_threads;
    cfg.pool_bytes = cli.pool_gb << 30;
    Run_Stats st;
    const bool ok = run_pipeline(cfg, *src, fasta, stats_os, st);
    if (!ok) { std::cerr << "error: " << st.error << "\n"; return EXIT_FAILURE; }
    for (const auto& d : st.dev)
    {
        std::ostringstream s;
        s << "gpu " << d.device << ": reads=" << d.reads << " batches=" << d.batches << " train_rounds=" << d.train_rounds
          << " fwbw_events=" << d.fwbw_events << " train_kernel_ms=" << d.train_kernel_ms << " viterbi_events=" << d.viterbi_events
          << " viterbi_kernel_ms=" << d.viterbi_kernel_ms << " init_s=" << d.init_s << " train_s=" << d.train_s
          << " basecall_s=" << d.basecall_s << " train_call_s=" << d.train_call_s << " viterbi_call_s=" << d.viterbi_call_s
          << " wait_s=" << d.wait_s << " emission_ms=" << d.emission_ms << " fwbw_ms=" << d.fwbw_ms
          << " pm_stats_ms=" << d.pm_stats_ms << " st_stats_ms=" << d.st_stats_ms;
        log_line(2, lvl, s.str());
    }
    log_line(2, lvl, "processed " + std::to_string(st.reads) + " reads in " + std::to_string(st.wall_s) + " seconds");
    if (!cli.summary_json.empty())
    {
        std::ofstream js(cli.summary_json);
        js << stats_json(cfg, st) << std::endl;
    }
    return EXIT_SUCCESS;
}
