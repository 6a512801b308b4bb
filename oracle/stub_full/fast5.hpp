// TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's fast5.hpp (which needs libhdf5, absent from this image) so
// that the UNMODIFIED src/nanocall/nanocall.cpp and Fast5_Summary.hpp compile into oracle/_ref/nanocall_ref.
// fast5::File here serves EventDetection events from a raw event-table file ("NCRW0001": the same container
// nanocall-b200 reads, see nanocall_b200/host/reads.cpp) instead of an HDF5 file; everything downstream of
// File::get_eventdetection_events -- truncation, abasic level, hairpin detection, trimming, filtering, time base,
// training, selection, Viterbi, FASTA, --stats -- is the reference's own code.
// Interface mirrored: src/fast5/src/fast5.hpp:55-80 (event entry, event parameters), :89-101 (model entry),
// :300-485 (File accessors used by Fast5_Summary.hpp:154-184,280,505-525 and nanocall.cpp:212-247,904).
#ifndef NC_ORACLE_STUB_FULL_FAST5_HPP
#define NC_ORACLE_STUB_FULL_FAST5_HPP
#include <array>
#include <cstdint>
#include <cstring>
#include <exception>
#include <fstream>
#include <list>
#include <string>
#include <vector>
#define MAX_K_LEN 8
namespace hdf5_tools
{
class Exception : public std::exception
{
public:
    explicit Exception(const std::string& m) : _m(m) {}
    const char* what() const noexcept { return _m.c_str(); }
private:
    std::string _m;
};
}
namespace fast5
{
struct EventDetection_Event_Entry { double mean, stdv; long long start, length; };
struct EventDetection_Event_Parameters
{
    std::string read_id;
    long long read_number, scaling_used, start_mux, start_time, duration;
    double median_before;
    unsigned abasic_found;
};
struct Model_Entry { std::array< char, MAX_K_LEN > kmer; double level_mean, level_stdv, sd_mean, sd_stdv; };
struct Model_Parameters { double scale, shift, drift, var, scale_sd, var_sd; };
class File
{
public:
    File() : _open(false), _rate(0) {}
    explicit File(const std::string& fn, bool = false) : _open(false), _rate(0) { open(fn); }
    static bool is_valid_file(const std::string& fn)
    {
        std::ifstream is(fn.c_str(), std::ios::binary);
        char magic[8];
        return is.read(magic, 8) && std::memcmp(magic, "NCRW0001", 8) == 0;
    }
    static int get_object_count() { return 0; }
    void open(const std::string& fn, bool = false)
    {
        std::ifstream is(fn.c_str(), std::ios::binary);
        char magic[8];
        uint32_t n_reads = 0, id_len = 0, n = 0;
        if (!is.read(magic, 8) || std::memcmp(magic, "NCRW0001", 8) != 0) throw hdf5_tools::Exception(fn + ": not a raw event table");
        is.read(reinterpret_cast< char* >(&n_reads), 4);
        is.read(reinterpret_cast< char* >(&id_len), 4);
        if (!is || n_reads < 1 || id_len > 4096) throw hdf5_tools::Exception(fn + ": corrupt header");
        _id.resize(id_len);
        is.read(&_id[0], id_len);
        is.read(reinterpret_cast< char* >(&_rate), 8);
        is.read(reinterpret_cast< char* >(&n), 4);
        _ev.resize(n);
        is.read(reinterpret_cast< char* >(_ev.data()), (std::streamsize)n * sizeof(EventDetection_Event_Entry));
        if (!is) throw hdf5_tools::Exception(fn + ": truncated");
        _open = true;
    }
    bool is_open() const { return _open; }
    bool have_sampling_rate() const { return _rate > 0; }
    double get_sampling_rate() const { return _rate; }
    bool have_eventdetection_events(const std::string& = std::string()) const { return _open; }
    EventDetection_Event_Parameters get_eventdetection_event_params(const std::string& = std::string()) const
    {
        EventDetection_Event_Parameters p = EventDetection_Event_Parameters();
        p.read_id = _id;
        return p;
    }
    std::vector< EventDetection_Event_Entry > get_eventdetection_events(const std::string& = std::string()) const { return _ev; }
    std::list< std::string > get_basecall_group_list() const { return std::list< std::string >(); }
    bool have_basecall_model(bool) const { return false; }
    Model_Parameters get_basecall_model_params(bool) const { return Model_Parameters(); }
    std::vector< Model_Entry > get_basecall_model(bool) const { return std::vector< Model_Entry >(); }
    // --write-fast5 needs HDF5: not served by the stand-in
    template < typename... A > void add_basecall_seq(A&&...) const { throw hdf5_tools::Exception("write-fast5 is not available in the oracle build"); }
    template < typename... A > void add_basecall_events(A&&...) const { throw hdf5_tools::Exception("write-fast5 is not available in the oracle build"); }
    template < typename... A > void add_basecall_model(A&&...) const { throw hdf5_tools::Exception("write-fast5 is not available in the oracle build"); }
    template < typename... A > void add_basecall_model_params(A&&...) const { throw hdf5_tools::Exception("write-fast5 is not available in the oracle build"); }
private:
    bool _open;
    double _rate;
    std::string _id;
    std::vector< EventDetection_Event_Entry > _ev;
};
}
#endif
