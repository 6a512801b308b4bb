// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the reference's fast5.hpp so that the
// reference's hot-path headers (Pore_Model.hpp, Event.hpp, Viterbi.hpp, ...) compile in a
// container without libhdf5.  Those headers only need three POD types, MAX_K_LEN and the
// names of three fast5::File methods used by loaders that the oracle never instantiates
// (reference: src/nanocall/Pore_Model.hpp:54-64,99-109,204-217; Viterbi.hpp:122).
#ifndef NC_ORACLE_STUB_FAST5_HPP
#define NC_ORACLE_STUB_FAST5_HPP
#include <array>
#include <string>
#include <vector>
#define MAX_K_LEN 8
namespace fast5
{
struct EventDetection_Event_Entry { double mean, stdv; long long start, length; };
struct Model_Entry { std::array< char, MAX_K_LEN > kmer; double level_mean, level_stdv, sd_mean, sd_stdv; };
struct Model_Parameters { double scale, shift, drift, var, scale_sd, var_sd; };
struct File
{
    bool have_basecall_model(bool) const { return false; }
    Model_Parameters get_basecall_model_params(bool) const { return Model_Parameters(); }
    std::vector< Model_Entry > get_basecall_model(bool) const { return std::vector< Model_Entry >(); }
};
}
#endif
