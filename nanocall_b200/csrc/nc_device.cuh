// Device helpers shared by the Viterbi and Forward/Backward kernels (sm_100a).
//
// Numerical contract: every value that reaches alpha/beta/backpointers is
// produced by the same sequence of IEEE binary32 operations the reference's x86-64 build executes
// -- no FMA contraction (-fmad=false), round-to-nearest division, denormals kept (-ftz=false).
// The only FMAs in this file are the explicit ones inside div_rn(), which implement a
// correctly-rounded division and therefore return the same bits as `divss`.
#ifndef NC_DEVICE_CUH
#define NC_DEVICE_CUH

#include "nc_internal.h"
#include <cuda_runtime.h>

namespace nc {

#define NC_NEG_INF (__int_as_float(0xff800000))

// a / b rounded to nearest, given rb = RN(1/b) (from __frcp_rn).  Markstein: q0 = RN(a*rb) is
// within 1 ulp of a/b, the residual a - q0*b is exact in one FMA, and RN(q0 + r*rb) is the
// correctly rounded quotient.  Valid while no intermediate over/underflows, which holds for the
// emission's operands (|a| < 2^20, 2^-7 < b < 2^7); checked against true division on 1.5e9 random
// and edge-case operand pairs (tools/check_div_rn.c) and by every Viterbi parity test.
__device__ __forceinline__ float div_rn(float a, float b, float rb)
{
    float q0 = __fmul_rn(a, rb);
    float r = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(r, rb, q0);
}

// Scaled per-state emission constants (Pore_Model_State::scale, Pore_Model.hpp:126-138), plus the
// derived values the kernels keep in registers.
struct StateParams
{
    float mu;    // level_mean * scale + shift
    float sg;    // level_stdv * var
    float rsg;   // RN(1 / sg)
    float nls;   // -(log_level_stdv + log var)
    float eta;   // sd_mean * scale_sd
    float reta;  // RN(1 / eta)
    float lam;   // sd_lambda * var_sd
    float c1;    // (log_sd_lambda + log var_sd) - log_2pi
};

__device__ __forceinline__ StateParams scale_state(float level_mean, float level_stdv, float sd_mean,
                                                   float sd_lambda, float log_level_stdv, float log_sd_lambda,
                                                   const DevJob& J, float log_2pi)
{
    StateParams p;
    p.mu = __fadd_rn(__fmul_rn(level_mean, J.scale), J.shift);
    p.sg = __fmul_rn(level_stdv, J.var);
    p.rsg = __frcp_rn(p.sg);
    p.nls = -__fadd_rn(log_level_stdv, J.log_var);
    p.eta = __fmul_rn(sd_mean, J.scale_sd);
    p.reta = __frcp_rn(p.eta);
    p.lam = __fmul_rn(sd_lambda, J.var_sd);
    p.c1 = __fsub_rn(__fadd_rn(log_sd_lambda, J.log_var_sd), log_2pi);
    return p;
}

// Event as staged in shared memory: x = corrected mean (Event.hpp:81), y = stdv,
// ly3 = 3 * log_stdv, ry = RN(1 / y)
// log_pr_corrected_emission = log_normal_pdf + log_invgauss_pdf (Pore_Model.hpp:24-40,145-149):
//   a = (x - mu) / sg;              ln = -log_sg - (log_2pi + a*a) / 2
//   b = (y - eta) / eta;            li = (log_lam - log_2pi - 3*log_y - lam*b*b / y) / 2
// The divisions by 2 are exact scalings, written as * 0.5f.
__device__ __forceinline__ float emission(const StateParams& p, float x, float y, float ly3, float ry, float log_2pi)
{
    float a = div_rn(__fsub_rn(x, p.mu), p.sg, p.rsg);
    float ln = __fsub_rn(p.nls, __fmul_rn(__fadd_rn(log_2pi, __fmul_rn(a, a)), 0.5f));
    float b = div_rn(__fsub_rn(y, p.eta), p.eta, p.reta);
    float u = div_rn(__fmul_rn(__fmul_rn(p.lam, b), b), y, ry);
    float li = __fmul_rn(__fsub_rn(__fsub_rn(p.c1, ly3), u), 0.5f);
    return __fadd_rn(ln, li);
}

// 6-bit overlap mask of an edge i -> j: bit 0 = (i == j), bit l = suffix(i, 6-l) == prefix(j, 6-l)
// (the conditions State_Transitions::get_trans_prob tests, State_Transitions.hpp:128-141)
__device__ __forceinline__ unsigned trans_mask(unsigned i, unsigned j)
{
    unsigned m = (i == j) ? 1u : 0u;
#pragma unroll
    for (unsigned l = 1; l < 6; ++l)
        m |= ((i & ((1u << (2 * (6 - l))) - 1u)) == (j >> (2 * l))) ? (1u << l) : 0u;
    return m;
}

// Kmer::min_skip (Kmer.hpp:51-68)
__device__ __forceinline__ unsigned min_skip(unsigned k1, unsigned k2)
{
    if (k1 == k2) return 0;
#pragma unroll
    for (unsigned k = 5; k > 0; --k)
        if ((k1 & ((1u << (2 * k)) - 1u)) == (k2 >> (2 * (6 - k)))) return 6 - k;
    return 6;
}

// Backpointer byte -> predecessor state.  0..15: two-step predecessor (bb<<8)|(j>>4);
// 16..19: one-step predecessor (b<<10)|(j>>2); 20: j itself.
__device__ __forceinline__ unsigned bp_decode(unsigned code, unsigned j)
{
    if (code < 16u) return (code << 8) | (j >> 4);
    if (code < 20u) return ((code - 16u) << 10) | (j >> 2);
    return j;
}

} // namespace nc

#endif
