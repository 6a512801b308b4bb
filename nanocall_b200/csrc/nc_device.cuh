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


// The Viterbi kernel's form of the same emission, one multiply and one add shorter per state.  Multiplying by
// 2 or 0.5 commutes exactly with IEEE rounding (no over/underflow at these magnitudes), so with
//   ah = RN((x-mu)/sg)/2  (= div_rn with divisor 2*sg and reciprocal rsg/2),   q = RN(ah*ah) = RN(a*a)/4
//   RN(RN(log_2pi + RN(a*a)) * 0.5)           == RN(q*2 + log_2pi/2)                 -- one FMA, 2q exact
//   RN(RN(RN(c1 - ly3) - u) * 0.5)            == RN(RN(c1/2 - ly3/2) - u/2),  u/2 = div_rn(lam*b*b, 2y, ry/2)
// every intermediate is the reference's value scaled by an exact power of two; the result has the same bits.
struct StateParamsH
{
    float mu, sg2, rsgh, nls, eta, reta, lam, c1h;
};
__device__ __forceinline__ StateParamsH halve(const StateParams& p)
{
    StateParamsH h;
    h.mu = p.mu; h.sg2 = __fadd_rn(p.sg, p.sg); h.rsgh = __fmul_rn(p.rsg, 0.5f); h.nls = p.nls;
    h.eta = p.eta; h.reta = p.reta; h.lam = p.lam; h.c1h = __fmul_rn(p.c1, 0.5f);
    return h;
}
// y2 = 2y, ly3h = (3 log y)/2, ryh = RN(1/y)/2, hl2pi = log_2pi/2
__device__ __forceinline__ float emission_h(const StateParamsH& p, float x, float y, float y2, float ly3h, float ryh, float hl2pi)
{
    float ah = div_rn(__fsub_rn(x, p.mu), p.sg2, p.rsgh);
    float ln = __fsub_rn(p.nls, __fmaf_rn(__fmul_rn(ah, ah), 2.0f, hl2pi));
    float b = div_rn(__fsub_rn(y, p.eta), p.eta, p.reta);
    float uh = div_rn(__fmul_rn(__fmul_rn(p.lam, b), b), y2, ryh);
    float li = __fsub_rn(__fsub_rn(p.c1h, ly3h), uh);
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


// expf with the bits of glibc's expf (sysdeps/ieee754/flt-32/e_expf.c, the ARM optimized-routines algorithm):
// double arithmetic, N = 32 table of 2^(i/N), cubic polynomial.  The table is 2^(i/32) correctly rounded to
// double minus (i << 47), regenerated from its definition (tools/gen_exp2f_table.py).  A C twin of this routine
// agrees with libm's expf on every one of 23e6 sampled inputs (tools/check_expf.c); the trainer's posteriors
// p = exp(alpha + beta - logZ) therefore carry the reference's bits (Parameter_Trainer.hpp:278).
__device__ const unsigned long long nc_exp2f_tab[32] = {
0x3ff0000000000000ULL,
0x3fefd9b0d3158574ULL,
0x3fefb5586cf9890fULL,
0x3fef9301d0125b51ULL,
0x3fef72b83c7d517bULL,
0x3fef54873168b9aaULL,
0x3fef387a6e756238ULL,
0x3fef1e9df51fdee1ULL,
0x3fef06fe0a31b715ULL,
0x3feef1a7373aa9cbULL,
0x3feedea64c123422ULL,
0x3feece086061892dULL,
0x3feebfdad5362a27ULL,
0x3feeb42b569d4f82ULL,
0x3feeab07dd485429ULL,
0x3feea47eb03a5585ULL,
0x3feea09e667f3bcdULL,
0x3fee9f75e8ec5f74ULL,
0x3feea11473eb0187ULL,
0x3feea589994cce13ULL,
0x3feeace5422aa0dbULL,
0x3feeb737b0cdc5e5ULL,
0x3feec49182a3f090ULL,
0x3feed503b23e255dULL,
0x3feee89f995ad3adULL,
0x3feeff76f2fb5e47ULL,
0x3fef199bdd85529cULL,
0x3fef3720dcef9069ULL,
0x3fef5818dcfba487ULL,
0x3fef7c97337b9b5fULL,
0x3fefa4afa2a490daULL,
0x3fefd0765b6e4540ULL
};

__device__ __forceinline__ float nc_expf(float x)
{
    if (!(x > -0x1.9fe368p6f)) return (x != x) ? x : 0.0f;  // underflow (and -inf) -> 0
    if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);
    const double InvLn2N = 0x1.71547652b82fep+0 * 32;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
    double z = __dmul_rn(InvLn2N, (double)x);
    double kd = __dadd_rn(z, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(z, kd);
    unsigned long long t = __ldg(nc_exp2f_tab + (ki & 31u));
    t += ki << 47;
    const double s = __longlong_as_double((long long)t);
    z = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(z, r2), y);
    y = __dmul_rn(y, s);
    return (float)y;
}


// logf with the bits of glibc 2.39's logf (sysdeps/ieee754/flt-32/e_logf.c): 16-entry {1/c, log c} table, cubic in
// double.  The 36 doubles are glibc's __logf_data (16 x {invc, logc}, ln2, poly[3]) dumped from libm-2.39.a
// (tools/logf_data.inc); the C twin tools/check_logf.c agrees with libm's logf on 35e6 sampled floats.  Used for
// Event::log_stdv (Event.hpp:43) when the caller does not supply it and for Parameter_Trainer.hpp:512.
__device__ const unsigned long long nc_logf_data[36] = {
0x3ff661ec79f8f3beULL,
0xbfd57bf7808caadeULL,
0x3ff571ed4aaf883dULL,
0xbfd2bef0a7c06ddbULL,
0x3ff49539f0f010b0ULL,
0xbfd01eae7f513a67ULL,
0x3ff3c995b0b80385ULL,
0xbfcb31d8a68224e9ULL,
0x3ff30d190c8864a5ULL,
0xbfc6574f0ac07758ULL,
0x3ff25e227b0b8ea0ULL,
0xbfc1aa2bc79c8100ULL,
0x3ff1bb4a4a1a343fULL,
0xbfba4e76ce8c0e5eULL,
0x3ff12358f08ae5baULL,
0xbfb1973c5a611cccULL,
0x3ff0953f419900a7ULL,
0xbfa252f438e10c1eULL,
0x3ff0000000000000ULL,
0x0000000000000000ULL,
0x3fee608cfd9a47acULL,
0x3faaa5aa5df25984ULL,
0x3feca4b31f026aa0ULL,
0x3fbc5e53aa362eb4ULL,
0x3feb2036576afce6ULL,
0x3fc526e57720db08ULL,
0x3fe9c2d163a1aa2dULL,
0x3fcbc2860d224770ULL,
0x3fe886e6037841edULL,
0x3fd1058bc8a07ee1ULL,
0x3fe767dcf5534862ULL,
0x3fd4043057b6ee09ULL,
0x3fe62e42fefa39efULL,
0xbfd00ea348b88334ULL,
0x3fd5575b0be00b6aULL,
0xbfdffffef20a4123ULL
};

__device__ __forceinline__ float nc_logf(float x)
{
    unsigned ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
    {
        if (ix * 2u == 0u) return NC_NEG_INF;                       // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                            // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return __int_as_float(0x7fc00000);  // x < 0 or NaN
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));                // subnormal: normalise
        ix -= 23u << 23;
    }
    const unsigned tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int)tmp >> 23;
    const unsigned iz = ix - (tmp & 0xff800000u);
    const double invc = __longlong_as_double((long long)__ldg(nc_logf_data + 2 * i));
    const double logc = __longlong_as_double((long long)__ldg(nc_logf_data + 2 * i + 1));
    const double ln2 = __longlong_as_double((long long)__ldg(nc_logf_data + 32));
    const double a0 = __longlong_as_double((long long)__ldg(nc_logf_data + 33));
    const double a1 = __longlong_as_double((long long)__ldg(nc_logf_data + 34));
    const double a2 = __longlong_as_double((long long)__ldg(nc_logf_data + 35));
    const double z = (double)__uint_as_float(iz);
    const double r = __dsub_rn(__dmul_rn(z, invc), 1.0);
    const double y0 = __dadd_rn(logc, __dmul_rn((double)k, ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(a1, r), a2);
    y = __dadd_rn(__dmul_rn(a0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return (float)y;
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
