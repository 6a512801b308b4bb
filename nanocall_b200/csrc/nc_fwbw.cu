// K2: log-space Forward/Backward (Forward_Backward.hpp:46-135) and the sufficient statistics of
// Parameter_Trainer (Parameter_Trainer.hpp:230-532) for batches of training sequences.
//
// Bit-exactness.  alpha, beta and log Pr[data] reproduce the reference bit for bit.  That needs
//   * p7_FLogsum literally (logsum.hpp:141-154): max, min, the test `min == -inf || max-min >= 15.999f`,
//     index (int)((max-min)*1000.f), the 16000-entry table built on the host in double (logsum.hpp:113-127);
//   * the reference's accumulation ORDER: from -inf, over the ascending merged predecessor list from_v(j)
//     (forward) / successor list to_v(j) (backward), each edge with its exact weight.
// The lists are never materialised.  Predecessors of j sorted by index fall into 16 "slots" (index >> 8):
// slot s holds the two-step predecessor (s<<8)|(j>>4), the one-step predecessor (b<<10)|(j>>2) when
// s == 4b + (j>>10), and j itself when s == j>>8; inside a slot the order is that of the low bytes.
// Successors of j are two contiguous blocks, [(j&255)<<4, +16) and [(j&1023)<<2, +4), plus j.  An index that
// occurs twice is ONE edge in the reference (std::set union, State_Transitions.hpp:205-209) whose weight
// carries every matching overlap term; here the lower-class duplicate is replaced by -inf, and
// p7_FLogsum(x, -inf) == x exactly, so the chain is the reference's chain.
//
// Kernels (one wave = the sequences whose E/alpha/beta slabs fit the scratch pool):
//   emission_kernel   E[i][j] = log_pr_corrected_emission(j, e_i) for every sequence (grid: seq x event tiles)
//   fwbw_kernel       one CTA per sequence: forward, backward, log Pr[data]; alpha/beta to the slabs
//   pm_stats_kernel   per event the six posterior-weighted sums of train_pm_params (:263-296)
//   st_stats_kernel   one CTA per (group, strand): the three log-space accumulators of train_st_params (:471-514),
//                     folded sequentially in the reference's (sequence, event, k-mer) order
// The 3x3 solve, the clamps and exp() of the final ratios run on the host in nc_train.cpp.
#include "nc_device.cuh"
#include "nc_kernels.h"

namespace nc {

namespace {

constexpr int FB_THREADS = 512;
constexpr int FB_SPT = 8;
constexpr int COL_PAD_SHIFT = 4;  // phys(n) = n + 4 * (n >> 4): conflict-free 16-float block reads
constexpr int COL_FLOATS = NC_N_STATES + 4 * (NC_N_STATES >> COL_PAD_SHIFT);  // 5120

__device__ __forceinline__ int cphys(int n) { return n + ((n >> COL_PAD_SHIFT) << 2); }

// p7_FLogsum (logsum.hpp:141-154): max + tbl[(int)((max - min) * 1000.f)], or max when min == -inf or
// max - min >= 15.999f.  Same bits with 9 instructions instead of 14 (the kernels are bound by the ALU pipe, where
// FSETP/FSEL/SEL run at half rate):
//   * max - min == |a - b| exactly (rounding is symmetric), used as an operand modifier: no min is formed;
//   * min == -inf implies |a - b| == +inf >= 15.999f, so that test is subsumed; when both are -inf the
//     difference is NaN, the comparison is false and -inf + tbl[.] == -inf == max;
//   * the index is clamped through fminf(|a - b|, 15.999f) (15.999f * 1000.f == 15999.f): in range without a select,
//     and only changed where the result is discarded.
// NaN inputs (which the reference only meets on invalid events) are not reproduced.
__device__ __forceinline__ float flogsum(float a, float b, const float* __restrict__ tbl)
{
    const float d = fabsf(__fsub_rn(a, b));
    const float mx = fmaxf(a, b);
    const int idx = __float2int_rz(__fmul_rn(fminf(d, 15.999f), 1000.0f));
    const float r = __fadd_rn(mx, tbl[idx]);
    return (d >= 15.999f) ? mx : r;
}

// Six sequential float sums at once: acc_s = sum over j = 0..4095 of T[s][j], s = 0..5, each the serial loop
// `for j: acc += T[s][j]` itself -- lane s < 6 of one warp carries chain s, the terms stream from shared memory as
// float4 one group ahead of the additions.  No screening: on training data most terms are live anyway (posteriors
// are broad while the scaling is still off; measured ~3000 of 4096), and the chain of 4096 dependent FADDs
// (~16 k cycles) is then the whole cost -- 5.6 k warp-instructions per event instead of ~70 k for six warps that
// screen and compact (profiles/r1_training_kernels.md).  Rows are PM_ROW = 4096 + 4 floats apart so the six lanes'
// float4 loads hit different banks.
constexpr int FOLD_LIST = 80;
constexpr int ST_CHUNK = FB_THREADS - 3 * 32;   // 416 k-mers per chunk of st_stats_kernel
constexpr int PM_ROW = NC_N_STATES + 4;
__device__ __forceinline__ float fold_six_sums_in_order(const float* __restrict__ T, const int lane)
{
    const float* R = T + (lane < 6 ? lane : 0) * PM_ROW;   // lanes >= 6 shadow chain 0
    float acc = 0.0f;
    float4 v = *reinterpret_cast< const float4* >(R);
#pragma unroll 4
    for (int j = 0; j < (int)NC_N_STATES; j += 4)
    {
        const float4 nv = *reinterpret_cast< const float4* >(R + j + 4);   // the last one reads the row's padding
        acc = __fadd_rn(acc, v.x);
        acc = __fadd_rn(acc, v.y);
        acc = __fadd_rn(acc, v.z);
        acc = __fadd_rn(acc, v.w);
        v = nv;
    }
    return acc;
}

// The log-space twin: acc = p7_FLogsum(acc, x_k) over k = 0..n-1 in order, n a multiple of 32, x_k = get(k).
// A term with x == -inf or acc - x >= 15.999 returns acc unchanged now and for every larger acc (the running value
// never decreases), so it is screened out; the survivors are folded from the per-warp list, padded with -inf.
template < typename Get >
__device__ __forceinline__ float fold_logsum_in_order(float acc, const int n, Get get, float* __restrict__ lst,
                                                      const float* __restrict__ tbl, const int lane)
{
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < n; base += 32)
    {
        const float x = get(base + lane);
        const bool live = !((x == NC_NEG_INF) || (acc > x && __fsub_rn(acc, x) >= 15.999f));
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (m == 0) continue;
        const int cnt = __popc(m);
        if (live) lst[__popc(m & lt)] = x;
        if (lane < 4) lst[cnt + lane] = NC_NEG_INF;
        __syncwarp();
        for (int k = 0; k < cnt; k += 2)
        {
            const float2 v = *reinterpret_cast< const float2* >(lst + k);
            acc = flogsum(acc, v.x, tbl);
            acc = flogsum(acc, v.y, tbl);
        }
        __syncwarp();
    }
    return acc;
}

struct StSmem
{
    float term[2][3][ST_CHUNK];   // two buffers: computed by warps 3..15, folded by warps 0..2
    float lst[3][FOLD_LIST];
};

struct FbSmem
{
    float tbl[16000];
    float col[2][COL_FLOATS];
    float red[FB_THREADS / 32];
    float lst[FOLD_LIST];
    unsigned item;
};

} // namespace

// ------------------------------------------------------------------------------------------------
// E[seq][i][j]
__global__ void __launch_bounds__(FB_THREADS) emission_kernel(const FbArgs a)
{
    const unsigned seq = blockIdx.y;
    const FbSeq& Q = a.seqs[seq];
    const unsigned i0 = blockIdx.x * FB_EV_TILE;
    if (i0 >= Q.n_events) return;
    const DevJob& J = a.jobs[Q.job];
    const int t = threadIdx.x;
    const unsigned j0 = FB_SPT * t;
    const float* M = a.models + (size_t)J.model * MODEL_FLOATS;
    float* E = a.scratch + Q.slab + 0 * (size_t)Q.n_events * NC_N_STATES;
    const unsigned i1 = min(i0 + FB_EV_TILE, Q.n_events);
#pragma unroll 1
    for (int half = 0; half < 2; ++half)
    {
        StateParams P[4];
        const unsigned jb = j0 + 4 * half;
        {
            const float4 lm = __ldg(reinterpret_cast< const float4* >(M + 0 * NC_N_STATES + jb));
            const float4 ls = __ldg(reinterpret_cast< const float4* >(M + 1 * NC_N_STATES + jb));
            const float4 sm = __ldg(reinterpret_cast< const float4* >(M + 2 * NC_N_STATES + jb));
            const float4 sl = __ldg(reinterpret_cast< const float4* >(M + 3 * NC_N_STATES + jb));
            const float4 ll = __ldg(reinterpret_cast< const float4* >(M + 4 * NC_N_STATES + jb));
            const float4 lsl = __ldg(reinterpret_cast< const float4* >(M + 5 * NC_N_STATES + jb));
            P[0] = scale_state(lm.x, ls.x, sm.x, sl.x, ll.x, lsl.x, J, a.log_2pi);
            P[1] = scale_state(lm.y, ls.y, sm.y, sl.y, ll.y, lsl.y, J, a.log_2pi);
            P[2] = scale_state(lm.z, ls.z, sm.z, sl.z, ll.z, lsl.z, J, a.log_2pi);
            P[3] = scale_state(lm.w, ls.w, sm.w, sl.w, ll.w, lsl.w, J, a.log_2pi);
        }
        for (unsigned i = i0; i < i1; ++i)
        {
            const unsigned long long e = Q.ev_off + i;
            const float stdv = __ldg(a.stdv + e);
            const float y = (stdv == 0.0f) ? 0.01f : stdv;                                  // Event.hpp:39-42
            const float x = __fsub_rn(__ldg(a.mean + e), __fmul_rn(J.drift, __ldg(a.start + e)));  // Event.hpp:81
            const float ly3 = __fmul_rn(3.0f, a.log_stdv ? __ldg(a.log_stdv + e) : nc_logf(y));
            const float ry = __frcp_rn(y);
            float4 o;
            o.x = emission(P[0], x, y, ly3, ry, a.log_2pi);
            o.y = emission(P[1], x, y, ly3, ry, a.log_2pi);
            o.z = emission(P[2], x, y, ly3, ry, a.log_2pi);
            o.w = emission(P[3], x, y, ly3, ry, a.log_2pi);
            *reinterpret_cast< float4* >(E + (size_t)i * NC_N_STATES + jb) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_THREADS, 2) fwbw_kernel(const FbArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FbSmem& sm = *reinterpret_cast< FbSmem* >(smem_raw);
    const int t = threadIdx.x;
    const int lane = t & 31;
    for (int q = t; q < 16000; q += FB_THREADS) sm.tbl[q] = a.logsum_tbl[q];
    const float* tbl = sm.tbl;

    for (;;)
    {
        __syncthreads();
        if (t == 0) sm.item = atomicAdd(a.next_item, 1u);
        __syncthreads();
        const unsigned seq = sm.item;
        if (seq >= a.n_seqs) break;
        const FbSeq& Q = a.seqs[seq];
        const DevJob& J = a.jobs[Q.job];
        const unsigned n = Q.n_events;
        const float* E = a.scratch + Q.slab;
        float* AL = a.scratch + Q.slab + 1 * (size_t)n * NC_N_STATES;
        float* BE = a.scratch + Q.slab + 2 * (size_t)n * NC_N_STATES;

        // =========================== forward (Forward_Backward.hpp:58-89); thread owns j = 8t .. 8t+7
        {
            const unsigned j0 = FB_SPT * t;
            const unsigned g = t >> 1;
            const float wT = J.lut[trans_mask(g, j0) & 0x3cu];
            float wO[2];
            wO[0] = J.lut[trans_mask(2 * t, j0) & 0x3eu];
            wO[1] = J.lut[trans_mask(2 * t + 1, j0 + 4) & 0x3eu];
            const int c = t >> 7;        // one-step predecessors sit in slots s with (s & 3) == c   (warp-uniform)
            const int sS = t >> 5;       // j >> 8: slot of the self predecessor                       (warp-uniform)
            const int kT = t >> 1;       // low byte of the two-step predecessors
            // column 0
            {
                const float4 e0 = *reinterpret_cast< const float4* >(E + j0);
                const float4 e1 = *reinterpret_cast< const float4* >(E + j0 + 4);
                float4 a0 = make_float4(__fsub_rn(e0.x, a.log_n_states), __fsub_rn(e0.y, a.log_n_states),
                                        __fsub_rn(e0.z, a.log_n_states), __fsub_rn(e0.w, a.log_n_states));
                float4 a1 = make_float4(__fsub_rn(e1.x, a.log_n_states), __fsub_rn(e1.y, a.log_n_states),
                                        __fsub_rn(e1.z, a.log_n_states), __fsub_rn(e1.w, a.log_n_states));
                *reinterpret_cast< float4* >(sm.col[0] + cphys(j0)) = a0;
                *reinterpret_cast< float4* >(sm.col[0] + cphys(j0 + 4)) = a1;
                *reinterpret_cast< float4* >(AL + j0) = a0;
                *reinterpret_cast< float4* >(AL + j0 + 4) = a1;
            }
            __syncthreads();
            int cur = 0;
            for (unsigned i = 1; i < n; ++i)
            {
                const float* A = sm.col[cur];
                float vT[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) vT[s] = __fadd_rn(wT, A[cphys((s << 8) | (int)g)]);
                float vO[2][4];
#pragma unroll
                for (int b = 0; b < 4; ++b)
                {
                    const float2 o = *reinterpret_cast< const float2* >(A + cphys((b << 10) + 2 * t));
                    vO[0][b] = __fadd_rn(wO[0], o.x);
                    vO[1][b] = __fadd_rn(wO[1], o.y);
                }
                float* An = sm.col[cur ^ 1];
                const float* Ei = E + (size_t)i * NC_N_STATES;
                // one state at a time (NOT unrolled): with the 16 slots unrolled inside, an unrolled k loop makes
                // the kernel 290 KB of SASS and every column refetches it through the instruction cache
                // (stall_no_instruction was 5 cycles per issue).  The other warps of the two resident CTAs hide
                // the latency of the now sequential chains.
#pragma unroll 1
                for (int k = 0; k < FB_SPT; ++k)
                {
                    const int hh = k >> 2;
                    const unsigned j = j0 + k;
                    const float own_k = A[cphys(j)];
                    const float e_k = __ldg(Ei + j);
                    const int kO = (2 * t + hh) & 255;
                    const int kS = j & 255;
                    const bool oBefT = kO < kT, oEqT = kO == kT;
                    const bool sEqT = kS == kT, sEqO = kS == kO;
                    const float vS = __fadd_rn(J.lut[trans_mask(j, j)], own_k);
                    float acc = NC_NEG_INF;
#pragma unroll
                    for (int s = 0; s < 16; ++s)
                    {
                        const bool hasO = (s & 3) == c;
                        const bool hasS = s == sS;
                        if (!hasO && !hasS) acc = flogsum(acc, vT[s], tbl);
                        else
                        {
                            const bool dropT = (hasO && oEqT) || (hasS && sEqT);
                            const bool dropO = !hasO || (hasS && sEqO);
                            const float xT = dropT ? NC_NEG_INF : vT[s];
                            const float xO = dropO ? NC_NEG_INF : (hh ? vO[1][s >> 2] : vO[0][s >> 2]);
                            const bool oFirst = hasO && oBefT;
                            const float first = oFirst ? xO : xT;
                            const float second = oFirst ? xT : xO;
                            if (!hasS)
                            {
                                acc = flogsum(acc, first, tbl);
                                acc = flogsum(acc, second, tbl);
                            }
                            else
                            {
                                const int pos = (kT < kS ? 1 : 0) + ((hasO && kO < kS) ? 1 : 0);
                                acc = flogsum(acc, pos == 0 ? vS : NC_NEG_INF, tbl);
                                acc = flogsum(acc, first, tbl);
                                acc = flogsum(acc, pos == 1 ? vS : NC_NEG_INF, tbl);
                                acc = flogsum(acc, second, tbl);
                                acc = flogsum(acc, pos == 2 ? vS : NC_NEG_INF, tbl);
                            }
                        }
                    }
                    An[cphys(j)] = __fadd_rn(e_k, acc);
                }
                // the thread's 8 new values, read back as two float4, go to the slab with vector stores
                *reinterpret_cast< float4* >(AL + (size_t)i * NC_N_STATES + j0) = *reinterpret_cast< const float4* >(An + cphys(j0));
                *reinterpret_cast< float4* >(AL + (size_t)i * NC_N_STATES + j0 + 4) = *reinterpret_cast< const float4* >(An + cphys(j0 + 4));
                cur ^= 1;
                __syncthreads();
            }
            // log Pr[data]: sequential fold of the last column, ascending j (Forward_Backward.hpp:129-134).
            // One warp: each lane screens 32 values against the running sum -- a term with sum - x >= 15.999
            // leaves p7_FLogsum's result unchanged now and for every larger sum -- and only the rest is folded.
            if (t < 32)
            {
                const float* A = sm.col[cur];
                const float acc = fold_logsum_in_order(NC_NEG_INF, (int)NC_N_STATES, [&](int j) { return A[cphys(j)]; }, sm.lst, tbl, lane);
                if (lane == 0) a.log_pr_data[seq] = acc;
            }
            __syncthreads();
        }

        // =========================== backward (Forward_Backward.hpp:93-125); thread owns j = t + 512k
        {
            const int tb = (t & 255) << 4;                 // two-step successors tb .. tb+15 (same for all 8 states)
            const float wTb = J.lut[trans_mask(t, tb) & 0x3cu];
            float wOb[2];
            int ob[2];
            ob[0] = t << 2;                                // one-step successors of j with (j & 1023) == t
            ob[1] = (t + 512) << 2;                        //                                      == t + 512
            wOb[0] = J.lut[trans_mask(t, ob[0]) & 0x3eu];
            wOb[1] = J.lut[trans_mask(t + 512, ob[1]) & 0x3eu];
            // beta[n-1] = 0
            {
#pragma unroll
                for (int k = 0; k < FB_SPT; ++k)
                {
                    const int j = t + FB_THREADS * k;
                    sm.col[0][cphys(j)] = 0.0f;
                    BE[(size_t)(n - 1) * NC_N_STATES + j] = 0.0f;
                }
            }
            __syncthreads();
            int cur = 0;
            for (unsigned ip1 = n - 1; ip1 > 0; --ip1)
            {
                const unsigned i = ip1 - 1;
                const float* Bn = sm.col[cur];
                const float* En = E + (size_t)ip1 * NC_N_STATES;
                float vT[16];
#pragma unroll
                for (int v = 0; v < 4; ++v)
                {
                    const float4 e = __ldg(reinterpret_cast< const float4* >(En + tb) + v);
                    const float4 b = *reinterpret_cast< const float4* >(Bn + cphys(tb) + 4 * v);
                    vT[4 * v + 0] = __fadd_rn(__fadd_rn(wTb, e.x), b.x);
                    vT[4 * v + 1] = __fadd_rn(__fadd_rn(wTb, e.y), b.y);
                    vT[4 * v + 2] = __fadd_rn(__fadd_rn(wTb, e.z), b.z);
                    vT[4 * v + 3] = __fadd_rn(__fadd_rn(wTb, e.w), b.w);
                }
                float* Bc = sm.col[cur ^ 1];
#pragma unroll
                for (int f = 0; f < 2; ++f)
                {
                    float vO[4];
                    {
                        const float4 e = __ldg(reinterpret_cast< const float4* >(En + ob[f]));
                        const float4 b = *reinterpret_cast< const float4* >(Bn + cphys(ob[f]));
                        vO[0] = __fadd_rn(__fadd_rn(wOb[f], e.x), b.x);
                        vO[1] = __fadd_rn(__fadd_rn(wOb[f], e.y), b.y);
                        vO[2] = __fadd_rn(__fadd_rn(wOb[f], e.z), b.z);
                        vO[3] = __fadd_rn(__fadd_rn(wOb[f], e.w), b.w);
                    }
                    // merged, ordered list of the 20 block successors of this family
                    const bool oIn = (ob[f] >> 4) == (tb >> 4);
                    const bool oBef = !oIn && ob[f] < tb;
                    const int c4 = (ob[f] & 15) >> 2;
                    float L[20];
#pragma unroll
                    for (int q = 0; q < 20; ++q)
                    {
                        // position q holds: oBef ? (q < 4 ? O[q] : T[q-4]) : (q < 16 ? T[q] : O[q-16])
                        float asT, asO;
                        if (q < 4) { asO = vO[q]; asT = vT[q]; }
                        else if (q < 16) { asO = vT[q - 4]; asT = vT[q]; }
                        else { asO = vT[q - 4]; asT = vO[q - 16]; }
                        float val = oBef ? asO : asT;
                        if (q < 16)
                        {
                            // oIn: the block is T with entries 4*c4 .. 4*c4+3 carrying the one-step weight
                            const bool rep = oIn && ((q >> 2) == c4);
                            val = rep ? vO[q & 3] : val;
                        }
                        else val = oIn ? NC_NEG_INF : val;
                        L[q] = val;
                    }
#pragma unroll 1
                    for (int kk = 0; kk < 4; ++kk)
                    {
                        const int k = 2 * kk + f;
                        const int j = t + FB_THREADS * k;
                        const float vS = __fadd_rn(__fadd_rn(J.lut[trans_mask(j, j)], __ldg(En + j)), Bn[cphys(j)]);
                        const bool sInT = (j >> 4) == (tb >> 4);
                        const bool sInO = (j >> 2) == (ob[f] >> 2);
                        // position of j inside the list when it coincides with a block entry, else -1
                        int mpos = -1;
                        if (sInT) mpos = (oBef ? 4 : 0) + (j & 15);
                        else if (sInO) mpos = (oBef ? 0 : 16) + (j & 3);
                        // otherwise: number of blocks entirely below j
                        const int pos = (mpos >= 0) ? -1 : ((j > tb ? 1 : 0) + ((!oIn && j > ob[f]) ? 1 : 0));
                        // p7_FLogsum(-inf, x) == x: the first fold is a select
                        float acc = (pos == 0) ? vS : NC_NEG_INF;
                        if (mpos < 0)
                        {
                            // common case: j is not one of its own block successors, the list is used as it is
#pragma unroll
                            for (int q = 0; q < 20; ++q)
                            {
                                if (q == 4) acc = flogsum(acc, (oBef && pos == 1) ? vS : NC_NEG_INF, tbl);
                                if (q == 16) acc = flogsum(acc, (!oBef && pos == 1) ? vS : NC_NEG_INF, tbl);
                                acc = flogsum(acc, L[q], tbl);
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int q = 0; q < 20; ++q)
                            {
                                acc = flogsum(acc, (q == mpos) ? vS : L[q], tbl);   // (pos == -1: no fold between the blocks)
                            }
                        }
                        acc = flogsum(acc, pos == 2 ? vS : NC_NEG_INF, tbl);
                        Bc[cphys(j)] = acc;
                        BE[(size_t)i * NC_N_STATES + j] = acc;
                    }
                }
                cur ^= 1;
                __syncthreads();
            }
        }
    }
}

size_t fwbw_smem_bytes() { return sizeof(FbSmem); }
size_t st_stats_smem_bytes() { return sizeof(StSmem); }

// ------------------------------------------------------------------------------------------------
// train_pm_params' inner sums (Parameter_Trainer.hpp:263-296): per event
//   s0 = sum_j p/sigma^2, s1 = sum_j p*mu/sigma^2, s2 = sum_j p*mu^2/sigma^2,
//   l0 = sum_j p*lambda,  l1 = sum_j p*lambda/eta,  l2 = sum_j p*lambda/eta^2      (UNSCALED model)
// with p = exp(alpha + beta - logZ).  The 3x3 system built from these sums is ill-conditioned (level means are
// 58 +- 6 pA), so the float rounding of the reference's SEQUENTIAL j = 0..4095 accumulation is visible in the trained
// shift/scale/var at the 1e-4 level: the order is reproduced exactly.  All terms are >= 0, so the running sum never
// shift/scale/var at the 1e-4 level: the order is reproduced exactly, by running the six serial loops themselves in six
// lanes of one warp (fold_six_sums_in_order).
__global__ void __launch_bounds__(FB_THREADS) pm_stats_kernel(const FbArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* term = reinterpret_cast< float* >(smem_raw);  // [6][PM_ROW]
    const unsigned seq = blockIdx.y;
    const FbSeq& Q = a.seqs[seq];
    const unsigned i0 = blockIdx.x * FB_EV_TILE;
    if (i0 >= Q.n_events) return;
    const unsigned n = Q.n_events;
    const DevJob& J = a.jobs[Q.job];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned j0 = FB_SPT * t;
    const float* M = a.models + (size_t)J.model * MODEL_FLOATS;
    const float* AL = a.scratch + Q.slab + 1 * (size_t)n * NC_N_STATES;
    const float* BE = a.scratch + Q.slab + 2 * (size_t)n * NC_N_STATES;
    const float logz = a.log_pr_data[seq];
    float mu[FB_SPT], sg2[FB_SPT], lam[FB_SPT], eta[FB_SPT];
#pragma unroll
    for (int k = 0; k < FB_SPT; ++k)
    {
        mu[k] = __ldg(M + 0 * NC_N_STATES + j0 + k);
        const float sg = __ldg(M + 1 * NC_N_STATES + j0 + k);
        sg2[k] = __fmul_rn(sg, sg);
        eta[k] = __ldg(M + 2 * NC_N_STATES + j0 + k);
        lam[k] = __ldg(M + 3 * NC_N_STATES + j0 + k);
    }
    const unsigned i1 = min(i0 + FB_EV_TILE, n);
    for (unsigned i = i0; i < i1; ++i)
    {
        float al[FB_SPT], be[FB_SPT];
        *reinterpret_cast< float4* >(al) = *reinterpret_cast< const float4* >(AL + (size_t)i * NC_N_STATES + j0);
        *reinterpret_cast< float4* >(al + 4) = *reinterpret_cast< const float4* >(AL + (size_t)i * NC_N_STATES + j0 + 4);
        *reinterpret_cast< float4* >(be) = *reinterpret_cast< const float4* >(BE + (size_t)i * NC_N_STATES + j0);
        *reinterpret_cast< float4* >(be + 4) = *reinterpret_cast< const float4* >(BE + (size_t)i * NC_N_STATES + j0 + 4);
#pragma unroll
        for (int k = 0; k < FB_SPT; ++k)
        {
            const float p = nc_expf(__fsub_rn(__fadd_rn(al[k], be[k]), logz));
            const float ts0 = __fdiv_rn(p, sg2[k]);
            const float ts1 = __fmul_rn(ts0, mu[k]);
            const float tl0 = __fmul_rn(p, lam[k]);
            const float tl1 = __fdiv_rn(tl0, eta[k]);
            term[0 * PM_ROW + j0 + k] = ts0;
            term[1 * PM_ROW + j0 + k] = ts1;
            term[2 * PM_ROW + j0 + k] = __fmul_rn(ts1, mu[k]);
            term[3 * PM_ROW + j0 + k] = tl0;
            term[4 * PM_ROW + j0 + k] = tl1;
            term[5 * PM_ROW + j0 + k] = __fdiv_rn(tl1, eta[k]);
        }
        __syncthreads();
        if (warp == 0)
        {
            const float acc = fold_six_sums_in_order(term, lane);
            if (lane < 6) a.pm_stats[(Q.ev_out + i) * 6 + lane] = acc;
        }
        __syncthreads();
    }
}

size_t pm_stats_smem_bytes() { return 6 * PM_ROW * sizeof(float); }

// ------------------------------------------------------------------------------------------------
// train_st_params' accumulators (Parameter_Trainer.hpp:434-517) for one (group, strand):
//   denom (+)= post(i,j1);  stay (+)= min(joint(j1->j1 | log p_stay), post);
//   skip (+)= log(exp(post) - exp(min(d01, post))),  d01 = stay' (+) the 4 one-step joints with log(p_step/4)
// over the strand's sequences in order, events i < n-1, the 2160 training k-mers in ascending order.
// Warps 3..15 compute the terms of 416 k-mers at a time into one of two buffers while warps 0..2 fold the previous
// chunk sequentially, one accumulator each (screening out terms that cannot change the running value, as in the
// logZ fold): the gathers and the expf/logf of chunk c+1 overlap the dependent p7_FLogsum chain of chunk c.
__global__ void __launch_bounds__(FB_THREADS) st_stats_kernel(const FbArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StSmem& ss = *reinterpret_cast< StSmem* >(smem_raw);
    // the p7_FLogsum table stays in global memory here: with 11 KB of shared memory per CTA four CTAs fit an SM
    // (the kernel is latency-bound by its sequential folds) and the 64 KB table lives in the L1 the carve-out leaves
    const float* __restrict__ tbl = a.logsum_tbl;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned grp = blockIdx.x;
    const unsigned st = blockIdx.y;
    const FbGroup& G = a.groups[grp];
    const float log_p_stay = G.log_p_stay[st];
    const float log_p_step_4 = G.log_p_step_4[st];
    const unsigned n_km = a.n_train_kmers;

    // the (sequence, event, chunk) items in the reference's order; every thread walks the same two cursors
    struct Cursor { unsigned sq, i, base; };
    auto skip = [&](Cursor& c) { while (c.sq < G.seq_end && (a.seqs[c.sq].strand != st || a.seqs[c.sq].n_events < 2)) ++c.sq; };
    auto valid = [&](const Cursor& c) { return c.sq < G.seq_end; };
    auto advance = [&](Cursor& c) {
        c.base += ST_CHUNK;
        if (c.base >= n_km)
        {
            c.base = 0;
            if (++c.i + 1 >= a.seqs[c.sq].n_events) { c.i = 0; ++c.sq; skip(c); }
        }
    };
    auto compute = [&](const Cursor& c, float (*term)[ST_CHUNK]) {
        const int ct = t - 3 * 32;   // 0..415
        const FbSeq& Q = a.seqs[c.sq];
        const unsigned n = Q.n_events;
        const float* E = a.scratch + Q.slab;
        const float* AL = a.scratch + Q.slab + 1 * (size_t)n * NC_N_STATES;
        const float* BE = a.scratch + Q.slab + 2 * (size_t)n * NC_N_STATES;
        const float logz = a.log_pr_data[c.sq];
        const float* Ai = AL + (size_t)c.i * NC_N_STATES;
        const float* Bi = BE + (size_t)c.i * NC_N_STATES;
        const float* Bn = BE + (size_t)(c.i + 1) * NC_N_STATES;
        const float* En = E + (size_t)(c.i + 1) * NC_N_STATES;
        float t_denom = NC_NEG_INF, t_stay = NC_NEG_INF, t_skip = NC_NEG_INF;
        if (c.base + ct < n_km)
        {
            const unsigned j1 = a.train_kmers[c.base + ct];
            const float al = Ai[j1];
            const float log_p_j1 = __fsub_rn(__fadd_rn(al, Bi[j1]), logz);
            // joint(i, j1, j2, lt) = alpha + lt + emission(j2, e_{i+1}) + beta(i+1, j2) - logZ   (:457-469)
            float jj = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(al, log_p_stay), En[j1]), Bn[j1]), logz);
            if (jj > log_p_j1) jj = log_p_j1;
            float s2 = flogsum(NC_NEG_INF, jj, tbl);
            const unsigned nb = (j1 & 1023u) << 2;   // the four one-step successors are one aligned float4
            const float4 e4 = *reinterpret_cast< const float4* >(En + nb);
            const float4 b4 = *reinterpret_cast< const float4* >(Bn + nb);
            const float ev[4] = { e4.x, e4.y, e4.z, e4.w }, bv[4] = { b4.x, b4.y, b4.z, b4.w };
#pragma unroll
            for (int b = 0; b < 4; ++b)
            {
                const float jv = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(al, log_p_step_4), ev[b]), bv[b]), logz);
                s2 = flogsum(s2, jv, tbl);
            }
            if (s2 > log_p_j1) s2 = log_p_j1;
            const float p2 = __fsub_rn(nc_expf(log_p_j1), nc_expf(s2));
            t_denom = log_p_j1;
            t_stay = jj;
            t_skip = nc_logf(p2);
        }
        term[0][ct] = t_denom;
        term[1][ct] = t_stay;
        term[2][ct] = t_skip;
    };

    Cursor cc = { G.seq_begin, 0, 0 };
    skip(cc);
    if (valid(cc) && warp >= 3) compute(cc, ss.term[0]);
    __syncthreads();
    Cursor fc = cc;
    advance(cc);
    int buf = 0;
    float acc = NC_NEG_INF;   // warps 0..2: denom, stay, skip
    while (valid(fc))
    {
        if (warp >= 3)
        {
            if (valid(cc)) compute(cc, ss.term[buf ^ 1]);
        }
        else
        {
            // NaN terms (log of a negative difference cannot occur: d01 <= post) are folded like the reference would
            const float* Tw = ss.term[buf][warp];
            acc = fold_logsum_in_order(acc, ST_CHUNK, [&](int k) { return Tw[k]; }, ss.lst[warp], tbl, lane);
        }
        __syncthreads();
        fc = cc;
        advance(cc);
        buf ^= 1;
    }
    if (warp < 3 && lane == 0) a.st_stats[(grp * 2 + st) * 3 + warp] = acc;   // denom, stay, skip
}

} // namespace nc
