// K2: log-space Forward/Backward (Forward_Backward.hpp:46-135) and the sufficient statistics of
// Parameter_Trainer (Parameter_Trainer.hpp:230-532) for batches of training sequences.
//
// Bit-exactness.  alpha, beta and log Pr[data] reproduce the reference bit for bit.  That needs
//   * p7_FLogsum literally (logsum.hpp:141-154) -- here in seven instructions with the same bits (nc_fwbw_core.cuh);
//   * the reference's accumulation ORDER: from -inf, over the ascending merged predecessor list from_v(j)
//     (forward) / successor list to_v(j) (backward), each edge with its exact weight.
// The lists are never materialised; the per-thread column code (shared chain prefixes, merged duplicate edges) lives
// in nc_fwbw_core.cuh as __host__ __device__ functions and is checked against the oracle on the host as well
// (tests/test_fwbw_emu.py).
//
// Kernels (one wave = the sequences whose E/alpha/beta slabs fit the scratch pool):
//   emission_kernel   E[i][j] = log_pr_corrected_emission(j, e_i) for every sequence (grid: seq x event tiles)
//   fwbw_kernel       one CTA per sequence: forward, log Pr[data], backward; alpha/beta to the slabs
//   pm_stats_kernel   per event the six posterior-weighted sums of train_pm_params (:263-296), each the reference's
//                     serial j = 0..4095 loop, 96 of them (16 events x 6 sums) in the lanes of three warps
//   st_stats_kernel   one CTA per (group, strand): the three log-space accumulators of train_st_params (:471-514),
//                     folded in the reference's (sequence, event, k-mer) order, 32 terms per step (speculate the table
//                     indices from the running value, verify, repeat: exactly the sequential result)
// The 3x3 solve, the clamps and exp() of the final ratios run on the host in nc_train.cu.
#include "nc_device.cuh"
#include "nc_fwbw_core.cuh"
#include "nc_kernels.h"

namespace nc {

namespace {

using fb::cphys;
using fb::flogsum;

constexpr int FB_THREADS = fb::THREADS;
constexpr int FB_SPT = fb::SPT;
constexpr int COL_FLOATS = fb::COL_FLOATS;
constexpr int FOLD_LIST = 80;
constexpr unsigned FULL = 0xffffffffu;

// acc = p7_FLogsum(acc, x_k) over k = 0..n-1 in order, n a multiple of 32, x_k = get(k).
// A term with x == -inf or acc - x >= 15.999 returns acc unchanged now and for every larger acc (the running value
// never decreases), so it is screened out; the survivors are folded from the per-warp list, padded with -inf.
// (log Pr[data]: one call per sequence; the trainer's long chains use fold32 below.)
template < typename Get, typename TB >
__device__ __forceinline__ float fold_logsum_in_order(float acc, const int n, Get get, float* __restrict__ lst,
                                                      const TB& tbl, const int lane)
{
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < n; base += 32)
    {
        const float x = get(base + lane);
        const bool live = !((x == NC_NEG_INF) || (acc > x && __fsub_rn(acc, x) >= 15.999f));
        const unsigned m = __ballot_sync(FULL, live);
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

// One step of a long in-order p7_FLogsum chain: acc (warp-uniform) absorbs the 32 terms x (one per lane, lane order).
// While acc >= x the step is acc' = acc + tbl[idx(acc - x)]: the lanes look their increments up with the value the
// chain had at the START of the block, the live lanes' increments are compacted into a warp-private list, every lane
// then forms its own running value a_k = acc + t_0 + ... + t_{k-1} (sequential float adds over the live terms below
// it, all lanes in lock step: a chain of 4-cycle FADDs fed by broadcast loads), and recomputes its index from a_k.
// When every index is confirmed the a_k are by induction the sequential chain's values and the last live lane's
// a + t is the result; otherwise the recomputed indices are the next guess (the first unconfirmed lane is certainly
// right then).  After three rounds, or when a term exceeds the running value (start of a chain), the block is folded
// term by term.  lst: 40 floats, 16-byte aligned, private to the warp.
#ifndef NC_FB_OPAQUE
#define NC_FB_OPAQUE 1
#endif
#ifndef NC_ST_EXP
#define NC_ST_EXP 0   // timing experiments only: 1 = no phase-2 arithmetic, 2 = no fold; anything but 0 is not a product build
#endif
#ifndef NC_ST_BLOCK
#define NC_ST_BLOCK 32   // 64: fold64 when that many terms are published (measured: a chain alone 2.60 -> 2.45 ms, but a full wave 19.9 -> 21.3 ms: more instructions per term)
#endif
#ifndef NC_ST_PRE
#define NC_ST_PRE 1
#endif
#if NC_ST_EXP == 9
__device__ unsigned long long g_dbg[8];   // blocks, speculative blocks, exact rounds, sequential fallbacks, live items
#define FOLD_DBG(k, v) do { if (lane == 0) atomicAdd(&g_dbg[k], (unsigned long long)(v)); } while (0)
#else
#define FOLD_DBG(k, v) do { } while (0)
#endif
template < typename TB >
__device__ __forceinline__ float fold32(float acc, const float x, const TB& tbl, const int lane, float* __restrict__ lst)
{
    const bool live = !((x == NC_NEG_INF) || (acc > x && __fsub_rn(acc, x) >= 15.999f));
    const unsigned m = __ballot_sync(FULL, live);
    if (m == 0) return acc;
    const bool spec_ok = __all_sync(FULL, !live || acc >= x);   // (false for acc == -inf with a finite term, and for NaN)
    FOLD_DBG(0, 1); FOLD_DBG(4, __popc(m));
    if (spec_ok)
    {
        FOLD_DBG(1, 1);
        const float INF = __int_as_float(0x7f800000);   // dead lanes: d = +inf selects the zero entry
        const int L = __popc(m);
        const int r = __popc(m & ((1u << lane) - 1u));   // live terms below this lane
        const int last = 31 - __clz((int)m);
        unsigned e = tbl.addr(live ? __fsub_rn(acc, x) : INF);
        // A cheap refinement of the guess before the exact chain (NC_ST_PRE of them; measured: 1.75 exact rounds per block
        // without, 1.04 with one, 1.006 with two, and one is the fastest): a parallel prefix of the increments (rounded
        // differently from the sequential sum, which only matters within an ulp of an index boundary) gives every lane an
        // estimate of its running value and with it a better index.  The exact rounds below verify whatever comes out.
#pragma unroll
        for (int pre = 0; pre < NC_ST_PRE; ++pre)
        {
            float p = live ? tbl.load(e) : 0.0f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const float v = __shfl_up_sync(FULL, p, d);
                if (lane >= d) p = __fadd_rn(p, v);
            }
            const float before = __shfl_up_sync(FULL, p, 1);                  // increments of the lanes below
            const float est = __fadd_rn(acc, lane ? before : 0.0f);
            e = tbl.addr(live ? fmaxf(__fsub_rn(est, x), 0.0f) : INF);
        }
#pragma unroll 1
        for (int round = 0; round < 3; ++round)
        {
            FOLD_DBG(2, 1);
            const float t = tbl.load(e);
            if (live) lst[r] = t;
            if (lane < 8) lst[L + lane] = 0.0f;
            __syncwarp();
            float a = acc;
            float4 v = *reinterpret_cast< const float4* >(lst);
            for (int k = 0; k < L; k += 4)
            {
                // the next four increments are fetched before these are added: the chain of FADDs never waits for a load
                // (a lane stops at its own position by adding zeros: the selects are off the chain, which is FADD after FADD)
                const float4 vn = *reinterpret_cast< const float4* >(lst + k + 4);
                const float t0 = (k + 0 < r) ? v.x : 0.0f, t1 = (k + 1 < r) ? v.y : 0.0f;
                const float t2 = (k + 2 < r) ? v.z : 0.0f, t3 = (k + 3 < r) ? v.w : 0.0f;
                a = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, t0), t1), t2), t3);
                v = vn;
            }
            __syncwarp();
            const unsigned e2 = tbl.addr(live ? __fsub_rn(a, x) : INF);
            const bool ok = !live || (a >= x && e2 == e);
            if (__all_sync(FULL, ok)) return __shfl_sync(FULL, __fadd_rn(a, t), last);
            e = e2;
        }
    }
    FOLD_DBG(3, 1);
    for (unsigned mm = m; mm; mm &= mm - 1)
    {
        const int k = __ffs(mm) - 1;
        acc = flogsum(acc, __shfl_sync(FULL, x, k), tbl);
    }
    return acc;
}

// The same step for 64 terms (lane l holds terms l and 32 + l).  Most of a step's time is fixed -- two table lookups that
// wait for the slowest lane, the votes, the list -- so twice the terms per step is what makes the chain faster; the
// per-term work (one FADD on the chain, the captures off it) is the same.  lst: 72 floats.
template < typename TB >
__device__ __forceinline__ float fold64(float acc, const float x0, const float x1, const TB& tbl, const int lane, float* __restrict__ lst)
{
    const bool live0 = !((x0 == NC_NEG_INF) || (acc > x0 && __fsub_rn(acc, x0) >= 15.999f));
    const bool live1 = !((x1 == NC_NEG_INF) || (acc > x1 && __fsub_rn(acc, x1) >= 15.999f));
    const unsigned m0 = __ballot_sync(FULL, live0), m1 = __ballot_sync(FULL, live1);
    if ((m0 | m1) == 0u) return acc;
    const bool spec_ok = __all_sync(FULL, (!live0 || acc >= x0) && (!live1 || acc >= x1));
    if (spec_ok)
    {
        const float INF = __int_as_float(0x7f800000);
        const unsigned lt = (1u << lane) - 1u;
        const int L0 = __popc(m0), L = L0 + __popc(m1);
        const int r0 = __popc(m0 & lt), r1 = L0 + __popc(m1 & lt);   // live terms before each of this lane's two
        unsigned e0 = tbl.addr(live0 ? __fsub_rn(acc, x0) : INF), e1 = tbl.addr(live1 ? __fsub_rn(acc, x1) : INF);
#pragma unroll
        for (int pre = 0; pre < NC_ST_PRE; ++pre)
        {
            float p0 = live0 ? tbl.load(e0) : 0.0f, p1 = live1 ? tbl.load(e1) : 0.0f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const float v0 = __shfl_up_sync(FULL, p0, d), v1 = __shfl_up_sync(FULL, p1, d);
                if (lane >= d) { p0 = __fadd_rn(p0, v0); p1 = __fadd_rn(p1, v1); }
            }
            const float tot0 = __shfl_sync(FULL, p0, 31);
            const float b0 = __shfl_up_sync(FULL, p0, 1), b1 = __shfl_up_sync(FULL, p1, 1);
            const float est0 = __fadd_rn(acc, lane ? b0 : 0.0f);
            const float est1 = __fadd_rn(acc, __fadd_rn(tot0, lane ? b1 : 0.0f));
            e0 = tbl.addr(live0 ? fmaxf(__fsub_rn(est0, x0), 0.0f) : INF);
            e1 = tbl.addr(live1 ? fmaxf(__fsub_rn(est1, x1), 0.0f) : INF);
        }
#pragma unroll 1
        for (int round = 0; round < 3; ++round)
        {
            const float t0 = tbl.load(e0), t1 = tbl.load(e1);
            if (live0) lst[r0] = t0;
            if (live1) lst[r1] = t1;
            if (lane < 8) lst[L + lane] = 0.0f;
            __syncwarp();
            // one chain over all live increments; the lane keeps the running value at its two positions
            float a = acc, a0 = acc, a1 = acc;
            float4 v = *reinterpret_cast< const float4* >(lst);
            for (int k = 0; k < L; k += 4)
            {
                const float4 vn = *reinterpret_cast< const float4* >(lst + k + 4);
                a0 = (k + 0 == r0) ? a : a0; a1 = (k + 0 == r1) ? a : a1; a = __fadd_rn(a, v.x);
                a0 = (k + 1 == r0) ? a : a0; a1 = (k + 1 == r1) ? a : a1; a = __fadd_rn(a, v.y);
                a0 = (k + 2 == r0) ? a : a0; a1 = (k + 2 == r1) ? a : a1; a = __fadd_rn(a, v.z);
                a0 = (k + 3 == r0) ? a : a0; a1 = (k + 3 == r1) ? a : a1; a = __fadd_rn(a, v.w);
                v = vn;
            }
            __syncwarp();
            const unsigned f0 = tbl.addr(live0 ? __fsub_rn(a0, x0) : INF), f1 = tbl.addr(live1 ? __fsub_rn(a1, x1) : INF);
            const bool ok = (!live0 || (a0 >= x0 && f0 == e0)) && (!live1 || (a1 >= x1 && f1 == e1));
            if (__all_sync(FULL, ok))
            {
                // the value after the last live increment: a holds acc + all L increments (the padding adds zeros)
                return __shfl_sync(FULL, a, 0);
            }
            e0 = f0; e1 = f1;
        }
    }
    for (unsigned mm = m0; mm; mm &= mm - 1)
    {
        const int k = __ffs(mm) - 1;
        acc = flogsum(acc, __shfl_sync(FULL, x0, k), tbl);
    }
    for (unsigned mm = m1; mm; mm &= mm - 1)
    {
        const int k = __ffs(mm) - 1;
        acc = flogsum(acc, __shfl_sync(FULL, x1, k), tbl);
    }
    return acc;
}

__device__ __forceinline__ void prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct FbSmem
{
    float tbl[fb::TBL_N];
    float col[2][COL_FLOATS];
    float lut[64];
    float lst[FOLD_LIST];
    unsigned item;
};

// st_stats_kernel's geometry.  The three fold warps of a CTA run dependent chains and every producer warp is a chain of
// memory latencies, so what counts is how many of them are in flight: CTAs of 256 threads (3 fold warps + 5 producer
// warps, each producer warp working on its own event), six per SM, the p7_FLogsum table read through L1 instead of
// 64 KB of shared memory per CTA.
constexpr int ST_THREADS = 256;
constexpr int ST_CTAS = 6;
constexpr int ST_PW = ST_THREADS / 32 - 3;     // producer warps
constexpr int ST_CH = 68;                      // training k-mers per lane and event (68 x 32 >= 2160), in 32-wide chunks
constexpr int ST_LIST = 512;                   // live items a producer warp lists at a time
constexpr int ST_RING = 1024;                  // items between the producers and the slowest fold warp (several events)
struct StSmem
{
    float ring[3][ST_RING];       // terms of the live items, in order, one ring per accumulator (slot = item index mod ST_RING)
    unsigned short list[ST_PW][ST_LIST];   // per producer warp: positions (in the training k-mer list) of live items, ascending
    unsigned e_claim;             // next event to hand to a producer warp
    unsigned e_counted;           // events whose live items have been counted (their item offsets are fixed)
    unsigned n_counted;           // live items of those events
    unsigned e_pub;               // events completely published
    unsigned pub;                 // items published so far (a prefix of the item sequence)
    unsigned done;                // set after the last publication
    unsigned cons[4];             // items consumed by each fold warp
    float acc_snap[4];            // running values the fold warps published last
    __align__(16) float lst[3][80];   // fold32's / fold64's warp-private lists
    unsigned seq_id[NC_MAX_TRAIN_SEQS];      // the strand's sequences with >= 2 events, in order
    unsigned seq_first[NC_MAX_TRAIN_SEQS + 1];  // first flattened (event) index of each
    unsigned n_seq;
};

// posteriors are scaled by 2^64 inside pm_stats_kernel: powers of two commute with every rounding as long as nothing
// underflows, and the scaled terms never do (p >= 2^-149 becomes >= 2^-85), so the six sums are the reference's sums
// times 2^64, bit for bit, except that terms the reference computes as DENORMALS (p/sigma^2 below 2^-126: they cannot
// influence a sum that contains one posterior above 2^-100) keep their full precision here
constexpr float PM_SCALE = 18446744073709551616.0f;
constexpr int PM_EV = 16;                 // events per CTA of pm_stats_kernel (= FB_EV_TILE)
constexpr int PM_CHAINS = PM_EV * 6;      // 96 serial sums = the lanes of three warps
constexpr int PM_ROW = PM_CHAINS + 2;     // row stride of the term buffer: even (the producers store float2), and the 16 lanes of a
                                          // half warp (consecutive states) cover all 32 banks
constexpr int PM_JT = 52;                 // states per tile: 13 producer warps x 32 = 8 event-pairs x 52 states
constexpr int PM_TILES = (NC_N_STATES + PM_JT - 1) / PM_JT;   // 79
static_assert(FB_EV_TILE == PM_EV, "pm_stats grid uses FB_EV_TILE");
static_assert((FB_THREADS - 96) == 8 * PM_JT, "13 producer warps = 8 x PM_JT");

} // namespace

// ------------------------------------------------------------------------------------------------
// E[seq][i][j].  One CTA = EM_TILE events of one sequence: the event scalars (drift-corrected mean, stdv, 3 log stdv,
// 1/stdv: a logf and a reciprocal each) are computed once per CTA by its first threads, not by every thread, and the
// scaled state constants of a thread (scale_state: divisions and logarithms) are spread over EM_TILE events.
__global__ void __launch_bounds__(FB_THREADS) emission_kernel(const FbArgs a)
{
    __shared__ float4 ev_s[EM_TILE];   // {x, y, 3 log y, 1/y}
    const unsigned seq = blockIdx.y;
    const FbSeq& Q = a.seqs[seq];
    const unsigned i0 = blockIdx.x * EM_TILE;
    if (i0 >= Q.n_events) return;
    const DevJob& J = a.jobs[Q.job];
    const int t = threadIdx.x;
    const unsigned j0 = 4 * t;   // a warp's float4 stores are 512 contiguous bytes (full sectors); second half: + 2048
    const float* M = a.models + (size_t)J.model * MODEL_FLOATS;
    float* E = a.scratch + Q.slab + 0 * (size_t)Q.n_events * NC_N_STATES;
    const unsigned i1 = min(i0 + EM_TILE, Q.n_events);
    if (i0 + t < i1)
    {
        const unsigned long long e = Q.ev_off + i0 + t;
        const float stdv = __ldg(a.stdv + e);
        const float y = (stdv == 0.0f) ? 0.01f : stdv;                                  // Event.hpp:39-42
        const float x = __fsub_rn(__ldg(a.mean + e), __fmul_rn(J.drift, __ldg(a.start + e)));  // Event.hpp:81
        const float ly3 = __fmul_rn(3.0f, a.log_stdv ? __ldg(a.log_stdv + e) : nc_logf(y));
        ev_s[t] = make_float4(x, y, ly3, __frcp_rn(y));
    }
    __syncthreads();
#pragma unroll 1
    for (int half = 0; half < 2; ++half)
    {
        StateParams P[4];
        const unsigned jb = j0 + (NC_N_STATES / 2) * half;
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
        float* Ei = E + (size_t)i0 * NC_N_STATES + jb;
        for (unsigned i = i0; i < i1; ++i, Ei += NC_N_STATES)
        {
            const float4 ev = ev_s[i - i0];
            float4 o;
            o.x = emission(P[0], ev.x, ev.y, ev.z, ev.w, a.log_2pi);
            o.y = emission(P[1], ev.x, ev.y, ev.z, ev.w, a.log_2pi);
            o.z = emission(P[2], ev.x, ev.y, ev.z, ev.w, a.log_2pi);
            o.w = emission(P[3], ev.x, ev.y, ev.z, ev.w, a.log_2pi);
            *reinterpret_cast< float4* >(Ei) = o;
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
    for (int q = t; q < fb::TBL_N; q += FB_THREADS) sm.tbl[q] = a.logsum_tbl[q];   // entry 15999 is 0 (nc_train.cu)
    fb::TblSmem tbl = fb::make_tbl_smem(sm.tbl);
#if NC_FB_OPAQUE
    // ptxas otherwise rebuilds the constant from the CTA's shared-window base (S2UR, UMOV, ULEA, IMAD, I2FP, FMUL, FADD:
    // seven instructions) in front of every group of folds instead of holding it in a register
    asm volatile("" : "+f"(tbl.magic));
#endif

    for (;;)
    {
        __syncthreads();
        if (t == 0) sm.item = atomicAdd(a.next_item, 1u);
        __syncthreads();
        const unsigned seq = sm.item;
        if (seq >= a.n_seqs) break;
        const FbSeq& Q = a.seqs[seq];
        if (Q.generic) continue;   // a custom transition table applies: fwbw_generic_kernel's sequence
        const DevJob& J = a.jobs[Q.job];
        const unsigned n = Q.n_events;
        const float* E = a.scratch + Q.slab;
        float* AL = a.scratch + Q.slab + 1 * (size_t)n * NC_N_STATES;
        float* BE = a.scratch + Q.slab + 2 * (size_t)n * NC_N_STATES;
        if (t < 64) sm.lut[t] = J.lut[t];
        __syncthreads();

        // =========================== forward (Forward_Backward.hpp:58-89); logical thread u owns j = 8u .. 8u+7
        {
            fb::FwdConst C;
            fb::fwd_const_init(C, fb::fwd_logical_thread(t), sm.lut);
#if NC_FB_OPAQUE
            // (the same for the per-thread order flags: recomputed from threadIdx inside the column loops, ~20 integer
            // instructions per one-step slot, unless their origin is hidden from the compiler)
            asm volatile("" : "+r"(C.flags), "+r"(C.pos), "+r"(C.sS), "+r"(C.c));
#endif
            const unsigned j0 = FB_SPT * (unsigned)C.u;
            const int p0 = cphys((int)j0), p1 = cphys((int)j0 + 4);
            float own[8];
            // column 0
            {
                const float4 e0 = *reinterpret_cast< const float4* >(E + j0);
                const float4 e1 = *reinterpret_cast< const float4* >(E + j0 + 4);
                own[0] = __fsub_rn(e0.x, a.log_n_states); own[1] = __fsub_rn(e0.y, a.log_n_states);
                own[2] = __fsub_rn(e0.z, a.log_n_states); own[3] = __fsub_rn(e0.w, a.log_n_states);
                own[4] = __fsub_rn(e1.x, a.log_n_states); own[5] = __fsub_rn(e1.y, a.log_n_states);
                own[6] = __fsub_rn(e1.z, a.log_n_states); own[7] = __fsub_rn(e1.w, a.log_n_states);
                const float4 a0 = make_float4(own[0], own[1], own[2], own[3]), a1 = make_float4(own[4], own[5], own[6], own[7]);
                *reinterpret_cast< float4* >(sm.col[0] + p0) = a0;
                *reinterpret_cast< float4* >(sm.col[0] + p1) = a1;
                *reinterpret_cast< float4* >(AL + j0) = a0;
                *reinterpret_cast< float4* >(AL + j0 + 4) = a1;
            }
            __syncthreads();
            int cur = 0;
            for (unsigned i = 1; i < n; ++i)
            {
                // the emissions were written for the whole wave before this kernel started (far more than L2 holds): the
                // next column's row is fetched into L2 now, so that its loads wait for L2 and not for DRAM
                if (t < 128 && i + 1 < n) prefetch_l2(E + (size_t)(i + 1) * NC_N_STATES + 32 * t);
                const float* Ei = E + (size_t)i * NC_N_STATES + j0;
                const float4 e0 = __ldg(reinterpret_cast< const float4* >(Ei));
                const float4 e1 = __ldg(reinterpret_cast< const float4* >(Ei + 4));
                const float e[8] = { e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w };
                fb::fwd_column(C, sm.col[cur], tbl, e, own);
                const float4 a0 = make_float4(own[0], own[1], own[2], own[3]), a1 = make_float4(own[4], own[5], own[6], own[7]);
                float* An = sm.col[cur ^ 1];
                *reinterpret_cast< float4* >(An + p0) = a0;
                *reinterpret_cast< float4* >(An + p1) = a1;
                float* Ao = AL + (size_t)i * NC_N_STATES + j0;
                *reinterpret_cast< float4* >(Ao) = a0;
                *reinterpret_cast< float4* >(Ao + 4) = a1;
                cur ^= 1;
                __syncthreads();
            }
            // log Pr[data]: sequential fold of the last column, ascending j (Forward_Backward.hpp:129-134), one warp
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
            fb::BwdConst C;
            fb::bwd_const_init(C, t, sm.lut);
#pragma unroll
            for (int k = 0; k < FB_SPT; ++k)
            {
                const unsigned code = fb::bwd_lane_code(t, k);
                const unsigned c0 = __shfl_sync(FULL, code, 0);
                if (__all_sync(FULL, code == c0)) C.paths |= c0 << (3 * k);
            }
#if NC_FB_OPAQUE
            asm volatile("" : "+r"(C.flags), "+r"(C.paths), "+r"(C.smask[0]), "+r"(C.smask[1]));
#endif
            // beta[n-1] = 0
#pragma unroll
            for (int k = 0; k < FB_SPT; ++k)
            {
                const int j = t + FB_THREADS * k;
                sm.col[0][cphys(j)] = 0.0f;
                BE[(size_t)(n - 1) * NC_N_STATES + j] = 0.0f;
            }
            __syncthreads();
            int cur = 0;
            for (unsigned ip1 = n - 1; ip1 > 0; --ip1)
            {
                if (t < 128 && ip1 >= 2) prefetch_l2(E + (size_t)(ip1 - 1) * NC_N_STATES + 32 * t);
                float* Bc = sm.col[cur ^ 1];
                float* Bo = BE + (size_t)(ip1 - 1) * NC_N_STATES;
                fb::bwd_column(C, sm.col[cur], E + (size_t)ip1 * NC_N_STATES, sm.lut, tbl,
                               [&](int j, float v) { Bc[cphys(j)] = v; Bo[j] = v; });
                cur ^= 1;
                __syncthreads();
            }
        }
    }
}

size_t fwbw_smem_bytes() { return sizeof(FbSmem); }
size_t st_stats_smem_bytes() { return sizeof(StSmem); }
unsigned st_stats_threads() { return ST_THREADS; }
unsigned st_stats_max_kmers() { return ST_CH * 32; }

// ------------------------------------------------------------------------------------------------
// train_pm_params' inner sums (Parameter_Trainer.hpp:263-296): per event
//   s0 = sum_j p/sigma^2, s1 = sum_j p*mu/sigma^2, s2 = sum_j p*mu^2/sigma^2,
//   l0 = sum_j p*lambda,  l1 = sum_j p*lambda/eta,  l2 = sum_j p*lambda/eta^2      (UNSCALED model)
// with p = exp(alpha + beta - logZ).  The 3x3 system built from these sums is ill-conditioned (level means are
// 58 +- 6 pA), so the float rounding of the reference's SEQUENTIAL j = 0..4095 accumulation is visible in the trained
// shift/scale/var at the 1e-4 level: every sum is the reference's serial loop itself.  A serial sum is a chain of 4096
// dependent FADDs (4 cycles each), so one CTA runs 96 of them at once -- 16 events x 6 sums, one per lane of warps
// 0..2 -- while warps 3..15 produce the terms of the next 52 states for all 16 events into the other half of a
// double-buffered tile (row stride 97 floats: the producers' lanes are consecutive states, the folding lanes
// consecutive sums; both conflict-free).
__global__ void __launch_bounds__(FB_THREADS) pm_stats_kernel(const FbArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* term = reinterpret_cast< float* >(smem_raw);  // [2][PM_JT][PM_ROW]
    const unsigned seq = blockIdx.y;
    const FbSeq& Q = a.seqs[seq];
    const unsigned i0 = blockIdx.x * PM_EV;
    if (i0 >= Q.n_events) return;
    const unsigned n = Q.n_events;
    const DevJob& J = a.jobs[Q.job];
    const int t = threadIdx.x;
    const float* AL = a.scratch + Q.slab + 1 * (size_t)n * NC_N_STATES;
    const float* BE = a.scratch + Q.slab + 2 * (size_t)n * NC_N_STATES;
    const float logz = a.log_pr_data[seq];
    const unsigned n_ev = min((unsigned)PM_EV, n - i0);

    if (t < PM_CHAINS)
    {
        // ---- folding lanes: chain c = 6 * event + sum
        float acc = 0.0f;
        for (int tile = 0; tile < PM_TILES; ++tile)
        {
            __syncthreads();   // tile `tile` is complete in buffer tile & 1
            const float* T = term + (size_t)(tile & 1) * PM_JT * PM_ROW + t;
            const int nj = min(PM_JT, (int)NC_N_STATES - tile * PM_JT);
#pragma unroll 4
            for (int jl = 0; jl < nj; ++jl) acc = __fadd_rn(acc, T[jl * PM_ROW]);
        }
        const unsigned ev = (unsigned)t / 6u, s = (unsigned)t % 6u;
        if (ev < n_ev) a.pm_stats[(Q.ev_out + i0 + ev) * 6 + s] = __fmul_rn(acc, 1.0f / PM_SCALE);
    }
    else
    {
        // ---- producers: thread p handles state tile*52 + (p % 52) for the events (p / 52) and (p / 52) + 8.
        // Everything that does not change from tile to tile is a pointer stepped by 52 states or a flag set up here, and
        // the state's constants (mu, sigma^2, 1/sigma^2, eta, 1/eta, lambda: two correctly rounded reciprocals) come
        // from a per-model table (FbArgs::pm_consts) instead of being rebuilt by every CTA for every pair of events.
        const int p = t - PM_CHAINS;
        const int jl = p % PM_JT, eg = p / PM_JT;   // eg = 0..7
        const bool on0 = (unsigned)eg < n_ev, on1 = (unsigned)eg + 8u < n_ev;
        const float* pA = AL + (size_t)(i0 + eg) * NC_N_STATES + jl;   // event eg; event eg + 8 lies 8 rows further
        const float* pB = BE + (size_t)(i0 + eg) * NC_N_STATES + jl;
        constexpr int ROW8 = 8 * NC_N_STATES;
        const float4* pC = a.pm_consts + ((size_t)J.model * NC_N_STATES + jl) * 2;
        float* T0 = term + jl * PM_ROW + 6 * eg;                       // this thread's six terms of event eg, buffer 0
        constexpr int BUF = PM_JT * PM_ROW;
        // alpha + beta of the NEXT tile are loaded while the current one is computed (the loads come from HBM: every
        // tile touches a new 208-byte piece of each of the 16 event rows)
        float al0 = 0.f, be0 = 0.f, al1 = 0.f, be1 = 0.f;
        if (on0) { al0 = __ldcs(pA); be0 = __ldcs(pB); }
        if (on1) { al1 = __ldcs(pA + ROW8); be1 = __ldcs(pB + ROW8); }
        auto terms = [&](const float4& c0, const float4& c1, float al, float be, float* R) {
            // the posterior, scaled by 2^64 (exact): every term below stays in the normal range, so the correctly rounded
            // divisions are three instructions (div_rn) instead of the IEEE division's denormal slow path, which two
            // thirds of the kernel's instructions used to be
            const float pst = __fmul_rn(nc_expf(__fsub_rn(__fadd_rn(al, be), logz)), PM_SCALE);
            const float ts0 = div_rn(pst, c0.y, c0.z);        // / sigma^2
            const float ts1 = __fmul_rn(ts0, c0.x);           // * mu
            const float ts2 = __fmul_rn(ts1, c0.x);
            const float tl0 = __fmul_rn(pst, c1.y);           // * lambda
            const float tl1 = div_rn(tl0, c0.w, c1.x);        // / eta
            const float tl2 = div_rn(tl1, c0.w, c1.x);
            *reinterpret_cast< float2* >(R) = make_float2(ts0, ts1);
            *reinterpret_cast< float2* >(R + 2) = make_float2(ts2, tl0);
            *reinterpret_cast< float2* >(R + 4) = make_float2(tl1, tl2);
        };
        const float2 Z2 = make_float2(0.f, 0.f);
        int j = jl;
        // the state constants of the next tile are requested as soon as this tile's terms are done (their first use was
        // 10 % of the kernel's stall samples); an L2 prefetch of alpha / beta two tiles ahead was measured too and is
        // slower (13.9 against 13.3 ms): the memory system, not the latency of one request, is what these loads wait for
        float4 c0 = __ldg(pC), c1 = __ldg(pC + 1);
#pragma unroll 2
        for (int tile = 0; tile < PM_TILES; ++tile, j += PM_JT, pA += PM_JT, pB += PM_JT, pC += 2 * PM_JT)
        {
            float aln0 = 0.f, ben0 = 0.f, aln1 = 0.f, ben1 = 0.f;
            if (j + PM_JT < (int)NC_N_STATES)
            {
                if (on0) { aln0 = __ldcs(pA + PM_JT); ben0 = __ldcs(pB + PM_JT); }
                if (on1) { aln1 = __ldcs(pA + PM_JT + ROW8); ben1 = __ldcs(pB + PM_JT + ROW8); }
            }
            float* R = T0 + (tile & 1) * BUF;
            if (j < (int)NC_N_STATES)
            {
                if (on0) terms(c0, c1, al0, be0, R);
                else { *reinterpret_cast< float2* >(R) = Z2; *reinterpret_cast< float2* >(R + 2) = Z2; *reinterpret_cast< float2* >(R + 4) = Z2; }
                if (on1) terms(c0, c1, al1, be1, R + 48);
                else { *reinterpret_cast< float2* >(R + 48) = Z2; *reinterpret_cast< float2* >(R + 50) = Z2; *reinterpret_cast< float2* >(R + 52) = Z2; }
                if (j + PM_JT < (int)NC_N_STATES) { c0 = __ldg(pC + 2 * PM_JT); c1 = __ldg(pC + 2 * PM_JT + 1); }
            }
            __syncthreads();   // publishes tile `tile`; the folding lanes are at most one tile behind
            al0 = aln0; be0 = ben0; al1 = aln1; be1 = ben1;
        }
    }
}

size_t pm_stats_smem_bytes() { return (size_t)2 * PM_JT * PM_ROW * sizeof(float); }

// ------------------------------------------------------------------------------------------------
// train_st_params' accumulators (Parameter_Trainer.hpp:434-517) for one (group, strand):
//   denom (+)= post(i,j1);  stay (+)= min(joint(j1->j1 | log p_stay), post);
//   skip (+)= log(exp(post) - exp(min(d01, post))),  d01 = stay' (+) the 4 one-step joints with log(p_step/4)
// over the strand's sequences in order, events i < n-1, the 2160 training k-mers in ascending order: ONE chain of
// (events x 2160) terms per accumulator, (+) = p7_FLogsum.
//
// A term more than 15.999 below the running value leaves it unchanged, and the three terms of a k-mer are bounded by its
// log posterior (stay and skip are clamped to it; log(exp(.)) round trips stay within a few ulps).  Only the fold is
// sequential; which items can matter, and their terms, can be worked out for several events at once:
//   producer warps  each takes the next event: (1) log posterior of every training k-mer, two loads each; a k-mer whose
//            log posterior is 16.25 below the smallest of the three running values the fold warps last published (stale
//            values are smaller, hence conservative) is dead for all three chains: one flag bit per k-mer, in registers;
//            (2) the number of live items fixes where the event's items go in the item sequence, once the events before
//            it are counted (e_counted / n_counted, in event order); (3) the full terms (five joint probabilities,
//            p7_FLogsum, two expf, one logf) of the live k-mers only -- typically a tenth -- 32 at a time into the rings;
//            (4) publication in event order (the warp whose event is the oldest unpublished one also publishes its
//            partial progress, so an event with more live items than the ring holds cannot block).
//   fold warps      one accumulator each, absorb the published items 32 per step (fold32) as they come.
// A chain of 200 events took 3.4 ms when the producers handled one event at a time (every event a chain of three DRAM
// latencies); the kernel's time is that latency times the number of CTA rounds of a wave.
// loads of the alpha / beta / emission rows: each is used once or twice, so they must not push the p7_FLogsum table
// (read through L1 by the fold warps, whose chain waits for every lookup) out of the cache
__device__ __forceinline__ float ld_stream(const float* p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS) st_stats_kernel(const FbArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StSmem& ss = *reinterpret_cast< StSmem* >(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned grp = blockIdx.x;
    const unsigned st = blockIdx.y;
    const FbGroup& G = a.groups[grp];
    if (t == 0)
    {
        unsigned ns = 0, first = 0;
        for (unsigned sq = G.seq_begin; sq < G.seq_end; ++sq)
            if (a.seqs[sq].strand == st && a.seqs[sq].n_events >= 2)
            {
                ss.seq_id[ns] = sq;
                ss.seq_first[ns] = first;
                first += a.seqs[sq].n_events - 1;
                ++ns;
            }
        ss.seq_first[ns] = first;
        ss.n_seq = ns;
        ss.e_claim = ss.e_counted = ss.n_counted = ss.e_pub = ss.pub = ss.done = 0u;
    }
    if (t < 4) { ss.acc_snap[t] = NC_NEG_INF; ss.cons[t] = 0u; }
    __syncthreads();
    fb::TblPtr tbl;
    tbl.p = a.logsum_tbl;
    const unsigned n_km = a.n_train_kmers;   // <= 32 ST_CH (checked on the host)
    const unsigned n_seq = ss.n_seq;
    const unsigned n_ev = ss.seq_first[n_seq];
    volatile unsigned* const v_pub = &ss.pub;
    volatile unsigned* const v_done = &ss.done;
    volatile unsigned* const v_cons = ss.cons;
    volatile float* const v_snap = ss.acc_snap;
    volatile unsigned* const v_e_counted = &ss.e_counted;
    volatile unsigned* const v_n_counted = &ss.n_counted;
    volatile unsigned* const v_e_pub = &ss.e_pub;

    if (warp >= 3)
    {
        // ================================================= producer warps, one event each at a time
        const int pw = warp - 3;
        const float log_p_stay = G.log_p_stay[st];
        const float log_p_step_4 = G.log_p_step_4[st];
        unsigned short* const list = ss.list[pw];
        const unsigned lt = (1u << lane) - 1u;
        if (n_ev == 0 && pw == 0 && lane == 0) *v_done = 1u;
#if NC_ST_EXP == 9
        long long tk[6] = { 0, 0, 0, 0, 0, 0 }; long long t_prev = clock64(); unsigned n_items_dbg = 0, n_ev_dbg = 0;
#define ST_TICK(k) do { const long long _n = clock64(); tk[k] += _n - t_prev; t_prev = _n; } while (0)
#else
#define ST_TICK(k) do { } while (0)
#endif
        for (;;)
        {
            unsigned ef = 0;
            if (lane == 0) ef = atomicAdd(&ss.e_claim, 1u);
            ef = __shfl_sync(FULL, ef, 0);
            if (ef >= n_ev) break;
            unsigned s = 0;
            while (s + 1 < n_seq && ef >= ss.seq_first[s + 1]) ++s;
            const unsigned sq = ss.seq_id[s], i = ef - ss.seq_first[s];
            const FbSeq& Q = a.seqs[sq];
            const unsigned n = Q.n_events;
            const float* E = a.scratch + Q.slab;
            const float* AL = E + 1 * (size_t)n * NC_N_STATES;
            const float* BE = E + 2 * (size_t)n * NC_N_STATES;
            const float logz = a.log_pr_data[sq];
            const float* Ai = AL + (size_t)i * NC_N_STATES;
            const float* Bi = BE + (size_t)i * NC_N_STATES;
            const float* Bn = BE + (size_t)(i + 1) * NC_N_STATES;
            const float* En = E + (size_t)(i + 1) * NC_N_STATES;
            // ---- (1) which k-mers are live: lane l looks at k-mers 32 c + l, flag bit c
            const float smin = fminf(fminf(v_snap[0], v_snap[1]), v_snap[2]);
            unsigned long long flo = 0ull;   // flag bits of chunks 0..63
            unsigned fhi = 0u;               // ... 64..67
            unsigned mine = 0;
#pragma unroll 1
            for (int c0 = 0; c0 < ST_CH; c0 += 8)
            {
                float av[8], bv[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                {
                    const unsigned km = (unsigned)(32 * (c0 + r) + lane);
                    if (c0 + r < ST_CH && km < n_km)
                    {
                        const unsigned j1 = __ldg(a.train_kmers + km);
                        av[r] = ld_stream(Ai + j1);
                        bv[r] = ld_stream(Bi + j1);
                    }
                    else av[r] = bv[r] = NC_NEG_INF;
                }
                unsigned f8 = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                {
                    const float lp = __fsub_rn(__fadd_rn(av[r], bv[r]), logz);
                    const bool live = !((lp == NC_NEG_INF) || (smin > lp && __fsub_rn(smin, lp) >= 16.25f));
                    f8 |= live ? (1u << r) : 0u;
                }
                mine += __popc(f8);
                if (c0 < 64) flo |= (unsigned long long)f8 << c0;
                else fhi |= f8;
            }
            const unsigned n_live = __reduce_add_sync(FULL, mine);
            ST_TICK(0);
            // ---- (2) the event's place in the item sequence
            unsigned base = 0;
            if (lane == 0)
            {
                while (*v_e_counted != ef) __nanosleep(100);
                __threadfence_block();
                base = *v_n_counted;
                *v_n_counted = base + n_live;
                __threadfence_block();
                *v_e_counted = ef + 1u;
            }
            base = __shfl_sync(FULL, base, 0);
            ST_TICK(1);
            // ---- (3) the terms of the live k-mers, ST_LIST listed at a time
            unsigned written = 0;
            int c = 0;
            while (written < n_live)
            {
                unsigned cnt = 0;
                for (; c < ST_CH && cnt + 32u <= (unsigned)ST_LIST; ++c)
                {
                    const bool f = (c < 64) ? ((flo >> c) & 1ull) != 0ull : ((fhi >> (c - 64)) & 1u) != 0u;
                    const unsigned m = __ballot_sync(FULL, f);
                    if (f) list[cnt + __popc(m & lt)] = (unsigned short)(32 * c + lane);
                    cnt += __popc(m);
                }
                __syncwarp();
                for (unsigned b0 = 0; b0 < cnt; b0 += 32)
                {
                    const unsigned q = b0 + (unsigned)lane;
                    if (q < cnt)
                    {
                        const unsigned j1 = __ldg(a.train_kmers + list[q]);
                        const unsigned nb = (j1 & 1023u) << 2;   // the four one-step successors are one aligned float4
                        const float al = ld_stream(Ai + j1), bi = ld_stream(Bi + j1), en = ld_stream(En + j1), bn = ld_stream(Bn + j1);
                        const float4 e4 = ld_stream4(En + nb);
                        const float4 b4 = ld_stream4(Bn + nb);
                        const float log_p_j1 = __fsub_rn(__fadd_rn(al, bi), logz);
                        float jj, t2;
#if NC_ST_EXP == 1
                        jj = log_p_j1 + (en + bn + e4.x + b4.x) * 0.0f;
                        t2 = jj;
#else
                        // joint(i, j1, j2, lt) = alpha + lt + emission(j2, e_{i+1}) + beta(i+1, j2) - logZ   (:457-469)
                        jj = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(al, log_p_stay), en), bn), logz);
                        if (jj > log_p_j1) jj = log_p_j1;
                        float s2 = flogsum(NC_NEG_INF, jj, tbl);
                        const float ev[4] = { e4.x, e4.y, e4.z, e4.w }, bv4[4] = { b4.x, b4.y, b4.z, b4.w };
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                        {
                            const float jv = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(al, log_p_step_4), ev[b]), bv4[b]), logz);
                            s2 = flogsum(s2, jv, tbl);
                        }
                        if (s2 > log_p_j1) s2 = log_p_j1;
                        const float p2 = __fsub_rn(nc_expf(log_p_j1), nc_expf(s2));
                        t2 = nc_logf(p2);
#endif
                        // the item's slot is free once every fold warp has consumed the item ST_RING before it
                        const unsigned g = base + written + q;
#if NC_ST_EXP == 9
                        if (lane == 0) ST_TICK(2);
#endif
                        while (g - min(min(v_cons[0], v_cons[1]), v_cons[2]) >= (unsigned)ST_RING)
                        {
                            // (an event with more live items than the ring holds: if it has become the oldest unpublished
                            // one while waiting here, the passes it has completed must be published for the ring to drain)
                            if (*v_e_pub == ef)
                            {
                                __threadfence_block();
                                *v_pub = base + written + b0;
                            }
                            __nanosleep(500);
                        }
#if NC_ST_EXP == 9
                        if (lane == 0) ST_TICK(3);
#endif
                        const unsigned slot = g & (ST_RING - 1);
                        ss.ring[0][slot] = log_p_j1;
                        ss.ring[1][slot] = jj;
                        ss.ring[2][slot] = t2;
                    }
                    __syncwarp();
                    // the oldest unpublished event publishes as it goes
                    if (lane == 0 && *v_e_pub == ef)
                    {
                        __threadfence_block();
                        *v_pub = base + written + min(b0 + 32u, cnt);
                    }
                }
                written += cnt;
                __syncwarp();
            }
            ST_TICK(2);
#if NC_ST_EXP == 9
            n_items_dbg += n_live; ++n_ev_dbg;
#endif
            // ---- (4) publication in event order
            if (lane == 0)
            {
                while (*v_e_pub != ef) __nanosleep(100);
                __threadfence_block();
                *v_pub = base + n_live;
                __threadfence_block();
                if (ef + 1u == n_ev) *v_done = 1u;
                *v_e_pub = ef + 1u;
            }
            __syncwarp();
            ST_TICK(4);
        }
#if NC_ST_EXP == 9
        if (lane == 0 && blockIdx.x == 0)
            printf("st %u producer %d: events %u items %u | phase1 %lld count-wait %lld terms %lld ring-wait %lld publish-wait %lld cycles\n", st, pw, n_ev_dbg,
                   n_items_dbg, tk[0], tk[1], tk[2], tk[3], tk[4]);
#endif
    }
    else
    {
        // ================================================= fold warps: denom, stay, skip
        // Each consumes its ring 32 items at a time as they are published (no CTA-wide barrier: the producers run ahead
        // by up to ST_RING items).  How the chain is cut into blocks depends on timing; its value does not (fold32 is
        // exact for any block).
        float acc = NC_NEG_INF;
        unsigned c = 0;   // items consumed
        const volatile float* Tw = ss.ring[warp];
#if NC_ST_EXP == 9
        long long tw = 0, tf = 0, t_prev = clock64(); unsigned steps = 0;
#endif
        for (;;)
        {
            unsigned avail = 0;
            if (lane == 0)
            {
                for (;;)
                {
                    const unsigned fin = *v_done;   // read before pub: if set, pub is final (both volatile: kept in order)
                    avail = *v_pub - c;
                    if (avail >= 32u || fin) break;
                    __nanosleep(200);
                }
            }
            avail = __shfl_sync(FULL, avail, 0);
#if NC_ST_EXP == 9
            { const long long n_ = clock64(); tw += n_ - t_prev; t_prev = n_; ++steps; }
#endif
            if (avail == 0u) break;          // (only when the producers are done)
            unsigned nb;
#if NC_ST_BLOCK == 64
            if (avail >= 64u)
            {
                nb = 64u;
                const float x0 = Tw[(c + (unsigned)lane) & (ST_RING - 1)], x1 = Tw[(c + 32u + (unsigned)lane) & (ST_RING - 1)];
                acc = fold64(acc, x0, x1, tbl, lane, ss.lst[warp]);
            }
            else
#endif
            {
                nb = avail < 32u ? avail : 32u;
                const float x = ((unsigned)lane < nb) ? Tw[(c + (unsigned)lane) & (ST_RING - 1)] : NC_NEG_INF;
#if NC_ST_EXP == 2
                float mx = x;
                for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, d));
                acc = fmaxf(acc, mx);
#else
                acc = fold32(acc, x, tbl, lane, ss.lst[warp]);
#endif
            }
            c += nb;
            if (lane == 0) { v_cons[warp] = c; v_snap[warp] = acc; }
#if NC_ST_EXP == 9
            { const long long n_ = clock64(); tf += n_ - t_prev; t_prev = n_; }
#endif
        }
#if NC_ST_EXP == 9
        if (lane == 0 && blockIdx.x == 0) printf("st %u fold %d: items %u steps %u | wait %lld fold %lld cycles | all CTAs so far: blocks %llu speculative %llu exact rounds %llu fallbacks %llu live items %llu\n", st, warp, c, steps, tw, tf, g_dbg[0], g_dbg[1], g_dbg[2], g_dbg[3], g_dbg[4]);
#endif
        if (lane == 0) a.st_stats[(grp * 2 + st) * 3 + warp] = acc;   // denom, stay, skip
    }
}

} // namespace nc
