// Per-thread column updates of the log-space Forward/Backward (Forward_Backward.hpp:58-125), written as
// __host__ __device__ code: the kernels in nc_fwbw.cu call them per CUDA thread, tests/emu/fwbw_emu.cu calls them
// on the host for 512 emulated threads and compares every alpha/beta bit with the oracle (no GPU needed), which
// pins the ORDER logic below independently of the device.
//
// Order.  The reference folds p7_FLogsum over the ascending merged predecessor list from_v(j) (forward) or
// successor list to_v(j) (backward), from -inf.  Predecessors of j sorted by index fall into 16 slots (index >> 8):
// slot s holds the two-step predecessor T_s = (s<<8)|(j>>4), the one-step predecessor O_b = (b<<10)|(j>>2) when
// s == 4b + (j>>10), and j itself when s == j>>8; inside a slot the order is that of the low bytes.  An index that
// occurs twice is ONE edge (std::set union, State_Transitions.hpp:205-209) carrying every matching overlap term:
// the lower-class duplicate is replaced by -inf and p7_FLogsum(x, -inf) == x exactly.
//
// Shared prefixes.  The four states 4m..4m+3 have the same T and O predecessors with the same weights; their
// chains differ only from the slot of the self predecessor (j>>8) on.  A thread owns 8 consecutive states = two
// such groups, folds the slots below its self slot ONCE per group and continues per state: 13.1 folds per state
// instead of 21 (the minimum over all prefix sharing is 13.1, tools/fold_trie.py).  The self slot is j>>8 =
// (logical thread)>>5, uniform in a warp, so the three loops (prefix, self slot, suffix) have warp-uniform trip
// counts.  Backward: the four states t + 512f + 1024q (q = 0..3) share both successor blocks [16(j&255), +16) and
// [4(j&1023), +4); their merged 20-entry chain is folded once and every state whose index lies above both blocks
// (or between them) continues from a snapshot of it: 17.6 folds per state instead of 21.
#ifndef NC_FWBW_CORE_CUH
#define NC_FWBW_CORE_CUH

#include <math.h>

#if defined(__CUDACC__)
#define NC_HD __host__ __device__ __forceinline__
#else
#define NC_HD inline
// plain C++ build (the host emulation under tests/emu): the two vector types the column code loads
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
#endif

namespace nc {
namespace fb {

constexpr int THREADS = 512;
constexpr int SPT = 8;
constexpr int N_STATES = 4096;
constexpr int COL_FLOATS = N_STATES + 4 * (N_STATES >> 4);   // 5120: 4 floats of padding after every 16
constexpr int TBL_N = 16000;

NC_HD float neg_inf()
{
#ifdef __CUDA_ARCH__
    return __int_as_float(0xff800000);
#else
    return -INFINITY;
#endif
}
NC_HD float fadd(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
NC_HD float fsub(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
NC_HD float fmul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

// physical slot of column entry n: conflict-free 16-float block reads and 32-lane consecutive reads
NC_HD int cphys(int n) { return n + ((n >> 4) << 2); }

// p7_FLogsum (logsum.hpp:141-154): max + tbl[(int)((max - min) * 1000.f)], or max when min == -inf or
// max - min >= 15.999f.  Same bits in fewer instructions:
//   * max - min == |a - b| exactly (rounding is symmetric): no min is formed;
//   * (int)(d * 1000.f) == 15999 exactly when d >= 15.999f (checked over every float in [15.9, 16.1)), so with the
//     index clamped through fminf(d, 15.999f) and a table whose entry 15999 is 0 (the kernels' copy; the reference
//     never reads that entry) the "return max" cases are max + 0.f: no compare, no select;
//   * min == -inf gives d == +inf (clamped); both -inf gives d == NaN, fminf(NaN, 15.999f) == 15.999f, -inf + 0.f.
// NaN inputs (which the reference only meets on invalid events) are not reproduced.
// The table is a functor d -> tbl[(int)(min(d, 15.999f) * 1000.f)]:
// (addr(d) = which entry, load(addr) = its value: the trainer's speculative fold compares entries before loading)
struct TblPtr   // any address space, index through a float -> int conversion
{
    const float* p;
    NC_HD unsigned addr(float d) const
    {
#ifdef __CUDA_ARCH__
        return (unsigned)__float2int_rz(fmul(fminf(d, 15.999f), 1000.0f));
#else
        const float dc = (d != d) ? 15.999f : fminf(d, 15.999f);
        return (unsigned)(int)fmul(dc, 1000.0f);
#endif
    }
    NC_HD float load(unsigned a) const
    {
#ifdef __CUDA_ARCH__
        // read-only for the lifetime of the kernel; kept in L1 against the streaming loads around it
        float v;
        asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p + a));
        return v;
#else
        return p[a];
#endif
    }
    NC_HD float operator()(float d) const { return load(addr(d)); }
};
#ifdef __CUDACC__
// Table in shared memory, truncation without the conversion pipe (F2I runs on the quarter-rate XU pipe): for
// 0 <= x < 2^21, RZ(x + 2^21 + B/4) has floor(4x) + B in its mantissa (ulp 1/4, B the table's shared-window address,
// a multiple of 4), and (floor(4x) + B) & ~3 == B + 4 floor(x) is the address of the entry: FMNMX, FMUL, FADD.RZ,
// LOP3, LDS.
struct TblSmem
{
    float magic;   // 2^21 + B/4
    __device__ __forceinline__ unsigned addr(float d) const
    {
        const float x = __fmul_rn(fminf(d, 15.999f), 1000.0f);
        return (unsigned)__float_as_int(__fadd_rz(x, magic)) & 0x007ffffcu;
    }
    __device__ __forceinline__ float load(unsigned a) const
    {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        return v;
    }
    __device__ __forceinline__ float operator()(float d) const { return load(addr(d)); }
};
__device__ __forceinline__ TblSmem make_tbl_smem(const float* tbl_in_smem)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(tbl_in_smem);
    if (b >= (1u << 20) || (b & 3u)) __trap();   // (cluster launches put the CTA rank above bit 24: not used here)
    TblSmem t;
    t.magic = 2097152.0f + 0.25f * (float)b;     // exact: b/4 is an integer below 2^18
    return t;
}
#endif
template < typename TB >
NC_HD float flogsum(float a, float b, const TB& tbl)
{
    const float d = fabsf(fsub(a, b));
    const float mx = fmaxf(a, b);
    return fadd(mx, tbl(d));
}

// N independent p7_FLogsum folds written stage by stage (all differences, all maxima, all table addresses, all loads,
// all sums): the chains of a column are long dependent sequences, and issued one after the other each fold costs its
// full latency (~50 cycles: four ALU steps, a shared-memory load, an add); interleaved, N folds cost little more than
// one.  Same operations on the same operands as N calls of flogsum.
template < int N, typename TB >
NC_HD void flogsum_n(float (&acc)[N], const float (&x)[N], const TB& tbl)
{
    float d[N], mx[N];
    unsigned ad[N];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < N; ++k) d[k] = fabsf(fsub(acc[k], x[k]));
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < N; ++k) mx[k] = fmaxf(acc[k], x[k]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < N; ++k) ad[k] = tbl.addr(d[k]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < N; ++k) d[k] = tbl.load(ad[k]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < N; ++k) acc[k] = fadd(mx[k], d[k]);
}
template < int N, typename TB >
NC_HD void flogsum_n(float (&acc)[N], const float x, const TB& tbl)
{
    float xs[N];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < N; ++k) xs[k] = x;
    flogsum_n< N >(acc, xs, tbl);
}

// 6-bit overlap mask of an edge i -> j (State_Transitions::get_trans_prob, State_Transitions.hpp:128-141)
NC_HD unsigned tmask(unsigned i, unsigned j)
{
    unsigned m = (i == j) ? 1u : 0u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (unsigned l = 1; l < 6; ++l)
        m |= ((i & ((1u << (2 * (6 - l))) - 1u)) == (j >> (2 * l))) ? (1u << l) : 0u;
    return m;
}

// ---------------------------------------------------------------------------------------------- forward
// Warp w of the CTA runs logical warp FWD_WARP_PERM[w]: the work of a logical warp falls with its self slot
// (168 folds per thread and column for slot 0, 56 for slot 15), and warps w, w+4, w+8, w+12 share a scheduler:
// the permutation gives every scheduler the same total.
NC_HD int fwd_logical_thread(int t)
{
    const int w = t >> 5;
    const int q = w & 3, r = w >> 2;          // r = 0..3
    const int lw = (r & 1) ? (4 * r + 3 - q) : (4 * r + q);   // 0 1 2 3 | 7 6 5 4 | 8 9 10 11 | 15 14 13 12
    return (lw << 5) | (t & 31);
}

struct FwdConst
{
    int u;        // logical thread: owns states 8u .. 8u+7
    int sS;       // self slot = (8u+k) >> 8 = u >> 5                  (warp-uniform)
    int c;        // one-step predecessors sit in slots s with (s & 3) == c; c = u >> 7  (warp-uniform)
    int offT;     // cphys(u >> 1): T_s is column entry offT + 320 s
    int offO;     // cphys(2u): the pair (O_b of half 0, of half 1) is the float2 at offO + 1280 b
    float wT;     // weight of the two-step edges (mask bits 2..5)
    float wO[2];  // weight of the one-step edges of half h (mask bits 1..5; carries bit 2 when O_b == T_s)
    float wS[8];  // weight of the self edge (every bit that matches)
    unsigned flags;
    // flags: bit h      oBefT[h]: O_b precedes T_s in their slot     (low byte (2u+h)&255 < u>>1)
    //        bit 2+h    oEqT[h]:  O_b == T_s (one merged edge, carried by O)
    //        bit 8+k    sFirst[k]: self precedes T in the self slot   (low byte (8u+k)&255 < u>>1)
    //        bit 16+k   sEqT[k]:  self == T_sS (merged edge, carried by self)
    //        bit 24+k   sEqO[k]:  self == O (merged edge, carried by self)       -- only in warps whose self slot is an O slot
    unsigned pos;  // 2 bits per state: entries of the self slot that precede self (0..2)  -- same warps
};

NC_HD void fwd_const_init(FwdConst& C, int u, const float* __restrict__ lut)
{
    C.u = u;
    C.sS = u >> 5;
    C.c = u >> 7;
    C.offT = cphys(u >> 1);
    C.offO = cphys(2 * u);
    const unsigned j0 = 8u * (unsigned)u;
    const unsigned kT = (unsigned)u >> 1;
    C.wT = lut[tmask(kT, j0) & 0x3cu];
    unsigned fl = 0, pos = 0;
    for (int h = 0; h < 2; ++h)
    {
        const unsigned m = 2u * (unsigned)u + (unsigned)h;          // j >> 2 of the half
        C.wO[h] = lut[tmask(m, j0 + 4u * (unsigned)h) & 0x3eu];
        const unsigned kO = m & 255u;
        if (kO < kT) fl |= 1u << h;
        if (kO == kT) fl |= 1u << (2 + h);
    }
    const bool selfO = (C.sS & 3) == C.c;
    for (int k = 0; k < 8; ++k)
    {
        const unsigned j = j0 + (unsigned)k;
        C.wS[k] = lut[tmask(j, j)];
        const unsigned kS = j & 255u, kO = (2u * (unsigned)u + (unsigned)(k >> 2)) & 255u;
        if (kS < kT) fl |= 1u << (8 + k);
        if (kS == kT) fl |= 1u << (16 + k);
        if (selfO && kS == kO) fl |= 1u << (24 + k);
        pos |= ((kT < kS ? 1u : 0u) + ((selfO && kO < kS) ? 1u : 0u)) << (2 * k);
    }
    C.flags = fl;
    C.pos = pos;
}

// One column: A = alpha[i-1] (padded layout), own[] = this thread's alpha[i-1][8u..8u+7] on entry and
// alpha[i][8u..8u+7] on return, e[] = emissions of column i for the same states.
template < typename TB >
NC_HD void fwd_column(const FwdConst& C, const float* __restrict__ A, const TB& tbl,
                      const float (&e)[8], float (&own)[8])
{
    const float NI = neg_inf();
    const unsigned F = C.flags;
    const bool oBef0 = F & 1u, oBef1 = F & 2u, oEq0 = F & 4u, oEq1 = F & 8u;
    const float* pT = A + C.offT;   // T_s, advanced by 320 floats per slot
    const float* pO = A + C.offO;   // the O pair of the next one-step slot, advanced by 1280 floats per such slot
    int so = C.c;                   // the next one-step slot
    float P[2] = { NI, NI };
    // ---- slots below the self slot: one chain per half.  The loops run from one-step slot to one-step slot (so is the
    // next one): plain counted loops over the two-step-only slots in between, without a test per slot
    int s = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    while (so < C.sS)
    {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (; s < so; ++s, pT += 320) flogsum_n< 2 >(P, fadd(C.wT, *pT), tbl);
        {
            const float xT = fadd(C.wT, *pT);
            const float2 o = *reinterpret_cast< const float2* >(pO);
            const float xO0 = fadd(C.wO[0], o.x), xO1 = fadd(C.wO[1], o.y);
            const float a0 = oBef0 ? xO0 : (oEq0 ? NI : xT), b0 = oBef0 ? xT : xO0;
            const float a1 = oBef1 ? xO1 : (oEq1 ? NI : xT), b1 = oBef1 ? xT : xO1;
            const float xa[2] = { a0, a1 }, xb[2] = { b0, b1 };
            flogsum_n< 2 >(P, xa, tbl);
            flogsum_n< 2 >(P, xb, tbl);
        }
        ++s; pT += 320; so += 4; pO += 1280;
    }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; s < C.sS; ++s, pT += 320) flogsum_n< 2 >(P, fadd(C.wT, *pT), tbl);
    // ---- the self slot: per state
    float acc[8];
    {
        const float xT = fadd(C.wT, *pT);
        pT += 320;
        if (C.sS == so)
        {
            // T, O and self share the slot: five folds, the three that are not self's position fold -inf
            const float2 o = *reinterpret_cast< const float2* >(pO);
            so += 4;
            pO += 1280;
            const float xO0 = fadd(C.wO[0], o.x), xO1 = fadd(C.wO[1], o.y);
            float f0[8], f1[8], f2[8], f3[8], f4[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < 8; ++k)
            {
                const bool h = k >> 2;
                const bool oBef = h ? oBef1 : oBef0, oEq = h ? oEq1 : oEq0;
                const bool sEqT = (F >> (16 + k)) & 1u, sEqO = (F >> (24 + k)) & 1u;
                const float xO = sEqO ? NI : (h ? xO1 : xO0);
                const float xTd = (oEq || sEqT) ? NI : xT;
                const unsigned p = (C.pos >> (2 * k)) & 3u;
                const float vS = fadd(C.wS[k], own[k]);
                f0[k] = p == 0 ? vS : NI;
                f1[k] = oBef ? xO : xTd;
                f2[k] = p == 1 ? vS : NI;
                f3[k] = oBef ? xTd : xO;
                f4[k] = p == 2 ? vS : NI;
                acc[k] = P[h];
            }
            flogsum_n< 8 >(acc, f0, tbl);
            flogsum_n< 8 >(acc, f1, tbl);
            flogsum_n< 8 >(acc, f2, tbl);
            flogsum_n< 8 >(acc, f3, tbl);
            flogsum_n< 8 >(acc, f4, tbl);
        }
        else
        {
            float e1[8], e2[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < 8; ++k)
            {
                const bool sFirst = (F >> (8 + k)) & 1u, sEq = (F >> (16 + k)) & 1u;
                const float vS = fadd(C.wS[k], own[k]);
                e1[k] = (sFirst || sEq) ? vS : xT;
                e2[k] = sEq ? NI : (sFirst ? xT : vS);
                acc[k] = P[k >> 2];
            }
            flogsum_n< 8 >(acc, e1, tbl);
            flogsum_n< 8 >(acc, e2, tbl);
        }
    }
    // ---- slots above the self slot: per state, again from one-step slot to one-step slot
    s = C.sS + 1;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    while (so < 16)
    {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (; s < so; ++s, pT += 320) flogsum_n< 8 >(acc, fadd(C.wT, *pT), tbl);
        {
            const float xT = fadd(C.wT, *pT);
            const float2 o = *reinterpret_cast< const float2* >(pO);
            const float xO0 = fadd(C.wO[0], o.x), xO1 = fadd(C.wO[1], o.y);
            const float a0 = oBef0 ? xO0 : (oEq0 ? NI : xT), b0 = oBef0 ? xT : xO0;
            const float a1 = oBef1 ? xO1 : (oEq1 ? NI : xT), b1 = oBef1 ? xT : xO1;
            const float xa[8] = { a0, a0, a0, a0, a1, a1, a1, a1 }, xb[8] = { b0, b0, b0, b0, b1, b1, b1, b1 };
            flogsum_n< 8 >(acc, xa, tbl);
            flogsum_n< 8 >(acc, xb, tbl);
        }
        ++s; pT += 320; so += 4; pO += 1280;
    }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; s < 16; ++s, pT += 320) flogsum_n< 8 >(acc, fadd(C.wT, *pT), tbl);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k) own[k] = fadd(e[k], acc[k]);
}

// ---------------------------------------------------------------------------------------------- backward
// Backward ownership.  A family is the four states j = j0 + 1024 q, q = 0..3, with j0 = lo + 256 rho (lo = j & 255,
// rho = 0..3): they share the two-step successor block [tb, tb+16), tb = 16 lo, and the one-step block [ob, ob+4),
// ob = 4 j0.  The work of a family (folds per column) depends on where its states lie relative to the two blocks:
// 31 (lo < 32, rho = 2) to 111 (lo >= 224, rho = 3) for a 32-lane range of lo.  Every thread runs two families, and
// every warp the 32 lanes of two (lo range, rho) items of complementary cost (BWD_ITEMS: 141..153 folds per warp and
// column; with the two families of the same lo, as the states t + 512 k would give, it is 80..204, and the column
// barrier waits for the slowest warp).  The price is that the two families of a thread no longer share their
// two-step block: four more 16-byte loads from each of the two columns and 32 more additions per thread and column.
NC_HD void bwd_item(int t, int f, int& lo, int& rho)
{
    // warp -> { lo range, rho } of family 0 and of family 1 (tools/fwbw_balance.py)
    const unsigned char BWD_ITEMS[16][4] = { {0, 2, 7, 3}, {0, 3, 5, 2}, {1, 0, 2, 1}, {1, 1, 0, 0}, {0, 1, 7, 2}, {2, 2, 6, 1},
                                             {1, 2, 6, 0}, {2, 3, 7, 1}, {1, 3, 7, 0}, {3, 0, 5, 3}, {3, 1, 4, 1}, {4, 2, 6, 3},
                                             {2, 0, 6, 2}, {3, 2, 4, 0}, {4, 3, 5, 1}, {3, 3, 5, 0} };
    const int w = t >> 5;
    lo = 32 * (int)BWD_ITEMS[w][2 * f] + (t & 31);
    rho = (int)BWD_ITEMS[w][2 * f + 1];
}
struct BwdConst
{
    int j0[2];       // first state of family f (the others: + 1024 q)
    int tb[2];       // first two-step successor of family f
    int ob[2];       // first one-step successor of family f
    float wTb[2];    // weight of the two-step edges
    float wOb[2];    // weight of the one-step edges
    unsigned smask[2];  // 6-bit self masks of the family's states: state q at bits 6 q
    unsigned flags;  // bit f: oIn[f] (one-step block inside the two-step block), bit 2+f: oBef[f] (one-step block first),
                     // bits 4+2f..5+2f: c4[f] = position (0..3) of the one-step block inside the two-step block
    unsigned paths;  // 3 bits per state k = 2 q + f (warp-uniform): 0 = generic chain, 1 = self after both blocks,
                     // 2 = self between the blocks, one-step block first, 3 = self between, two-step block first,
                     // 4 = self before both blocks
};

// per-lane classification of state k = 2 q + f of thread t: the kernel (and the emulation) turn it into `paths` with a vote
NC_HD unsigned bwd_lane_code(int t, int k)
{
    const int f = k & 1, q = k >> 1;
    int lo, rho;
    bwd_item(t, f, lo, rho);
    const int j0 = lo + 256 * rho, j = j0 + 1024 * q;
    const int tb = lo << 4, ob = j0 << 2;
    const bool oIn = (ob >> 4) == (tb >> 4);
    const bool oBef = !oIn && ob < tb;
    const bool sInT = (j >> 4) == (tb >> 4), sInO = (j >> 2) == (ob >> 2);
    if (oIn || sInT || sInO) return 0;
    const int pos = (j > tb ? 1 : 0) + (j > ob ? 1 : 0);
    if (pos == 2) return 1;
    if (pos == 1) return oBef ? 2 : 3;
    return 4;
}

NC_HD void bwd_const_init(BwdConst& C, int t, const float* __restrict__ lut)
{
    unsigned fl = 0;
    for (int f = 0; f < 2; ++f)
    {
        int lo, rho;
        bwd_item(t, f, lo, rho);
        C.j0[f] = lo + 256 * rho;
        C.tb[f] = lo << 4;
        C.ob[f] = C.j0[f] << 2;
        C.wTb[f] = lut[tmask((unsigned)C.j0[f], (unsigned)C.tb[f]) & 0x3cu];
        C.wOb[f] = lut[tmask((unsigned)C.j0[f], (unsigned)C.ob[f]) & 0x3eu];
        const bool oIn = (C.ob[f] >> 4) == (C.tb[f] >> 4);
        const bool oBef = !oIn && C.ob[f] < C.tb[f];
        if (oIn) fl |= 1u << f;
        if (oBef) fl |= 1u << (2 + f);
        fl |= (unsigned)((C.ob[f] & 15) >> 2) << (4 + 2 * f);
        C.smask[f] = 0;
        for (int q = 0; q < 4; ++q)
        {
            const unsigned j = (unsigned)(C.j0[f] + 1024 * q);
            C.smask[f] |= tmask(j, j) << (6 * q);
        }
    }
    C.flags = fl;
    C.paths = 0;   // filled by the caller from bwd_lane_code with a warp vote
}

// One column: Bn = beta[i+1] (padded layout), En = emissions of column i+1 (plain layout, global memory),
// lut = the job's 64 transition weights; store(j, beta[i][j]) for the thread's eight states.
template < typename TB, typename Store >
NC_HD void bwd_column(const BwdConst& C, const float* __restrict__ Bn, const float* __restrict__ En,
                      const float* __restrict__ lut, const TB& tbl, Store store)
{
    const float NI = neg_inf();
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int f = 0; f < 2; ++f)
    {
        float vT[16];
        {
            const float4* e4 = reinterpret_cast< const float4* >(En + C.tb[f]);
            const float4* b4 = reinterpret_cast< const float4* >(Bn + cphys(C.tb[f]));
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int v = 0; v < 4; ++v)
            {
#ifdef __CUDA_ARCH__
                const float4 e = __ldg(e4 + v);
#else
                const float4 e = e4[v];
#endif
                const float4 b = b4[v];
                vT[4 * v + 0] = fadd(fadd(C.wTb[f], e.x), b.x);
                vT[4 * v + 1] = fadd(fadd(C.wTb[f], e.y), b.y);
                vT[4 * v + 2] = fadd(fadd(C.wTb[f], e.z), b.z);
                vT[4 * v + 3] = fadd(fadd(C.wTb[f], e.w), b.w);
            }
        }
        float vO[4];
        {
#ifdef __CUDA_ARCH__
            const float4 e = __ldg(reinterpret_cast< const float4* >(En + C.ob[f]));
#else
            const float4 e = *reinterpret_cast< const float4* >(En + C.ob[f]);
#endif
            const float4 b = *reinterpret_cast< const float4* >(Bn + cphys(C.ob[f]));
            vO[0] = fadd(fadd(C.wOb[f], e.x), b.x);
            vO[1] = fadd(fadd(C.wOb[f], e.y), b.y);
            vO[2] = fadd(fadd(C.wOb[f], e.z), b.z);
            vO[3] = fadd(fadd(C.wOb[f], e.w), b.w);
        }
        const bool oIn = (C.flags >> f) & 1u, oBef = (C.flags >> (2 + f)) & 1u;
        const int c4 = (int)((C.flags >> (4 + 2 * f)) & 3u);
        // merged, ordered list of the 20 block successors of this family:
        // position q holds  oBef ? (q < 4 ? O[q] : T[q-4]) : (q < 16 ? T[q] : O[q-16]);
        // oIn: the list is T with entries 4 c4 .. 4 c4 + 3 carrying the one-step weight, then four -inf
        float L[20];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int q = 0; q < 20; ++q)
        {
            float asT, asO;
            if (q < 4) { asO = vO[q]; asT = vT[q]; }
            else if (q < 16) { asO = vT[q - 4]; asT = vT[q]; }
            else { asO = vT[q - 4]; asT = vO[q - 16]; }
            float val = oBef ? asO : asT;
            if (q < 16) val = (oIn && ((q >> 2) == c4)) ? vO[q & 3] : val;
            else val = oIn ? NI : val;
            L[q] = val;
        }
        // the chain without self, folded once: snapshots after the first block and after both
        float snap4, snap16, m20;
        {
            float m = L[0];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int q = 1; q < 4; ++q) m = flogsum(m, L[q], tbl);
            snap4 = m;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int q = 4; q < 16; ++q) m = flogsum(m, L[q], tbl);
            snap16 = m;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int q = 16; q < 20; ++q) m = flogsum(m, L[q], tbl);
            m20 = m;
        }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (int kk = 0; kk < 4; ++kk)
        {
            const int k = 2 * kk + f;
            const int j = C.j0[f] + 1024 * kk;
            const float wS = lut[(C.smask[f] >> (6 * kk)) & 63u];
#ifdef __CUDA_ARCH__
            const float eS = __ldg(En + j);
#else
            const float eS = En[j];
#endif
            const float vS = fadd(fadd(wS, eS), Bn[cphys(j)]);
            const unsigned path = (C.paths >> (3 * k)) & 7u;
            float acc;
            if (path == 1) acc = flogsum(m20, vS, tbl);
            else if (path == 2)
            {
                acc = flogsum(snap4, vS, tbl);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int q = 4; q < 20; ++q) acc = flogsum(acc, L[q], tbl);
            }
            else if (path == 3)
            {
                acc = flogsum(snap16, vS, tbl);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int q = 16; q < 20; ++q) acc = flogsum(acc, L[q], tbl);
            }
            else if (path == 4)
            {
                acc = vS;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int q = 0; q < 20; ++q) acc = flogsum(acc, L[q], tbl);
            }
            else
            {
                // generic: self at its place in the list.  mpos = its position when it coincides with a block entry
                // (it then replaces that entry: one merged edge), else pos = number of blocks entirely below it
                const bool sInT = (j >> 4) == (C.tb[f] >> 4), sInO = (j >> 2) == (C.ob[f] >> 2);
                int mpos = -1;
                if (sInT) mpos = (oBef ? 4 : 0) + (j & 15);
                else if (sInO) mpos = (oBef ? 0 : 16) + (j & 3);
                const int pos = (mpos >= 0) ? -1 : ((j > C.tb[f] ? 1 : 0) + ((!oIn && j > C.ob[f]) ? 1 : 0));
                acc = (pos == 0) ? vS : NI;
                const float ins4 = (oBef && pos == 1) ? vS : NI, ins16 = (!oBef && pos == 1) ? vS : NI;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int q = 0; q < 20; ++q)
                {
                    if (q == 4) acc = flogsum(acc, ins4, tbl);
                    if (q == 16) acc = flogsum(acc, ins16, tbl);
                    acc = flogsum(acc, (q == mpos) ? vS : L[q], tbl);
                }
                acc = flogsum(acc, pos == 2 ? vS : NI, tbl);
            }
            store(j, acc);
        }
    }
}

} // namespace fb
} // namespace nc

#endif
