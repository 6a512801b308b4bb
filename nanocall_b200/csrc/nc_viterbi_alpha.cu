// K1a: Viterbi in "alpha-column" form -- the fast path of nc_viterbi_packed.
//
// Same contract as viterbi_kernel (nc_viterbi.cu; replaces Viterbi<float,6>::fill, Viterbi.hpp:44-150), different
// division of labour between the forward pass and the traceback, chosen from what binds on B200:
//
//   * viterbi_kernel finds arg max AND max in the forward pass and streams 1 B/state/event of backpointers.  Its
//     forward loop is bound by issue slots: ~140 of its ~390 instructions per thread-event are the compare / select /
//     index bookkeeping of the arg max (FSETP/FSEL/SEL run at half rate on sm_100a, tools/ubench), while HBM idles
//     at 7 % of its bandwidth.
//   * this kernel computes ONLY the max in the forward pass and streams the alpha column itself (4 B/state/event).
//     The arg max is evaluated during the traceback, for the one state per column that lies on the path, from the
//     stored columns: 21 loads per event instead of 4096 x 21 compares.  The forward loop drops to the emission
//     (FMA pipe) plus ~25 max instructions per thread-event; HBM write traffic rises to 16 KiB/event, which is the
//     idle resource (measured in profiles/).
//
// Exactness.  max_k RN(w + a_k) == RN(w + max_k a_k) because rounding is monotone, so the max over the 16 two-step
// (4 one-step) predecessors is taken on the raw alphas and the class weight added once.  The value of alpha[i][j]
// is therefore bit-identical to the reference's (Viterbi.hpp:78-90), and the traceback re-derives each backpointer
// from those bits with the reference's rule: first maximum in ascending predecessor order (strict '>').
//
// Mapping: as viterbi_kernel -- one persistent CTA of 512 threads per SM, thread t owns states 8t..8t+7, previous
// column double-buffered in shared memory, one mbarrier phase per column.
//
// What binds (tools/ubench/emis.cu, profiles/): the fused emission alone keeps the FMA pipe busy ~730 cycles per
// column and SM sub-partition, everything else must hide behind it.  So (1) the loop body carries no integer
// arithmetic on the FMA pipe (IMAD runs there at half rate): every shared-memory access is `base register +
// immediate`, the loop is unrolled by two so buffer and barrier parity are compile-time constants; (2) the emission
// is split: state pairs 2,3 of event i are computed next to the predecessor loads / max tree / shuffle of column i
// (filling that chain's latency bubbles), pairs 0,1 of event i+1 between the barrier arrive and the barrier wait
// (filling the barrier bubble); (3) the emission runs as packed FADD2/FMUL2/FFMA2 to halve its issue slots.
#include "nc_vit_common.cuh"

#include <type_traits>

namespace nc {

using namespace vit;

namespace {


__device__ __forceinline__ void st_cs_v8(float* p, const float (&v)[8])
{
    // one 256-bit streaming store: every lane writes a full 32-byte sector, a warp 1 KiB contiguous
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}


// ---- packed binary32 pairs (sm_100a FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per issue slot.
// The FMA pipe still spends two cycles on them; what they save is issue bandwidth, which the emission (19 FP
// operations per state and event) would otherwise monopolise.  Element-wise they are the scalar operations, so the
// emission keeps the reference's bits (tests/test_viterbi_gpu.py runs every case through this kernel).
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_of(f2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi_of(f2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Emission constants of two adjacent states, subtrahends stored negated so every step is an add / mul / fma
// (RN(a - b) == RN(a + (-b)) and RN(-(x)) == -RN(x): same bits as emission_h in nc_device.cuh).
// Event operands: X = x, Y = y, Y2 = 2y, NLY = -(3 log y)/2, NRY = -RN(1/y)/2, each broadcast to both halves
struct EvPairs { f2 X, Y, Y2, NLY, NRY; };
// log_pr_corrected_emission for two states (Pore_Model.hpp:24-40,145-149), step for step emission_h:
//   ah = RN((x-mu)/(2 sg)) by Markstein division;  ln = nls - (2 ah^2 + log_2pi/2)
//   b  = RN((y-eta)/eta);  uh = RN(lam b b / (2y));  li = (c1/2 - (3 log y)/2) - uh;   e = ln + li
// nls and c1h are passed separately: the kernel keeps them in shared memory to make room in the register file.
struct PairRegs { f2 nmu, nsg2, rsgh, neta, reta, lam; };
__device__ __forceinline__ f2 emission2(const PairRegs& p, f2 nls, f2 c1h, const EvPairs& e, f2 m2, f2 nhl2pi)
{
    const f2 t1 = add2(e.X, p.nmu);
    const f2 q0 = mul2(t1, p.rsgh);
    const f2 r = fma2(q0, p.nsg2, t1);
    const f2 ah = fma2(r, p.rsgh, q0);
    const f2 ns = fma2(mul2(ah, ah), m2, nhl2pi);     // -(2 ah^2 + log_2pi/2)
    const f2 ln = add2(nls, ns);
    const f2 t2 = add2(e.Y, p.neta);
    const f2 p0 = mul2(t2, p.reta);
    const f2 r2 = fma2(p0, p.neta, t2);
    const f2 b = fma2(r2, p.reta, p0);
    const f2 l2 = mul2(mul2(p.lam, b), b);
    const f2 nq = mul2(l2, e.NRY);                    // -RN(l2 * ry/2)
    const f2 r3 = fma2(nq, e.Y2, l2);
    const f2 nuh = fma2(r3, e.NRY, nq);               // -uh
    const f2 li = add2(add2(c1h, e.NLY), nuh);
    return add2(ln, li);
}
// event as staged for this kernel: {x, y, -(3 log y)/2, -RN(1/y)/2}
__device__ __forceinline__ float4 ev_slot(const float4& p) { return make_float4(p.x, p.y, -p.z, -p.w); }
__device__ __forceinline__ EvPairs ev_pairs(const float4& s)
{
    EvPairs e;
    const float y2 = __fadd_rn(s.y, s.y);
    e.X = pk(s.x, s.x); e.Y = pk(s.y, s.y); e.Y2 = pk(y2, y2); e.NLY = pk(s.z, s.z); e.NRY = pk(s.w, s.w);
    return e;
}
// shared-memory access by 32-bit shared address: one base register + immediate offset in SASS, no generic-address
// arithmetic.  volatile: ordered against the barrier asm statements.
__device__ __forceinline__ float lds32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds64(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ f2 lds64p(unsigned a) { f2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds128(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_if(unsigned bar, unsigned pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\n@p mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n}" ::"r"(bar), "r"(pred) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NC_WAITA:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NC_DONEA;\n"
        "bra NC_WAITA;\n"
        "NC_DONEA:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

struct __align__(16) SmemA
{
    float alpha[2][NC_N_STATES + ALPHA_PAD];
    float4 ev[2 * CH];             // two chunks of staged events (slot = event index & 255)
    f2 prm[2][SPT / 2][THREADS];   // nls / c1h of every state pair, slot [.][pair][thread]: conflict-free LDS.64
    float red_v[THREADS / 32];
    int red_j[THREADS / 32];
    unsigned job;
    int final_state;
    unsigned long long col_bar;    // mbarrier: one phase per event column
};

// One traceback step: the predecessor of state s, given the alpha column of the previous event.
// Candidates as in the forward pass (two-step class weight w2(g), one-step class weight w1(h), exact self weight);
// a predecessor that belongs to two classes appears twice with the same index and a weight <= its exact one, which
// changes neither the maximum nor the lowest index attaining it (see nc_viterbi.cu).  Ties -> lowest index, which
// is what the reference's strict '>' over the ascending from_v yields (Viterbi.hpp:78-89).
__device__ __forceinline__ unsigned tb_step(const float* __restrict__ Ap, unsigned s, const float* lut)
{
    // column layout in the slab is group-major, G[g*16 + bb] = alpha[(bb<<8)|g]: the 16 two-step predecessors of s
    // are one 64-byte line, its 4 one-step predecessors share another, the self predecessor sits in a third
    const unsigned g = s >> 4, h = s >> 2;
    float v2[16], v1[4];
    const float4* q = reinterpret_cast< const float4* >(Ap + (g << 4));
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const float4 x = __ldcg(q + k);
        v2[4 * k] = x.x; v2[4 * k + 1] = x.y; v2[4 * k + 2] = x.z; v2[4 * k + 3] = x.w;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) v1[b] = __ldcg(Ap + ((h & 255u) << 4) + 4 * b + (h >> 8));   // alpha[(b<<10)|h]
    const float v0 = __ldcg(Ap + ((s & 255u) << 4) + (s >> 8));                               // alpha[s]
    const float w2 = lut[trans_mask(g, s) & 0x3cu];
    const float w1 = lut[trans_mask(h, s) & 0x3eu];
    const float w0 = lut[trans_mask(s, s)];
    float best = __fadd_rn(w2, v2[0]);
    unsigned bp = g;
#pragma unroll
    for (int bb = 1; bb < 16; ++bb)
    {
        const float c = __fadd_rn(w2, v2[bb]);
        if (c > best) { best = c; bp = ((unsigned)bb << 8) | g; }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
        const float c = __fadd_rn(w1, v1[b]);
        const unsigned p = ((unsigned)b << 10) | h;
        if (c > best || (c == best && p < bp)) { best = c; bp = p; }
    }
    {
        const float c = __fadd_rn(w0, v0);
        if (c > best || (c == best && s < bp)) { best = c; bp = s; }
    }
    return bp;
}

__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void forward_cta(const VitArgs& a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemA& sm = *reinterpret_cast< SmemA* >(smem_raw);

    const int t = threadIdx.x;
    const int lane = t & 31;
    const int warp = t >> 5;
    const unsigned j0 = SPT * t;
    const unsigned g = t >> 1;
    const bool keep = a.states != nullptr;   // path probability only: nothing to trace back, nothing stored
    unsigned jobs_done = 0;
    const float log_2pi = a.log_2pi;
    const float hl2pi = __fmul_rn(0.5f, a.log_2pi);
    const unsigned bar = smem_u32(&sm.col_bar);
    if (t == 0) mbar_init(&sm.col_bar, THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    for (;;)
    {
        if (t == 0) sm.job = atomicAdd(a.next_job, 1u);
        __syncthreads();
        const unsigned q = sm.job;
        if (q >= a.n_jobs) break;
        const unsigned job_idx = a.order[q];
        const DevJob& J = a.jobs[job_idx];
        const unsigned n = J.n_events;
        const unsigned long long off = J.ev_off;
        // this job's slab: the CTA owns two and alternates; the previous user of the slab (job k-2 of this CTA) must
        // have been traced back, i.e. the slab released (k >> 1) times so far
        const unsigned slab_id = 2u * (blockIdx.x - a.n_tb) + (jobs_done & 1u);
        float* const acol = reinterpret_cast< float* >(a.bp_pool + (size_t)slab_id * a.slab_bytes);
        if (keep && jobs_done >= 2)
        {
            if (t == 0)
            {
                const long long c0 = clock64();
                unsigned ns = 64;
                while (ld_acquire_u32(a.slab_free + slab_id) < (jobs_done >> 1)) { __nanosleep(ns); if (ns < 4096) ns *= 2; }
                if (a.stats) atomicAdd(a.stats + 1, (unsigned long long)(clock64() - c0));
            }
            __syncthreads();
        }
        const long long fwd_c0 = clock64();

        // ---------------- prologue: scaled model constants (as state pairs) and transition weights
        PairRegs P[SPT / 2];
        f2 ws[SPT / 2];
        const unsigned prm_nls = smem_u32(&sm.prm[0][0][t]), prm_c1h = smem_u32(&sm.prm[1][0][t]);
        constexpr unsigned PRM_PAIR = THREADS * sizeof(f2);   // byte stride between the pairs of one thread
        {
            const float* M = a.models + (size_t)J.model * MODEL_FLOATS;
            float lm[SPT], ls[SPT], sdm[SPT], sdl[SPT], lls[SPT], lsl[SPT];
#pragma unroll
            for (int v = 0; v < SPT / 4; ++v)
            {
                *reinterpret_cast< float4* >(lm + 4 * v) = __ldg(reinterpret_cast< const float4* >(M + 0 * NC_N_STATES + j0) + v);
                *reinterpret_cast< float4* >(ls + 4 * v) = __ldg(reinterpret_cast< const float4* >(M + 1 * NC_N_STATES + j0) + v);
                *reinterpret_cast< float4* >(sdm + 4 * v) = __ldg(reinterpret_cast< const float4* >(M + 2 * NC_N_STATES + j0) + v);
                *reinterpret_cast< float4* >(sdl + 4 * v) = __ldg(reinterpret_cast< const float4* >(M + 3 * NC_N_STATES + j0) + v);
                *reinterpret_cast< float4* >(lls + 4 * v) = __ldg(reinterpret_cast< const float4* >(M + 4 * NC_N_STATES + j0) + v);
                *reinterpret_cast< float4* >(lsl + 4 * v) = __ldg(reinterpret_cast< const float4* >(M + 5 * NC_N_STATES + j0) + v);
            }
#pragma unroll
            for (int k = 0; k < SPT; k += 2)
            {
                const StateParamsH x = halve(scale_state(lm[k], ls[k], sdm[k], sdl[k], lls[k], lsl[k], J, log_2pi));
                const StateParamsH y = halve(scale_state(lm[k + 1], ls[k + 1], sdm[k + 1], sdl[k + 1], lls[k + 1], lsl[k + 1], J, log_2pi));
                PairRegs& p = P[k / 2];
                p.nmu = pk(-x.mu, -y.mu); p.nsg2 = pk(-x.sg2, -y.sg2); p.rsgh = pk(x.rsgh, y.rsgh);
                p.neta = pk(-x.eta, -y.eta); p.reta = pk(x.reta, y.reta); p.lam = pk(x.lam, y.lam);
                sm.prm[0][k / 2][t] = pk(x.nls, y.nls);      // read back by this thread only
                sm.prm[1][k / 2][t] = pk(x.c1h, y.c1h);
                ws[k / 2] = pk(J.lut[trans_mask(j0 + k, j0 + k)], J.lut[trans_mask(j0 + k + 1, j0 + k + 1)]);
            }
        }
        // two-step weight of group g: mask bits 2..5 (bit 2 always set); one-step weight of h: bits 1..5
        const float w2 = J.lut[trans_mask(g, j0) & 0x3cu];
        const float w1a = J.lut[trans_mask(2 * t, j0) & 0x3eu];
        const float w1b = J.lut[trans_mask(2 * t + 1, j0 + 4) & 0x3eu];
        const f2 M2 = pk(-2.0f, -2.0f), NH = pk(-hl2pi, -hl2pi);

        // ---------------- first chunk of events, column 0 (Viterbi.hpp:57-67)
        if (t < CH) sm.ev[t] = ev_slot(ev_pack(ev_load(a, off, t, n), J.drift));
        __syncthreads();
        f2 a_own[SPT / 2];
        {
            const EvPairs E = ev_pairs(sm.ev[0]);
            const f2 nlog_n = pk(-a.log_n_states, -a.log_n_states);
            float a0[SPT];
#pragma unroll
            for (int k = 0; k < SPT / 2; ++k)
            {
                a_own[k] = add2(emission2(P[k], sm.prm[0][k][t], sm.prm[1][k][t], E, M2, NH), nlog_n);
                a0[2 * k] = lo_of(a_own[k]);
                a0[2 * k + 1] = hi_of(a_own[k]);
            }
            float* A = sm.alpha[0];
            *reinterpret_cast< float4* >(A + phys(j0)) = make_float4(a0[0], a0[1], a0[2], a0[3]);
            *reinterpret_cast< float4* >(A + phys(j0 + 4)) = make_float4(a0[4], a0[5], a0[6], a0[7]);
        }
        // emission of event 1 for state pairs 0,1: carried into the loop (the loop computes it one event ahead)
        f2 e01[2];
        {
            const EvPairs E = ev_pairs(sm.ev[1]);
            e01[0] = emission2(P[0], sm.prm[0][0][t], sm.prm[1][0][t], E, M2, NH);
            e01[1] = emission2(P[1], sm.prm[0][1][t], sm.prm[1][1][t], E, M2, NH);
        }
        __syncthreads();

        // ---------------- columns 1..n-1 (Viterbi.hpp:72-96), max only
        constexpr unsigned COL_BYTES = (NC_N_STATES + ALPHA_PAD) * sizeof(float);
        const unsigned alpha0 = smem_u32(&sm.alpha[0][0]);
        const int half = t & 1;
        const unsigned two_b = alpha0 + 4u * (unsigned)phys(((8 * half) << 8) + (int)g);  // this thread's 8 two-step predecessors
        const unsigned one_b = alpha0 + 4u * (unsigned)phys(2 * t);                        // + (b<<10) + ((b>>1)<<4) floats, b = 0..3
        const unsigned wr_lo = alpha0 + 4u * (unsigned)phys(j0), wr_hi = alpha0 + 4u * (unsigned)phys(j0 + 4);
        const unsigned ev_b = smem_u32(&sm.ev[0]);
        EvRegs pre = { 0.f, 1.f, 0.f, 0.f };
        float* gcol = acol + j0;                                // this thread's 8 slots of column i-1 (group-major)
        const unsigned lane0 = (lane == 0) ? 1u : 0u;

        auto column = [&](auto cur_tag, const unsigned i) {
            constexpr unsigned RD = decltype(cur_tag)::value * COL_BYTES;          // column i-1
            constexpr unsigned WR = (1 - decltype(cur_tag)::value) * COL_BYTES;    // column i
            if constexpr (decltype(cur_tag)::value == 0)   // i is odd in this half: the staging points are odd
            {
                const unsigned ic = i & (CH - 1);
                if (ic == 1 && t < CH) pre = ev_load(a, off, (i - 1) + CH + t, n);
                if (ic == 17 && t < CH) sm.ev[((((i - 1) / CH) + 1) & 1) * CH + t] = ev_slot(ev_pack(pre, J.drift));
            }
            // ---- part 1: column i-1 is complete.  Predecessor loads first, then the rest of this event's emission.
            float c2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) c2[k] = lds32(two_b + RD + (k << 10));
            // these 8 values are slots 8t..8t+7 of column i-1 in the slab's group-major layout: stream them out now
            if (keep) st_cs_v8(gcol, c2);
            gcol += NC_N_STATES;
            float2 o[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) o[b] = lds64(one_b + RD + (b << 12) + ((b >> 1) << 6));
            const float4 ev_i = lds128(ev_b + ((i & (2 * CH - 1)) << 4));
            f2 e23[2];
            {
                const EvPairs E = ev_pairs(ev_i);
                e23[0] = emission2(P[2], lds64p(prm_nls + 2 * PRM_PAIR), lds64p(prm_c1h + 2 * PRM_PAIR), E, M2, NH);
                e23[1] = emission2(P[3], lds64p(prm_nls + 3 * PRM_PAIR), lds64p(prm_c1h + 3 * PRM_PAIR), E, M2, NH);
            }
            // two-step class: this thread's 8 of the group's 16 predecessors, the partner (lane ^ 1) holds the rest
            const float m2 = fmaxf(fmaxf(fmaxf(c2[0], c2[1]), fmaxf(c2[2], c2[3])), fmaxf(fmaxf(c2[4], c2[5]), fmaxf(c2[6], c2[7])));
            const float m2o = __shfl_xor_sync(0xffffffffu, m2, 1);
            // one-step class for h = 2t and 2t+1
            const float v1a = __fadd_rn(w1a, fmaxf(fmaxf(o[0].x, o[1].x), fmaxf(o[2].x, o[3].x)));
            const float v1b = __fadd_rn(w1b, fmaxf(fmaxf(o[0].y, o[1].y), fmaxf(o[2].y, o[3].y)));
            f2 vs[SPT / 2];
#pragma unroll
            for (int k = 0; k < SPT / 2; ++k) vs[k] = add2(ws[k], a_own[k]);   // self candidates
            const float v2 = __fadd_rn(w2, fmaxf(m2, m2o));
            const float va = fmaxf(v1a, v2), vb = fmaxf(v1b, v2);
            a_own[0] = add2(pk(fmaxf(lo_of(vs[0]), va), fmaxf(hi_of(vs[0]), va)), e01[0]);
            a_own[1] = add2(pk(fmaxf(lo_of(vs[1]), va), fmaxf(hi_of(vs[1]), va)), e01[1]);
            a_own[2] = add2(pk(fmaxf(lo_of(vs[2]), vb), fmaxf(hi_of(vs[2]), vb)), e23[0]);
            a_own[3] = add2(pk(fmaxf(lo_of(vs[3]), vb), fmaxf(hi_of(vs[3]), vb)), e23[1]);
            float an[SPT];
#pragma unroll
            for (int k = 0; k < SPT / 2; ++k) { an[2 * k] = lo_of(a_own[k]); an[2 * k + 1] = hi_of(a_own[k]); }
            sts128(wr_lo + WR, an[0], an[1], an[2], an[3]);
            sts128(wr_hi + WR, an[4], an[5], an[6], an[7]);
            // ---- column i is published; part 2 runs in the barrier's shadow: pairs 0,1 of event i+1
            __syncwarp();
            mbar_arrive_if(bar, lane0);
            {
                const EvPairs E = ev_pairs(lds128(ev_b + (((i + 1) & (2 * CH - 1)) << 4)));
                e01[0] = emission2(P[0], lds64p(prm_nls), lds64p(prm_c1h), E, M2, NH);
                e01[1] = emission2(P[1], lds64p(prm_nls + PRM_PAIR), lds64p(prm_c1h + PRM_PAIR), E, M2, NH);
            }
            mbar_wait_u32(bar, decltype(cur_tag)::value);
        };
        {
            unsigned i = 1;
            for (; i + 1 < n; i += 2)
            {
                column(std::integral_constant< int, 0 >{}, i);
                column(std::integral_constant< int, 1 >{}, i + 1);
            }
            if (i < n)
            {
                column(std::integral_constant< int, 0 >{}, i);
                // an empty phase keeps the number of barrier phases per job even (parity == buffer index)
                __syncwarp();
                mbar_arrive_if(bar, lane0);
                mbar_wait_u32(bar, 1);
            }
        }
        if (keep)
        {
            // the last column (buffer (n-1) & 1), in the same group-major order
            float c2[8];
            const unsigned RDL = ((n - 1) & 1u) * COL_BYTES;
#pragma unroll
            for (int k = 0; k < 8; ++k) c2[k] = lds32(two_b + RDL + (k << 10));
            st_cs_v8(gcol, c2);
        }
        float a_fin[SPT];
#pragma unroll
        for (int k = 0; k < SPT / 2; ++k) { a_fin[2 * k] = lo_of(a_own[k]); a_fin[2 * k + 1] = hi_of(a_own[k]); }

        // ---------------- fill_state_seq: argmax over the last column, strict '>' ascending j (Viterbi.hpp:123-133)
        {
            float bv = a_fin[0];
            int bj = j0;
#pragma unroll
            for (int k = 1; k < SPT; ++k)
                if (a_fin[k] > bv) { bv = a_fin[k]; bj = j0 + k; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                float ov = __shfl_down_sync(0xffffffffu, bv, d);
                int oj = __shfl_down_sync(0xffffffffu, bj, d);
                if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
            }
            if (lane == 0) { sm.red_v[warp] = bv; sm.red_j[warp] = bj; }
            __syncthreads();   // also makes this CTA's global alpha stores visible to its own traceback loads
            if (t == 0)
            {
                float fv = sm.red_v[0];
                int fj = sm.red_j[0];
                for (int w = 1; w < THREADS / 32; ++w)
                    if (sm.red_v[w] > fv) { fv = sm.red_v[w]; fj = sm.red_j[w]; }
                sm.final_state = fj;
                a.path_logprob[job_idx] = fv;
            }
            __syncthreads();
        }

        // ---------------- hand the traceback to a service warp (other CTAs of this grid) and go on with the next job.
        // The alpha columns of this job stay in the slab until the service warp releases it; the CTA alternates
        // between its two slabs, so the forward pass of job k+1 overlaps the traceback of job k.
        if (keep)
        {
            __threadfence();        // every thread: its alpha stores are visible device-wide before the ticket is
            __syncthreads();
            if (t == 0)
            {
                const unsigned slot = atomicAdd(a.tb_tail, 1u);
                TbTicket& tk = a.tickets[slot];
                tk.job = job_idx;
                tk.slab = slab_id;
                tk.final_state = (unsigned)sm.final_state;
                __threadfence();
                st_release_u32(&tk.ready, 1u);
            }
        }
        ++jobs_done;
        if (t == 0 && a.stats) atomicAdd(a.stats + 0, (unsigned long long)(clock64() - fwd_c0));
    }
}

// ---------------- traceback service (Viterbi.hpp:134-150).  One warp per job: the chain s[c-1] = pred(s[c]) is cut
// into <= 32 blocks, one per lane.  A lane starts from a GUESS of its block's end state, obtained by walking back
// TB_SPEC_DEPTH columns from an arbitrary state (survivor paths coalesce quickly); all lanes walk in parallel;
// guesses are verified against the state the successor block actually reached and wrong blocks are re-walked until
// nothing changes -- exactly the sequential traceback, in ~(n/32 + depth) dependent steps.  Every step evaluates
// the arg max from the stored alpha column (tb_step).
__device__ void traceback_service(const VitArgs& a)
{
    const int lane = threadIdx.x & 31;
    for (;;)
    {
        unsigned h = 0;
        if (lane == 0) h = atomicAdd(a.tb_head, 1u);
        h = __shfl_sync(0xffffffffu, h, 0);
        if (h >= a.n_jobs) return;
        TbTicket& tk = a.tickets[h];
        const long long w0 = clock64();
        if (lane == 0)
        {
            unsigned ns = 64;
            while (ld_acquire_u32(&tk.ready) == 0u) { __nanosleep(ns); if (ns < 2048) ns *= 2; }
        }
        __syncwarp();
        const long long w1 = clock64();
        unsigned passes = 0, steps = 0;
        const unsigned job_idx = __ldcg(&tk.job), slab_id = __ldcg(&tk.slab), final_state = __ldcg(&tk.final_state);
        const DevJob& J = a.jobs[job_idx];
        const unsigned n = J.n_events;
        const float* lut = J.lut;
        const float* acol = reinterpret_cast< const float* >(a.bp_pool + (size_t)slab_id * a.slab_bytes);
        unsigned short* out_s = a.states + J.ev_off;
        const unsigned T = n - 1;  // transitions: column c in 1..T is entered from column c-1
        if (T == 0)
        {
            if (lane == 0) out_s[0] = (unsigned short)final_state;
        }
        else
        {
            unsigned B = (T + 31) / 32;
            if (B < (unsigned)TB_MIN_BLOCK) B = TB_MIN_BLOCK;
            const unsigned nb = (T + B - 1) / B;
            const unsigned lo = (unsigned)lane * B;
            const unsigned hi = (lo + B < T) ? lo + B : T;
            const bool active = (unsigned)lane < nb;
            unsigned end_s = final_state, start_s = 0;
            if (active && hi != T)
            {
                unsigned c = hi + TB_SPEC_DEPTH;
                unsigned s = 0;
                if (c >= T) { c = T; s = final_state; }
                steps += c - hi;
                for (; c > hi; --c) s = tb_step(acol + (size_t)(c - 1) * NC_N_STATES, s, lut);
                end_s = s;
            }
            bool dirty = active;  // first pass: every lane walks its block
            for (;;)
            {
                ++passes;
                if (dirty)
                {
                    unsigned s = end_s;
                    steps += hi - lo;
                    out_s[hi] = (unsigned short)s;
                    for (unsigned c = hi; c > lo; --c)
                    {
                        s = tb_step(acol + (size_t)(c - 1) * NC_N_STATES, s, lut);
                        if (c - 1 > lo || lane == 0) out_s[c - 1] = (unsigned short)s;
                    }
                    start_s = s;
                }
                __syncwarp();
                const unsigned next_start = __shfl_down_sync(0xffffffffu, start_s, 1);
                dirty = active && (unsigned)lane + 1 < nb && end_s != next_start;
                if (!__any_sync(0xffffffffu, dirty)) break;
                if (dirty) end_s = next_start;
            }
        }
        __threadfence();   // states visible to the lanes that derive the moves; slab reads are complete
        __syncwarp();
        if (lane == 0) atomicAdd(a.slab_free + slab_id, 1u);
        if (a.stats)
        {
            for (int d = 16; d > 0; d >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, d);
            if (lane == 0)
            {
                atomicAdd(a.stats + 2, (unsigned long long)(clock64() - w1));
                atomicAdd(a.stats + 3, (unsigned long long)(w1 - w0));
                atomicAdd(a.stats + 4, (unsigned long long)passes);
                atomicAdd(a.stats + 5, (unsigned long long)steps);
                atomicAdd(a.stats + 6, 1ull);
            }
        }
        // ---------------- fill_move_seq (Viterbi.hpp:144-150)
        if (a.moves != nullptr)
        {
            unsigned char* out_m = a.moves + J.ev_off;
            for (unsigned i = lane; i < n; i += 32)
                out_m[i] = (i == 0) ? 0 : (unsigned char)min_skip(__ldcg(out_s + i - 1), __ldcg(out_s + i));
        }
    }
}

} // namespace

__global__ void __launch_bounds__(VIT_THREADS, 1) viterbi_alpha_kernel(const VitArgs a)
{
    // service CTAs take the lowest block indices so they are resident before any forward CTA can wait on them
    if (blockIdx.x < a.n_tb) traceback_service(a);
    else forward_cta(a);
}

size_t viterbi_alpha_smem_bytes() { return sizeof(SmemA); }

} // namespace nc
