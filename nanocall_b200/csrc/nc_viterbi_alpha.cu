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
// column double-buffered in shared memory, emission of event i+1 computed behind the split-phase column barrier.
#include "nc_vit_common.cuh"

namespace nc {

using namespace vit;

namespace {

struct __align__(16) SmemA
{
    float alpha[2][NC_N_STATES + ALPHA_PAD];
    float4 ev[2][CH];
    float lut[64];                 // transition log-weights of the current job (traceback)
    float red_v[THREADS / 32];
    int red_j[THREADS / 32];
    unsigned short tb_end[THREADS];
    unsigned short tb_start[THREADS];
    unsigned job;
    int final_state;
    unsigned long long col_bar;    // mbarrier: one phase per event column
};

__device__ __forceinline__ void st_cs_v8(float* p, const float (&v)[8])
{
    // one 256-bit streaming store: every lane writes a full 32-byte sector, a warp 1 KiB contiguous
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// One traceback step: the predecessor of state s, given the alpha column of the previous event.
// Candidates as in the forward pass (two-step class weight w2(g), one-step class weight w1(h), exact self weight);
// a predecessor that belongs to two classes appears twice with the same index and a weight <= its exact one, which
// changes neither the maximum nor the lowest index attaining it (see nc_viterbi.cu).  Ties -> lowest index, which
// is what the reference's strict '>' over the ascending from_v yields (Viterbi.hpp:78-89).
__device__ __forceinline__ unsigned tb_step(const float* __restrict__ Ap, unsigned s, const float* lut)
{
    const unsigned g = s >> 4, h = s >> 2;
    float v2[16], v1[4];
#pragma unroll
    for (int bb = 0; bb < 16; ++bb) v2[bb] = __ldcg(Ap + (bb << 8) + g);
#pragma unroll
    for (int b = 0; b < 4; ++b) v1[b] = __ldcg(Ap + (b << 10) + h);
    const float v0 = __ldcg(Ap + s);
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

} // namespace

__global__ void __launch_bounds__(VIT_THREADS, 1) viterbi_alpha_kernel(const VitArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemA& sm = *reinterpret_cast< SmemA* >(smem_raw);

    const int t = threadIdx.x;
    const int lane = t & 31;
    const int warp = t >> 5;
    const unsigned j0 = SPT * t;
    const unsigned g = t >> 1;
    float* const acol = reinterpret_cast< float* >(a.bp_pool + (size_t)blockIdx.x * a.slab_bytes);
    const bool keep = a.states != nullptr;   // path probability only: nothing to trace back, nothing stored
    const float log_2pi = a.log_2pi;
    const float hl2pi = __fmul_rn(0.5f, a.log_2pi);
    if (t == 0) mbar_init(&sm.col_bar, THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned col_phase = 0;

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

        // ---------------- prologue: scaled model constants and transition weights into registers
        StateParamsH P[SPT];
        float ws[SPT];
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
            for (int k = 0; k < SPT; ++k)
            {
                P[k] = halve(scale_state(lm[k], ls[k], sdm[k], sdl[k], lls[k], lsl[k], J, log_2pi));
                ws[k] = J.lut[trans_mask(j0 + k, j0 + k)];
            }
            if (t < 64) sm.lut[t] = J.lut[t];
        }
        // two-step weight of group g: mask bits 2..5 (bit 2 always set); one-step weight of h: bits 1..5
        const float w2 = J.lut[trans_mask(g, j0) & 0x3cu];
        float w1[2];
        w1[0] = J.lut[trans_mask(2 * t, j0) & 0x3eu];
        w1[1] = J.lut[trans_mask(2 * t + 1, j0 + 4) & 0x3eu];

        // ---------------- first chunk of events, column 0 (Viterbi.hpp:57-67)
        if (t < CH) sm.ev[0][t] = ev_pack(ev_load(a, off, t, n), J.drift);
        __syncthreads();
        float a_own[SPT];
        {
            const float4 E = sm.ev[0][0];
#pragma unroll
            for (int k = 0; k < SPT; ++k)
                a_own[k] = __fsub_rn(emission_h(P[k], E.x, E.y, __fadd_rn(E.y, E.y), E.z, E.w, hl2pi), a.log_n_states);
            float* A = sm.alpha[0];
            *reinterpret_cast< float4* >(A + phys(j0)) = make_float4(a_own[0], a_own[1], a_own[2], a_own[3]);
            *reinterpret_cast< float4* >(A + phys(j0 + 4)) = make_float4(a_own[4], a_own[5], a_own[6], a_own[7]);
            if (keep) st_cs_v8(acol + j0, a_own);
        }
        __syncthreads();
        float e_cur[SPT];
        {
            const float4 E = sm.ev[0][1 & (CH - 1)];
            const float y2 = __fadd_rn(E.y, E.y);
#pragma unroll
            for (int k = 0; k < SPT; ++k) e_cur[k] = emission_h(P[k], E.x, E.y, y2, E.z, E.w, hl2pi);
        }

        // ---------------- columns 1..n-1 (Viterbi.hpp:72-96), max only
        const int half = t & 1;
        const int two_off = phys(((8 * half) << 8) + (int)g);  // first of this thread's 8 two-step predecessors
        const int one_off = 2 * t;                              // (b<<10) + 2t, b = 0..3
        int cur = 0;
        EvRegs pre = { 0.f, 1.f, 0.f, 0.f };
        float* gcol = acol + NC_N_STATES + j0;                  // this thread's 8 slots of column i
        for (unsigned i = 1; i < n; ++i)
        {
            const unsigned ic = i & (CH - 1);
            if (ic == 1 && t < CH) pre = ev_load(a, off, (i - 1) + CH + t, n);
            if (ic == 17 && t < CH) sm.ev[(((i - 1) / CH) + 1) & 1][t] = ev_pack(pre, J.drift);

            const float4 E = sm.ev[((i + 1) / CH) & 1][(i + 1) & (CH - 1)];  // event i+1 (staged >= 1 barrier ago)
            const float* A = sm.alpha[cur];
            const float y2 = __fadd_rn(E.y, E.y);

            // two-step class: max over this thread's 8 of the group's 16 predecessors, the partner holds the rest
            float c2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) c2[k] = A[two_off + (k << 8)];
            float m2 = fmaxf(fmaxf(fmaxf(c2[0], c2[1]), fmaxf(c2[2], c2[3])), fmaxf(fmaxf(c2[4], c2[5]), fmaxf(c2[6], c2[7])));
            m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
            const float v2 = __fadd_rn(w2, m2);

            // one-step class for h = 2t and 2t+1
            float2 o[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) o[b] = *reinterpret_cast< const float2* >(A + phys((b << 10) + one_off));
            const float m1a = fmaxf(fmaxf(o[0].x, o[1].x), fmaxf(o[2].x, o[3].x));
            const float m1b = fmaxf(fmaxf(o[0].y, o[1].y), fmaxf(o[2].y, o[3].y));
            float v12[2];
            v12[0] = fmaxf(__fadd_rn(w1[0], m1a), v2);
            v12[1] = fmaxf(__fadd_rn(w1[1], m1b), v2);

            // self candidate, emission
#pragma unroll
            for (int k = 0; k < SPT; ++k)
                a_own[k] = __fadd_rn(fmaxf(__fadd_rn(ws[k], a_own[k]), v12[k >> 2]), e_cur[k]);

            float* An = sm.alpha[cur ^ 1];
            *reinterpret_cast< float4* >(An + phys(j0)) = make_float4(a_own[0], a_own[1], a_own[2], a_own[3]);
            *reinterpret_cast< float4* >(An + phys(j0 + 4)) = make_float4(a_own[4], a_own[5], a_own[6], a_own[7]);
            if (keep) st_cs_v8(gcol, a_own);
            gcol += NC_N_STATES;
            cur ^= 1;
            // column i is published: arrive now, wait after the next event's emission
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.col_bar);
#pragma unroll
            for (int k = 0; k < SPT; ++k) e_cur[k] = emission_h(P[k], E.x, E.y, y2, E.z, E.w, hl2pi);
            mbar_wait(&sm.col_bar, col_phase & 1u);
            ++col_phase;
        }

        // ---------------- fill_state_seq: argmax over the last column, strict '>' ascending j (Viterbi.hpp:123-133)
        {
            float bv = a_own[0];
            int bj = j0;
#pragma unroll
            for (int k = 1; k < SPT; ++k)
                if (a_own[k] > bv) { bv = a_own[k]; bj = j0 + k; }
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

        // ---------------- traceback (Viterbi.hpp:134-142): blocked and speculative as in viterbi_kernel, but every
        // step evaluates the arg max from the stored column i-1 (tb_step) instead of decoding a stored byte.
        if (keep)
        {
            const unsigned T = n - 1;  // transitions: column c in 1..T is entered from column c-1
            unsigned short* out_s = a.states + off;
            if (T == 0)
            {
                if (t == 0) out_s[0] = (unsigned short)sm.final_state;
            }
            else
            {
                unsigned B = (T + THREADS - 1) / THREADS;
                if (B < (unsigned)TB_MIN_BLOCK) B = TB_MIN_BLOCK;
                const unsigned nb = (T + B - 1) / B;
                const unsigned lo = (unsigned)t * B;
                const unsigned hi = (lo + B < T) ? lo + B : T;
                const bool active = (unsigned)t < nb;
                if (active)
                {
                    unsigned s;
                    if (hi == T) s = sm.final_state;
                    else
                    {
                        unsigned c = hi + TB_SPEC_DEPTH;
                        if (c >= T) { c = T; s = sm.final_state; }
                        else s = 0;
                        for (; c > hi; --c) s = tb_step(acol + (size_t)(c - 1) * NC_N_STATES, s, sm.lut);
                    }
                    sm.tb_end[t] = (unsigned short)s;
                }
                __syncthreads();
                bool dirty = active;  // first pass: everyone walks
                for (;;)
                {
                    if (dirty)
                    {
                        unsigned s = sm.tb_end[t];
                        out_s[hi] = (unsigned short)s;
                        for (unsigned c = hi; c > lo; --c)
                        {
                            s = tb_step(acol + (size_t)(c - 1) * NC_N_STATES, s, sm.lut);
                            if (c - 1 > lo || t == 0) out_s[c - 1] = (unsigned short)s;
                        }
                        sm.tb_start[t] = (unsigned short)s;
                    }
                    __syncthreads();
                    dirty = false;
                    if (active && (unsigned)t + 1 < nb && sm.tb_end[t] != sm.tb_start[t + 1]) dirty = true;
                    const int any = __syncthreads_or(dirty ? 1 : 0);
                    if (!any) break;
                    if (dirty) sm.tb_end[t] = sm.tb_start[t + 1];
                    __syncthreads();
                }
            }
            // ---------------- fill_move_seq (Viterbi.hpp:144-150)
            __syncthreads();
            if (a.moves != nullptr)
            {
                unsigned char* out_m = a.moves + off;
                for (unsigned i = t; i < n; i += THREADS)
                    out_m[i] = (i == 0) ? 0 : (unsigned char)min_skip(out_s[i - 1], out_s[i]);
            }
        }
        __syncthreads();
    }
}

size_t viterbi_alpha_smem_bytes() { return sizeof(SmemA); }

} // namespace nc
