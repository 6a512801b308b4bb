// K1: Viterbi max-plus recursion over the 4096-state 6-mer HMM, emission fused, packed uint8
// backpointers, on-device traceback.  Replaces Viterbi<float,6>::fill (Viterbi.hpp:44-150) as
// called from basecall_strand (nanocall.cpp:645-690) for a whole batch of (read-strand, model) jobs.
//
// Mapping.  One persistent CTA of 512 threads per SM pulls jobs (longest first) from a global
// counter.  Thread t owns the 8 consecutive states j = 8t..8t+7 for every event of the job and
// keeps their scaled emission constants (8 x 9 floats) in registers.  The previous alpha column
// lives in shared memory, double-buffered, one __syncthreads per event.
//
// Predecessor structure.  from_v(j) = {j} U {(b<<10)|(j>>2)} U
// {(bb<<8)|(j>>4)}.  The log-weight of an edge is a function of its overlap mask; for an edge
// that is only a two-step edge the mask depends on g = j>>4 alone, for a one-step edge (that is
// not the self loop) on h = j>>2 alone.  So
//   * the 16 two-step candidates w2(g) + alpha[(bb<<8)|g] are shared by the 16 states of a group:
//     the two threads that own a group take 8 candidates each and exchange with one shuffle;
//   * the 4 one-step candidates w1(h) + alpha[(b<<10)|h] are shared by the 4 states with that h;
//   * the self candidate uses the exact weight ws(j).
// A predecessor that belongs to two classes (e.g. a two-step predecessor that is also a one-step
// predecessor) shows up twice, once with its exact weight and once with a weight that is <= the
// exact one, for the SAME predecessor index: the extra candidate can never change the maximum or
// the index that attains it, so the result equals the reference's merged-set scan bit for bit.
// Ties resolve to the LOWEST predecessor index, as the reference's strict '>' over the ascending
// from_v does (Viterbi.hpp:78-89).
#include "nc_vit_common.cuh"

namespace nc {

using namespace vit;

namespace {

struct __align__(16) Smem
{
    float alpha[2][NC_N_STATES + ALPHA_PAD];
    float ws[NC_N_STATES + ALPHA_PAD];   // self-loop log-weights of the current job, same slots as alpha
    float4 ev[2][CH];
    float red_v[THREADS / 32];
    int red_j[THREADS / 32];
    unsigned short tb_end[THREADS];
    unsigned short tb_start[THREADS];
    unsigned job;
    int final_state;
    unsigned long long col_bar;   // mbarrier: one phase per event column
};

} // namespace

__global__ void __launch_bounds__(VIT_THREADS, 1) viterbi_kernel(const VitArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast< Smem* >(smem_raw);

    const int t = threadIdx.x;
    const int lane = t & 31;
    const int warp = t >> 5;
    const unsigned j0 = SPT * t;
    const unsigned g = t >> 1;
    unsigned char* const bp = a.bp_pool + (size_t)blockIdx.x * a.slab_bytes;
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
        if (a.landed)
        {
            if (t == 0) wait_events_landed(a, off, n);
            __syncthreads();
        }

        // ---------------- prologue: scaled model constants and transition weights into registers
        StateParamsH P[SPT];
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
                sm.ws[phys(j0 + k)] = J.lut[trans_mask(j0 + k, j0 + k)];
            }
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
        }
        __syncthreads();
        // emission of event 1, computed ahead: inside the loop the emission of event i+1 is issued next to the
        // max-plus recursion of event i (independent work: the FP pipes run it while the ALU pipe does compare/select)
        float e_cur[SPT];
        {
            const float4 E = sm.ev[0][1 & (CH - 1)];
            const float y2 = __fadd_rn(E.y, E.y);
#pragma unroll
            for (int k = 0; k < SPT; ++k) e_cur[k] = emission_h(P[k], E.x, E.y, y2, E.z, E.w, hl2pi);
        }

        // ---------------- columns 1..n-1 (Viterbi.hpp:72-96)
        const int half = t & 1;
        const int two_off = phys(((8 * half) << 8) + (int)g);  // first of this thread's 8 two-step predecessors
        const int one_off = 2 * t;                              // (b<<10) + 2t, b = 0..3
        int cur = 0;
        EvRegs pre = { 0.f, 1.f, 0.f, 0.f };
        for (unsigned i = 1; i < n; ++i)
        {
            const unsigned ic = i & (CH - 1);
            // prefetch the next chunk of events: loads issued at the start of a chunk, stored 16 events later
            if (ic == 1 && t < CH) pre = ev_load(a, off, (i - 1) + CH + t, n);
            if (ic == 17 && t < CH) sm.ev[(((i - 1) / CH) + 1) & 1][t] = ev_pack(pre, J.drift);

            const float4 E = sm.ev[((i + 1) / CH) & 1][(i + 1) & (CH - 1)];  // event i+1 (staged >= 1 barrier ago)
            float ws[SPT];
            *reinterpret_cast< float4* >(ws) = *reinterpret_cast< const float4* >(sm.ws + phys(j0));
            *reinterpret_cast< float4* >(ws + 4) = *reinterpret_cast< const float4* >(sm.ws + phys(j0 + 4));
            const float* A = sm.alpha[cur];
            const float y2 = __fadd_rn(E.y, E.y);

            // two-step candidates: 8 of the group's 16, ascending bb, strict '>'
            // (tournament instead of a linear scan: the right operand wins only if strictly greater, so the lowest
            //  index still wins ties, and the dependent chain is 3 compares deep instead of 8)
            float c2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) c2[k] = __fadd_rn(w2, A[two_off + (k << 8)]);
            float m01 = c2[0], m23 = c2[2], m45 = c2[4], m67 = c2[6];
            int i01 = 0, i23 = 2, i45 = 4, i67 = 6;
            if (c2[1] > m01) { m01 = c2[1]; i01 = 1; }
            if (c2[3] > m23) { m23 = c2[3]; i23 = 3; }
            if (c2[5] > m45) { m45 = c2[5]; i45 = 5; }
            if (c2[7] > m67) { m67 = c2[7]; i67 = 7; }
            if (m23 > m01) { m01 = m23; i01 = i23; }
            if (m67 > m45) { m45 = m67; i45 = i67; }
            float v2 = m01;
            int bb = i01;
            if (m45 > v2) { v2 = m45; bb = i45; }
            bb += 8 * half;
            {
                float ov = __shfl_xor_sync(0xffffffffu, v2, 1);
                int ob = __shfl_xor_sync(0xffffffffu, bb, 1);
                // the even thread holds bb 0..7: it wins ties
                bool take = half ? (ov >= v2) : (ov > v2);
                v2 = take ? ov : v2;
                bb = take ? ob : bb;
            }
            const int p2 = (bb << 8) | (int)g;

            // one-step candidates for h = 2t and 2t+1: ascending b, strict '>'
            float v1[2];
            int b1[2];
            {
                float ca[4], cb[4];
#pragma unroll
                for (int b = 0; b < 4; ++b)
                {
                    float2 o = *reinterpret_cast< const float2* >(A + phys((b << 10) + one_off));
                    ca[b] = __fadd_rn(w1[0], o.x);
                    cb[b] = __fadd_rn(w1[1], o.y);
                }
                float xa = ca[0], ya = ca[2], xb = cb[0], yb = cb[2];
                int ia = 0, ja = 2, ib = 0, jb = 2;
                if (ca[1] > xa) { xa = ca[1]; ia = 1; }
                if (ca[3] > ya) { ya = ca[3]; ja = 3; }
                if (cb[1] > xb) { xb = cb[1]; ib = 1; }
                if (cb[3] > yb) { yb = cb[3]; jb = 3; }
                if (ya > xa) { xa = ya; ia = ja; }
                if (yb > xb) { xb = yb; ib = jb; }
                v1[0] = xa; b1[0] = ia; v1[1] = xb; b1[1] = ib;
            }
            // merge two-step and one-step per h; equal values -> lower predecessor index
            float v12[2];
            int p12[2], c12[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
            {
                int p1 = (b1[hh] << 10) | (2 * t + hh);
                bool take1 = (v1[hh] > v2) | ((v1[hh] == v2) & (p1 < p2));
                v12[hh] = take1 ? v1[hh] : v2;
                p12[hh] = take1 ? p1 : p2;
                c12[hh] = take1 ? (16 + b1[hh]) : bb;
            }
            // self candidate, emission, new alpha, backpointer code.  The 4 codes of an h-group start as the group's
            // code replicated into every byte; a state whose self loop wins gets its byte replaced by 20 through a
            // byte mask, so packing costs one select per state and three logic ops per word.
            float a_new[SPT];
            unsigned code_w[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
            {
                unsigned m = 0;
                const int tie = p12[hh] - (int)j0 - 4 * hh;  // state kk of the group wins a tie iff kk < tie
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                {
                    const int k = 4 * hh + kk;
                    float vs = __fadd_rn(ws[k], a_own[k]);
                    bool takes = (vs > v12[hh]) | ((vs == v12[hh]) & (kk < tie));
                    float best = takes ? vs : v12[hh];
                    m |= takes ? (0xffu << (8 * kk)) : 0u;
                    a_new[k] = __fadd_rn(best, e_cur[k]);
                }
                const unsigned base = (unsigned)c12[hh] * 0x01010101u;
                code_w[hh] = (base & ~m) | (0x14141414u & m);
            }
            const unsigned code_lo = code_w[0], code_hi = code_w[1];
            float* An = sm.alpha[cur ^ 1];
            *reinterpret_cast< float4* >(An + phys(j0)) = make_float4(a_new[0], a_new[1], a_new[2], a_new[3]);
            *reinterpret_cast< float4* >(An + phys(j0 + 4)) = make_float4(a_new[4], a_new[5], a_new[6], a_new[7]);
            // 8 backpointer bytes per thread, a warp writes 256 contiguous bytes
            __stcs(reinterpret_cast< uint2* >(bp + (size_t)i * NC_N_STATES + j0), make_uint2(code_lo, code_hi));
#pragma unroll
            for (int k = 0; k < SPT; ++k) a_own[k] = a_new[k];
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
            __syncthreads();
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

        // ---------------- traceback (Viterbi.hpp:134-142), blocked and speculative.
        // The chain s[i-1] = pred(bp[i][s[i]]) is cut into nb blocks of B transitions; block b starts from a
        // GUESS of its end state, obtained by walking back from column end+D starting at an arbitrary state
        // (survivor paths coalesce within a few tens of events).  Every guess is then verified against the
        // state its successor block actually reached; a block whose guess was wrong is re-walked, so the
        // result is exactly the sequential traceback.
        if (a.states != nullptr)
        {
            const unsigned T = n - 1;  // transitions = columns 1..T carry backpointers
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
                        for (; c > hi; --c) s = bp_decode(__ldcg(bp + (size_t)c * NC_N_STATES + s), s);
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
                            s = bp_decode(__ldcg(bp + (size_t)c * NC_N_STATES + s), s);
                            if (c - 1 > lo || t == 0) out_s[c - 1] = (unsigned short)s;
                        }
                        sm.tb_start[t] = (unsigned short)s;
                    }
                    __syncthreads();
                    dirty = false;
                    if (active && (unsigned)t + 1 < nb && sm.tb_end[t] != sm.tb_start[t + 1])
                    {
                        dirty = true;
                    }
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

size_t viterbi_smem_bytes() { return sizeof(Smem); }

} // namespace nc
