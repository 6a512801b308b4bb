// Kernels for an ARBITRARY state-transition table (nanocall --trans: State_Transitions::operator>>,
// State_Transitions.hpp:237-252; used wherever the transition parameters are the defaults, nanocall.cpp:651-661,
// Parameter_Trainer.hpp:118-131).  The fast kernels derive the stay/step/skip structure and its weights from bit
// patterns; a table read from a file can hold any edges in any order, so these kernels walk stored lists:
//   from lists (forward, Viterbi): predecessors of j in ascending source order, ties in file order -- the order
//       State_Transitions::update_fields builds from_v in (:79-99);
//   to lists (backward): successors of j in file order (to_v as read).
// Same arithmetic and the same tie rule as the reference (Viterbi.hpp:78-90: strict >, first maximum in list order;
// Forward_Backward.hpp:74-125: p7_FLogsum folds from -inf in list order).  One CTA per job / sequence, the previous
// column in shared memory, thread t owns states t + 512k.  This is the compatibility path, not the tuned one: a custom
// table is only in force until training has moved a read's transition parameters off the defaults.
#include "nc_device.cuh"
#include "nc_fwbw_core.cuh"
#include "nc_kernels.h"

namespace nc {

namespace {
constexpr int G_THREADS = 512;
constexpr int G_SPT = 8;
constexpr unsigned FULL = 0xffffffffu;

struct GenVitSmem
{
    float col[2][NC_N_STATES];
    float red_v[G_THREADS / 32];
    int red_j[G_THREADS / 32];
    unsigned job;
    int final_state;
};

struct GenFbSmem
{
    float tbl[fb::TBL_N];
    float col[2][NC_N_STATES];
    float lst[80];
    unsigned item;
};
} // namespace

// ------------------------------------------------------------------------------------------------ Viterbi
__global__ void __launch_bounds__(G_THREADS, 1) viterbi_generic_kernel(const VitArgs a, const GenTrans g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GenVitSmem& sm = *reinterpret_cast< GenVitSmem* >(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned short* const bp = reinterpret_cast< unsigned short* >(a.bp_pool + (size_t)blockIdx.x * a.slab_bytes);
    for (;;)
    {
        __syncthreads();
        if (t == 0) sm.job = atomicAdd(a.next_job, 1u);
        __syncthreads();
        const unsigned q = sm.job;
        if (q >= a.n_jobs) break;
        const unsigned job_idx = a.order[q];
        const DevJob& J = a.jobs[job_idx];
        const unsigned n = J.n_events;
        const unsigned long long off = J.ev_off;
        const float* M = a.models + (size_t)J.model * MODEL_FLOATS;
        StateParams P[G_SPT];
#pragma unroll
        for (int k = 0; k < G_SPT; ++k)
        {
            const int j = t + G_THREADS * k;
            P[k] = scale_state(__ldg(M + 0 * NC_N_STATES + j), __ldg(M + 1 * NC_N_STATES + j), __ldg(M + 2 * NC_N_STATES + j),
                               __ldg(M + 3 * NC_N_STATES + j), __ldg(M + 4 * NC_N_STATES + j), __ldg(M + 5 * NC_N_STATES + j), J, a.log_2pi);
        }
        int cur = 0;
        for (unsigned i = 0; i < n; ++i)
        {
            const float stdv = __ldg(a.stdv + off + i);
            const float y = (stdv == 0.0f) ? 0.01f : stdv;                                        // Event.hpp:39-42
            const float x = __fsub_rn(__ldg(a.mean + off + i), __fmul_rn(J.drift, __ldg(a.start + off + i)));   // Event.hpp:81
            const float ly3 = __fmul_rn(3.0f, a.log_stdv ? __ldg(a.log_stdv + off + i) : nc_logf(y));
            const float ry = __frcp_rn(y);
            const float* A = sm.col[cur];
            float* An = sm.col[cur ^ 1];
#pragma unroll 1
            for (int k = 0; k < G_SPT; ++k)
            {
                const int j = t + G_THREADS * k;
                const float em = emission(P[k], x, y, ly3, ry, a.log_2pi);
                if (i == 0)
                {
                    An[j] = __fsub_rn(em, a.log_n_states);                                        // Viterbi.hpp:60
                    continue;
                }
                float best = NC_NEG_INF;
                unsigned arg = NC_N_STATES;
                const unsigned e1 = __ldg(g.from_off + j + 1);
                for (unsigned e = __ldg(g.from_off + j); e < e1; ++e)
                {
                    const unsigned p = __ldg(g.from_idx + e);
                    const float v = __fadd_rn(__ldg(g.from_lp + e), A[p]);                        // :83
                    if (v > best) { best = v; arg = p; }                                          // :84-88
                }
                An[j] = __fadd_rn(best, em);                                                      // :90
                bp[(size_t)i * NC_N_STATES + j] = (unsigned short)arg;
            }
            cur ^= 1;
            __syncthreads();
        }
        // ---- arg max of the last column: strict >, ascending j (Viterbi.hpp:124-133)
        {
            const float* A = sm.col[cur];
            float bv = NC_NEG_INF;
            int bj = (int)NC_N_STATES;
#pragma unroll
            for (int k = 0; k < G_SPT; ++k)
            {
                const int j = t + G_THREADS * k;
                const float v = A[j];
                if (v > bv || (v == bv && j < bj)) { bv = v; bj = j; }
            }
            for (int d = 16; d; d >>= 1)
            {
                const float ov = __shfl_xor_sync(FULL, bv, d);
                const int oj = __shfl_xor_sync(FULL, bj, d);
                if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
            }
            if (lane == 0) { sm.red_v[warp] = bv; sm.red_j[warp] = bj; }
            __syncthreads();
            if (t == 0)
            {
                float fv = sm.red_v[0];
                int fj = sm.red_j[0];
                for (int w = 1; w < G_THREADS / 32; ++w)
                    if (sm.red_v[w] > fv || (sm.red_v[w] == fv && sm.red_j[w] < fj)) { fv = sm.red_v[w]; fj = sm.red_j[w]; }
                a.path_logprob[job_idx] = fv;
                sm.final_state = fj;
                // ---- traceback (:134-141)
                if (a.states)
                {
                    unsigned short* out_s = a.states + off;
                    unsigned s = (unsigned)fj;
                    for (unsigned i = n - 1; i > 0; --i)
                    {
                        out_s[i] = (unsigned short)s;
                        s = bp[(size_t)i * NC_N_STATES + (s & (NC_N_STATES - 1))];
                    }
                    out_s[0] = (unsigned short)s;
                }
            }
            __syncthreads();
            if (a.states && a.moves)
            {
                __threadfence_block();
                const unsigned short* out_s = a.states + off;
                for (unsigned i = t; i < n; i += G_THREADS)
                    a.moves[off + i] = (unsigned char)(i ? min_skip(out_s[i - 1], out_s[i]) : 0u);   // :143-149
            }
        }
    }
}

size_t viterbi_generic_smem_bytes() { return sizeof(GenVitSmem); }

// ------------------------------------------------------------------------------------------------ Forward/Backward
__global__ void __launch_bounds__(G_THREADS, 2) fwbw_generic_kernel(const FbArgs a, const GenTrans g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GenFbSmem& sm = *reinterpret_cast< GenFbSmem* >(smem_raw);
    const int t = threadIdx.x, lane = t & 31;
    for (int q = t; q < fb::TBL_N; q += G_THREADS) sm.tbl[q] = a.logsum_tbl[q];
    const fb::TblSmem tbl = fb::make_tbl_smem(sm.tbl);
    for (;;)
    {
        __syncthreads();
        if (t == 0) sm.item = atomicAdd(a.next_item + 1, 1u);
        __syncthreads();
        const unsigned seq = sm.item;
        if (seq >= a.n_seqs) break;
        const FbSeq& Q = a.seqs[seq];
        if (!Q.generic) continue;
        const unsigned n = Q.n_events;
        const float* E = a.scratch + Q.slab;
        float* AL = a.scratch + Q.slab + 1 * (size_t)n * NC_N_STATES;
        float* BE = a.scratch + Q.slab + 2 * (size_t)n * NC_N_STATES;
        // ---- forward (Forward_Backward.hpp:58-89)
#pragma unroll
        for (int k = 0; k < G_SPT; ++k)
        {
            const int j = t + G_THREADS * k;
            const float v = __fsub_rn(E[j], a.log_n_states);
            sm.col[0][j] = v;
            AL[j] = v;
        }
        __syncthreads();
        int cur = 0;
        for (unsigned i = 1; i < n; ++i)
        {
            const float* A = sm.col[cur];
            float* An = sm.col[cur ^ 1];
#pragma unroll 1
            for (int k = 0; k < G_SPT; ++k)
            {
                const int j = t + G_THREADS * k;
                float acc = NC_NEG_INF;
                const unsigned e1 = __ldg(g.from_off + j + 1);
                for (unsigned e = __ldg(g.from_off + j); e < e1; ++e)
                    acc = fb::flogsum(acc, __fadd_rn(__ldg(g.from_lp + e), A[__ldg(g.from_idx + e)]), tbl);
                const float v = __fadd_rn(__ldg(E + (size_t)i * NC_N_STATES + j), acc);
                An[j] = v;
                AL[(size_t)i * NC_N_STATES + j] = v;
            }
            cur ^= 1;
            __syncthreads();
        }
        // ---- log Pr[data] (:129-134): sequential fold of the last column
        if (t < 32)
        {
            const float* A = sm.col[cur];
            float acc = NC_NEG_INF;
            const unsigned lt = (1u << lane) - 1u;
            for (int base = 0; base < (int)NC_N_STATES; base += 32)
            {
                const float x = A[base + lane];
                const bool live = !((x == NC_NEG_INF) || (acc > x && __fsub_rn(acc, x) >= 15.999f));
                const unsigned m = __ballot_sync(FULL, live);
                if (m == 0) continue;
                const int cnt = __popc(m);
                if (live) sm.lst[__popc(m & lt)] = x;
                __syncwarp();
                for (int k = 0; k < cnt; ++k) acc = fb::flogsum(acc, sm.lst[k], tbl);
                __syncwarp();
            }
            if (lane == 0) a.log_pr_data[seq] = acc;
        }
        __syncthreads();
        // ---- backward (:93-125)
#pragma unroll
        for (int k = 0; k < G_SPT; ++k)
        {
            const int j = t + G_THREADS * k;
            sm.col[0][j] = 0.0f;
            BE[(size_t)(n - 1) * NC_N_STATES + j] = 0.0f;
        }
        __syncthreads();
        cur = 0;
        for (unsigned ip1 = n - 1; ip1 > 0; --ip1)
        {
            const float* Bn = sm.col[cur];
            float* Bc = sm.col[cur ^ 1];
            const float* En = E + (size_t)ip1 * NC_N_STATES;
#pragma unroll 1
            for (int k = 0; k < G_SPT; ++k)
            {
                const int j = t + G_THREADS * k;
                float acc = NC_NEG_INF;
                const unsigned e1 = __ldg(g.to_off + j + 1);
                for (unsigned e = __ldg(g.to_off + j); e < e1; ++e)
                {
                    const unsigned v = __ldg(g.to_idx + e);
                    acc = fb::flogsum(acc, __fadd_rn(__fadd_rn(__ldg(g.to_lp + e), __ldg(En + v)), Bn[v]), tbl);
                }
                Bc[j] = acc;
                BE[(size_t)(ip1 - 1) * NC_N_STATES + j] = acc;
            }
            cur ^= 1;
            __syncthreads();
        }
    }
}

size_t fwbw_generic_smem_bytes() { return sizeof(GenFbSmem); }

} // namespace nc
