// TEST INFRASTRUCTURE: host emulation of the Forward/Backward kernels' per-thread column code
// (nanocall_b200/csrc/nc_fwbw_core.cuh) for one CTA of 512 threads, so that the ORDER logic (shared prefixes,
// merged duplicate edges, warp-uniform path selection) is checked bit for bit against the oracle without a GPU.
// Built by tests/test_fwbw_emu.py:  g++ -O2 -ffp-contract=off -shared -fPIC
#include "nc_fwbw_core.cuh"

#include <cstdint>
#include <cstring>
#include <vector>

using namespace nc::fb;

extern "C" {

// alpha[n][4096] from E[n][4096] (Forward_Backward.hpp:58-89); tbl = the kernels' table (entry 15999 zeroed)
int emu_forward(const float* lut, const float* tbl, const float* E, uint32_t n, float log_n_states, float* alpha)
{
    std::vector< float > col[2] = { std::vector< float >(COL_FLOATS, 0.f), std::vector< float >(COL_FLOATS, 0.f) };
    std::vector< FwdConst > C(THREADS);
    std::vector< float > own(THREADS * 8);
    for (int t = 0; t < THREADS; ++t) fwd_const_init(C[t], fwd_logical_thread(t), lut);
    // every logical thread exactly once
    std::vector< int > seen(THREADS, 0);
    for (int t = 0; t < THREADS; ++t) ++seen[C[t].u];
    for (int t = 0; t < THREADS; ++t) if (seen[t] != 1) return -1;
    for (int t = 0; t < THREADS; ++t)
        for (int k = 0; k < 8; ++k)
        {
            const int j = 8 * C[t].u + k;
            const float a = E[j] - log_n_states;
            own[8 * t + k] = a;
            col[0][cphys(j)] = a;
            alpha[j] = a;
        }
    int cur = 0;
    for (uint32_t i = 1; i < n; ++i)
    {
        for (int t = 0; t < THREADS; ++t)
        {
            float e[8], o[8];
            for (int k = 0; k < 8; ++k) { e[k] = E[(size_t)i * N_STATES + 8 * C[t].u + k]; o[k] = own[8 * t + k]; }
            fwd_column(C[t], col[cur].data(), TblPtr{ tbl }, e, o);
            for (int k = 0; k < 8; ++k)
            {
                own[8 * t + k] = o[k];
                col[cur ^ 1][cphys(8 * C[t].u + k)] = o[k];
                alpha[(size_t)i * N_STATES + 8 * C[t].u + k] = o[k];
            }
        }
        cur ^= 1;
    }
    return 0;
}

// beta[n][4096] from E[n][4096] (Forward_Backward.hpp:93-125)
int emu_backward(const float* lut, const float* tbl, const float* E, uint32_t n, float* beta, uint32_t* path_hist)
{
    std::vector< float > col[2] = { std::vector< float >(COL_FLOATS, 0.f), std::vector< float >(COL_FLOATS, 0.f) };
    std::vector< BwdConst > C(THREADS);
    for (int t = 0; t < THREADS; ++t) bwd_const_init(C[t], t, lut);
    for (int w = 0; w < THREADS / 32; ++w)
        for (int k = 0; k < 8; ++k)
        {
            const unsigned c0 = bwd_lane_code(32 * w, k);
            bool all = true;
            for (int l = 1; l < 32; ++l) all = all && bwd_lane_code(32 * w + l, k) == c0;
            const unsigned path = all ? c0 : 0u;
            if (path_hist) ++path_hist[path];
            for (int l = 0; l < 32; ++l) C[32 * w + l].paths |= path << (3 * k);
        }
    for (int j = 0; j < N_STATES; ++j) { col[0][cphys(j)] = 0.f; beta[(size_t)(n - 1) * N_STATES + j] = 0.f; }
    int cur = 0;
    for (uint32_t ip1 = n - 1; ip1 > 0; --ip1)
    {
        const uint32_t i = ip1 - 1;
        float* out = beta + (size_t)i * N_STATES;
        float* nxt = col[cur ^ 1].data();
        for (int t = 0; t < THREADS; ++t)
            bwd_column(C[t], col[cur].data(), E + (size_t)ip1 * N_STATES, lut, TblPtr{ tbl },
                       [&](int j, float v) { out[j] = v; nxt[cphys(j)] = v; });
        cur ^= 1;
    }
    return 0;
}

}
