// Helpers shared by the two Viterbi kernels (backpointer form and alpha-column form): event staging,
// the split-phase column barrier, the shared-memory slot swizzle.
#ifndef NC_VIT_COMMON_CUH
#define NC_VIT_COMMON_CUH

#include "nc_device.cuh"
#include "nc_kernels.h"

namespace nc {
namespace vit {

constexpr int THREADS = VIT_THREADS;   // 512
constexpr int SPT = 8;                 // states per thread
constexpr int CH = 128;                // events staged per chunk
constexpr int ALPHA_PAD = 16;          // bank swizzle: upper half of the column shifted by 16 floats
constexpr int TB_SPEC_DEPTH = 64;      // speculative look-back of the blocked traceback
constexpr int TB_MIN_BLOCK = 64;

// physical slot of alpha[j]: (1) the upper half of the column is shifted by 16 floats so the two threads of a
// group (two-step candidates bb 0..7 / 8..15) hit different banks; (2) bit 2 is flipped when bit 5 is set so the two
// STS.128 a thread issues for its 8 states (32-byte lane stride) are conflict-free.
__device__ __forceinline__ int phys(int j) { return (j ^ (((j >> 5) & 1) << 2)) + ((j >> 11) << 4); }


// stage one chunk of events: x = mean - drift*start (Event.hpp:81), y = stdv (0 -> 0.01,
// Event.hpp:39-42), (3*log_stdv)/2, RN(1/y)/2   (the halves are exact scalings, see emission_h)
struct EvRegs { float mean, stdv, start, lstd; };

// Split-phase CTA barrier on an mbarrier: one lane per warp arrives once the warp's alpha stores are done
// (__syncwarp orders them before the release), everybody waits on the phase parity later.  Between arrive and
// wait a warp runs the next event's emission, so a warp that finishes its recursion early keeps issuing useful
// work instead of idling at a bar.sync while the slowest warp catches up.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NC_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NC_DONE;\n"
        "bra NC_WAIT;\n"
        "NC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Host-memory calls stream the event arrays to the device while the kernels already run (nc_viterbi_packed): `landed`
// counts the events whose copy has completed, in multiples of 32 so that every 128-byte line a job touches is whole
// before the job starts.  One thread polls; the caller synchronises the CTA afterwards.
__device__ __forceinline__ void wait_events_landed(const VitArgs& a, unsigned long long off, unsigned n)
{
    unsigned long long need = (off + n + 31ull) & ~31ull;
    if (need > a.ev_total) need = a.ev_total;
    unsigned ns = 256;
    for (;;)
    {
        unsigned long long have;
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(have) : "l"(a.landed) : "memory");
        if (have >= need) break;
        __nanosleep(ns);
        if (ns < 8192) ns *= 2;
    }
    __threadfence();
}

__device__ __forceinline__ EvRegs ev_load(const VitArgs& a, unsigned long long off, unsigned i, unsigned n)
{
    EvRegs r;
    if (i < n)
    {
        // read once, and possibly written by the copy engine during this launch: L2 only
        r.mean = __ldcg(a.mean + off + i);
        r.stdv = __ldcg(a.stdv + off + i);
        r.start = __ldcg(a.start + off + i);
        r.lstd = a.log_stdv ? __ldcg(a.log_stdv + off + i) : nc_logf(r.stdv == 0.0f ? 0.01f : r.stdv);
    }
    else { r.mean = 0.f; r.stdv = 1.f; r.start = 0.f; r.lstd = 0.f; }
    return r;
}
__device__ __forceinline__ float4 ev_pack(const EvRegs& r, float drift)
{
    float y = (r.stdv == 0.0f) ? 0.01f : r.stdv;
    float x = __fsub_rn(r.mean, __fmul_rn(drift, r.start));
    return make_float4(x, y, __fmul_rn(0.5f, __fmul_rn(3.0f, r.lstd)), __fmul_rn(0.5f, __frcp_rn(y)));
}


} // namespace vit
} // namespace nc

#endif
