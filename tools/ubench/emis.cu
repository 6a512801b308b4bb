// Cost of the fused emission on sm_100a: scalar (emission_h x 8 states) vs packed (emission2 x 4 pairs), 16 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nanocall_b200/csrc/nc_device.cuh"
using namespace nc;
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_of(f2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi_of(f2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct PairParams { f2 nmu, nsg2, rsgh, nls, neta, reta, lam, c1h; };
struct EvPairs { f2 X, Y, Y2, NLY, NRY; };
__device__ __forceinline__ f2 emission2(const PairParams& p, const EvPairs& e, f2 m2, f2 nhl2pi)
{
    const f2 t1 = add2(e.X, p.nmu);
    const f2 q0 = mul2(t1, p.rsgh);
    const f2 r = fma2(q0, p.nsg2, t1);
    const f2 ah = fma2(r, p.rsgh, q0);
    const f2 ns = fma2(mul2(ah, ah), m2, nhl2pi);
    const f2 ln = add2(p.nls, ns);
    const f2 t2 = add2(e.Y, p.neta);
    const f2 p0 = mul2(t2, p.reta);
    const f2 r2 = fma2(p0, p.neta, t2);
    const f2 b = fma2(r2, p.reta, p0);
    const f2 l2 = mul2(mul2(p.lam, b), b);
    const f2 nq = mul2(l2, e.NRY);
    const f2 r3 = fma2(nq, e.Y2, l2);
    const f2 nuh = fma2(r3, e.NRY, nq);
    const f2 li = add2(add2(p.c1h, e.NLY), nuh);
    return add2(ln, li);
}
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, const float* in)
{
    float x = in[0], y = in[1], ly = in[2], ry = in[3], hl = in[4];
    float acc[8];
    StateParamsH P[8];
    PairParams Q[4];
    for (int k = 0; k < 8; ++k)
    {
        P[k].mu = in[8 + k] + threadIdx.x * 1e-3f; P[k].sg2 = 2.f + in[k]; P[k].rsgh = 0.25f + in[k] * 1e-3f; P[k].nls = in[k];
        P[k].eta = 1.f + in[k]; P[k].reta = 1.f / P[k].eta; P[k].lam = 2.f + in[k]; P[k].c1h = in[k] + 1.f;
        acc[k] = 0.f;
    }
    for (int k = 0; k < 4; ++k)
    {
        const StateParamsH &a = P[2 * k], &b = P[2 * k + 1];
        Q[k].nmu = pk(-a.mu, -b.mu); Q[k].nsg2 = pk(-a.sg2, -b.sg2); Q[k].rsgh = pk(a.rsgh, b.rsgh); Q[k].nls = pk(a.nls, b.nls);
        Q[k].neta = pk(-a.eta, -b.eta); Q[k].reta = pk(a.reta, b.reta); Q[k].lam = pk(a.lam, b.lam); Q[k].c1h = pk(a.c1h, b.c1h);
    }
    const f2 M2 = pk(-2.f, -2.f), NH = pk(-hl, -hl);
    f2 acc2[4] = { 0, 0, 0, 0 };
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
        if (MODE == 0)
        {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = __fadd_rn(fmaxf(acc[k], x), emission_h(P[k], x, y, __fadd_rn(y, y), ly, ry, hl));
        }
        else
        {
            EvPairs E;
            E.X = pk(x, x); E.Y = pk(y, y); E.Y2 = pk(y + y, y + y); E.NLY = pk(-ly, -ly); E.NRY = pk(-ry, -ry);
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                f2 e = emission2(Q[k], E, M2, NH);
                acc2[k] = add2(pk(fmaxf(lo_of(acc2[k]), x), fmaxf(hi_of(acc2[k]), x)), e);
            }
        }
        x = __fadd_rn(x, 1e-3f); y = __fadd_rn(y, 1e-4f);
    }
    long long t1 = clock64();
    float s = 0;
    for (int k = 0; k < 8; ++k) s += acc[k];
    for (int k = 0; k < 4; ++k) s += lo_of(acc2[k]) + hi_of(acc2[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, float* out, long long* cyc, const float* in)
{
    k<MODE><<<148, threads>>>(out, cyc, in); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, in); cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    printf("%-8s threads=%d cycles/iter(all warps of the SM, 8 states/thread)=%.1f  => per scheduler per 4096-state column: %.1f  %s\n",
           name, threads, m / ITER, m / ITER * (512.0 / threads), e ? cudaGetErrorString(e) : "");
}
int main()
{
    float* out; long long* cyc; float* in;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&in, 256);
    float hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 0.5f + 0.01f * i;
    cudaMemcpy(in, hin, 256, cudaMemcpyHostToDevice);
    for (int th : { 128, 256, 512 }) { run<0>("scalar", th, out, cyc, in); run<1>("packed", th, out, cyc, in); }
    return 0;
}
