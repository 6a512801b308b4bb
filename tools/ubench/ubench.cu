// Instruction-throughput microbenchmarks for sm_100a (B200): which pipe binds the Viterbi recursion.
// Each test: 1 CTA per SM, NW warps per SMSP, 8 independent chains per thread, ITER x 32 ops.
// Reports warp-instructions issued per cycle per SM sub-partition (SMSP).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITER 2048
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

#define OP_FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c1));
#define OP_FMUL(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c1));
#define OP_FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c1), "f"(c2));
#define OP_FFMA4(i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(x[(i + 1) & 7]), "f"(x[(i + 2) & 7]), "f"(x[(i + 3) & 7]));
#define OP_FMNMX(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c1));
#define OP_FMNMX3(i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c1), "f"(c2));
#define OP_SETSEL(i) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(x[i]) : "f"(c1), "f"(c2));
#define OP_IADD(i) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(k1));
#define OP_LOP(i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(k1));
#define OP_IMAD(i) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(k1), "r"(k2));
#define OP_FFMA2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(d1), "l"(d2));
#define OP_FFMA2R(i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(w[i]) : "l"(w[(i + 1) & 7]), "l"(w[(i + 2) & 7]), "l"(w[(i + 3) & 7]));
#define OP_FADD2R(i) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(w[i]) : "l"(w[(i + 1) & 7]), "l"(w[(i + 2) & 7]));
#define OP_MIX_FFMA2R_FMNMX3(i) OP_FFMA2R(i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(c2), "f"(c1));
#define OP_MIX_FFMA2R_LDS128(i) OP_FFMA2R(i) OP_LDS128(i)
#define OP_MIX_FADD2_FMNMX(i) OP_FADD2(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_FADD2(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(d1));
#define OP_FMUL2(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(d1));
#define OP_LDS32(i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[i]) : "r"(sa + 128 * i));
#define OP_LDS64(i) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[i]), "=f"(y[i]) : "r"(sa2 + 256 * i));
#define OP_LDS128(i) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[i]), "=f"(y[i]), "=f"(z[i]), "=f"(v[i]) : "r"(sa4 + 512 * i));
#define OP_MIX_FADD_FMNMX(i) OP_FADD(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_MIX_FFMA_FMNMX(i) OP_FFMA(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_MIX_2FADD_FMNMX(i) OP_FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(c2)); asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_MIX_3FADD_FMNMX(i) OP_FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(c2)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(c2)); asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_MIX_FFMA2_FMNMX(i) OP_FFMA2(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_MIX_FFMA2_FADD(i) OP_FFMA2(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
#define OP_MIX_FFMA2_2FMNMX(i) OP_FFMA2(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2)); asm volatile("max.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(c1));
#define OP_MIX_FADD_LDS(i) OP_FADD(i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[i]) : "r"(sa + 128 * i));
#define OP_MIX_FADD_IMAD(i) OP_FADD(i) OP_IMAD(i)
#define OP_MIX_FMNMX_IADD(i) OP_FMNMX(i) OP_IADD(i)
#define OP_MIX_FMNMX_IMAD(i) OP_FMNMX(i) OP_IMAD(i)

#define DEFK(NAME, OPS_PER, OP)                                                                              \
    __global__ void __launch_bounds__(1024, 1) k_##NAME(float* out, long long* cyc, const float* in)         \
    {                                                                                                        \
        extern __shared__ float smf[];                                                                       \
        for (int i = threadIdx.x; i < 8192; i += blockDim.x) smf[i] = in[i & 7];                            \
        __syncthreads();                                                                                     \
        float c1 = in[0], c2 = in[1];                                                                        \
        int k1 = (int)in[2], k2 = (int)in[3];                                                                \
        float x[8], y[8], z[8], v[8];                                                                        \
        int u[8];                                                                                            \
        unsigned long long w[8], d1, d2;                                                                     \
        unsigned sa = (unsigned)__cvta_generic_to_shared(smf) + 4 * (threadIdx.x & 31);                      \
        unsigned sa2 = (unsigned)__cvta_generic_to_shared(smf) + 8 * (threadIdx.x & 31);                     \
        unsigned sa4 = (unsigned)__cvta_generic_to_shared(smf) + 16 * (threadIdx.x & 31);                    \
        for (int i = 0; i < 8; ++i) { x[i] = in[i & 3] + i; y[i] = x[i] + 1; z[i] = y[i] + 1; v[i] = z[i] + 1; u[i] = i + k1; \
            asm volatile("mov.b64 %0, {%1,%2};" : "=l"(w[i]) : "f"(x[i]), "f"(y[i])); }                      \
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d1) : "f"(c1), "f"(c1));                                \
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d2) : "f"(c2), "f"(c2));                                \
        __syncthreads();                                                                                     \
        long long t0 = clock64();                                                                            \
        _Pragma("unroll 1") for (int it = 0; it < ITER; ++it) { REP8(OP) REP8(OP) REP8(OP) REP8(OP) }       \
        long long t1 = clock64();                                                                            \
        float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(w[i])); \
            s += x[i] + y[i] + z[i] + v[i] + u[i] + lo + hi; }                                               \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                                      \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                     \
    }                                                                                                        \
    static void run_##NAME(int nw, float* out, long long* cyc, const float* in)                              \
    {                                                                                                        \
        cudaFuncSetAttribute(k_##NAME, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);                  \
        k_##NAME<<< 148, nw * 128, 32768 >>>(out, cyc, in);                                                  \
        cudaDeviceSynchronize();                                                                             \
        k_##NAME<<< 148, nw * 128, 32768 >>>(out, cyc, in);                                                  \
        cudaError_t e = cudaDeviceSynchronize();                                                             \
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                             \
        double m = 0; for (int i = 0; i < 148; ++i) m += (double)h[i]; m /= 148;                             \
        double ipc = (double)nw * ITER * 32.0 * OPS_PER / m;                                                 \
        printf("%-18s nw/smsp=%d  ops/iter=%d  cycles=%.0f  warp-instr/clk/SMSP=%.3f  %s\n", #NAME, nw, OPS_PER, m, ipc, e ? cudaGetErrorString(e) : ""); \
    }

DEFK(fadd, 1, OP_FADD)
DEFK(fmul, 1, OP_FMUL)
DEFK(ffma, 1, OP_FFMA)
DEFK(ffma4, 1, OP_FFMA4)
DEFK(fmnmx, 1, OP_FMNMX)
DEFK(fmnmx3, 1, OP_FMNMX3)
DEFK(setsel, 2, OP_SETSEL)
DEFK(iadd, 1, OP_IADD)
DEFK(lop, 1, OP_LOP)
DEFK(imad, 1, OP_IMAD)
DEFK(ffma2, 1, OP_FFMA2)
DEFK(fadd2, 1, OP_FADD2)
DEFK(fmul2, 1, OP_FMUL2)
DEFK(ffma2r, 1, OP_FFMA2R)
DEFK(fadd2r, 1, OP_FADD2R)
DEFK(mix_ffma2r_fmnmx3, 2, OP_MIX_FFMA2R_FMNMX3)
DEFK(mix_ffma2r_lds128, 2, OP_MIX_FFMA2R_LDS128)
DEFK(mix_fadd2_fmnmx, 2, OP_MIX_FADD2_FMNMX)
DEFK(lds32, 1, OP_LDS32)
DEFK(lds64, 1, OP_LDS64)
DEFK(lds128, 1, OP_LDS128)
DEFK(mix_fadd_fmnmx, 2, OP_MIX_FADD_FMNMX)
DEFK(mix_ffma_fmnmx, 2, OP_MIX_FFMA_FMNMX)
DEFK(mix_2fadd_fmnmx, 3, OP_MIX_2FADD_FMNMX)
DEFK(mix_3fadd_fmnmx, 4, OP_MIX_3FADD_FMNMX)
DEFK(mix_ffma2_fmnmx, 2, OP_MIX_FFMA2_FMNMX)
DEFK(mix_ffma2_fadd, 2, OP_MIX_FFMA2_FADD)
DEFK(mix_ffma2_2fmnmx, 3, OP_MIX_FFMA2_2FMNMX)
DEFK(mix_fadd_lds, 2, OP_MIX_FADD_LDS)
DEFK(mix_fadd_imad, 2, OP_MIX_FADD_IMAD)
DEFK(mix_fmnmx_iadd, 2, OP_MIX_FMNMX_IADD)
DEFK(mix_fmnmx_imad, 2, OP_MIX_FMNMX_IMAD)

int main()
{
    float* out; long long* cyc; float* in;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&in, 64);
    float hin[16] = { 1.0f, 0.0f, 3.0f, 5.0f, 1.5f, 2.5f, 3.5f, 4.5f };
    cudaMemcpy(in, hin, 64, cudaMemcpyHostToDevice);
    int nws[3] = { 2, 4, 8 };
    for (int q = 0; q < 3; ++q)
    {
        int nw = nws[q];
        run_fadd(nw, out, cyc, in); run_fmul(nw, out, cyc, in); run_ffma(nw, out, cyc, in); run_ffma4(nw, out, cyc, in);
        run_fmnmx(nw, out, cyc, in); run_fmnmx3(nw, out, cyc, in); run_setsel(nw, out, cyc, in);
        run_iadd(nw, out, cyc, in); run_lop(nw, out, cyc, in); run_imad(nw, out, cyc, in);
        run_ffma2(nw, out, cyc, in); run_fadd2(nw, out, cyc, in); run_fmul2(nw, out, cyc, in);
        run_ffma2r(nw, out, cyc, in); run_fadd2r(nw, out, cyc, in); run_mix_ffma2r_fmnmx3(nw, out, cyc, in);
        run_mix_ffma2r_lds128(nw, out, cyc, in); run_mix_fadd2_fmnmx(nw, out, cyc, in);
        run_lds32(nw, out, cyc, in); run_lds64(nw, out, cyc, in); run_lds128(nw, out, cyc, in);
        run_mix_fadd_fmnmx(nw, out, cyc, in); run_mix_ffma_fmnmx(nw, out, cyc, in); run_mix_2fadd_fmnmx(nw, out, cyc, in);
        run_mix_3fadd_fmnmx(nw, out, cyc, in); run_mix_ffma2_fmnmx(nw, out, cyc, in); run_mix_ffma2_fadd(nw, out, cyc, in);
        run_mix_ffma2_2fmnmx(nw, out, cyc, in); run_mix_fadd_lds(nw, out, cyc, in); run_mix_fadd_imad(nw, out, cyc, in);
        run_mix_fmnmx_iadd(nw, out, cyc, in); run_mix_fmnmx_imad(nw, out, cyc, in);
        printf("\n");
    }
    return 0;
}
