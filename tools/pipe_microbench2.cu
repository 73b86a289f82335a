/*
 * pipe_microbench2.cu -- cost of single instruction classes on an sm_100a SM sub-partition, alone and next to IMAD:
 * integer multiply without addend, multiply-add with an immediate factor, dp2a / dp4a, int -> float conversion,
 * min/max, and the packed FP32 forms.  Same harness as pipe_microbench.cu (8 independent chains per thread, 4 warps per
 * sub-partition); prints cycles per warp-instruction.
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_microbench2 tools/pipe_microbench2.cu
 */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { IMAD3, IMUL, IMADI, IDP2, IDP4, I2F, FMNMX, FFMA, FFMA2, LEA_, SHF_, NONE };

template <Op OP> __device__ __forceinline__ void one(int &x, float &f, unsigned long long &p, int m, int a, float mf, float af, unsigned long long mp, unsigned long long ap)
{
    if (OP == IMAD3) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a));
    if (OP == IMUL)  asm volatile("mul.lo.s32 %0, %0, %1;" : "+r"(x) : "r"(m));
    if (OP == IMADI) asm volatile("mad.lo.s32 %0, %0, 257, %1;" : "+r"(x) : "r"(a));
    if (OP == IDP2)  asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a));
    if (OP == IDP4)  asm volatile("dp4a.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a));
    if (OP == I2F)   { float t; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(x)); x ^= __float_as_int(t); }
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f) : "f"(mf));
    if (OP == FFMA)  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(mf), "f"(af));
    if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(mp), "l"(ap));
    if (OP == LEA_)  asm volatile("{ .reg .b32 t; shl.b32 t, %0, 2; add.s32 %0, t, %1; }" : "+r"(x) : "r"(a));
    if (OP == SHF_)  asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(a));
}

template <Op A, Op B, int NA, int NB>
__global__ void __launch_bounds__(1024) k(int iters, int seed, int *sink, long long *cycles)
{
    int xa[8], xb[8], m = seed | 1, a = seed + 3;
    float fa[8], fb[8], mf = 1.0f + 1e-7f * seed, af = 1e-9f * seed;
    unsigned long long pa[8], pb[8], mp, ap;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        xa[i] = threadIdx.x + i; xb[i] = threadIdx.x * 3 + i; fa[i] = 1.0f + i; fb[i] = 2.0f + i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa[i]) : "f"(fa[i]), "f"(fb[i]));
        pb[i] = pa[i] + 1;
    }
    asm("mov.b64 %0, {%1, %1};" : "=l"(mp) : "f"(mf));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ap) : "f"(af));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        constexpr int M = NA > NB ? NA : NB;
#pragma unroll
        for (int i = 0; i < M; i++) {
            if (i < NA) one<A>(xa[i & 7], fa[i & 7], pa[i & 7], m, a, mf, af, mp, ap);
            if (i < NB) one<B>(xb[i & 7], fb[i & 7], pb[i & 7], m, a, mf, af, mp, ap);
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += xa[i] + xb[i] + (int)fa[i] + (int)fb[i] + (int)pa[i] + (int)pb[i];
    if (s == 0x7fffffff) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <Op A, Op B, int NA, int NB>
static void run(const char *what)
{
    int *sink; long long *cyc, h[148];
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, sizeof(h));
    const int iters = 2000, threads = 512;
    k<A, B, NA, NB><<<148, threads>>>(10, 1, sink, cyc);
    k<A, B, NA, NB><<<148, threads>>>(iters, 1, sink, cyc);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148.0 * iters * 4;
    printf("%-40s  %2d + %2d instr: %6.2f cycles per warp-iteration, %5.2f per instruction\n", what, NA, NB, avg, avg / (NA + NB));
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    run<IMAD3, NONE, 16, 0>("IMAD a*b+c");
    run<IMUL, NONE, 16, 0>("IMAD a*b (no addend)");
    run<IMADI, NONE, 16, 0>("IMAD a*imm+c");
    run<IDP2, NONE, 16, 0>("IDP.2A");
    run<IDP4, NONE, 16, 0>("IDP.4A");
    run<I2F, NONE, 16, 0>("I2FP.F32.S32 (+ LOP3)");
    run<FMNMX, NONE, 16, 0>("FMNMX");
    run<LEA_, NONE, 16, 0>("shl + add (LEA?)");
    run<SHF_, NONE, 16, 0>("SHF");
    run<IMAD3, IDP2, 16, 16>("IMAD + IDP.2A");
    run<IMAD3, IDP4, 16, 16>("IMAD + IDP.4A");
    run<IMAD3, I2F, 16, 16>("IMAD + I2FP(+LOP3)");
    run<IMAD3, FMNMX, 16, 16>("IMAD + FMNMX");
    run<IMAD3, LEA_, 16, 16>("IMAD + shl/add");
    run<IMAD3, SHF_, 16, 16>("IMAD + SHF");
    run<IMAD3, IMUL, 16, 16>("IMAD + IMUL");
    run<IMAD3, IMADI, 16, 16>("IMAD + IMAD imm");
    run<FFMA2, IDP2, 16, 16>("FFMA2 + IDP.2A");
    run<FFMA, IDP2, 16, 16>("FFMA + IDP.2A");
    run<SHF_, IDP2, 16, 16>("SHF + IDP.2A");
    return 0;
}
