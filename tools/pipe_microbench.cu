/*
 * pipe_microbench.cu -- which SM sub-partition resource do the epilogue's instruction classes occupy on sm_100a?
 *
 * Every kernel runs an unrolled mix of independent dependency chains per thread:
 *   I  integer multiply-adds      (mad.lo.s32            -> SASS IMAD)
 *   F  scalar FP32 multiply-adds  (fma.rn.f32            -> SASS FFMA)
 *   P  packed FP32 multiply-adds  (fma.rn.f32x2          -> SASS FFMA2)
 *   A  ALU-pipe operations        (shf.r.wrap / lop3-xor -> SASS SHF / LOP3)
 * with W warps per SM sub-partition, and reports cycles per iteration per SM sub-partition, i.e. the issue / pipe cost of
 * the mix.  Comparing mixes tells which classes share a pipe (costs add) and which overlap (cost = max).
 *
 *   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_microbench tools/pipe_microbench.cu
 */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int I, int F, int P, int A>
__global__ void __launch_bounds__(1024) mix_kernel(int iters, int seed, int *sink, long long *cycles)
{
    int xi[8], mi = seed | 1, ai = seed + 3;
    float xf[8], mf = 1.0f + 1e-7f * seed, af = 1e-9f * seed;
    unsigned long long xp[8], mp, ap;
    unsigned xa[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        xi[k] = threadIdx.x + k; xf[k] = 1.0f + k; xa[k] = threadIdx.x * 7 + k;
        asm("mov.b64 %0, {%1, %2};" : "=l"(xp[k]) : "f"(xf[k]), "f"(xf[k] + 0.5f));
    }
    asm("mov.b64 %0, {%1, %1};" : "=l"(mp) : "f"(mf));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ap) : "f"(af));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        /* interleave the classes so that every pipe always has something independent to take */
        constexpr int M = (I > F ? I : F) > (P > A ? P : A) ? (I > F ? I : F) : (P > A ? P : A);
#pragma unroll
        for (int k = 0; k < M; k++) {
            if (k < I) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(xi[k & 7]) : "r"(mi), "r"(ai));
            if (k < F) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xf[k & 7]) : "f"(mf), "f"(af));
            if (k < P) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xp[k & 7]) : "l"(mp), "l"(ap));
            if (k < A) {
                if (k & 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(xa[k & 7]) : "r"(xa[(k + 1) & 7]), "r"(ai));
                else asm volatile("xor.b32 %0, %0, %1;" : "+r"(xa[k & 7]) : "r"(ai));
            }
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(xp[k]));
        s += xi[k] + (int)xf[k] + (int)lo + (int)hi + (int)xa[k];
    }
    if (s == 0x7fffffff) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int I, int F, int P, int A>
static void run(const char *what, int warps_per_smsp)
{
    int *sink; long long *cyc, h[148];
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, sizeof(h));
    const int iters = 2000, threads = 128 * warps_per_smsp;
    mix_kernel<I, F, P, A><<<148, threads>>>(10, 1, sink, cyc);
    mix_kernel<I, F, P, A><<<148, threads>>>(iters, 1, sink, cyc);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148.0 * iters;
    const int n = I + F + P + A;
    printf("%-34s I=%2d F=%2d P=%2d A=%2d  warps/SMSP=%d  cycles/iter/SMSP=%8.1f  per warp-iteration=%7.2f  per instr=%5.2f\n", what, I, F, P, A,
           warps_per_smsp, avg, avg / warps_per_smsp, avg / warps_per_smsp / n);
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    for (int w = 4; w <= 8; w += 4) {
        run<16, 0, 0, 0>("IMAD only", w);
        run<0, 16, 0, 0>("FFMA only", w);
        run<0, 0, 16, 0>("FFMA2 only", w);
        run<0, 0, 0, 16>("ALU only", w);
        run<16, 16, 0, 0>("IMAD + FFMA", w);
        run<16, 0, 8, 0>("IMAD + FFMA2 (same flops)", w);
        run<16, 0, 16, 0>("IMAD + FFMA2", w);
        run<16, 0, 0, 16>("IMAD + ALU", w);
        run<0, 16, 0, 16>("FFMA + ALU", w);
        run<0, 0, 16, 16>("FFMA2 + ALU", w);
        run<16, 16, 0, 16>("IMAD + FFMA + ALU", w);
        run<16, 0, 8, 16>("IMAD + FFMA2 (same flops) + ALU", w);
        run<8, 16, 0, 16>("IMAD/2 + FFMA + ALU", w);
        run<8, 0, 8, 16>("IMAD/2 + FFMA2 (same flops) + ALU", w);
        run<12, 4, 8, 16>("epilogue-like: 12 I, 4 F, 8 P, 16 A", w);
        run<12, 20, 0, 16>("epilogue-like all scalar: 12 I, 20 F, 16 A", w);
        run<12, 12, 4, 16>("epilogue-like half packed: 12 I, 12 F, 4 P, 16 A", w);
    }
    return 0;
}
