/*
 * tc_microbench.cu -- diagnostics only (not part of the product library): how fast does one SM retire
 * tcgen05.mma kind::i8 M=128 instructions as a function of N, the accumulator rotation, operand placement
 * (K-major no-swizzle "slab" layout as used by tc_engine.cu) and signedness switching?
 *   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tsl-sdr_b200/csrc -o tools/tc_microbench tools/tc_microbench.cu
 */
#include "tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace tslb200;

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

struct Case { int N, nacc, a_adv16, b_adv16, sw, nmma, a_rows16 /* LBO of A >> 4 */, b_lbo16, twice_mid; };

__global__ void __launch_bounds__(128, 1) mb_kernel(Case c, int reps, long long *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::fence_proxy_async(); ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        const uint64_t dA0 = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(smem), c.a_rows16 * 16, 128);
        const uint64_t dB0 = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(smem) + 120 * 1024, c.b_lbo16 * 16, 128);
        const uint32_t id0 = ptx::idesc_i8(128, c.N, true, true), id1 = ptx::idesc_i8(128, c.N, false, false);
        long long best = 1ll << 60;
        /* accumulator slots of the 4 MMAs of one step */
        const uint32_t s0 = 0, s1 = (c.nacc > 1 ? 1 : 0) * c.N, s2 = (c.nacc == 4 ? 1 : (c.nacc >= 3 ? 2 : 0)) * c.N, s3 = (c.nacc == 4 ? 2 : (c.nacc == 2 ? 1 : 0)) * c.N;
        for (int r = 0; r < reps; r++) {
            __syncwarp();
            const long long t0 = clock64();
            for (int i = 0; i < c.nmma; i += 4) {
                const uint64_t da = dA0 + (uint64_t)((uint32_t)((i >> 2) & 3) * c.a_adv16);
                const uint64_t db = dB0 + (uint64_t)((uint32_t)((i >> 2) & 3) * c.b_adv16);
                const uint32_t i1 = c.sw ? id1 : id0;
                if (elect_one()) {
                    ptx::mma_i8(tmem + s0, da, db, id0, 1);
                    ptx::mma_i8(tmem + s1, da, db + 1, i1, 1);
                    ptx::mma_i8(tmem + s2, da + c.a_adv16, db, id0, 1);
                    ptx::mma_i8(tmem + s3, da + c.a_adv16, db + 1, i1, 1);
                }
            }
            if (tid == 0) ptx::mma_commit(&bar);
            ptx::mbar_wait(&bar, r & 1);
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (tid == 0) out[blockIdx.x] = best;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main()
{
    long long *d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    /* N, nacc, a_adv16, b_adv16, sw, nmma, a_lbo16, b_lbo16, twice_mid */
    const Case cases[] = {
        { 64, 1, 0, 0, 0, 64, 128, 65, 0 },       /* fixed operands, one accumulator */
        { 64, 1, 256, 130, 0, 64, 128, 65, 0 },   /* operands advance like the engine's K loop */
        { 64, 3, 256, 130, 0, 64, 128, 65, 0 },   /* 3 accumulators round robin */
        { 64, 4, 256, 130, 0, 64, 128, 65, 1 },   /* 0,1,1,2 pattern */
        { 64, 4, 256, 130, 1, 64, 128, 65, 1 },   /* + signedness switching */
        { 64, 1, 256, 130, 1, 64, 128, 65, 0 },   /* one acc, signedness switching */
        { 128, 1, 256, 130, 0, 64, 128, 129, 0 }, /* N = 128 */
        { 128, 3, 256, 130, 0, 64, 128, 129, 0 },
        { 256, 1, 256, 130, 0, 64, 128, 257, 0 }, /* N = 256 */
        { 256, 2, 256, 130, 0, 64, 128, 257, 0 },
        { 32, 1, 256, 130, 0, 64, 128, 65, 0 },
        { 80, 1, 256, 162, 0, 64, 128, 81, 0 },
        { 80, 2, 256, 162, 0, 64, 128, 81, 0 },
        { 96, 2, 256, 162, 0, 64, 128, 97, 0 },
        { 112, 2, 256, 162, 0, 64, 128, 113, 0 },
        { 64, 1, 256, 130, 0, 64, 128, 64, 0 },   /* B slab stride a multiple of 128 B */
        { 64, 1, 256, 130, 0, 64, 128, 72, 0 },
    };
    for (const Case &c : cases) {
        mb_kernel<<<148, 128, 200 * 1024>>>(c, 20, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(148); cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        long long mn = h[0], mx = h[0]; for (auto v : h) { if (v < mn) mn = v; if (v > mx) mx = v; }
        printf("N=%3d nacc=%d a_adv=%4d b_adv=%4d sw=%d b_lbo16=%3d mid2=%d : %6.1f .. %6.1f cycles/MMA (ideal %d)\n", c.N, c.nacc, c.a_adv16 * 16,
               c.b_adv16 * 16, c.sw, c.b_lbo16, c.twice_mid, (double)mn / c.nmma, (double)mx / c.nmma, 128 * c.N / 256);
    }
    return 0;
}
