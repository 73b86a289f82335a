/*
 * tc_microbench2.cu -- diagnostics only (not part of the product library).  Three questions for the round-2 kernel:
 *   1. how fast does one SM retire tcgen05.mma kind::i8 M=128 with the A operand (tap image) in shared memory
 *      ("SS") against the same MMA with A resident in TMEM ("TS"), for the N the engine could use;
 *   2. does a 3-D TMA tensor map with byte strides (4D, 16) over the raw cs16 stream -- dims (16 B, block-row, K slab)
 *      -- encode, and does it deliver the K-major "slab" layout [slab][row][16 B] the MMA wants; cycles per tile;
 *   3. issue cost of fma.rn.f32x2 (FFMA2) against scalar FFMA in an issue-bound loop next to integer multiply-adds.
 *   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tsl-sdr_b200/csrc -o tools/tc_microbench2 tools/tc_microbench2.cu
 */
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace tslb200;

__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

struct Case { int N, ts, nacc, nmma, b_lbo16; };

__global__ void __launch_bounds__(128, 1) mma_kernel(Case c, int reps, long long *out)
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
    /* A chunks in TMEM: columns [384, 512), 16 chunks of 8 columns; every warp fills its 32 lanes */
    for (int ch = 0; ch < 16; ch++) {
        const uint32_t ta = tmem + 384 + 8 * ch + ((uint32_t)(32 * warp) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(ta), "r"(tid * 3 + ch), "r"(tid + 1), "r"(tid + 2), "r"(tid + 3), "r"(tid + 4), "r"(tid + 5), "r"(tid + 6), "r"(tid + 7) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (warp == 0) {
        const uint64_t dA0 = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(smem), 2048, 128);
        const uint64_t dB0 = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(smem) + 120 * 1024, c.b_lbo16 * 16, 128);
        const uint32_t id0 = ptx::idesc_i8(128, c.N, true, true);
        long long best = 1ll << 60;
        for (int r = 0; r < reps; r++) {
            __syncwarp();
            const long long t0 = clock64();
            if (ptx::elect_one()) {
                for (int i = 0; i < c.nmma; i++) {
                    const uint32_t acc = tmem + (uint32_t)(i % c.nacc) * c.N;
                    const uint64_t db = dB0 + (uint64_t)((uint32_t)(i & 7) * 2 * c.b_lbo16);
                    if (c.ts) mma_i8_ts(acc, tmem + 384 + 8 * (i & 15), db, id0, 1);
                    else ptx::mma_i8(acc, dA0 + (uint64_t)((uint32_t)(i & 15) * 256), db, id0, 1);
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

/* ---- 2: TMA slab load ---- */
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int D, int R, int nslab, int tiles,
                                                      long long *out, uint32_t *dump)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    __syncthreads();
    const uint32_t bytes = (uint32_t)R * nslab * 16;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int t = 0; t < tiles; t++) {
            const int row0 = (blockIdx.x * tiles + t) * 64;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(&bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(ptx::smem_u32(smem)), "l"(&map), "r"(0), "r"(row0), "r"(0), "r"(ptx::smem_u32(&bar)) : "memory");
            ptx::mbar_wait(&bar, t & 1);
        }
        t1 = clock64();
        out[blockIdx.x] = (t1 - t0) / tiles;
    }
    __syncthreads();
    if (blockIdx.x == 0 && dump)
        for (int i = tid; i < (int)(bytes / 4); i += 128) dump[i] = reinterpret_cast<uint32_t *>(smem)[i];
}

/* ---- 3: FFMA2 ---- */
template <int MODE>
__global__ void __launch_bounds__(512, 1) ffma_kernel(float *out, int iters, long long *cyc)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-4f;
    int k0 = threadIdx.x, k1 = k0 + 1, k2 = k0 + 2, k3 = k0 + 3;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        } else {
            unsigned long long p0, p1, p2, p3, mm, cc;
            asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
            asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
            asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
            asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
            asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
        }
        /* integer side work (ALU pipe), as in the epilogue */
        k0 = (k0 ^ (k1 >> 3)) + 7; k1 = (k1 ^ (k2 >> 5)) + 3; k2 = (k2 ^ (k3 >> 7)) + 1; k3 = (k3 ^ (k0 >> 2)) + 5;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(k0 + k1 + k2 + k3);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    long long *d; cudaMalloc(&d, 148 * 8);
    std::vector<long long> h(148);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const Case cases[] = {
        { 64, 0, 2, 64, 65 }, { 64, 1, 2, 64, 65 }, { 80, 0, 2, 64, 81 }, { 80, 1, 2, 64, 81 }, { 128, 0, 2, 64, 129 }, { 128, 1, 2, 64, 129 },
        { 256, 0, 1, 64, 257 }, { 256, 1, 1, 64, 257 }, { 64, 0, 1, 64, 65 }, { 64, 1, 1, 64, 65 }, { 64, 0, 4, 64, 65 }, { 64, 1, 4, 64, 65 },
    };
    for (const Case &c : cases) {
        mma_kernel<<<148, 128, 200 * 1024>>>(c, 20, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        long long mn = h[0], mx = h[0]; for (auto v : h) { if (v < mn) mn = v; if (v > mx) mx = v; }
        printf("MMA i8 M=128 N=%3d A=%s nacc=%d : %6.1f .. %6.1f cycles/MMA (floor %d)\n", c.N, c.ts ? "tmem" : "smem", c.nacc,
               (double)mn / c.nmma, (double)mx / c.nmma, 128 * c.N / 256);
    }

    /* ---- TMA ---- */
    {
        EncodeTiled enc = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr);
        const int D = 100, R = 65, nslab = (4 * D + 15) / 16;     /* raw interleaved bytes: 400 B per block-row = 25 slabs */
        const size_t rows = 1 << 18;
        int *src; cudaMalloc(&src, rows * D * 4 + 4096);
        std::vector<int> hs(rows * D);
        for (size_t i = 0; i < hs.size(); i++) hs[i] = (int)i;
        cudaMemcpy(src, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
        for (int variant = 0; variant < 2 && enc; variant++) {
            CUtensorMap map;
            /* variant 0: dims (4 words, rows, slabs) strides (4D, 16) -> smem [slab][row][16 B];  variant 1: plain 2-D rows */
            cuuint64_t dims3[3] = { 4, rows, (cuuint64_t)nslab }, str3[2] = { (cuuint64_t)4 * D, 16 };
            cuuint32_t box3[3] = { 4, (cuuint32_t)R, (cuuint32_t)nslab }, es[3] = { 1, 1, 1 };
            cuuint64_t dims2[2] = { (cuuint64_t)D, rows }, str2[1] = { (cuuint64_t)4 * D };
            cuuint32_t box2[2] = { (cuuint32_t)D, (cuuint32_t)R };
            CUresult rc = variant == 0
                ? enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, src, dims3, str3, box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                : enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, src, dims2, str2, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("TMA variant %d (%s): encode rc=%d\n", variant, variant == 0 ? "3-D slab order, strides (4D,16)" : "2-D row major", (int)rc);
            if (rc != CUDA_SUCCESS) continue;
            uint32_t *dump; cudaMalloc(&dump, 64 * 1024);
            cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (variant == 0) {
                tma_kernel<<<148, 128, 64 * 1024>>>(map, D, R, nslab, 24, d, dump);
                cudaError_t e = cudaDeviceSynchronize();
                printf("  launch: %s\n", cudaGetErrorString(e));
                if (e == cudaSuccess) {
                    cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
                    long long mn = h[0], mx = h[0]; for (auto v : h) { if (v < mn) mn = v; if (v > mx) mx = v; }
                    std::vector<uint32_t> hd(R * nslab * 4);
                    cudaMemcpy(hd.data(), dump, hd.size() * 4, cudaMemcpyDeviceToHost);
                    int bad = 0;
                    for (int j = 0; j < nslab; j++) for (int m = 0; m < R; m++) for (int w = 0; w < 4; w++)
                        if (hd[(j * R + m) * 4 + w] != (uint32_t)(m * D + 4 * j + w)) bad++;
                    printf("  slab layout check: %d mismatches of %d words; %lld .. %lld cycles per %d-byte tile (serialised, one in flight)\n",
                           bad, R * nslab * 4, mn, mx, R * nslab * 16);
                }
            }
            cudaFree(dump);
        }
        if (!enc) printf("TMA: cuTensorMapEncodeTiled not found\n");
    }

    /* ---- FFMA2 ---- */
    {
        float *o; cudaMalloc(&o, 148 * 512 * 4);
        for (int mode = 0; mode < 2; mode++) {
            if (mode == 0) ffma_kernel<0><<<148, 512>>>(o, 4096, d); else ffma_kernel<1><<<148, 512>>>(o, 4096, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("ffma error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
            printf("%s: %.2f cycles per iteration (8 FMA + 12 int ops per thread, 16 warps/SM)\n", mode ? "fma.rn.f32x2" : "fma.rn.f32   ", (double)h[0] / 4096);
        }
    }
    return 0;
}
