/*
 * tc_selftest.cu -- unit kernel that pins down the tcgen05 kind::i8 conventions the tensor-core engine relies
 * on (K-major no-swizzle "slab" layout, 16-byte row shifts of the B operand, mixed operand signedness,
 * accumulation into one TMEM accumulator, int32 wrap-around, tcgen05.ld lane/column addressing).
 * Exposed for tests as gpuchan_tc_selftest(); not on the data path.
 */
#include "../../include/tslb200_gpuchan.h"
#include "tc_ptx.cuh"

#include <cuda_runtime.h>

using namespace tslb200;

namespace {

/* smem slab layout: slab j (16 bytes of K) x rows, 16 B per row:  offset = j * rows * 16 + r * 16 */
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const uint8_t *__restrict__ A0, const uint8_t *__restrict__ B0, const uint8_t *__restrict__ A1,
                   const uint8_t *__restrict__ B1, int Kp, int R, int N, int shift0, int shift1, int a0_signed,
                   int b0_signed, int a1_signed, int b1_signed, int *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int nslab = Kp / 16;
    uint8_t *sA0 = smem, *sA1 = sA0 + (size_t)Kp * 128;
    uint8_t *sB0 = sA1 + (size_t)Kp * 128, *sB1 = sB0 + (size_t)Kp * R;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 128 * Kp; i += 128) {
        const int r = i / Kp, k = i % Kp;
        sA0[(k / 16) * 128 * 16 + r * 16 + (k % 16)] = A0[i];
        sA1[(k / 16) * 128 * 16 + r * 16 + (k % 16)] = A1[i];
    }
    for (int i = tid; i < R * Kp; i += 128) {
        const int r = i / Kp, k = i % Kp;
        sB0[(k / 16) * R * 16 + r * 16 + (k % 16)] = B0[i];
        sB1[(k / 16) * R * 16 + r * 16 + (k % 16)] = B1[i];
    }
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_base_s, 64);
    ptx::fence_proxy_async();               /* generic-proxy smem writes -> visible to the tensor core (async proxy) */
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        const uint32_t id0 = ptx::idesc_i8(128, N, a0_signed, b0_signed);
        const uint32_t id1 = ptx::idesc_i8(128, N, a1_signed, b1_signed);
        for (int kk = 0; kk < Kp / 32; kk++) {
            const uint64_t da = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sA0) + kk * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t db = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sB0) + kk * 2 * R * 16 + shift0 * 16, R * 16, 128);
            ptx::mma_i8(tmem_base, da, db, id0, kk > 0);
        }
        for (int kk = 0; kk < Kp / 32; kk++) {
            const uint64_t da = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sA1) + kk * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t db = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sB1) + kk * 2 * R * 16 + shift1 * 16, R * 16, 128);
            ptx::mma_i8(tmem_base, da, db, id1, 1);
        }
        ptx::mma_commit(&bar);
    }
    (void)nslab;
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        int v[16];
        ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 16; i++) out[(size_t)tid * N + c0 + i] = v[i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

} // namespace

/* D[128][N] (int32) = A0 (128 x Kp bytes) * B0[shift0 .. shift0+N)^T + A1 * B1[shift1 ..)^T; rows of B are Kp bytes.
 * Signedness per operand; all pointers are HOST pointers. Returns 0 or a negative GPUCHAN_E_* code. */
extern "C" int gpuchan_tc_selftest(const uint8_t *A0, const uint8_t *B0, const uint8_t *A1, const uint8_t *B1, int Kp, int R,
                                   int N, int shift0, int shift1, int a0_signed, int b0_signed, int a1_signed, int b1_signed,
                                   int32_t *out)
{
    if (!A0 || !B0 || !A1 || !B1 || !out || Kp % 32 || N % 16 || N > 64 || shift0 + N > R || shift1 + N > R) return GPUCHAN_E_BADARGS;
    uint8_t *dA0 = nullptr, *dA1 = nullptr, *dB0 = nullptr, *dB1 = nullptr;
    int *dout = nullptr;
    const size_t a_bytes = (size_t)128 * Kp, b_bytes = (size_t)R * Kp;
    cudaError_t e = cudaSuccess;
#define ST(x) do { if (e == cudaSuccess) e = (x); } while (0)
    ST(cudaMalloc(&dA0, a_bytes)); ST(cudaMalloc(&dA1, a_bytes)); ST(cudaMalloc(&dB0, b_bytes)); ST(cudaMalloc(&dB1, b_bytes));
    ST(cudaMalloc(&dout, (size_t)128 * N * sizeof(int)));
    ST(cudaMemcpy(dA0, A0, a_bytes, cudaMemcpyHostToDevice)); ST(cudaMemcpy(dA1, A1, a_bytes, cudaMemcpyHostToDevice));
    ST(cudaMemcpy(dB0, B0, b_bytes, cudaMemcpyHostToDevice)); ST(cudaMemcpy(dB1, B1, b_bytes, cudaMemcpyHostToDevice));
    const size_t smem = 2 * a_bytes + 2 * b_bytes;
    ST(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (e == cudaSuccess) {
        tc_selftest_kernel<<<1, 128, smem>>>(dA0, dB0, dA1, dB1, Kp, R, N, shift0, shift1, a0_signed, b0_signed, a1_signed,
                                             b1_signed, dout);
        ST(cudaGetLastError());
        ST(cudaDeviceSynchronize());
    }
    ST(cudaMemcpy(out, dout, (size_t)128 * N * sizeof(int), cudaMemcpyDeviceToHost));
#undef ST
    cudaFree(dA0); cudaFree(dA1); cudaFree(dB0); cudaFree(dB1); cudaFree(dout);
    return e == cudaSuccess ? GPUCHAN_OK : GPUCHAN_E_CUDA;
}
