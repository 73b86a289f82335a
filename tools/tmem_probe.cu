/*
 * tmem_probe.cu -- diagnostics only: which (TMEM lane, column) does each thread of a warp receive from
 * tcgen05.ld.16x32bx2 as a function of the address' lane field and the half-split offset?
 * TMEM is filled with tcgen05.st.32x32b (cell value = lane * 1000 + column).
 *   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I tsl-sdr_b200/csrc -o tools/tmem_probe tools/tmem_probe.cu
 */
#include "tc_ptx.cuh"
#include <cstdio>
using namespace tslb200;

__global__ void __launch_bounds__(128, 1) probe(int *out)
{
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) ptx::tmem_alloc(&tmem_base_s, 64);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    /* fill: thread (warp, lane) owns TMEM lane 32*warp + lane; 64 columns */
    for (int c = 0; c < 64; c++) {
        const int v = (32 * warp + lane) * 1000 + c;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    /* case 0: lane field +0, columns 4.., split 24;  case 1: lane field +16, same */
    for (int cs = 0; cs < 2; cs++) {
        int v[8];
        const uint32_t a = tmem + lane_base + ((uint32_t)(16 * cs) << 16) + 4;
        asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], 24;"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(a) : "memory");
        ptx::tmem_ld_wait();
        for (int i = 0; i < 8; i++) out[((cs * 4 + warp) * 32 + lane) * 8 + i] = v[i];
    }
    {
        int v;
        const uint32_t a = tmem + lane_base + 7;
        asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x1.b32 {%0}, [%1], 8;" : "=r"(v) : "r"(a) : "memory");
        ptx::tmem_ld_wait();
        out[8192 + tid] = v;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

int main()
{
    int *d, h[8192 + 128];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0xff, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int cs = 0; cs < 2; cs++)
        for (int warp = 0; warp < 4; warp += 3)
            for (int lane = 0; lane < 32; lane += 1) {
                printf("case %d warp %d lane %2d:", cs, warp, lane);
                for (int i = 0; i < 8; i++) printf(" %6d", h[((cs * 4 + warp) * 32 + lane) * 8 + i]);
                printf("\n");
            }
    printf("x1 split 8, col 7:");
    for (int t = 0; t < 128; t++) printf(" %d", h[8192 + t]);
    printf("\n");
    return 0;
}
