/*
 * tc_ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the tensor-core engine uses:
 * mbarrier, TMA bulk copy (cp.async.bulk -> SASS UBLKCP), L2 bulk prefetch (cp.async.bulk.prefetch.L2 -> SASS UBLKPF), TMEM allocation, tcgen05.mma kind::i8
 * (SASS UTCIMMA), tcgen05.commit, tcgen05.ld (SASS LDTM).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tslb200 {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

/* exactly one lane of the (converged) warp gets true; lets the compiler issue tcgen05 instructions straight from
 * uniform registers instead of looping over the active lanes */
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

/* ---- mbarrier ---- */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
/* for the single-lane producer / MMA roles: do not burn issue slots the epilogue warps need */
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 200000u)
{
    uint32_t ok = 0;
    while (!ok) {       /* try_wait with a suspend-time hint: the warp sleeps in hardware until the phase flips */
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
    }
}

/* for roles that run ahead of the pipeline: poll rarely, leave the issue slots to the warps doing arithmetic */
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

/* ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP): global -> shared, completion counted in bytes on an mbarrier ---- */
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* dst (shared) and src (global) 16-byte aligned, bytes a multiple of 16 */
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ask L2 to fetch [p, p + bytes) (16-byte aligned, multiple of 16) ahead of the loads that will want it */
__device__ __forceinline__ void prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

/* ---- TMEM ---- */
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)     /* whole warp, ncols = 2^k >= 32 */
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)       /* whole warp */
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

/* ---- descriptors ---- */
/* K-major, no swizzle ("interleave"): in 16-byte units the operand is ((8,n),2):((1,SBO),LBO):
 * 8 rows of a core matrix are 16 B apart, 8-row groups SBO apart, the two 16-byte K halves LBO apart. */
__device__ __forceinline__ uint64_t smem_desc_kmajor_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                 /* descriptor version for sm_100 */
    return d;                               /* base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE) */
}

/* kind::i8 instruction descriptor: S32 accumulate, K-major A and B, no saturation (wraps like int32) */
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, bool a_signed, bool b_signed)
{
    return (2u << 4)                                /* c_format = S32 */
         | ((a_signed ? 1u : 0u) << 7)              /* a_format: 0 = u8, 1 = s8 */
         | ((b_signed ? 1u : 0u) << 10)             /* b_format */
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

/* all previously issued tcgen05.mma of this thread arrive on the mbarrier when they complete */
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

/* TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns */
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
/* TMEM -> registers, shape 16x32bx2 with a half-split offset of 16 columns: threads 0-15 of the warp receive
 * columns [c, c + 8) of TMEM lanes L .. L + 15, threads 16-31 columns [c + 16, c + 24) of the SAME 16 lanes
 * (L = lane field of taddr: the warp's 32-lane slice, optionally + 16).  Measured with tools/tmem_probe.cu.
 * This hands every thread the columns it owns, for the real and for the imaginary rows, without any shuffle or
 * select. */
__device__ __forceinline__ void tmem_ld8_split16(uint32_t taddr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], 16;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1_split16(uint32_t taddr, int &v)
{
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x1.b32 {%0}, [%1], 16;" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

} // namespace ptx
} // namespace tslb200
