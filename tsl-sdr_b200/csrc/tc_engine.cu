/*
 * tc_engine.cu -- exact tensor-core engine for the channel bank (GPUCHAN_ENGINE_TC).
 *
 * The per-channel complex FIR of filter/direct_fir.c:366-385, acc[k] = sum_i c[i] * x[kD+i] (int32, wraps), is a
 * Toeplitz contraction shared by all channels.  Writing i = qD + i' turns it into Q = ceil(T/D) small GEMMs over
 * the SAME block-row matrix X[m][:] = x[mD .. (m+1)D) taken at row offsets q:
 *
 *     acc[row, k] = sum_q  A_q[row, :] . X[k + q, :]
 *
 * rows = (channel, re|im): the re row holds (c_re, -c_im) interleaved, the im row (c_im, c_re), so that X is the
 * raw interleaved I,Q stream.  int16 x int16 is made exact on the int8 tensor cores by splitting both operands into
 * int8 pieces; the int32 partial accumulators are recombined modulo 2^32 in the epilogue, which is bit-identical
 * to the reference's wrapping int32 sum.  Samples are always split by radix (x = 256*hi + lo, hi signed, lo
 * unsigned).  Taps are split either
 *   SUM   (every |tap entry| <= 508, the usual narrow low-pass): v = v1 + v2 (+ v3 + v4) with every term in int8; the
 *         later terms are non-zero only around the centre of the filter, so only those K chunks issue further pairs
 *         of MMAs; two accumulators (weights 2^8 and 1) per tile, or
 *   RADIX (any int16 taps): v = 256*hi + lo, four products per K chunk into three accumulators (2^16, 2^8, 1).
 * Only the K chunks a block-row really covers are issued (the last block-row of the filter is usually short).
 *
 * Kernel tc_fir_fm_kernel: persistent, warp specialised (24 or 28 warps):
 *   warps 0-15  epilogue, two sets of 8: drain TMEM (LDTM, shape 16x32bx2: both components of a thread's own columns)
 *               straight into registers, recombine the limbs, and run the exact rq / derotator recurrence /
 *               discriminator (fm_math.cuh); every thread owns 16 consecutive outputs of one channel = two 16-byte PCM
 *               stores.  No shared-memory staging, no shuffles and no CTA-wide barriers: a tile's 16 lead-in columns
 *               make it self-contained.  With one channel group per CTA the sets take alternate tiles; with two groups
 *               per CTA (TcPlan::gpc = 2) set s owns group s and takes every tile.
 *   next 6 or 10 transform: read the raw cs16 tile from HBM/L2 and split it into two byte planes (hi s8 / lo u8) in
 *               "slab" order [16-byte K slab][block-row][16 B] (rows zero padded to Kp = round_up(2D, 32) bytes)
 *               directly in an NB-stage shared-memory ring, one warp group per stage;
 *   last 2      issue the tile's MMA program (tcgen05.mma kind::i8, SASS UTCIMMA) into an NT-stage TMEM ring, each
 *               warp the MMAs of its own limb accumulators.  With two groups per CTA every sample stage is used twice
 *               in a row (virtual tile 2 * tile + group: the other group's tap image, the next accumulator stage).
 *   The tap images (A operand, constant for the whole kernel) come in by one TMA bulk copy per CTA.
 * The B operand needs no im2col: with K-major / no-swizzle descriptors the Q row shifts are just +16 B on the
 * operand start address (validated by tc_selftest.cu).
 */
#include "tc_engine.cuh"
#include "tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace tslb200 {

namespace {

constexpr int EPI_WARPS = 16;
/* Transform warps (raw cs16 -> byte planes in smem) are a kernel template parameter XF.  They form one group per sample
 * stage (warp w -> group w % NB); a group fills its stage on its own, so NB tiles' loads are in flight at any time.  One
 * group per stage also keeps every barrier wait at most one phase behind (a parity wait cannot tell phases two apart).
 * Measured on B200 (profiles/r02_xf_warps_experiment.txt): when a transformed tile serves two channel groups (TcPlan::gpc
 * = 2) 6 warps are enough and leave the epilogue 80 registers (256 channels x 127 taps: 4 / 5 / 6 / 8 warps 0.2548 / 0.2520 /
 * 0.2493 / 0.2559 ms); with one group per CTA more warps are faster (8 against 6: 64 channels 0.088 against 0.090 ms; 1024
 * channels x 255 taps, D = 200: 0.822 against 0.861; 256 x 512 taps: 0.348 against 0.367) and 10 warps -- 28 in the CTA, still
 * 72 registers per thread -- faster again (64 channels 0.0828 against 0.0868, 256 x 512 taps 0.3483 against 0.3530); 5 warps
 * in groups of 2, 2, 1 lose 25 %. */
constexpr int XF_PAIRED = 6, XF_ONE_GROUP = 10;
/* Warp roles, lowest warp index first: epilogue | transform | MMA issuer. */
constexpr int EPI_WARP0 = 0;            /* must be a multiple of 4: warp w may only read TMEM lanes 32*(w%4).. */
constexpr int XF_WARP0 = EPI_WARPS;     /* first transform warp */
constexpr int MMA_WARPS = 2;            /* MMA issuers (2 = each owns a disjoint set of limb accumulators) */
__host__ __device__ constexpr int tc_threads(int xf) { return 32 * (xf + MMA_WARPS + EPI_WARPS); }
constexpr int NB_MAX = 4, NT_MAX = 3;
/* TC_DIAG (diagnostics builds only, `make diag`; results are wrong by design): bit 0 = no epilogue arithmetic (the PCM
 * stores write the raw accumulator words), bit 1 = no transform (the MMAs read whatever the sample stages hold), bit 2 = one
 * MMA per issuing warp and tile.  Timing such builds attributes the tile time to the three roles. */
#ifndef TC_DIAG
#define TC_DIAG 0
#endif
/* nanoseconds between polls of a role's barrier wait (measured on B200: shorter intervals spend issue slots the epilogue
 * needs, longer ones add latency to the hand-offs; hardware-suspended try_wait was no faster) */
#ifndef TC_SLEEP_EPI
#define TC_SLEEP_EPI 200
#define TC_SLEEP_XF 1000
#define TC_SLEEP_MMA 100
#endif
constexpr uint32_t SLEEP_EPI = TC_SLEEP_EPI, SLEEP_XF = TC_SLEEP_XF, SLEEP_MMA = TC_SLEEP_MMA;

/* ---------------------------------------------------------------------------------------------- */
struct TcKernelParams {
    InWindow in;            /* raw interleaved int16 I,Q stream of this submit: [carry | fresh] */
    int D;
    const uint8_t *tap_img;
    const int *incr, *ckpt, *last_in;
    const uint32_t *mu, *lambda;
    const int *cyc;
    int cyc_pitch;
    unsigned long long k_base;
    int *carry_out;
    long long carry_from;
    int carry_keep;
    int *last_out;
    const float2 *atan_tab;
    short *pcm;
    int *iq_out;
    long long pitch;
    long long K;            /* outputs per channel of this submit */
    int n_tiles;            /* tiles per CTA */
    int total_tiles;
    int C, G, Kp, Q, R;
    int gpc;                /* channel groups per CTA: 1, or 2 sharing every transformed sample tile (one epilogue set each) */
    int nb_stages, prog_len, prog_split, prog_regular;
    int atan_copies;        /* interleaved copies of the arctangent table in shared memory: 16 or 1 */
    uint32_t a_group_bytes, b_stage_bytes;
    AtanParams atan;
    TcMma prog[TC_PROG_MAX];
};

/* 8 consecutive raw samples starting at a (4-byte aligned, inside the fresh buffer with >= 12 samples of slack):
 * three aligned 16-byte loads, then a word rotation by o = (address / 4) mod 4. */
__device__ __forceinline__ void load8_unaligned(const int *a, uint32_t (&w)[8])
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a);
    const uint4 *al = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
    const uint32_t o = (uint32_t)(addr >> 2) & 3u;
    const uint4 v0 = __ldg(al), v1 = __ldg(al + 1);
    uint4 v2 = make_uint4(0, 0, 0, 0);
    if (o) v2 = __ldg(al + 2);
    const bool o2 = (o & 2u) != 0, o1 = (o & 1u) != 0;
    /* stage 1: drop 2 words if o2 */
    const uint32_t t0 = o2 ? v0.z : v0.x, t1 = o2 ? v0.w : v0.y, t2 = o2 ? v1.x : v0.z, t3 = o2 ? v1.y : v0.w,
                   t4 = o2 ? v1.z : v1.x, t5 = o2 ? v1.w : v1.y, t6 = o2 ? v2.x : v1.z, t7 = o2 ? v2.y : v1.w,
                   t8 = o2 ? v2.z : v2.x;
    /* stage 2: drop 1 word if o1 */
    w[0] = o1 ? t1 : t0; w[1] = o1 ? t2 : t1; w[2] = o1 ? t3 : t2; w[3] = o1 ? t4 : t3;
    w[4] = o1 ? t5 : t4; w[5] = o1 ? t6 : t5; w[6] = o1 ? t7 : t6; w[7] = o1 ? t8 : t7;
}

/* packed sample = bytes (lo(re), hi(re), lo(im), hi(im)): split 8 samples into a hi and a lo 16-byte slab entry */
__device__ __forceinline__ void split_store(const uint32_t (&w)[8], uint8_t *hi_dst, uint8_t *lo_dst)
{
    uint4 lo, hi;
    lo.x = __byte_perm(w[0], w[1], 0x6420); hi.x = __byte_perm(w[0], w[1], 0x7531);
    lo.y = __byte_perm(w[2], w[3], 0x6420); hi.y = __byte_perm(w[2], w[3], 0x7531);
    lo.z = __byte_perm(w[4], w[5], 0x6420); hi.z = __byte_perm(w[4], w[5], 0x7531);
    lo.w = __byte_perm(w[6], w[7], 0x6420); hi.w = __byte_perm(w[6], w[7], 0x7531);
    *reinterpret_cast<uint4 *>(hi_dst) = hi;
    *reinterpret_cast<uint4 *>(lo_dst) = lo;
}


/* ATAN16: 16 interleaved copies of the arctangent table in shared memory (else one) */
template <int MODE, bool KEEP_IQ, bool FMA, bool ATAN16, int XF_WARPS>
__global__ void __launch_bounds__(tc_threads(XF_WARPS), 1) tc_fir_fm_kernel(const __grid_constant__ TcKernelParams p)
{
    constexpr int XF_THREADS = 32 * XF_WARPS;
    constexpr int MMA_WARP = EPI_WARPS + XF_WARPS;              /* warp index of the first MMA issuer */
    constexpr int TC_THREADS = tc_threads(XF_WARPS);
    constexpr int ACCS = (MODE == TC_MODE_SUM) ? 2 : 3;
    constexpr int NT = 512 / (ACCS * TC_ACC_STRIDE);            /* TMEM stages: 3 (SUM) or 2 (RADIX) */
    constexpr uint32_t STAGE_COLS = ACCS * TC_ACC_STRIDE;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t b_full[NB_MAX], b_empty[NB_MAX], t_full[NT_MAX], t_empty[NT_MAX], a_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ TcMma prog_s[TC_PROG_MAX];

    uint8_t *sA = smem;                                         /* [a_chunks][2 slabs][128][16] */
    uint8_t *sB = smem + (size_t)p.gpc * p.a_group_bytes;       /* [NB stages][2 planes][nslab][R][16] */
    /* arctangent table, entry-major with atan_copies (16 or 1) interleaved copies: entry i of copy c sits at
     * (i * copies + c) * 8 bytes, so that with 16 copies lane l (copy l & 15) always reads bank pair l & 15 -- every
     * table load is two conflict-free wavefronts.  (One shared copy cost 12.8 wavefronts per load on average: the
     * look-ups alone kept the shared-memory data pipe 30 % busy next to the tensor core's operand reads.) */
    float2 *sT = reinterpret_cast<float2 *>(sB + (size_t)p.nb_stages * p.b_stage_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);       /* same value, but provably warp-uniform for the compiler */
    const int nslab = p.Kp >> 4;
    const int NB = p.nb_stages;
    const int gpc = p.gpc;
    const int cta_groups = p.G / gpc;                           /* CTAs per tile range */
    const int g = (blockIdx.x % cta_groups) * gpc;              /* (first) channel group of this CTA */
    const int chunk = blockIdx.x / cta_groups;
    const int tile0 = chunk * p.n_tiles;                        /* first tile of this CTA */
    int my_tiles = p.total_tiles - tile0;
    if (my_tiles > p.n_tiles) my_tiles = p.n_tiles;
    if (my_tiles < 0) my_tiles = 0;

    /* L2 prefetch of the TC_OUT new block-rows of a tile (the stream is read from HBM exactly once) */
    auto prefetch_tile = [&](int t, bool with_lead_in = false) {
        if (t >= tile0 + my_tiles) return;
        long long s_a = ((long long)TC_OUT * t + (p.Q - 1)) * (long long)p.D - p.in.carry_len;
        long long s_b = s_a + (long long)TC_OUT * p.D;
        const long long n_fresh = p.in.total - p.in.carry_len;
        if (with_lead_in) { s_a -= (long long)(TC_LEAD + p.Q - 1) * p.D; if (s_a < 0) s_a = 0; }   /* rows the previous tile would have brought in */
        if (s_a < 0 || s_a >= n_fresh) return;
        if (s_b > n_fresh) s_b = n_fresh;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p.in.fresh + s_a) & ~(uintptr_t)15;
        const uintptr_t b = reinterpret_cast<uintptr_t>(p.in.fresh + s_b) & ~(uintptr_t)15;
        if (b > a) ptx::prefetch_l2(reinterpret_cast<const void *>(a), (uint32_t)(b - a));
    };
    /* the first tiles of this CTA are read cold: start their HBM fetches before anything else */
    if (tid < NB) prefetch_tile(tile0 + tid, tid == 0);

    /* ---- one-time setup ---- */
    if (tid == 0) {
        for (int s = 0; s < NB_MAX; s++) { ptx::mbar_init(&b_full[s], (uint32_t)((XF_WARPS - s + p.nb_stages - 1) / p.nb_stages)); ptx::mbar_init(&b_empty[s], (uint32_t)(MMA_WARPS * p.gpc)); }
        for (int s = 0; s < NT_MAX; s++) { ptx::mbar_init(&t_full[s], MMA_WARPS); ptx::mbar_init(&t_empty[s], EPI_WARPS / 2); }
        ptx::mbar_init(&a_full, 1);
        ptx::fence_mbar_init();
        /* this group's tap image (the A operand of every MMA of the kernel, 36 - 140 KB) comes in by TMA: one bulk copy
         * straight into the operand's shared-memory layout, landing while the other threads fill the tables below; the MMA
         * warps wait for it before their first instruction */
        ptx::mbar_expect_tx(&a_full, (uint32_t)gpc * p.a_group_bytes);
        ptx::bulk_g2s(sA, p.tap_img + (size_t)g * p.a_group_bytes, (uint32_t)gpc * p.a_group_bytes, &a_full);
    }
    {
        for (int i = tid; i < 256 * p.atan_copies; i += TC_THREADS) sT[i] = p.atan_tab[i / p.atan_copies];
        for (int i = tid; i < p.prog_len; i += TC_THREADS) prog_s[i] = p.prog[i];
    }
    if (warp_u == MMA_WARP) ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp_u >= XF_WARP0 && warp_u < XF_WARP0 + XF_WARPS) {
        const int xall = tid - 32 * XF_WARP0;       /* 0 .. XF_THREADS-1 */
        const int xw = warp_u - XF_WARP0;
        const int grp = xw % NB;                    /* this warp's group = the stage it fills: tiles grp, grp + NB, ... of the CTA */
        const int XG_THREADS = 32 * ((XF_WARPS - grp + NB - 1) / NB);      /* threads in my group */
        const int xt = 32 * (xw / NB) + lane;       /* 0 .. XG_THREADS-1 inside the group */
        /* ================= transform: raw cs16 samples -> hi/lo byte planes of the smem ring =================
         * Plane row m of tile t = stream samples [(TC_OUT*t - TC_LEAD + m) * D, +D); item (m, j) is one 16-byte slab
         * entry = 8 complex samples = 32 raw bytes.  Consecutive threads take consecutive j: 32-byte pieces of one
         * contiguous run, so the global reads coalesce.  Entries past D in a row multiply zero taps, so the fast path
         * does not mask them.
         * A tile costs a handful of dependent global-load round trips whatever the number of threads on it (measured:
         * ~4000 cycles with all 8 warps on one tile, the whole kernel waiting on it), so the 8 warps form one group
         * per stage that fills it on its own: NB tiles' loads are in flight at any time. */
        const int items = p.R * nslab;
        const uint32_t slab_bytes = (uint32_t)p.R * 16;
        if (xt == 0) prefetch_tile(tile0 + grp + NB);
        for (int it = grp; it < my_tiles; it += NB) {
            const int s = grp, ph = (it / NB) & 1;
            if (xt == 0) prefetch_tile(tile0 + it + 2 * NB);
            ptx::mbar_wait_backoff(&b_empty[s], ph ^ 1, SLEEP_XF);
            uint8_t *dst = sB + (size_t)s * p.b_stage_bytes;
            const long long row_base = (long long)TC_OUT * (tile0 + it) - TC_LEAD;
            const long long s_first = row_base * (long long)p.D;
            const long long s_last = s_first + (long long)(p.R - 1) * p.D + 8 * nslab + 12;     /* one past the furthest word read */
            if (TC_DIAG & 2) {
            } else if (s_first >= p.in.carry_len + 4 && s_last <= p.in.total) {
                /* a thread keeps one slab column j and walks down the rows: addresses advance by constants, and the
                 * loads of consecutive rows are independent, so several stay in flight */
                const int *base = p.in.fresh + (s_first - p.in.carry_len);
                const uint32_t o_tile = (uint32_t)(reinterpret_cast<uintptr_t>(base) >> 2) & 3u;
                const int row_lanes = XG_THREADS / nslab;
                const int j = xt % nslab, rl = xt / nslab;
                uint8_t *d_hi = dst + (size_t)j * slab_bytes + rl * 16;
                const size_t lo_off = (size_t)nslab * slab_bytes;
                if (rl >= row_lanes) {
                } else if ((p.D & 3) == 0) {
                    /* rows start a multiple of 4 samples apart: every 8-sample item of the tile has the same
                     * misalignment o_tile against 16 bytes, so the word rotation is resolved at compile time */
                    const uint4 *al = reinterpret_cast<const uint4 *>(reinterpret_cast<uintptr_t>(base) & ~(uintptr_t)15) + ((rl * p.D) >> 2) + 2 * j;
                    const int al_step = (row_lanes * p.D) >> 2;
                    auto run = [&](auto o_tag) {
                        constexpr int O = decltype(o_tag)::value;
#pragma unroll 4
                        for (int m = rl; m < p.R; m += row_lanes) {
                            const uint4 v0 = __ldg(al), v1 = __ldg(al + 1);
                            uint4 v2 = make_uint4(0, 0, 0, 0);
                            if (O) v2 = __ldg(al + 2);
                            const uint32_t c[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
                            uint32_t w[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) w[u] = c[O + u];
                            split_store(w, d_hi, d_hi + lo_off);
                            al += al_step;
                            d_hi += row_lanes * 16;
                        }
                    };
                    if (o_tile == 0) run(std::integral_constant<int, 0>{});
                    else if (o_tile == 1) run(std::integral_constant<int, 1>{});
                    else if (o_tile == 2) run(std::integral_constant<int, 2>{});
                    else run(std::integral_constant<int, 3>{});
                } else {
                    const int *src = base + rl * p.D + 8 * j;
#pragma unroll 4
                    for (int m = rl; m < p.R; m += row_lanes) {
                        uint32_t w[8];
                        load8_unaligned(src, w);
                        split_store(w, d_hi, d_hi + lo_off);
                        src += row_lanes * p.D;
                        d_hi += row_lanes * 16;
                    }
                }
            } else {
                for (int item = xt; item < items; item += XG_THREADS) {
                    const int m = item / nslab, j = item - m * nslab;
                    const long long s0 = (row_base + m) * (long long)p.D + 8 * j;
                    uint32_t w[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) w[u] = (8 * j + u < p.D) ? (uint32_t)in_sample(p.in, s0 + u) : 0u;
                    split_store(w, dst + (size_t)j * slab_bytes + m * 16, dst + (size_t)(nslab + j) * slab_bytes + m * 16);
                }
            }
            ptx::fence_proxy_async();       /* generic-proxy stores -> visible to the tensor core's async proxy */
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&b_full[s]);
        }
        /* keep the tail of the window the next submit still needs (fewer than T samples) */
        if (blockIdx.x == 0 && p.carry_out)
            for (int i = xall; i < p.carry_keep; i += XF_THREADS) p.carry_out[i] = in_sample(p.in, p.carry_from + i);
    } else if (warp_u >= MMA_WARP) {
        /* ================= MMA issuers =================
         * Two warps, each issuing the MMAs of its own accumulators (SUM: the 2^8 / the 2^0 limb; RADIX: limbs
         * {2^16, 2^0} / {2^8}), so no ordering is needed between them; both commit to the same barriers.
         * Within a warp the first MMA into an accumulator (accumulate = 0) goes out before the others; beyond that
         * the order of accumulation does not matter. */
        const int mw = warp_u - MMA_WARP;
        const int i0 = (MMA_WARPS == 1 || mw == 0) ? 0 : p.prog_split, i1 = (MMA_WARPS == 1 || mw != 0) ? p.prog_len : p.prog_split;
        const int n_mine = i1 - i0;
        const uint32_t a_base = ptx::smem_u32(sA) >> 4, b_base0 = ptx::smem_u32(sB) >> 4;
        constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);      /* SBO = 128 B, descriptor version 1 (bit 46) */
        /* NFIX > 0: the warp's whole program (at most NFIX MMAs) stays in registers across the tile loop -- read once
         * from the parameter block with warp-uniform indices, so ptxas keeps it on the uniform datapath and an MMA
         * costs two uniform adds and the UTCIMMA itself.  With everything else compiled out the tile pipeline ran
         * at 3000 cycles per tile on the lane loop (183 cycles per MMA, mostly R2UR latency) against 55 cycles of
         * tensor-pipe time per MMA.  NFIX = 0: longer programs keep the lane loop. */
        auto mma_role = [&](auto nfix_tag) {
            constexpr int NFIX = decltype(nfix_tag)::value;
            constexpr int NARR = NFIX > 0 ? NFIX : 1;
            /* per MMA only the two operand offsets differ; accumulator and instruction descriptor alternate between
             * the values of entry 0 (even entries) and entry 1 (odd entries) -- checked by tc_make_plan */
            uint32_t fa[NARR], fb[NARR];
            uint32_t d_even = 0, d_odd = 0, i_even = 0, i_odd = 0;
            if (NFIX > 0) {
#pragma unroll
                for (int i = 0; i < NARR; i++) {
                    const TcMma m = p.prog[i < n_mine ? i0 + i : i0];
                    fa[i] = m.a_lo + a_base; fb[i] = m.b_lo;
                }
                const TcMma m0 = p.prog[i0], m1 = p.prog[n_mine > 1 ? i0 + 1 : i0];
                d_even = m0.d_acc & 0xffffu; i_even = m0.idesc; d_odd = m1.d_acc & 0xffffu; i_odd = m1.idesc;
            }
            const bool odd_is_new_acc = d_odd != d_even;
            const int passes = (n_mine + 31) >> 5;
            TcMma m0 = { 0, 0, 0, 0 };
            const bool have0 = i0 + lane < i1;
            if (NFIX == 0 && have0) m0 = prog_s[i0 + lane];
            int sb = 0, phb = 0, st = 0, pht = 0;
            ptx::mbar_wait(&a_full, 0);                 /* the tap image has landed (also keeps the CTA alive until it has) */
            /* With two channel groups per CTA every sample stage is used twice in a row -- "virtual tile" v = 2 * tile + group:
             * same B operand, the other group's tap image, the next accumulator stage -- and released after the second use. */
            const uint32_t a_group16 = p.a_group_bytes >> 4;
            const int n_virtual = my_tiles * gpc;
            for (int v = 0; v < n_virtual; v++) {
                const bool second = gpc == 2 && (v & 1);
                ptx::mbar_wait_backoff(&b_full[sb], phb, SLEEP_MMA);
                ptx::mbar_wait_backoff(&t_empty[st], pht ^ 1, SLEEP_MMA);
                ptx::tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)st * STAGE_COLS;
                const uint32_t b_base = b_base0 + (((uint32_t)sb * p.b_stage_bytes) >> 4);
                const uint32_t a_off = second ? a_group16 : 0u;
                if (NFIX > 0) {
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int i = 0; i < NARR; i++) {
                            if (i < ((TC_DIAG & 4) ? 1 : n_mine))
                                ptx::mma_i8(acc + ((i & 1) ? d_odd : d_even), ((uint64_t)DESC_HI << 32) | (uint64_t)(fa[i] + a_off),
                                            ((uint64_t)DESC_HI << 32) | (uint64_t)(fb[i] + b_base), (i & 1) ? i_odd : i_even,
                                            (i == 0 || (i == 1 && odd_is_new_acc)) ? 0u : 1u);
                        }
                    }
                    __syncwarp();
                } else {
                    /* lane-resident program: lane l of a pass owns one MMA and issues it itself (ptxas emits an
                     * ELECT / R2UR.BROADCAST / UTCIMMA loop over the active lanes).  The first MMA into an
                     * accumulator (accumulate = 0) goes out before the others. */
                    for (int ps = 0; ps < passes; ps++) {
                        TcMma m = m0;
                        bool have = have0;
                        if (ps > 0) {                       /* more than 32 MMAs per warp and tile: fetch the next 32 */
                            have = i0 + 32 * ps + lane < i1;
                            if (have) m = prog_s[i0 + 32 * ps + lane];
                        }
                        const uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(m.a_lo + a_base + a_off);
                        const uint64_t db = ((uint64_t)DESC_HI << 32) | (uint64_t)(m.b_lo + b_base);
                        const uint32_t d = acc + (m.d_acc & 0xffffu);
                        const bool first = (m.d_acc >> 31) == 0;
                        if (have && first) ptx::mma_i8(d, da, db, m.idesc, 0u);
                        __syncwarp();
                        if (have && !first) ptx::mma_i8(d, da, db, m.idesc, 1u);
                        __syncwarp();
                    }
                }
                if (ptx::elect_one()) {
                    ptx::mma_commit(&b_empty[sb]);      /* smem stage may be refilled once these MMAs have read it */
                    ptx::mma_commit(&t_full[st]);       /* accumulators complete */
                }
                __syncwarp();
                if (gpc == 1 || second) { if (++sb == NB) { sb = 0; phb ^= 1; } }
                if (++st == NT) { st = 0; pht ^= 1; }
            }
        };
        if (!p.prog_regular) mma_role(std::integral_constant<int, 0>{});
        else if (n_mine <= 12) mma_role(std::integral_constant<int, 12>{});
        else if (n_mine <= 24) mma_role(std::integral_constant<int, 24>{});
        else mma_role(std::integral_constant<int, 0>{});
    } else {
        /* ================= epilogue: TMEM -> registers -> derotate -> discriminate -> PCM =================
         * Accumulator row 32*s + i (i < 16) is the real part of channel 16*s + i, row 32*s + 16 + i its imaginary
         * part.  The 16 warps form two SETS of 8 that take alternate tiles, so one set's barrier wait and TMEM drain
         * overlap the other set's arithmetic (with all 16 warps in lock step on one tile the issue ports idled
         * through every drain).  Inside a set, warp (s, half) turns columns TC_LEAD + 32*half + [0, 32) of its 16
         * channels into PCM: lanes 0-15 the first 16 columns, lanes 16-31 the last 16, as two consecutive 8-output
         * blocks.  The 16x32bx2 load shape hands lanes 0-15 and 16-31 different columns of the SAME 16 TMEM lanes,
         * once for the real rows and once for the imaginary rows, so every thread receives both components of its
         * own channel x 8 outputs -- no shuffles, no selects.
         * The arithmetic (fm_math.cuh "v2") is arranged for the pipe split of sm_100: ncu showed the first
         * epilogue bound by the ALU pipe (67 % busy, FMA pipe 24 %), so selects/compares became multiply-adds. */
        const int e = warp_u - EPI_WARP0;           /* warp_u: provably warp-uniform, so the TMEM addresses stay in uniform registers */
        const int set = e >> 3;
        const int slice = warp_u & 3;
        const int half = (e >> 2) & 1;
        const bool hi = lane >= 16;
        const int ch = 16 * slice + (lane & 15);
        const int blk0 = 4 * half + (hi ? 2 : 0);   /* first of this thread's two 8-output blocks of a tile */
        const uint32_t lane_re = (uint32_t)(32 * slice) << 16, lane_im = (uint32_t)(32 * slice + 16) << 16;
        /* one group per CTA: the two sets take alternate tiles; two groups: set s owns group g + s and takes every tile */
        const int t_first = gpc == 2 ? 0 : set, t_step = gpc == 2 ? 1 : 2;
        const int c = (g + (gpc == 2 ? set : 0)) * TC_CH + ch;
        const bool live = c < p.C;
        const int K32 = (int)p.K;                   /* outputs per channel of this submit (tc_launch_fir_fm checks the range) */
        const int iw = live ? __ldg(p.incr + c) : 0;
        const int i4_re = 4 * lo16(iw), i4_im = 4 * hi16(iw), ni4_im = -i4_im;
        short *const pcm_c = p.pcm + (size_t)c * p.pitch;
        int *const iq_c = KEEP_IQ ? p.iq_out + (size_t)c * p.pitch : nullptr;
        const float z_thr = p.atan.z_small_thr;
        constexpr int ATAN_SHIFT = ATAN16 ? 7 : 3;                  /* bytes between entries of one copy: 8 * copies */
        const uint32_t atan_smem = ptx::smem_u32(sT) + 8u * ((uint32_t)lane & (uint32_t)(p.atan_copies - 1)) - (0x4B000000u << ATAN_SHIFT);
        /* Steady state (every channel on its limit cycle): the phase of the output before my first one comes from the
         * channel's cycle table; it advances by t_step * TC_OUT outputs from one of my tiles to the next. */
        const bool table_mode = p.ckpt == nullptr;
        uint32_t lam = 1, tph = 0, tstep = 0;
        const int *tab = nullptr;
        if (table_mode && live) {
            lam = __ldg(p.lambda + c);
            const unsigned long long g0 = p.k_base + (unsigned long long)TC_OUT * (tile0 + t_first) + 8 * blk0 - 1 - __ldg(p.mu + c);
            tph = (uint32_t)(g0 % lam);
            tstep = (uint32_t)(t_step * TC_OUT) % lam;
            tab = p.cyc + (size_t)c * p.cyc_pitch;
        }
        /* derotator phase word of tile `it` of this CTA: the phase of the output before my first block (of output 0
         * itself for the very first block of the stream segment); fetched one tile ahead so that its latency hides
         * behind the arithmetic of the current tile */
        auto phase_word = [&](int it) -> int {
            if (!live) return 0;
            const int tile = tile0 + it;
            if (!table_mode && TC_OUT * tile + 8 * blk0 >= K32) return 0;   /* block past the last output: its
                                                                                          checkpoint was never written */
            if (table_mode) {
                uint32_t ix = tph;
                tph += tstep;
                if (tph >= lam) tph -= lam;
                if (tile == 0 && blk0 == 0) { ix++; if (ix == lam) ix = 0; }
                return __ldg(tab + ix);
            }
            return __ldg(p.ckpt + ((size_t)tile * TC_SUB + blk0) * p.C + c);
        };
        int cwk_next = t_first < my_tiles ? phase_word(t_first) : 0;

        for (int it = t_first; it < my_tiles; it += t_step) {
            const int tile = tile0 + it;
            const int v = gpc == 2 ? 2 * it + set : it;         /* virtual tile: which accumulator stage holds it */
            const int st = v % NT, pht = (v / NT) & 1;
            const int cwk = cwk_next;
            const bool tile_inside = TC_OUT * (tile + 1) < K32;
            ptx::mbar_wait_backoff(&t_full[st], pht, SLEEP_EPI);
            ptx::tc_fence_after();
            const uint32_t col0 = tmem_base + (uint32_t)st * STAGE_COLS + TC_LEAD + 32 * half;
            /* limb weights: SUM (2^8, 1), RADIX (2^16, 2^8, 1); times 4 and + 0x8000 for the rounding shift */
            auto comb = [](int a0, int a1, int a2) -> int {
                if (ACCS == 3) return (int)((unsigned)a2 * 4u + ((unsigned)a1 * 1024u + ((unsigned)a0 * 262144u + 0x8000u)));
                return (int)((unsigned)a1 * 4u + ((unsigned)a0 * 1024u + 0x8000u));
            };
            /* 8 columns starting at `col` (+16 for lanes 16-31), both components, every limb accumulator:
             * x = 4 * acc + 0x8000 (mod 2^32), top 16 bits = rq14(acc) */
            auto drain8 = [&](uint32_t col, int (&x_re)[8], int (&x_im)[8]) {
                int a0r[8], a1r[8], a2r[8], a0i[8], a1i[8], a2i[8];
                ptx::tmem_ld8_split16(col + lane_re, a0r);
                ptx::tmem_ld8_split16(col + lane_im, a0i);
                ptx::tmem_ld8_split16(col + lane_re + TC_ACC_STRIDE, a1r);
                ptx::tmem_ld8_split16(col + lane_im + TC_ACC_STRIDE, a1i);
                if (ACCS == 3) {
                    ptx::tmem_ld8_split16(col + lane_re + 2 * TC_ACC_STRIDE, a2r);
                    ptx::tmem_ld8_split16(col + lane_im + 2 * TC_ACC_STRIDE, a2i);
                }
                ptx::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    x_re[u] = comb(a0r[u], a1r[u], ACCS == 3 ? a2r[u] : 0);
                    x_im[u] = comb(a0i[u], a1i[u], ACCS == 3 ? a2i[u] : 0);
                }
            };

            int r_re = lo16(cwk), r_im = hi16(cwk);
            int p_re = 0, p_im = 0;
#pragma unroll 1
            for (int b = 0; b < 2; b++) {
                int x_re[8], x_im[8];
                const int kfirst = TC_OUT * tile + 8 * (blk0 + b);     /* output index of the block's first column (a submit has < 2^31 outputs) */
                if (b == 0) {
                    /* the column before the first block: the discriminator's previous sample */
                    int l0r, l1r, l2r = 0, l0i, l1i, l2i = 0;
                    ptx::tmem_ld1_split16(col0 - 1 + lane_re, l0r);
                    ptx::tmem_ld1_split16(col0 - 1 + lane_im, l0i);
                    ptx::tmem_ld1_split16(col0 - 1 + lane_re + TC_ACC_STRIDE, l1r);
                    ptx::tmem_ld1_split16(col0 - 1 + lane_im + TC_ACC_STRIDE, l1i);
                    if (ACCS == 3) {
                        ptx::tmem_ld1_split16(col0 - 1 + lane_re + 2 * TC_ACC_STRIDE, l2r);
                        ptx::tmem_ld1_split16(col0 - 1 + lane_im + 2 * TC_ACC_STRIDE, l2i);
                    }
                    drain8(col0, x_re, x_im);
                    if (it + t_step < my_tiles) cwk_next = phase_word(it + t_step);
                    if (kfirst == 0) {
                        /* very first output of the submit: y[-1] is carried state; the phase word = phase of output 0 */
                        const int lw = live ? __ldg(p.last_in + c) : 0;
                        p_re = lo16(lw); p_im = hi16(lw);
                    } else {
                        /* previous output = the column before my block; the phase word is its phase */
                        derotate_v2(comb(l0r, l1r, l2r) >> 16, comb(l0i, l1i, l2i) >> 16, r_re, r_im, p_re, p_im);
                        rot_step_v2(r_re, r_im, i4_re, i4_im, ni4_im);
                    }
                } else {
                    drain8(col0 + 8, x_re, x_im);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&t_empty[st]);  /* TMEM stage is free again */
                }

                /* ---- one channel x 8 consecutive outputs ---- */
                /* tile_inside (warp-uniform): every block of this tile is complete and none holds the submit's last output --
                 * all tiles but the last one or two skip the per-thread edge bookkeeping */
                const int nvalid = tile_inside ? 8 : ((K32 - kfirst > 8) ? 8 : K32 - kfirst);
                if (TC_DIAG & 1) {
                    if (live && kfirst + 8 < K32) {
                        int acc = 0;
#pragma unroll
                        for (int u = 0; u < 8; u++) acc ^= x_re[u] + x_im[u];
                        *reinterpret_cast<uint4 *>(pcm_c + kfirst) = make_uint4((uint32_t)acc, (uint32_t)x_re[0], (uint32_t)x_im[1], (uint32_t)r_re);
                    }
                } else if (live && nvalid > 0) {
                    /* EDGE = this block holds the submit's last output (or is cut short by it): also track y[K-1] */
                    auto block8 = [&](auto edge_tag) {
                        constexpr bool EDGE = decltype(edge_tag)::value;
                        int l_re = 0, l_im = 0;
                        float phi[8];
                        /* phase 1: derotate, discriminate -> 8 angles, in two groups of four outputs = two packed pairs whose
                         * arctangent stages run side by side (fm_math.cuh v3) */
#pragma unroll
                        for (int g4 = 0; g4 < 2; g4++) {
                            int sre[4], sim[4];
#pragma unroll
                            for (int v = 0; v < 4; v++) {
                                const int u = 4 * g4 + v;
                                int y_re, y_im;
                                derotate_v2(x_re[u] >> 16, x_im[u] >> 16, r_re, r_im, y_re, y_im);
                                rot_step_v2(r_re, r_im, i4_re, i4_im, ni4_im);
                                sre[v] = (int)((unsigned)y_re * (unsigned)p_re + (unsigned)y_im * (unsigned)p_im);    /* y * conj(prev) */
                                sim[v] = (int)((unsigned)y_im * (unsigned)p_re - (unsigned)y_re * (unsigned)p_im);
                                if (KEEP_IQ) { if (!EDGE || u < nvalid) iq_c[kfirst + u] = pack16(y_re, y_im); }
                                p_re = y_re; p_im = y_im;
                                if (EDGE) { if (u == nvalid - 1) { l_re = y_re; l_im = y_im; } }
                            }
                            Atan2Pair as[2];
                            float ex[2][2], ey[2][2];
#pragma unroll
                            for (int q = 0; q < 2; q++) atan2p_stage1(sim[2 * q], sre[2 * q], sim[2 * q + 1], sre[2 * q + 1], as[q]);
#pragma unroll
                            for (int q = 0; q < 2; q++) atan2p_stage2<ATAN_SHIFT>(as[q], atan_smem, ex[q], ey[q]);
#pragma unroll
                            for (int q = 0; q < 2; q++)
                                atan2p_stage3<FMA>(sim[2 * q], sre[2 * q], sim[2 * q + 1], sre[2 * q + 1], as[q], ex[q], ey[q], z_thr,
                                                   phi[4 * g4 + 2 * q], phi[4 * g4 + 2 * q + 1]);
                        }
                        /* phase 2: angles -> PCM */
                        int pcm[8];
                        float margin = 1.0f;
#pragma unroll
                        for (int q = 0; q < 4; q++) pcm_from_phi_pair(phi[2 * q], phi[2 * q + 1], margin, pcm[2 * q], pcm[2 * q + 1]);
                        if (margin < 0.0f) {    /* about 1 block in 4000: some output sits on a float rounding boundary -> FP64 */
#pragma unroll
                            for (int u = 0; u < 8; u++) pcm[u] = pcm_from_phi_exact(__fmul_rn(phi[u], 16384.0f));
                        }
                        uint32_t out[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) out[u] = ((uint32_t)pcm[2 * u] & 0xffffu) | ((uint32_t)pcm[2 * u + 1] << 16);
                        if (!EDGE || nvalid == 8) {
                            *reinterpret_cast<uint4 *>(pcm_c + kfirst) = make_uint4(out[0], out[1], out[2], out[3]);
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; u++)
                                if (u < nvalid) pcm_c[kfirst + u] = (short)(out[u >> 1] >> (16 * (u & 1)));
                        }
                        /* the thread that produced the submit's last output hands y[K-1] to the next submit */
                        if (EDGE) { if (kfirst + nvalid == K32) p.last_out[c] = pack16(l_re, l_im); }
                    };
                    if (tile_inside || kfirst + 8 < K32) block8(std::false_type{});
                    else block8(std::true_type{});
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp_u == MMA_WARP) ptx::tmem_dealloc(tmem_base, 512);
}

} // namespace

/* ---------------------------------------------------------------------------------------------- */
/* entry (row, q, k_elem) of the tap matrix: row = 32*s + 16*is_im + i  <->  channel 16*s + i of the group */
static inline int tap_entry(const int16_t *c_re, const int16_t *c_im, int T, int D, int c, int row_is_im, int q, int k_elem)
{
    const int ip = k_elem >> 1, comp = k_elem & 1;
    const int i = q * D + ip;
    if (ip >= D || i >= T) return 0;
    const int re = c_re[(size_t)c * T + i], im = c_im[(size_t)c * T + i];
    if (!row_is_im) return comp ? -im : re;     /* re row: (c_re, -c_im) against (s_re, s_im) */
    return comp ? re : im;                      /* im row: (c_im,  c_re) */
}

static inline int row_channel(int g, int row, int &is_im)
{
    const int s = row >> 5, w = row & 31;
    is_im = w >> 4;
    return g * TC_CH + 16 * s + (w & 15);
}

TcPlan tc_make_plan(int T, int D, int C, const int16_t *c_re, const int16_t *c_im, int smem_max)
{
    TcPlan pl;
    pl.T = T; pl.D = D; pl.C = C;
    pl.Kp = ((2 * D + 31) / 32) * 32;
    pl.Q = (T + D - 1) / D;
    pl.R = TC_N + pl.Q - 1;
    pl.G = (C + TC_CH - 1) / TC_CH;
    int maxabs = 0;
    for (int c = 0; c < C; c++)
        for (int i = 0; i < T; i++) {
            const int re = c_re[(size_t)c * T + i], im = c_im[(size_t)c * T + i];
            if (im == -32768) { pl.why = "a tap component equals -32768"; return pl; }
            if (abs(re) > maxabs) maxabs = abs(re);
            if (abs(im) > maxabs) maxabs = abs(im);
        }
    const int sum_terms = (maxabs + 126) / 127;         /* int8 terms needed to write every entry as a sum */
    pl.mode = (sum_terms <= 4) ? TC_MODE_SUM : TC_MODE_RADIX;
    pl.accs = (pl.mode == TC_MODE_SUM) ? 2 : 3;
    pl.nt_stages = 512 / (pl.accs * TC_ACC_STRIDE);

    /* K chunks (32 bytes = 16 complex taps) each block-row q really covers */
    struct QK { int q, kk; };
    std::vector<QK> qk;
    for (int q = 0; q < pl.Q; q++) {
        const int taps = (T - q * D < D) ? T - q * D : D;
        const int n = (2 * taps + 31) / 32;
        for (int kk = 0; kk < n; kk++) qk.push_back({ q, kk });
    }
    /* SUM mode: the largest |entry| of every chunk anywhere in the bank decides how many terms it needs */
    std::vector<int> chunk_max(qk.size(), 0);
    if (pl.mode == TC_MODE_SUM) {
        for (size_t n = 0; n < qk.size(); n++)
            for (int c = 0; c < C; c++)
                for (int im = 0; im < 2; im++)
                    for (int b = 0; b < 32; b++) {
                        const int v = abs(tap_entry(c_re, c_im, T, D, c, im, qk[n].q, qk[n].kk * 32 + b));
                        if (v > chunk_max[n]) chunk_max[n] = v;
                    }
    }
    const uint32_t slab16 = (uint32_t)pl.R, nslab = (uint32_t)pl.Kp / 16;
    const uint32_t plane_lo16 = nslab * slab16;
    if ((size_t)2 * nslab * slab16 >= 8192) { pl.why = "sample tile too large for the MMA program encoding"; return pl; }
    std::vector<bool> started(pl.accs, false);
    std::vector<TcMma> part[2];         /* per issuing warp; acc 0 and 2 -> warp 0, acc 1 -> warp 1 */
    auto emit = [&](int chunk_idx, bool b_lo, int q, int kk, int acc, int idesc) {
        TcMma m;
        const uint32_t a_off16 = (uint32_t)chunk_idx * 256;
        const uint32_t b_off16 = (b_lo ? plane_lo16 : 0) + (uint32_t)kk * 2 * slab16 + (uint32_t)q;
        m.a_lo = a_off16 | ((2048u >> 4) << 16);                    /* LBO of the tap image: 2048 B between the K halves */
        m.b_lo = b_off16 | ((uint32_t)pl.R << 16);                  /* LBO of a sample stage: R * 16 B */
        m.d_acc = (uint32_t)acc * TC_ACC_STRIDE | ((started[acc] ? 1u : 0u) << 31);
        m.idesc = ptx::idesc_i8(128, TC_N, !(idesc & 2), !(idesc & 1));
        started[acc] = true;
        part[acc & 1].push_back(m);
    };
    if (pl.mode == TC_MODE_SUM) {
        /* term 0 everywhere, further terms where an entry exceeds 127 * term; acc 0 = weight 2^8 (x hi, signed),
         * acc 1 = weight 1 (x lo, unsigned) */
        for (int term = 0; term < (sum_terms > 0 ? sum_terms : 1); term++)
            for (size_t n = 0; n < qk.size(); n++) {
                if (term > 0 && chunk_max[n] <= 127 * term) continue;
                const int ci = (int)pl.chunks.size();
                pl.chunks.push_back({ qk[n].q, qk[n].kk, term });
                emit(ci, false, qk[n].q, qk[n].kk, 0, 0);
                emit(ci, true, qk[n].q, qk[n].kk, 1, 1);
            }
    } else {
        /* term 0 = low byte (unsigned), term 1 = high byte (signed); accs: 0 = 2^16, 1 = 2^8, 2 = 1 */
        for (size_t n = 0; n < qk.size(); n++) {
            const int lo = (int)pl.chunks.size();
            pl.chunks.push_back({ qk[n].q, qk[n].kk, 0 });
            const int hi = (int)pl.chunks.size();
            pl.chunks.push_back({ qk[n].q, qk[n].kk, 1 });
            emit(hi, false, qk[n].q, qk[n].kk, 0, 0);      /* hi(s) x hi(s) */
            emit(hi, true,  qk[n].q, qk[n].kk, 1, 1);      /* hi(s) x lo(u) */
            emit(lo, false, qk[n].q, qk[n].kk, 1, 2);      /* lo(u) x hi(s) */
            emit(lo, true,  qk[n].q, qk[n].kk, 2, 3);      /* lo(u) x lo(u) */
        }
    }
    pl.a_chunks = (int)pl.chunks.size();
    /* Both parts alternate between at most two (accumulator, instruction descriptor) pairs -- even entries one, odd
     * entries the other -- and only the first MMA into an accumulator does not accumulate.  The kernel relies on this
     * to keep nothing but the two operand offsets per MMA in uniform registers; verify it rather than assume it. */
    pl.prog_regular = true;
    for (int w = 0; w < 2; w++)
        for (size_t i = 0; i < part[w].size(); i++) {
            const TcMma &m = part[w][i], &ref = part[w][i & 1];
            const bool first = i == 0 || (i == 1 && (part[w][1].d_acc & 0xffffu) != (part[w][0].d_acc & 0xffffu));
            if ((m.d_acc & 0xffffu) != (ref.d_acc & 0xffffu) || m.idesc != ref.idesc || (m.d_acc >> 31) != (first ? 0u : 1u))
                pl.prog_regular = false;
        }
    pl.prog = part[0];
    pl.prog_split = (int)part[0].size();
    pl.prog.insert(pl.prog.end(), part[1].begin(), part[1].end());
    if (pl.prog.size() > (size_t)TC_PROG_MAX) { pl.why = "too many MMAs per tile (taps / decimation too large)"; return pl; }
    if (pl.a_chunks * 256 >= 16384) { pl.why = "tap image too large for the MMA program encoding"; return pl; }
    pl.a_group_bytes = (size_t)pl.a_chunks * 4096;
    pl.b_stage_bytes = (size_t)2 * pl.Kp * pl.R;
    const size_t static_smem = 2560 + 512;     /* MMA program, barriers, slack */
    /* Two channel groups per CTA whenever both tap images, two sample stages and the 16 arctangent copies fit: a sample
     * tile is then transformed once for both groups (each epilogue set owns one of them) -- the transform, which paces the
     * kernel as much as the epilogue does (diagnostics builds), costs half as much per output.
     * Otherwise one group per CTA with 16 interleaved copies of the arctangent table (32 KB) if at least 3 sample stages
     * still fit, else one copy (2 KB). */
    long long nb = 0;
    pl.gpc = 1;
    const char *gpc_env = getenv("GPUCHAN_TC_GPC");            /* measurement knob: GPUCHAN_TC_GPC=1 keeps one group per CTA */
    const bool may_pair = !(gpc_env && atoi(gpc_env) == 1);
    if (may_pair && pl.G % 2 == 0 && (pl.a_chunks * 2) * 256 < 16384) {
        const long long room = (long long)smem_max - (long long)static_smem - 2 * (long long)pl.a_group_bytes - 128 - 2048LL * 16;
        nb = room / (long long)pl.b_stage_bytes;
        if (nb > NB_MAX) nb = NB_MAX;
        if (nb >= 2) { pl.gpc = 2; pl.atan_copies = 16; }
    }
    if (pl.gpc == 1) {
        for (int copies = 16; copies >= 1; copies /= 16) {
            const long long room = (long long)smem_max - (long long)static_smem - (long long)pl.a_group_bytes - 128 - 2048LL * copies;
            nb = room / (long long)pl.b_stage_bytes;
            if (nb > NB_MAX) nb = NB_MAX;
            pl.atan_copies = copies;
            if (nb >= 3) break;
        }
    }
    if (nb < 2) { pl.why = "tap image + sample ring exceed shared memory"; return pl; }
    pl.nb_stages = (int)nb;
    pl.smem_bytes = (size_t)pl.gpc * pl.a_group_bytes + (size_t)pl.nb_stages * pl.b_stage_bytes + 2048 * (size_t)pl.atan_copies + 128;
    if ((size_t)pl.R * 16 >= (1u << 18)) { pl.why = "tile too tall for the descriptor"; return pl; }
    pl.ok = true;
    return pl;
}

void tc_build_tap_image(const TcPlan &pl, const int16_t *c_re, const int16_t *c_im, std::vector<uint8_t> &img)
{
    img.assign((size_t)pl.G * pl.a_group_bytes, 0);
    for (int g = 0; g < pl.G; g++)
        for (int ci = 0; ci < pl.a_chunks; ci++) {
            const TcPlan::Chunk &ck = pl.chunks[ci];
            uint8_t *dst = img.data() + (size_t)g * pl.a_group_bytes + (size_t)ci * 4096;
            for (int j2 = 0; j2 < 2; j2++)
                for (int row = 0; row < 128; row++) {
                    int is_im;
                    const int c = row_channel(g, row, is_im);
                    for (int b = 0; b < 16; b++) {
                        const int v = (c < pl.C) ? tap_entry(c_re, c_im, pl.T, pl.D, c, is_im, ck.q, ck.kk * 32 + j2 * 16 + b) : 0;
                        int piece;
                        if (pl.mode == TC_MODE_SUM) {
                            int rest = v;
                            piece = 0;
                            for (int k = 0; k <= ck.term; k++) {
                                piece = rest > 127 ? 127 : (rest < -127 ? -127 : rest);
                                rest -= piece;
                            }
                        } else {
                            piece = ck.term == 0 ? (v & 0xff) : ((v >> 8) & 0xff);
                        }
                        dst[((size_t)j2 * 128 + row) * 16 + b] = (uint8_t)(piece & 0xff);
                    }
                }
        }
}

TcGeom tc_geometry(const TcPlan &pl, long long K, int nr_sms)
{
    TcGeom gm;
    if (K <= 0) return gm;
    gm.total_tiles = (int)((K + TC_OUT - 1) / TC_OUT);
    long long ctas = nr_sms / (pl.G / pl.gpc);
    if (ctas < 1) ctas = 1;
    if (ctas > gm.total_tiles) ctas = gm.total_tiles;
    gm.n_tiles = (int)((gm.total_tiles + ctas - 1) / ctas);
    gm.chunks = (gm.total_tiles + gm.n_tiles - 1) / gm.n_tiles;
    return gm;
}

size_t tc_max_ckpt_tiles(const TcPlan &, long long max_K, int)
{
    return (size_t)(max_K / TC_OUT + 2);
}

template <int MODE, bool KEEP_IQ, bool FMA, bool ATAN16, int XF>
static cudaError_t launch_variant(const TcKernelParams &p, unsigned ctas, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(tc_fir_fm_kernel<MODE, KEEP_IQ, FMA, ATAN16, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_fir_fm_kernel<MODE, KEEP_IQ, FMA, ATAN16, XF><<<ctas, tc_threads(XF), smem, st>>>(p);
    return cudaGetLastError();
}

template <int MODE, bool KEEP_IQ, bool FMA>
static cudaError_t launch_variant3(const TcKernelParams &p, unsigned ctas, size_t smem, cudaStream_t st, bool atan16, bool many_groups)
{
    if (many_groups)
        return atan16 ? launch_variant<MODE, KEEP_IQ, FMA, true, XF_PAIRED>(p, ctas, smem, st) : launch_variant<MODE, KEEP_IQ, FMA, false, XF_PAIRED>(p, ctas, smem, st);
    return atan16 ? launch_variant<MODE, KEEP_IQ, FMA, true, XF_ONE_GROUP>(p, ctas, smem, st) : launch_variant<MODE, KEEP_IQ, FMA, false, XF_ONE_GROUP>(p, ctas, smem, st);
}

template <int MODE, bool KEEP_IQ>
static cudaError_t launch_variant2(const TcKernelParams &p, unsigned ctas, size_t smem, cudaStream_t st, bool fma, bool atan16, bool many_groups)
{
    return fma ? launch_variant3<MODE, KEEP_IQ, true>(p, ctas, smem, st, atan16, many_groups) : launch_variant3<MODE, KEEP_IQ, false>(p, ctas, smem, st, atan16, many_groups);
}

cudaError_t tc_launch_fir_fm(const TcPlan &pl, const TcBatch &b, cudaStream_t st)
{
    if (b.geom.chunks <= 0) return cudaSuccess;
    if (b.K >= (1ull << 31) - 2 * TC_OUT) return cudaErrorInvalidValue;     /* the kernel indexes a submit's outputs with int */
    TcKernelParams p;
    memset(&p, 0, sizeof(p));
    p.in = b.in; p.D = pl.D;
    p.tap_img = b.tap_img; p.incr = b.incr; p.ckpt = b.ckpt; p.last_in = b.last_in; p.last_out = b.last_out;
    p.mu = b.mu; p.lambda = b.lambda; p.cyc = b.cyc; p.cyc_pitch = b.cyc_pitch; p.k_base = b.k_base;
    p.carry_out = b.carry_out; p.carry_from = b.carry_from; p.carry_keep = b.carry_keep;
    p.atan_tab = b.atan_tab; p.pcm = b.pcm; p.iq_out = b.iq_out; p.pitch = b.pitch; p.K = (long long)b.K;
    p.n_tiles = b.geom.n_tiles; p.total_tiles = b.geom.total_tiles;
    p.C = pl.C; p.G = pl.G; p.Kp = pl.Kp; p.Q = pl.Q; p.R = pl.R; p.gpc = pl.gpc;
    p.nb_stages = pl.nb_stages; p.prog_len = (int)pl.prog.size(); p.prog_split = pl.prog_split;
    p.prog_regular = pl.prog_regular ? 1 : 0;
    p.a_group_bytes = (uint32_t)pl.a_group_bytes; p.b_stage_bytes = (uint32_t)pl.b_stage_bytes;
    p.atan = b.atan;
    memcpy(p.prog, pl.prog.data(), pl.prog.size() * sizeof(TcMma));
    /* persistent grid: one CTA per (tile range, channel group), at most one per SM */
    const unsigned ctas = (unsigned)b.geom.chunks * (unsigned)(pl.G / pl.gpc);
    const bool iq = b.iq_out != nullptr, fma = b.atan.use_fma != 0;
    p.atan_copies = pl.atan_copies;
    const size_t sm = pl.smem_bytes;
    /* 6 transform warps (80 registers for the epilogue) when a transformed tile serves two groups, 8 otherwise (measured:
     * profiles/r02_xf_warps_experiment.txt).  GPUCHAN_TC_XF=6|8 (8 = the one-group form, now 10 warps)
     * overrides the choice for measurements. */
    bool mg = pl.gpc == 2;
    if (const char *e = getenv("GPUCHAN_TC_XF")) { if (atoi(e) == 6) mg = true; else if (atoi(e) == 8) mg = false; }
    const bool a16 = pl.atan_copies == 16;
    if (pl.mode == TC_MODE_RADIX)
        return iq ? launch_variant2<TC_MODE_RADIX, true>(p, ctas, sm, st, fma, a16, mg) : launch_variant2<TC_MODE_RADIX, false>(p, ctas, sm, st, fma, a16, mg);
    return iq ? launch_variant2<TC_MODE_SUM, true>(p, ctas, sm, st, fma, a16, mg) : launch_variant2<TC_MODE_SUM, false>(p, ctas, sm, st, fma, a16, mg);
}

} // namespace tslb200
