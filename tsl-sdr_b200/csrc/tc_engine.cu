/*
 * tc_engine.cu -- exact tensor-core engine for the channel bank (GPUCHAN_ENGINE_TC).
 *
 * The per-channel complex FIR of filter/direct_fir.c:366-385, acc[k] = sum_i c[i] * x[kD+i] (int32, wraps), is a
 * Toeplitz contraction shared by all channels.  Writing i = qD + i' turns it into Q = ceil(T/D) small GEMMs over
 * the SAME block-row matrix X[m][:] = x[(m-1)D .. mD) taken at row offsets q:
 *
 *     acc[row, k] = sum_q  A_q[row, :] . X[k + 1 + q, :]
 *
 * rows = (channel, re|im): the re row holds (c_re, -c_im) interleaved, the im row (c_im, c_re), so that X is the
 * raw interleaved I,Q stream.  int16 x int16 is made exact on the int8 tensor cores by limb splitting
 * (v = 256*hi + lo, hi signed, lo unsigned): four kind::i8 products accumulate into three int32 TMEM accumulators
 * (weights 2^16, 2^8, 1) that are recombined modulo 2^32 in the epilogue -- bit-identical to the reference's
 * wrapping int32 sum.  When every tap entry fits in int8 (typical narrow low-pass at unit gain) one limb suffices.
 *
 * Kernel:
 *   tc_fir_fm_kernel        persistent, warp specialised (21 warps):
 *     warps 0-3   transform: read the raw cs16 tile once from HBM/L2 and split it into two byte planes (hi s8 / lo u8)
 *                 in "slab" order [16-byte K slab][block-row][16 B] (rows zero padded to Kp = round_up(2D, 32) bytes)
 *                 directly in a 2-stage shared-memory ring;
 *     warp  4     issues tcgen05.mma kind::i8 (SASS UTCIMMA) from precomputed descriptors into a 2-stage TMEM ring;
 *     warps 5-20  drain TMEM (LDTM), recombine the limbs, and run the exact epilogue: rq, derotator recurrence,
 *                 discriminator (fm_math.cuh), int16 PCM, coalesced stores through shared memory.
 * The B operand needs no im2col: with K-major / no-swizzle descriptors the Q row shifts are just +16 B on the
 * operand start address (validated by tc_selftest.cu).
 */
#include "tc_engine.cuh"
#include "tc_ptx.cuh"

#include <cstdio>
#include <cstring>

namespace tslb200 {

namespace {

constexpr int EPI_WARPS = 16;
constexpr int XF_WARPS = 4;             /* transform warps: raw cs16 -> byte planes in smem */
constexpr int XF_THREADS = 32 * XF_WARPS;
/* Warp roles, lowest warp index first: epilogue | transform | MMA issuer.  The SM's warp arbiter favours the highest
 * warp index, so the short, latency-critical roles (MMA issue, then the loads feeding it) sit on top and are never
 * starved by the 16 arithmetic-heavy epilogue warps. */
constexpr int EPI_WARP0 = 0;            /* first epilogue warp; any 4 consecutive warps cover all 4 TMEM lane slices */
constexpr int XF_WARP0 = EPI_WARPS;     /* first transform warp */
constexpr int MMA_WARP = EPI_WARPS + XF_WARPS;      /* warp index of the MMA issuer */
constexpr int TC_THREADS = 32 * (XF_WARPS + 1 + EPI_WARPS);
constexpr int MAX_KSTEPS = 64;          /* Q * (Kp/32) descriptors kept in shared memory */
constexpr int EPI_THREADS = 32 * EPI_WARPS;

/* ---------------------------------------------------------------------------------------------- */
struct TcKernelParams {
    InWindow in;            /* raw interleaved int16 I,Q stream of this submit: [carry | fresh] */
    int D;
    const uint8_t *tap_img;
    const int *incr, *ckpt, *last_in;
    int *last_out;
    const float2 *atan_tab;
    short *pcm;
    int *iq_out;
    long long pitch;
    long long K;            /* outputs per channel of this submit */
    long long L;            /* outputs per chunk (64 * n_tiles - 8) */
    int n_tiles;            /* tiles per chunk */
    int C, G, Kp, Q, R;
    float inv_nslab;
    uint32_t a_group_bytes, b_stage_bytes;
    AtanParams atan;
    long long *dbg;         /* optional per-role clock stamps of CTA 0 (bench diagnostics) */
    int dbg_flags;          /* diagnostics only: 1 = skip the epilogue arithmetic, 2 = skip the transform */
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

/*
 * Work decomposition.  The K outputs of a submit are cut into `chunks` contiguous ranges of L = 64*n - 8 outputs,
 * one per CTA of a channel group; a CTA walks its range in n tiles of 64 FIR outputs (columns).  Tile i of chunk j
 * covers outputs j*L - 8 + 64*i + [0, 64): the first 8 columns of a chunk belong to the previous chunk and are only
 * computed so that column 7 can serve as the discriminator's "previous sample" of the chunk's first output.  Inside a
 * chunk the previous sample crosses tiles through shared memory.  Every epilogue thread therefore owns 8 consecutive
 * outputs whose index is a multiple of 8: one 16-byte PCM store.
 */
template <int LIMBS, bool KEEP_IQ, bool FMA>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_fir_fm_kernel(const TcKernelParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t b_full[2], b_empty[2], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float2 atan_s[256];
    __shared__ int yprev_s[2][TC_CH];

    uint8_t *sA = smem;                                         /* [Q][LIMBS][nslab][128][16] */
    uint8_t *sB = smem + p.a_group_bytes;                       /* [2 stages][2 planes][nslab][R][16] */
    int *accbuf = reinterpret_cast<int *>(sB + 2 * (size_t)p.b_stage_bytes);   /* [64 columns][128 rows] recombined accumulators */
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);       /* same value, but provably warp-uniform for the compiler */
    const int nslab = p.Kp >> 4, nchunk = p.Kp >> 5;
    const int g = blockIdx.x % p.G;                             /* channel group of this CTA */
    const int chunk = blockIdx.x / p.G;
    const long long k0 = (long long)chunk * p.L;                /* first output of this chunk */
    const long long k1 = (k0 + p.L < p.K) ? k0 + p.L : p.K;     /* one past its last output */
    const int my_tiles = (k0 < p.K) ? (int)((k1 - k0 + 8 + TC_N - 1) / TC_N) : 0;

    /* ---- one-time setup ---- */
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tap_img + (size_t)g * p.a_group_bytes);
        uint4 *dst = reinterpret_cast<uint4 *>(sA);
        for (uint32_t i = tid; i < p.a_group_bytes / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
        for (int i = tid; i < 256; i += TC_THREADS) atan_s[i] = p.atan_tab[i];
    }
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&b_full[s], XF_WARPS); ptx::mbar_init(&b_empty[s], 1);
            ptx::mbar_init(&t_full[s], 1); ptx::mbar_init(&t_empty[s], EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == MMA_WARP) ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
#define DBG(role, it, slot) do { if (p.dbg && blockIdx.x == 0 && (it) < 32) p.dbg[((role) * 32 + (it)) * 8 + (slot)] = clock64(); } while (0)

    if (warp >= XF_WARP0 && warp < XF_WARP0 + XF_WARPS) {
        const int xt = tid - 32 * XF_WARP0;         /* 0 .. XF_THREADS-1 */
        /* ================= transform: raw cs16 samples -> hi/lo byte planes of the smem ring =================
         * Plane row m of a tile whose column 0 is output kt0 = stream samples [(kt0 + m) * D, +D); item (m, j) is one
         * 16-byte slab entry = 8 complex samples = 32 raw bytes.  Consecutive threads take consecutive j: 32-byte
         * pieces of one contiguous run, so the global reads coalesce.  Entries past D in a row multiply zero taps, so
         * the fast path does not mask them. */
        const int items = p.R * nslab;
        const uint32_t slab_bytes = (uint32_t)p.R * 16;
        /* L2 prefetch of the TC_N new block-rows of tile `t` (the stream is read exactly once, straight from HBM) */
        auto prefetch_tile = [&](int t) {
            if (t >= my_tiles) return;
            const long long s_a = (k0 - 8 + (long long)t * TC_N + (p.Q - 1)) * (long long)p.D - p.in.carry_len;
            long long s_b = s_a + (long long)TC_N * p.D;
            const long long n_fresh = p.in.total - p.in.carry_len;
            if (s_a < 0 || s_a >= n_fresh) return;
            if (s_b > n_fresh) s_b = n_fresh;
            const uintptr_t a = reinterpret_cast<uintptr_t>(p.in.fresh + s_a) & ~(uintptr_t)15;
            const uintptr_t b = reinterpret_cast<uintptr_t>(p.in.fresh + s_b) & ~(uintptr_t)15;
            if (b > a) ptx::prefetch_l2(reinterpret_cast<const void *>(a), (uint32_t)(b - a));
        };
        if (xt == 0) { prefetch_tile(1); prefetch_tile(2); }
        for (int it = 0; it < my_tiles; it++) {
            const int s = it & 1, ph = (it >> 1) & 1;
            if (xt == 0) { DBG(0, it, 0); prefetch_tile(it + 3); }
            ptx::mbar_wait_sleep(&b_empty[s], ph ^ 1, 64);
            if (xt == 0) DBG(0, it, 1);
            uint8_t *dst = sB + (size_t)s * p.b_stage_bytes;
            const long long row_base = k0 - 8 + (long long)it * TC_N;
            const long long s_first = row_base * (long long)p.D;
            const long long s_last = s_first + (long long)(p.R - 1) * p.D + 8 * nslab + 12;     /* one past the furthest word read */
            if (p.dbg_flags & 2) {
            } else if (s_first >= p.in.carry_len + 4 && s_last <= p.in.total) {
                const int *base = p.in.fresh + (s_first - p.in.carry_len);
#pragma unroll 2
                for (int item = xt; item < items; item += XF_THREADS) {
                    const int m = __float2int_rz(__fmul_rn((float)item + 0.5f, p.inv_nslab));
                    const int j = item - m * nslab;
                    uint32_t w[8];
                    load8_unaligned(base + m * p.D + 8 * j, w);
                    split_store(w, dst + (size_t)j * slab_bytes + m * 16, dst + (size_t)(nslab + j) * slab_bytes + m * 16);
                }
            } else {
                for (int item = xt; item < items; item += XF_THREADS) {
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
            if (xt == 0) DBG(0, it, 2);
        }
    } else if (warp_u == MMA_WARP) {
        /* ================= MMA issuer =================
         * The whole warp walks the (warp-uniform) loops so that descriptors live in uniform registers; only the
         * tcgen05 instructions themselves are issued by one lane. */
        const bool leader = lane == 0;
        const uint32_t id_ss = ptx::idesc_i8(128, TC_N, true, true);    /* A s8, B s8 */
        const uint32_t id_su = ptx::idesc_i8(128, TC_N, true, false);   /* A s8, B u8 */
        const uint32_t id_us = ptx::idesc_i8(128, TC_N, false, true);
        const uint32_t id_uu = ptx::idesc_i8(128, TC_N, false, false);
        const uint32_t slab16 = (uint32_t)p.R;                           /* slab_bytes >> 4 */
        const uint32_t a_mat16 = (uint32_t)p.Kp * 8;                     /* a_mat_bytes >> 4 */
        const uint64_t descA0 = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sA), 2048, 128);
        const uint64_t descB0 = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sB), slab16 * 16, 128);
        for (int it = 0; it < my_tiles; it++) {
            const int s = it & 1, ph = (it >> 1) & 1;
            if (leader) DBG(1, it, 0);
            ptx::mbar_wait_sleep(&b_full[s], ph, 200000);
            if (leader) DBG(1, it, 1);
            ptx::mbar_wait_sleep(&t_empty[s], ph ^ 1, 200000);
            if (leader) DBG(1, it, 2);
            ptx::tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)s * 256;         /* slots: +0 (2^16), +64 (2^8), +128 (1) */
            const uint64_t dB_hi0 = descB0 + (uint64_t)(((uint32_t)s * p.b_stage_bytes) >> 4);
            const uint64_t dB_lo0 = dB_hi0 + (uint64_t)((uint32_t)nslab * slab16);
            uint32_t accum = 0;
            for (int q = 0; q < p.Q; q++) {
                const uint64_t dA_q = descA0 + (uint64_t)((uint32_t)(q * LIMBS) * a_mat16);
                for (int kk = 0; kk < nchunk; kk++) {
                    const uint64_t dal = dA_q + (uint64_t)((uint32_t)kk * 256u);         /* 2 slabs of 128 x 16 B */
                    const uint64_t dbh = dB_hi0 + (uint64_t)((uint32_t)kk * 2u * slab16 + (uint32_t)q);
                    const uint64_t dbl = dB_lo0 + (uint64_t)((uint32_t)kk * 2u * slab16 + (uint32_t)q);
                    if (leader) {
                        if (LIMBS == 2) {
                            const uint64_t dah = dal + (uint64_t)a_mat16;
                            ptx::mma_i8(acc + 0,   dah, dbh, id_ss, accum);
                            ptx::mma_i8(acc + 64,  dah, dbl, id_su, accum);
                            ptx::mma_i8(acc + 64,  dal, dbh, id_us, 1);
                            ptx::mma_i8(acc + 128, dal, dbl, id_uu, accum);
                        } else {
                            ptx::mma_i8(acc + 64,  dal, dbh, id_ss, accum);
                            ptx::mma_i8(acc + 128, dal, dbl, id_su, accum);
                        }
                    }
                    accum = 1;
                }
            }
            if (leader) {
                DBG(1, it, 3);
                ptx::mma_commit(&b_empty[s]);       /* smem stage may be refilled once these MMAs have read it */
                ptx::mma_commit(&t_full[s]);        /* accumulators complete */
                DBG(1, it, 4);
            }
            __syncwarp();
        }
    } else {
        /* ================= epilogue: TMEM -> smem -> derotate -> discriminate -> PCM ================= */
        const int e = warp - EPI_WARP0;
        const int slice = warp & 3;                 /* TMEM lanes 32*slice .. +31 are the only ones this warp may read */
        const int quarter = e >> 2;                 /* which 16-column quarter of the tile this warp drains */
        const int row = 32 * slice + lane;          /* accumulator row: 2*channel + (0 = re, 1 = im) */
        const uint32_t lane_base = (uint32_t)(32 * slice) << 16;
        /* compute mapping: consecutive lanes = consecutive channels (conflict-free smem reads) */
        const int et = tid - 32 * EPI_WARP0;        /* 0 .. EPI_THREADS-1 */
        const int ch = et & 63;
        const int r = et >> 6;                      /* which 8-column block of the tile this thread turns into PCM */
        const int c = g * TC_CH + ch;
        const bool live = c < p.C;
        const int iw = live ? __ldg(p.incr + c) : 0;
        const int i_re = lo16(iw), i_im = hi16(iw);
        const int2 *acc2 = reinterpret_cast<const int2 *>(accbuf);
        short *const pcm_c = p.pcm + (size_t)c * p.pitch;
        int *const iq_c = KEEP_IQ ? p.iq_out + (size_t)c * p.pitch : nullptr;
        AtanParams ap = p.atan;
        ap.use_fma = FMA ? 1 : 0;

        for (int it = 0; it < my_tiles; it++) {
            const int s = it & 1, ph = (it >> 1) & 1;
            if (et == 0) DBG(2, it, 0);
            ptx::mbar_wait_sleep(&t_full[s], ph, 32);
            if (et == 0) DBG(2, it, 1);
            ptx::tc_fence_after();
            /* ---- phase 1: drain TMEM, recombine the limbs modulo 2^32, park in smem as [column][row] ---- */
            {
                const uint32_t acc = tmem_base + (uint32_t)s * 256 + lane_base;
                int hh[16], mid[16], ll[16];
                const int cc = 16 * quarter;
                if (LIMBS == 2) ptx::tmem_ld16(acc + 0 + cc, hh);
                ptx::tmem_ld16(acc + 64 + cc, mid);
                ptx::tmem_ld16(acc + 128 + cc, ll);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++)
                    accbuf[(cc + i) * 128 + row] = ll[i] + (mid[i] << 8) + (LIMBS == 2 ? (hh[i] << 16) : 0);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&t_empty[s]);   /* TMEM stage is free again */
            if (et == 0) DBG(2, it, 2);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");      /* all 64 columns parked */
            if (et == 0) DBG(2, it, 3);

            /* ---- phase 2: one channel x 8 columns per thread ---- */
            const long long kfirst = k0 - 8 + (long long)it * TC_N + 8 * r;     /* output index of this thread's first column */
            int nvalid = (k1 - kfirst > 8) ? 8 : (int)(k1 - kfirst);
            if (live && kfirst >= k0 && nvalid > 0 && !(p.dbg_flags & 1)) {
                const int cwk = __ldg(p.ckpt + (((size_t)chunk * p.n_tiles + it) * TC_SUB + r) * p.C + c);
                int r_re = lo16(cwk), r_im = hi16(cwk);
                int p_re, p_im;
                if (r == 0) {
                    /* the previous output is the last column of the previous tile; checkpoint = phase of column 0 */
                    const int w = yprev_s[it & 1][ch];
                    p_re = lo16(w); p_im = hi16(w);
                } else if (kfirst == 0) {
                    /* very first output of the submit: y[-1] is carried state; checkpoint = phase of output 0 */
                    const int lw = __ldg(p.last_in + c);
                    p_re = lo16(lw); p_im = hi16(lw);
                } else {
                    /* previous output = column 8r-1 of this tile; the checkpoint is its phase */
                    const int2 v = acc2[(8 * r - 1) * 64 + ch];
                    derotate(rq14(v.x), rq14(v.y), r_re, r_im, p_re, p_im);
                    rot_step(r_re, r_im, i_re, i_im);
                }
                const int2 *src = acc2 + (8 * r) * 64 + ch;
                uint32_t out[4];
                int l_re = 0, l_im = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int2 v = src[u * 64];
                    int y_re, y_im;
                    derotate(rq14(v.x), rq14(v.y), r_re, r_im, y_re, y_im);
                    rot_step(r_re, r_im, i_re, i_im);
                    const uint32_t pcm = (uint32_t)fm_pcm_bf(y_re, y_im, p_re, p_im, atan_s, ap) & 0xffffu;
                    if (u & 1) out[u >> 1] |= pcm << 16; else out[u >> 1] = pcm;
                    if (KEEP_IQ) { if (u < nvalid) iq_c[kfirst + u] = pack16(y_re, y_im); }
                    p_re = y_re; p_im = y_im;
                    if (u == nvalid - 1) { l_re = y_re; l_im = y_im; }
                }
                if (r == TC_SUB - 1) yprev_s[(it + 1) & 1][ch] = pack16(p_re, p_im);
                if (nvalid == 8) {
                    *reinterpret_cast<uint4 *>(pcm_c + kfirst) = make_uint4(out[0], out[1], out[2], out[3]);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (u < nvalid) pcm_c[kfirst + u] = (short)(out[u >> 1] >> (16 * (u & 1)));
                }
                /* the thread that produced the submit's last output hands y[K-1] to the next submit */
                if (kfirst + nvalid == p.K) p.last_out[c] = pack16(l_re, l_im);
            }
            if (et == 0) DBG(2, it, 4);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");      /* accbuf may be overwritten by the next tile */
            if (et == 0) DBG(2, it, 5);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) ptx::tmem_dealloc(tmem_base, 512);
}

} // namespace

/* ---------------------------------------------------------------------------------------------- */
static inline int tap_entry(const int16_t *c_re, const int16_t *c_im, int T, int D, int c, int row_is_im, int q, int k_elem, bool &valid)
{
    const int ip = k_elem >> 1, comp = k_elem & 1;
    const int i = q * D + ip;
    valid = true;
    if (ip >= D || i >= T) return 0;
    const int re = c_re[(size_t)c * T + i], im = c_im[(size_t)c * T + i];
    int v;
    if (!row_is_im) v = comp ? -im : re;       /* re row: (c_re, -c_im) against (s_re, s_im) */
    else            v = comp ? re : im;        /* im row: (c_im,  c_re) */
    if (v > 32767) valid = false;              /* -(-32768) does not fit: engine unavailable for this tap set */
    return v;
}

TcPlan tc_make_plan(int T, int D, int C, const int16_t *c_re, const int16_t *c_im, int smem_max)
{
    TcPlan pl;
    pl.T = T; pl.D = D; pl.C = C;
    pl.Kp = ((2 * D + 31) / 32) * 32;
    pl.Q = (T + D - 1) / D;
    pl.R = TC_N + pl.Q - 1;
    pl.G = (C + TC_CH - 1) / TC_CH;
    bool fits8 = true;
    for (int c = 0; c < C; c++)
        for (int i = 0; i < T; i++) {
            const int re = c_re[(size_t)c * T + i], im = c_im[(size_t)c * T + i];
            if (im == -32768) { pl.why = "a tap component equals -32768"; return pl; }
            if (re < -128 || re > 127 || im < -127 || im > 127) fits8 = false;
        }
    pl.limbs = fits8 ? 1 : 2;
    pl.a_group_bytes = (size_t)pl.Q * pl.limbs * pl.Kp * 128;
    pl.b_stage_bytes = (size_t)2 * pl.Kp * pl.R;
    pl.smem_bytes = pl.a_group_bytes + 2 * pl.b_stage_bytes + (size_t)TC_N * 128 * 4 + 128;
    const size_t static_smem = 5248 + 512 + 256;  /* atan table, descriptors, previous-sample hand-off, barriers */
    if (pl.smem_bytes + static_smem > (size_t)smem_max) { pl.why = "tap image + sample ring exceed shared memory"; return pl; }
    if ((size_t)pl.R * 16 >= (1u << 18)) { pl.why = "tile too tall for the descriptor"; return pl; }
    if (pl.Q * (pl.Kp / 32) > 64) { pl.why = "too many K steps (taps / decimation too large)"; return pl; }
    pl.ok = true;
    return pl;
}

void tc_build_tap_image(const TcPlan &pl, const int16_t *c_re, const int16_t *c_im, std::vector<uint8_t> &img)
{
    const int nslab = pl.Kp / 16;
    img.assign((size_t)pl.G * pl.a_group_bytes, 0);
    for (int g = 0; g < pl.G; g++)
        for (int q = 0; q < pl.Q; q++)
            for (int j = 0; j < nslab; j++)
                for (int row = 0; row < 128; row++) {
                    const int c = g * TC_CH + row / 2;
                    for (int b = 0; b < 16; b++) {
                        bool valid;
                        const int v = (c < pl.C) ? tap_entry(c_re, c_im, pl.T, pl.D, c, row & 1, q, 16 * j + b, valid) : 0;
                        const size_t base = (size_t)g * pl.a_group_bytes + (size_t)(q * pl.limbs) * pl.Kp * 128 +
                                            ((size_t)j * 128 + row) * 16 + b;
                        if (pl.limbs == 2) {
                            img[base] = (uint8_t)(v & 0xff);                                    /* limb 0: low byte, unsigned */
                            img[base + (size_t)pl.Kp * 128] = (uint8_t)((v >> 8) & 0xff);       /* limb 1: high byte, signed */
                        } else {
                            img[base] = (uint8_t)(v & 0xff);                                    /* the int8 value itself */
                        }
                    }
                }
}

TcGeom tc_geometry(const TcPlan &pl, long long K, int nr_sms)
{
    TcGeom gm;
    long long chunks = nr_sms / pl.G;
    if (chunks < 1) chunks = 1;
    if (K <= 0) { gm.chunks = 0; gm.n_tiles = 0; gm.L = TC_N - 8; return gm; }
    const long long per = (K + chunks - 1) / chunks;
    gm.n_tiles = (int)((per + 8 + TC_N - 1) / TC_N);
    gm.L = (long long)TC_N * gm.n_tiles - 8;
    gm.chunks = (int)((K + gm.L - 1) / gm.L);
    return gm;
}

size_t tc_max_ckpt_tiles(const TcPlan &pl, long long max_K, int nr_sms)
{
    long long chunks = nr_sms / pl.G;
    if (chunks < 1) chunks = 1;
    return (size_t)(max_K / TC_N + 2 * chunks + 2);
}

template <int LIMBS, bool KEEP_IQ, bool FMA>
static cudaError_t launch_variant(const TcKernelParams &p, unsigned ctas, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(tc_fir_fm_kernel<LIMBS, KEEP_IQ, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_fir_fm_kernel<LIMBS, KEEP_IQ, FMA><<<ctas, TC_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t tc_launch_fir_fm(const TcPlan &pl, const TcBatch &b, cudaStream_t st)
{
    if (b.geom.chunks <= 0) return cudaSuccess;
    TcKernelParams p;
    p.in = b.in; p.D = pl.D;
    p.tap_img = b.tap_img; p.incr = b.incr; p.ckpt = b.ckpt; p.last_in = b.last_in; p.last_out = b.last_out;
    p.atan_tab = b.atan_tab; p.pcm = b.pcm; p.iq_out = b.iq_out; p.pitch = b.pitch; p.K = (long long)b.K;
    p.L = b.geom.L; p.n_tiles = b.geom.n_tiles;
    p.C = pl.C; p.G = pl.G; p.Kp = pl.Kp; p.Q = pl.Q; p.R = pl.R;
    p.inv_nslab = 1.0f / (float)(pl.Kp / 16);
    p.a_group_bytes = (uint32_t)pl.a_group_bytes; p.b_stage_bytes = (uint32_t)pl.b_stage_bytes;
    p.atan = b.atan;
    p.dbg = b.dbg;
    p.dbg_flags = b.dbg_flags;
    /* persistent grid: one CTA per (chunk, channel group), at most one per SM */
    const unsigned ctas = (unsigned)b.geom.chunks * (unsigned)pl.G;
    const bool iq = b.iq_out != nullptr, fma = b.atan.use_fma != 0;
    const size_t sm = pl.smem_bytes;
    if (pl.limbs == 2) {
        if (iq) return fma ? launch_variant<2, true, true>(p, ctas, sm, st) : launch_variant<2, true, false>(p, ctas, sm, st);
        return fma ? launch_variant<2, false, true>(p, ctas, sm, st) : launch_variant<2, false, false>(p, ctas, sm, st);
    }
    if (iq) return fma ? launch_variant<1, true, true>(p, ctas, sm, st) : launch_variant<1, true, false>(p, ctas, sm, st);
    return fma ? launch_variant<1, false, true>(p, ctas, sm, st) : launch_variant<1, false, false>(p, ctas, sm, st);
}

} // namespace tslb200
