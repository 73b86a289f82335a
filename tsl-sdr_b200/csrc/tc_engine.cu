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
constexpr int MMA_WARP = XF_WARPS;      /* warp index of the MMA issuer */
constexpr int EPI_WARP0 = XF_WARPS + 1; /* first epilogue warp (EPI_WARP0 % 4 == 1: any 4 consecutive warps cover all TMEM slices) */
constexpr int TC_THREADS = 32 * (XF_WARPS + 1 + EPI_WARPS);
constexpr int MAX_KSTEPS = 64;          /* Q * (Kp/32) descriptors kept in shared memory */
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int PCM_PITCH = 66;           /* int16 per column row: 64 channels + 2 pad -> 33 words, conflict-free both ways */

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
    unsigned long long K;
    int nr_tiles, C, G, Kp, Q, R;
    uint32_t a_group_bytes, b_stage_bytes;
    AtanParams atan;
    long long *dbg;         /* optional per-role clock stamps of CTA 0 (bench diagnostics) */
};

template <int LIMBS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_fir_fm_kernel(const TcKernelParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t b_full[2], b_empty[2], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float2 atan_s[256];
    __shared__ __align__(8) uint64_t descA[MAX_KSTEPS * 2];       /* [q*nchunk + kk][limb] */
    __shared__ __align__(8) uint64_t descB[2 * 2 * MAX_KSTEPS];   /* [stage][plane][q*nchunk + kk] */

    uint8_t *sA = smem;                                         /* [Q][LIMBS][nslab][128][16] */
    uint8_t *sB = smem + p.a_group_bytes;                       /* [2 stages][2 planes][nslab][R][16] */
    int *accbuf = reinterpret_cast<int *>(sB + 2 * (size_t)p.b_stage_bytes);   /* [64 columns][128 rows] recombined accumulators */
    short *pcmbuf = reinterpret_cast<short *>(accbuf + TC_N * 128);             /* [64 columns][PCM_PITCH] int16 PCM of the tile */
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nslab = p.Kp >> 4, nchunk = p.Kp >> 5;
    const int g = blockIdx.x % p.G;                             /* channel group of this CTA */
    const int t_first = blockIdx.x / p.G, t_step = gridDim.x / p.G;

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
    {   /* operand descriptors are tile-invariant: build them once */
        const uint32_t a_mat_bytes = (uint32_t)p.Kp * 128, slab_bytes = (uint32_t)p.R * 16;
        const int nsteps = p.Q * nchunk;
        for (int i = tid; i < nsteps * 2; i += TC_THREADS) {
            const int step = i >> 1, limb = i & 1, q = step / nchunk, kk = step - q * nchunk;
            if (limb < LIMBS)
                descA[i] = ptx::smem_desc_kmajor_noswz(ptx::smem_u32(sA) + (uint32_t)(q * LIMBS + limb) * a_mat_bytes + kk * 2 * 2048, 2048, 128);
        }
        for (int i = tid; i < nsteps * 4; i += TC_THREADS) {
            const int step = i % nsteps, pl = (i / nsteps) & 1, st = i / (2 * nsteps);
            const int q = step / nchunk, kk = step - q * nchunk;
            const uint32_t base = ptx::smem_u32(sB) + (uint32_t)st * p.b_stage_bytes + (uint32_t)pl * nslab * slab_bytes;
            descB[(st * 2 + pl) * MAX_KSTEPS + step] = ptx::smem_desc_kmajor_noswz(base + kk * 2 * slab_bytes + q * 16, slab_bytes, 128);
        }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
#define DBG(role, it, slot) do { if (p.dbg && blockIdx.x == 0 && (it) < 32) p.dbg[((role) * 32 + (it)) * 8 + (slot)] = clock64(); } while (0)

    if (warp < XF_WARPS) {
        /* ================= transform: raw cs16 samples -> hi/lo byte planes of the smem ring =================
         * Plane row m of tile t = stream samples [(t*63 + m - 1) * D, +D); item (m, j) is one 16-byte slab entry
         * = 8 complex samples = 32 raw bytes.  Consecutive threads take consecutive j: 32-byte pieces of one
         * contiguous run, so the global reads coalesce. */
        const int items = p.R * nslab;
        const uint32_t slab_bytes = (uint32_t)p.R * 16;
        int it = 0;
        for (int t = t_first; t < p.nr_tiles; t += t_step, it++) {
            const int s = it & 1, ph = (it >> 1) & 1;
            if (tid == 0) DBG(0, it, 0);
            ptx::mbar_wait(&b_empty[s], ph ^ 1);
            if (tid == 0) DBG(0, it, 1);
            uint8_t *dst = sB + (size_t)s * p.b_stage_bytes;
            const long long row_base = (long long)t * TC_KP - 1;
            for (int item = tid; item < items; item += XF_THREADS) {
                const int m = item / nslab, j = item - m * nslab;
                const long long s0 = (row_base + m) * (long long)p.D + 8 * j;
                uint32_t w[8];
#pragma unroll
                for (int u = 0; u < 8; u++) w[u] = (8 * j + u < p.D) ? (uint32_t)in_sample(p.in, s0 + u) : 0u;
                uint4 lo, hi;       /* packed sample = bytes (lo(re), hi(re), lo(im), hi(im)) */
                lo.x = __byte_perm(w[0], w[1], 0x6420); hi.x = __byte_perm(w[0], w[1], 0x7531);
                lo.y = __byte_perm(w[2], w[3], 0x6420); hi.y = __byte_perm(w[2], w[3], 0x7531);
                lo.z = __byte_perm(w[4], w[5], 0x6420); hi.z = __byte_perm(w[4], w[5], 0x7531);
                lo.w = __byte_perm(w[6], w[7], 0x6420); hi.w = __byte_perm(w[6], w[7], 0x7531);
                *reinterpret_cast<uint4 *>(dst + (size_t)j * slab_bytes + m * 16) = hi;
                *reinterpret_cast<uint4 *>(dst + (size_t)(nslab + j) * slab_bytes + m * 16) = lo;
            }
            ptx::fence_proxy_async();       /* generic-proxy stores -> visible to the tensor core's async proxy */
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&b_full[s]);
            if (tid == 0) DBG(0, it, 2);
        }
    } else if (warp == MMA_WARP) {
        /* ================= MMA issuer ================= */
        if (lane == 0) {
            const uint32_t id_ss = ptx::idesc_i8(128, TC_N, true, true);    /* A s8, B s8 */
            const uint32_t id_su = ptx::idesc_i8(128, TC_N, true, false);   /* A s8, B u8 */
            const uint32_t id_us = ptx::idesc_i8(128, TC_N, false, true);
            const uint32_t id_uu = ptx::idesc_i8(128, TC_N, false, false);
            int it = 0;
            for (int t = t_first; t < p.nr_tiles; t += t_step, it++) {
                const int s = it & 1, ph = (it >> 1) & 1;
                DBG(1, it, 0);
                ptx::mbar_wait_sleep(&b_full[s], ph);
                DBG(1, it, 1);
                ptx::mbar_wait_sleep(&t_empty[s], ph ^ 1);
                DBG(1, it, 2);
                ptx::tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)s * 256;         /* slots: +0 (2^16), +64 (2^8), +128 (1) */
                const uint64_t *dbh_p = descB + (s * 2 + 0) * MAX_KSTEPS, *dbl_p = descB + (s * 2 + 1) * MAX_KSTEPS;
                const int nsteps = p.Q * nchunk;
                for (int st = 0; st < nsteps; st++) {
                    const uint64_t dbh = dbh_p[st], dbl = dbl_p[st];
                    const uint32_t accum = st > 0;
                    if (LIMBS == 2) {
                        const uint64_t dal = descA[2 * st], dah = descA[2 * st + 1];
                        ptx::mma_i8(acc + 0,   dah, dbh, id_ss, accum);
                        ptx::mma_i8(acc + 64,  dah, dbl, id_su, accum);
                        ptx::mma_i8(acc + 64,  dal, dbh, id_us, 1);
                        ptx::mma_i8(acc + 128, dal, dbl, id_uu, accum);
                    } else {
                        const uint64_t da = descA[2 * st];
                        ptx::mma_i8(acc + 64,  da, dbh, id_ss, accum);
                        ptx::mma_i8(acc + 128, da, dbl, id_su, accum);
                    }
                }
                DBG(1, it, 3);
                ptx::mma_commit(&b_empty[s]);       /* smem stage may be refilled once these MMAs have read it */
                ptx::mma_commit(&t_full[s]);        /* accumulators complete */
                DBG(1, it, 4);
            }
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
        const int r = et >> 6;                      /* TC_STEP-column range this thread turns into PCM */
        const int c0 = TC_STEP * r;
        const int c = g * TC_CH + ch;
        const bool live = c < p.C;
        const int iw = live ? __ldg(p.incr + c) : 0;
        const int i_re = lo16(iw), i_im = hi16(iw);
        const int2 *acc2 = reinterpret_cast<const int2 *>(accbuf);
        int *const iq_c = p.iq_out ? p.iq_out + (size_t)c * p.pitch : nullptr;
        const int K32 = (int)p.K;                   /* outputs per channel of one submit always fit 31 bits */

        int it = 0;
        for (int t = t_first; t < p.nr_tiles; t += t_step, it++) {
            const int s = it & 1, ph = (it >> 1) & 1;
            if (et == 0) DBG(2, it, 0);
            ptx::mbar_wait(&t_full[s], ph);
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

            /* ---- phase 2: one channel x TC_STEP columns per thread ---- */
            if (live) {
                const int cwk = __ldg(p.ckpt + ((size_t)t * TC_SUB + r) * p.C + c);
                int r_re = lo16(cwk), r_im = hi16(cwk);
                int p_re, p_im;
                int col = c0;
                if (r > 0 || t > 0) {
                    /* previous output (column c0-1, or the tile's leading column 0): checkpoint is its phase */
                    const int lead = (r > 0) ? c0 - 1 : 0;
                    const int2 v = acc2[lead * 64 + ch];
                    derotate(rq14(v.x), rq14(v.y), r_re, r_im, p_re, p_im);
                    rot_step(r_re, r_im, i_re, i_im);
                    if (r == 0) col = 1;
                } else {
                    /* very first column of the submit: y[k0-1] is carried state, checkpoint is column 1's phase */
                    const int lw = __ldg(p.last_in + c);
                    p_re = lo16(lw); p_im = hi16(lw);
                    col = 1;
                }
                const int kofs = t * TC_KP - 1;                         /* stream output index of column 0 */
                int col_end = c0 + TC_STEP;
                if (kofs + col_end > K32) col_end = K32 - kofs;         /* ragged last tile */
                const bool produced = col < col_end;
                const int2 *src = acc2 + col * 64 + ch;
                short *out = pcmbuf + col * PCM_PITCH + ch;
#pragma unroll 8
                for (; col < col_end; col++, src += 64, out += PCM_PITCH) {
                    const int2 v = *src;
                    int y_re, y_im;
                    derotate(rq14(v.x), rq14(v.y), r_re, r_im, y_re, y_im);
                    rot_step(r_re, r_im, i_re, i_im);
                    *out = (short)fm_pcm_bf(y_re, y_im, p_re, p_im, atan_s, p.atan);
                    if (iq_c) iq_c[kofs + col] = pack16(y_re, y_im);
                    p_re = y_re; p_im = y_im;
                }
                /* the thread that produced the submit's last output hands y[K-1] to the next submit */
                if (produced && kofs + col_end == K32) p.last_out[c] = pack16(p_re, p_im);
            }
            if (et == 0) DBG(2, it, 4);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");      /* accbuf free; the tile's PCM is complete in smem */
            if (et == 0) DBG(2, it, 5);
            /* ---- phase 3: coalesced copy-out; warp e owns channels 4e..4e+3, lanes run along time ---- */
            {
                const int kofs = t * TC_KP - 1;
                int ncol = K32 - kofs;                                  /* valid columns are 1 .. ncol-1 */
                if (ncol > TC_N) ncol = TC_N;
                const bool ok0 = lane >= 1 && lane < ncol, ok1 = lane + 32 < ncol;
#pragma unroll
                for (int j = 0; j < TC_CH / EPI_WARPS; j++) {
                    const int chn = e * (TC_CH / EPI_WARPS) + j;
                    const int cg = g * TC_CH + chn;
                    if (cg < p.C) {
                        short *dst = p.pcm + (size_t)cg * p.pitch + kofs;
                        if (ok0) dst[lane] = pcmbuf[lane * PCM_PITCH + chn];
                        if (ok1) dst[lane + 32] = pcmbuf[(lane + 32) * PCM_PITCH + chn];
                    }
                }
            }
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
    pl.smem_bytes = pl.a_group_bytes + 2 * pl.b_stage_bytes + (size_t)TC_N * 128 * 4 + (size_t)TC_N * 66 * 2 + 128;
    const size_t static_smem = 5248 + 256;  /* atan table, descriptors, barriers */
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

cudaError_t tc_launch_fir_fm(const TcPlan &pl, const TcBatch &b, int nr_sms, cudaStream_t st)
{
    TcKernelParams p;
    p.in = b.in; p.D = pl.D;
    p.tap_img = b.tap_img; p.incr = b.incr; p.ckpt = b.ckpt; p.last_in = b.last_in; p.last_out = b.last_out;
    p.atan_tab = b.atan_tab; p.pcm = b.pcm; p.iq_out = b.iq_out; p.pitch = b.pitch; p.K = b.K;
    p.nr_tiles = b.nr_tiles; p.C = pl.C; p.G = pl.G; p.Kp = pl.Kp; p.Q = pl.Q; p.R = pl.R;
    p.a_group_bytes = (uint32_t)pl.a_group_bytes; p.b_stage_bytes = (uint32_t)pl.b_stage_bytes;
    p.atan = b.atan;
    p.dbg = b.dbg;
    /* persistent grid: a multiple of G CTAs, at most one per SM, no more CTAs than (group, tile) pairs */
    long long ctas = (long long)(nr_sms / pl.G) * pl.G;
    if (ctas < pl.G) ctas = pl.G;
    const long long work = (long long)pl.G * b.nr_tiles;
    if (ctas > work) ctas = work;
    cudaError_t e;
    if (pl.limbs == 2) {
        e = cudaFuncSetAttribute(tc_fir_fm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
        if (e != cudaSuccess) return e;
        tc_fir_fm_kernel<2><<<(unsigned)ctas, TC_THREADS, pl.smem_bytes, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(tc_fir_fm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
        if (e != cudaSuccess) return e;
        tc_fir_fm_kernel<1><<<(unsigned)ctas, TC_THREADS, pl.smem_bytes, st>>>(p);
    }
    return cudaGetLastError();
}

} // namespace tslb200
