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
 * Kernels:
 *   tc_deinterleave_kernel  raw cs16 stream -> two byte planes (hi s8 / lo u8) in "slab" order
 *                           [16-byte K slab][block-row][16 B], zero padded to Kp = round_up(2D, 32) bytes per row
 *   tc_fir_fm_kernel        persistent, warp specialised: warp 0 streams sample tiles with cp.async.bulk (UBLKCP)
 *                           into a 2-stage smem ring, warp 1 issues tcgen05.mma kind::i8 (UTCIMMA) into a 2-stage
 *                           TMEM ring, warps 2-9 drain TMEM (LDTM) and run the exact epilogue: limb recombination,
 *                           rq, derotator recurrence, discriminator (fm_math.cuh), int16 PCM stores.
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
constexpr int TC_THREADS = 32 * (2 + EPI_WARPS);    /* producer warp, MMA warp, epilogue warps */
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int PCM_PITCH = 66;           /* int16 per column row: 64 channels + 2 pad -> 33 words, conflict-free both ways */

/* ---------------------------------------------------------------------------------------------- */
__global__ void tc_deinterleave_kernel(InWindow in, int D, int nslab, long long Mrows, uint8_t *__restrict__ plane_hi,
                                       uint8_t *__restrict__ plane_lo)
{
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;   /* block-row: samples [(m-1)D, mD) */
    const int j = blockIdx.y;                                               /* 16-byte slab = int16 elements 16j.. = samples 8j.. */
    if (m >= Mrows) return;
    const long long s0 = (m - 1) * (long long)D + 8 * j;
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; u++) w[u] = (8 * j + u < D) ? (uint32_t)in_sample(in, s0 + u) : 0u;
    uint4 lo, hi;
    /* packed sample = bytes (lo(re), hi(re), lo(im), hi(im)) */
    lo.x = __byte_perm(w[0], w[1], 0x6420); hi.x = __byte_perm(w[0], w[1], 0x7531);
    lo.y = __byte_perm(w[2], w[3], 0x6420); hi.y = __byte_perm(w[2], w[3], 0x7531);
    lo.z = __byte_perm(w[4], w[5], 0x6420); hi.z = __byte_perm(w[4], w[5], 0x7531);
    lo.w = __byte_perm(w[6], w[7], 0x6420); hi.w = __byte_perm(w[6], w[7], 0x7531);
    const size_t off = ((size_t)j * Mrows + m) * 16;
    *reinterpret_cast<uint4 *>(plane_lo + off) = lo;
    *reinterpret_cast<uint4 *>(plane_hi + off) = hi;
    (void)nslab;
}

/* ---------------------------------------------------------------------------------------------- */
struct TcKernelParams {
    const uint8_t *plane_hi, *plane_lo;
    long long Mrows;
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
            ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1);
            ptx::mbar_init(&t_full[s], 1); ptx::mbar_init(&t_empty[s], EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
#define DBG(role, it, slot) do { if (p.dbg && blockIdx.x == 0 && (it) < 32) p.dbg[((role) * 32 + (it)) * 8 + (slot)] = clock64(); } while (0)

    if (warp == 0) {
        /* ================= producer: sample tiles -> smem ring ================= */
        if (lane == 0) {
            int it = 0;
            for (int t = t_first; t < p.nr_tiles; t += t_step, it++) {
                const int s = it & 1, ph = (it >> 1) & 1;
                DBG(0, it, 0);
                ptx::mbar_wait_sleep(&b_empty[s], ph ^ 1);
                DBG(0, it, 1);
                ptx::mbar_arrive_expect_tx(&b_full[s], p.b_stage_bytes);
                uint8_t *dst = sB + (size_t)s * p.b_stage_bytes;
                const uint32_t slab_bytes = (uint32_t)p.R * 16;
                const size_t row0 = (size_t)t * TC_KP;
                for (int j = 0; j < nslab; j++) {
                    ptx::bulk_g2s(dst + (size_t)j * slab_bytes, p.plane_hi + ((size_t)j * p.Mrows + row0) * 16, slab_bytes, &b_full[s]);
                    ptx::bulk_g2s(dst + (size_t)(nslab + j) * slab_bytes, p.plane_lo + ((size_t)j * p.Mrows + row0) * 16, slab_bytes, &b_full[s]);
                }
                DBG(0, it, 2);
            }
        }
    } else if (warp == 1) {
        /* ================= MMA issuer ================= */
        if (lane == 0) {
            const uint32_t id_ss = ptx::idesc_i8(128, TC_N, true, true);    /* A s8, B s8 */
            const uint32_t id_su = ptx::idesc_i8(128, TC_N, true, false);   /* A s8, B u8 */
            const uint32_t id_us = ptx::idesc_i8(128, TC_N, false, true);
            const uint32_t id_uu = ptx::idesc_i8(128, TC_N, false, false);
            const uint32_t a_mat_bytes = (uint32_t)p.Kp * 128;              /* one (q, limb) tap matrix */
            const uint32_t slab_bytes = (uint32_t)p.R * 16;
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
                const uint32_t b_hi = ptx::smem_u32(sB + (size_t)s * p.b_stage_bytes);
                const uint32_t b_lo = b_hi + (uint32_t)nslab * slab_bytes;
                uint32_t first_hh = 0, first_mid = 0, first_ll = 0;
                for (int q = 0; q < p.Q; q++) {
                    const uint32_t a_q = ptx::smem_u32(sA) + (uint32_t)(q * LIMBS) * a_mat_bytes;
                    for (int kk = 0; kk < nchunk; kk++) {
                        const uint64_t dbh = ptx::smem_desc_kmajor_noswz(b_hi + kk * 2 * slab_bytes + q * 16, slab_bytes, 128);
                        const uint64_t dbl = ptx::smem_desc_kmajor_noswz(b_lo + kk * 2 * slab_bytes + q * 16, slab_bytes, 128);
                        if (LIMBS == 2) {
                            const uint64_t dah = ptx::smem_desc_kmajor_noswz(a_q + a_mat_bytes + kk * 2 * 2048, 2048, 128);
                            const uint64_t dal = ptx::smem_desc_kmajor_noswz(a_q + kk * 2 * 2048, 2048, 128);
                            ptx::mma_i8(acc + 0,   dah, dbh, id_ss, first_hh);  first_hh = 1;
                            ptx::mma_i8(acc + 64,  dah, dbl, id_su, first_mid); first_mid = 1;
                            ptx::mma_i8(acc + 64,  dal, dbh, id_us, 1);
                            ptx::mma_i8(acc + 128, dal, dbl, id_uu, first_ll);  first_ll = 1;
                        } else {
                            const uint64_t da = ptx::smem_desc_kmajor_noswz(a_q + kk * 2 * 2048, 2048, 128);
                            ptx::mma_i8(acc + 64,  da, dbh, id_ss, first_mid); first_mid = 1;
                            ptx::mma_i8(acc + 128, da, dbl, id_su, first_ll);  first_ll = 1;
                        }
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
        const int e = warp - 2;
        const int slice = warp & 3;                 /* TMEM lanes 32*slice .. +31 are the only ones this warp may read */
        const int quarter = e >> 2;                 /* which 16-column quarter of the tile this warp drains */
        const int row = 32 * slice + lane;          /* accumulator row: 2*channel + (0 = re, 1 = im) */
        const uint32_t lane_base = (uint32_t)(32 * slice) << 16;
        /* compute mapping: consecutive lanes = consecutive channels (conflict-free smem reads) */
        const int et = tid - 64;                    /* 0 .. EPI_THREADS-1 */
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
            if (tid == 64) DBG(2, it, 0);
            ptx::mbar_wait(&t_full[s], ph);
            if (tid == 64) DBG(2, it, 1);
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
            if (tid == 64) DBG(2, it, 2);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");      /* all 64 columns parked */
            if (tid == 64) DBG(2, it, 3);

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
            if (tid == 64) DBG(2, it, 4);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");      /* accbuf free; the tile's PCM is complete in smem */
            if (tid == 64) DBG(2, it, 5);
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
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
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
    const size_t static_smem = 2048 + 256;
    if (pl.smem_bytes + static_smem > (size_t)smem_max) { pl.why = "tap image + sample ring exceed shared memory"; return pl; }
    if ((size_t)pl.R * 16 >= (1u << 18)) { pl.why = "tile too tall for the descriptor"; return pl; }
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

cudaError_t tc_launch_deinterleave(const TcPlan &pl, const TcBatch &b, cudaStream_t st)
{
    dim3 grid((unsigned)((b.Mrows + 127) / 128), pl.Kp / 16);
    tc_deinterleave_kernel<<<grid, 128, 0, st>>>(b.in, pl.D, pl.Kp / 16, b.Mrows, b.plane_hi, b.plane_lo);
    return cudaGetLastError();
}

cudaError_t tc_launch_fir_fm(const TcPlan &pl, const TcBatch &b, int nr_sms, cudaStream_t st)
{
    TcKernelParams p;
    p.plane_hi = b.plane_hi; p.plane_lo = b.plane_lo; p.Mrows = b.Mrows;
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
