/*
 * gpuchan.cu -- B200 (sm_100a) channel bank: every narrow-band channel of a multifm receiver
 * computed from one shared wide-band int16 IQ stream.  C ABI in include/tslb200_gpuchan.h.
 *
 * Kernels in this file (exact-integer CUDA-core engine, "IMAD"):
 *   rot_cycle_detect_kernel   once per bank: transient length / period of each channel's derotator
 *   rot_prepass_kernel        per submit: derotator phase at the start of every output tile
 *   fir_fm_imad_kernel        per submit: mix+FIR+decimate (complex taps) -> derotate -> FM discriminator
 *   carry_save_kernel         per submit: keep the < T input samples the next submit still needs
 *
 * Reference semantics reproduced bit for bit (paths relative to the reference tree):
 *   filter/direct_fir.c:329-417  acc = sum_i c[i] * x[kD+i]   (int32, wraps), then rq -> derotate -> rq
 *   filter/direct_fir.c:152-172  rot <- rq(rot * incr)         (lossy int16 recurrence, sequential)
 *   multifm/fm_demod.c:36-85     s = y[k] * conj(y[k-1]);  pcm = (int16)(atan2(s)/pi * 16384)
 *
 * The complex tap x sample product uses the 3-multiplication form, which is exact modulo 2^32:
 *   (c + jd)(a + jb):  P1 = c(a+b), P2 = a(d-c), P3 = b(c+d);  re = P1 - P3, im = P1 + P2.
 */
#include "../../include/tslb200_gpuchan.h"
#include "fm_math.cuh"
#include "tc_engine.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <complex>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace tslb200;

/* ------------------------------------------------------------------------------------------ */
/* error plumbing                                                                             */
/* ------------------------------------------------------------------------------------------ */
static thread_local std::string g_last_error;

static int set_err(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return set_err(GPUCHAN_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                           cudaGetErrorString(_e));                                             \
    } while (0)

extern "C" const char *gpuchan_last_error(void) { return g_last_error.c_str(); }

/* ------------------------------------------------------------------------------------------ */
/* a1: host-side tap preparation (double + libm, like the reference)                          */
/* ------------------------------------------------------------------------------------------ */
extern "C" int gpuchan_prepare_taps(const double *lpf_taps, size_t nr_taps, int32_t offset_hz,
                                    uint32_t sample_rate_hz, double gain, int16_t *c_re, int16_t *c_im)
{
    if (!lpf_taps || !nr_taps || !sample_rate_hz || !c_re || !c_im) return set_err(GPUCHAN_E_BADARGS, "bad args");
    /* multifm/demod.c:210 -- f_offs = -2*pi*offset/fs ; :234 tap = gain * e^{j f_offs i} * lpf[i] ;
     * :242-243 (int16_t)(component * 2^14), C cast = truncation toward zero */
    const double w = -2.0 * M_PI * (double)offset_hz / (double)sample_rate_hz;
    for (size_t i = 0; i < nr_taps; i++) {
        const std::complex<double> rot = std::exp(std::complex<double>(0.0, w * (double)i));
        /* C99 (double * complex) * double evaluates component-wise: (gain*re)*lpf, (gain*im)*lpf */
        const double re = (gain * rot.real()) * lpf_taps[i];
        const double im = (gain * rot.imag()) * lpf_taps[i];
        c_re[i] = (int16_t)(re * 16384.0);
        c_im[i] = (int16_t)(im * 16384.0);
    }
    return GPUCHAN_OK;
}

extern "C" int gpuchan_derot_increment(int32_t offset_hz, uint32_t sample_rate_hz, uint32_t decimation, int16_t incr[2])
{
    if (!sample_rate_hz || !decimation || !incr) return set_err(GPUCHAN_E_BADARGS, "bad args");
    /* filter/direct_fir.c:73-77 */
    const double fwt0 = 2.0 * M_PI * (double)offset_hz / (double)sample_rate_hz;
    const std::complex<double> d = std::exp(std::complex<double>(0.0, -fwt0 * (double)decimation));
    incr[0] = (int16_t)(int32_t)(d.real() * 16384.0);
    incr[1] = (int16_t)(int32_t)(d.imag() * 16384.0);
    return GPUCHAN_OK;
}

extern "C" double gpuchan_db_to_gain(double db_gain)
{
    return pow(10.0, db_gain / 10.0);   /* multifm/receiver.c:220 (power-dB formula used as amplitude) */
}

/* ------------------------------------------------------------------------------------------ */
/* device code                                                                                */
/* ------------------------------------------------------------------------------------------ */
namespace {

constexpr int FIR_WARPS   = 8;       /* warps doing the FIR contraction */
constexpr int CTA_THREADS = (FIR_WARPS + 1) * 32;   /* + 1 warp expanding the derotator sequence */
constexpr int ROT_LMAX    = 4096;    /* longest derotator limit cycle we tabulate */
constexpr uint32_t ROT_BUDGET = 1u << 24;   /* cycle-detection step budget per channel */

/* -------------------------------------------------------------------------------------- */
/* Derotator recurrence analysis: the map rot -> rq(rot*incr) acts on a finite set, so every
 * orbit is eventually periodic.  Brent's algorithm finds transient mu and period lambda. */
__global__ void rot_cycle_detect_kernel(const int *__restrict__ incr, int nr_channels, uint32_t *__restrict__ mu_out,
                                        uint32_t *__restrict__ lambda_out, int *__restrict__ cyc /* [C][ROT_LMAX] */)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nr_channels) return;
    const int i_re = lo16(incr[c]), i_im = hi16(incr[c]);
    uint32_t steps = 0, power = 1, lam = 1;
    int t_re = 16384, t_im = 0, h_re = 16384, h_im = 0;
    rot_step(h_re, h_im, i_re, i_im);
    bool ok = true;
    while (t_re != h_re || t_im != h_im) {
        if (power == lam) { t_re = h_re; t_im = h_im; power <<= 1; lam = 0; }
        rot_step(h_re, h_im, i_re, i_im);
        lam++;
        if (++steps > ROT_BUDGET) { ok = false; break; }
    }
    if (!ok || lam > (uint32_t)ROT_LMAX) { mu_out[c] = 0xffffffffu; lambda_out[c] = 0; return; }
    t_re = 16384; t_im = 0; h_re = 16384; h_im = 0;
    for (uint32_t i = 0; i < lam; i++) rot_step(h_re, h_im, i_re, i_im);
    uint32_t mu = 0;
    while (t_re != h_re || t_im != h_im) {
        rot_step(t_re, t_im, i_re, i_im);
        rot_step(h_re, h_im, i_re, i_im);
        mu++;
    }
    for (uint32_t i = 0; i < lam; i++) {
        cyc[(size_t)c * ROT_LMAX + i] = pack16(t_re, t_im);
        rot_step(t_re, t_im, i_re, i_im);
    }
    mu_out[c] = mu; lambda_out[c] = lam;
}

/* -------------------------------------------------------------------------------------- */
/* Per submit: derotator phase checkpoints, ckpt[(t*sub + r)*C + c], one for every (tile t, sub-block r) an engine
 * starts a run of consecutive outputs at.  ckpt_local() gives the output index (relative to the submit) whose
 * phase is stored, or -1 for an entry nobody reads.  Entries are visited in increasing output order.
 *
 * IMAD engine (mode 0): tile t covers FIR outputs (columns) j = 0..KT-1 <-> output t*KP - 1 + j (KP = KT-1); column
 *   0 only feeds the discriminator's "previous sample".  r = 0 is the phase of column 0 (column 1 for the first tile
 *   of a submit, whose column 0 is the previous submit's last output), r > 0 the phase of column step*r-1.
 * Tensor-core engine (mode 1): tile t turns outputs 64*t + [0, 64) into PCM, 8 per thread (tc_engine.cuh); entry r
 *   is the phase of output 64*t + 8*r - 1, the previous sample of block r (clamped to output 0 for the very first
 *   output of the submit, whose previous sample is carried state). */
struct CkptGeom {
    int mode;
    int KP, sub, step;      /* mode 0 */
};

__device__ __forceinline__ long long ckpt_local(const CkptGeom &gm, int t, int r)
{
    if (gm.mode == 0) {
        const long long col = (r == 0) ? 0 : gm.step * r - 1;
        return (long long)t * gm.KP + col - 1 + ((t == 0 && r == 0) ? 1 : 0);
    }
    const long long v = (long long)TC_OUT * t + TC_STEP * r - 1;
    return v < 0 ? 0 : v;
}

__global__ void rot_prepass_kernel(const int *__restrict__ incr, int *__restrict__ rot_state, int nr_channels,
                                   const uint32_t *__restrict__ mu, const uint32_t *__restrict__ lambda,
                                   const int *__restrict__ cyc, unsigned long long k0, unsigned long long K,
                                   CkptGeom gm, int nr_tiles, int *__restrict__ ckpt, int transient_only)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nr_channels) return;
    const int i_re = lo16(incr[c]), i_im = hi16(incr[c]);
    const uint32_t m = mu[c], lam = lambda[c];
    const int *tab = cyc + (size_t)c * ROT_LMAX;
    int r_re = lo16(rot_state[c]), r_im = hi16(rot_state[c]);
    unsigned long long cur = k0;                 /* (r_re, r_im) == rot at output index cur */

    bool have_ph = false;                        /* on the cycle: phase index of output `cur`, advanced incrementally */
    uint32_t ph = 0;
    const uint32_t step8 = lam ? (uint32_t)TC_STEP % lam : 0;       /* the usual distance between two checkpoints */
    auto seek = [&](unsigned long long g) {
        if (lam != 0 && g >= (unsigned long long)m) {
            if (!have_ph) { ph = (uint32_t)((g - m) % lam); have_ph = true; }       /* one 64-bit division per channel */
            else {
                const unsigned long long d = g - cur;
                ph += (d == (unsigned long long)TC_STEP) ? step8 : (uint32_t)(d % lam);
                if (ph >= lam) ph -= lam;
            }
            const int w = tab[ph];
            r_re = lo16(w); r_im = hi16(w);
        } else {
            while (cur < g) { rot_step(r_re, r_im, i_re, i_im); cur++; }
        }
        cur = g;
    };

    /* checkpoints at or past the start of a tabulated cycle are independent look-ups: rot_prepass_table_kernel fills
     * them in parallel (transient_only), this thread only walks the transient */
    bool on_cycle = false;
    for (int t = 0; t < nr_tiles && !on_cycle; t++)
        for (int r = 0; r < gm.sub; r++) {
            const long long loc = ckpt_local(gm, t, r);
            if (loc < 0 || (unsigned long long)loc > K) continue;   /* unused, or past the last output of this submit: never
                                                   read, and the sequential walk must not overshoot the state we hand on */
            if (transient_only && lam != 0 && k0 + (unsigned long long)loc >= (unsigned long long)m) { on_cycle = true; break; }
            seek(k0 + (unsigned long long)loc);
            ckpt[((size_t)t * gm.sub + r) * nr_channels + c] = pack16(r_re, r_im);
        }
    seek(k0 + K);
    rot_state[c] = pack16(r_re, r_im);
}

/* Steady-state variant: every channel is past its transient (k0 >= mu) and has a tabulated cycle, so
 * every checkpoint is an independent table lookup. */
__global__ void rot_prepass_table_kernel(int *__restrict__ rot_state, int nr_channels, const uint32_t *__restrict__ mu,
                                         const uint32_t *__restrict__ lambda, const int *__restrict__ cyc,
                                         unsigned long long k0, unsigned long long K, CkptGeom gm,
                                         int nr_tiles, int *__restrict__ ckpt, int cycle_part_only)
{
    /* cycle_part_only: some channels are still in their transient (or have no tabulated cycle); fill only the
     * checkpoints at or past the start of a channel's cycle and leave rot_state to rot_prepass_kernel */
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nr_channels) return;
    const uint32_t m = mu[c], lam = lambda[c];
    if (lam == 0) return;
    const int *tab = cyc + (size_t)c * ROT_LMAX;
    const int t_begin = blockIdx.y * 16;
    const int t_end = min(nr_tiles, t_begin + 16);
    if (t_begin >= nr_tiles) return;
    bool have = false;
    unsigned long long prev = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; t++)
        for (int r = 0; r < gm.sub; r++) {
            const long long loc = ckpt_local(gm, t, r);
            if (loc < 0) continue;
            const unsigned long long g = k0 + (unsigned long long)loc;
            if (g < (unsigned long long)m) continue;
            if (!have) { ph = (uint32_t)((g - m) % lam); have = true; }       /* one 64-bit division per thread */
            else ph = (ph + (uint32_t)(g - prev)) % lam;
            prev = g;
            ckpt[((size_t)t * gm.sub + r) * nr_channels + c] = tab[ph];
        }
    if (blockIdx.y == 0 && !cycle_part_only) rot_state[c] = tab[(k0 + K - m) % lam];
}

/* rot_state[c] = phase at output index k, for banks whose checkpoints come straight from the table */
__global__ void rot_state_table_kernel(int *__restrict__ rot_state, int nr_channels, const uint32_t *__restrict__ mu,
                                       const uint32_t *__restrict__ lambda, const int *__restrict__ cyc, unsigned long long k)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nr_channels) rot_state[c] = cyc[(size_t)c * ROT_LMAX + (k - mu[c]) % lambda[c]];
}

/* -------------------------------------------------------------------------------------- */
struct FirFmParams {
    InWindow in;
    const int *taps;        /* [T][Cpad] packed (c_re | c_im << 16) */
    const int *incr;        /* [C] */
    const int *ckpt;        /* [tiles][C] */
    const int *last_in;     /* [C] packed y[k0-1] */
    int *last_out;          /* [C] */
    const float2 *atan_tab; /* [256] */
    short *pcm;             /* [C][pitch] */
    int *iq_out;            /* [C][pitch] or null */
    long long pitch;
    unsigned long long K;   /* outputs of this submit (per channel) */
    int T, D, C, Cpad;
    int first_stream;       /* 1 if k0 == 0: y[-1] = 0 comes from last_in (zero) */
    AtanParams atan;
};

template <int CPT, int R>
__global__ void __launch_bounds__(CTA_THREADS, 1) fir_fm_imad_kernel(const FirFmParams p)
{
    constexpr int CG = 32 * CPT;            /* channels per CTA */
    constexpr int KT = FIR_WARPS * R;       /* FIR outputs per tile */
    constexpr int KP = KT - 1;              /* PCM outputs per tile */
    constexpr int QP = KT + 1;              /* padded row pitch of qbuf / rotbuf */

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, D = p.D;
    const int W = KP * D + T;               /* input samples needed by the tile */
    int2 *win     = reinterpret_cast<int2 *>(smem_raw);                 /* [W] (a, b) */
    int *taps_s   = reinterpret_cast<int *>(win + W);                   /* [T][CG] packed */
    int *qbuf     = taps_s + T * CG;                                    /* [CG][QP] packed q */
    int *rotbuf   = qbuf + CG * QP;                                     /* [CG][QP] packed rot */
    float2 *atan_s = reinterpret_cast<float2 *>(rotbuf + CG * QP);      /* [256] */

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int c_base = blockIdx.y * CG;
    /* stream-relative index of FIR output j = 0 of this tile (may be -1 for tile 0) */
    const long long kf = (long long)tile * KP - 1;
    const long long s0 = kf * D;

    /* ---- stage: input window (unpacked to int32 pairs), taps, atan table ---- */
    for (int i = tid; i < W; i += CTA_THREADS) {
        const int w = in_sample(p.in, s0 + i);
        win[i] = make_int2(lo16(w), hi16(w));
    }
    for (int i = tid; i < T * CG; i += CTA_THREADS) {
        const int ti = i / CG, ch = i - ti * CG;
        const int c = c_base + ch;
        taps_s[i] = (c < p.Cpad) ? __ldg(p.taps + (size_t)ti * p.Cpad + c) : 0;
    }
    for (int i = tid; i < 256; i += CTA_THREADS) atan_s[i] = p.atan_tab[i];
    __syncthreads();

    if (warp < FIR_WARPS) {
        /* ---- FIR: lane = channel (CPT channels per thread), warp*R.. = R consecutive outputs ---- */
        int A1[CPT][R], A2[CPT][R], A3[CPT][R];
#pragma unroll
        for (int u = 0; u < CPT; u++)
#pragma unroll
            for (int r = 0; r < R; r++) { A1[u][r] = 0; A2[u][r] = 0; A3[u][r] = 0; }

        const int2 *wbase = win + (warp * R) * D;
        const int *tp = taps_s + lane;
#pragma unroll 2
        for (int i = 0; i < T; i++) {
            int tc[CPT], tdmc[CPT], tcpd[CPT];
#pragma unroll
            for (int u = 0; u < CPT; u++) {
                const int tw = tp[i * CG + 32 * u];
                const int c = lo16(tw), d = hi16(tw);
                tc[u] = c; tdmc[u] = d - c; tcpd[u] = c + d;
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int2 s = wbase[r * D + i];
                const int ab = s.x + s.y;
#pragma unroll
                for (int u = 0; u < CPT; u++) {
                    A1[u][r] += tc[u] * ab;
                    A2[u][r] += s.x * tdmc[u];
                    A3[u][r] += s.y * tcpd[u];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CPT; u++)
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int re = A1[u][r] - A3[u][r];
                const int im = A1[u][r] + A2[u][r];
                qbuf[(lane + 32 * u) * QP + warp * R + r] = pack16(rq14(re), rq14(im));
            }
    } else {
        /* ---- derotator sequence for this tile (sequential per channel) ---- */
#pragma unroll
        for (int u = 0; u < CPT; u++) {
            const int ch = lane + 32 * u, c = c_base + ch;
            if (c < p.C) {
                const int iw = __ldg(p.incr + c);
                const int i_re = lo16(iw), i_im = hi16(iw);
                const int cw = __ldg(p.ckpt + (size_t)tile * p.C + c);
                int r_re = lo16(cw), r_im = hi16(cw);
                const int j0 = (tile == 0) ? 1 : 0;
                for (int j = j0; j < KT; j++) {
                    rotbuf[ch * QP + j] = pack16(r_re, r_im);
                    rot_step(r_re, r_im, i_re, i_im);
                }
            }
        }
    }
    __syncthreads();

    /* ---- epilogue: derotate, discriminate, store; lanes run along time for coalesced stores ---- */
    const bool derot = true;
    for (int e = tid; e < CG * KP; e += CTA_THREADS) {
        const int ch = e / KP, jj = e - ch * KP + 1;
        const int c = c_base + ch;
        const long long k = (long long)tile * KP + (jj - 1);
        if (c >= p.C || (unsigned long long)k >= p.K) continue;
        const int qw = qbuf[ch * QP + jj], rw = rotbuf[ch * QP + jj];
        int y_re, y_im, p_re, p_im;
        if (derot) derotate(lo16(qw), hi16(qw), lo16(rw), hi16(rw), y_re, y_im);
        if (tile == 0 && jj == 1) {
            const int lw = __ldg(p.last_in + c);
            p_re = lo16(lw); p_im = hi16(lw);
        } else {
            const int qv = qbuf[ch * QP + jj - 1], rv = rotbuf[ch * QP + jj - 1];
            derotate(lo16(qv), hi16(qv), lo16(rv), hi16(rv), p_re, p_im);
        }
        const int pcm = fm_pcm(y_re, y_im, p_re, p_im, atan_s, p.atan);
        p.pcm[(size_t)c * p.pitch + k] = (short)pcm;
        if (p.iq_out) p.iq_out[(size_t)c * p.pitch + k] = pack16(y_re, y_im);
        if ((unsigned long long)k == p.K - 1) p.last_out[c] = pack16(y_re, y_im);
    }
}

__global__ void carry_save_kernel(InWindow in, long long from, int *__restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = in_sample(in, from + i);
}

/* 8-bit I,Q bytes -> packed int16 pairs, with the reference's conversions (see gpuchan_submit_bytes) */
__global__ void widen_bytes_kernel(const uchar2 *__restrict__ src, int *__restrict__ dst, size_t n, unsigned format)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uchar2 b = src[i];
        int re, im;
        if (format == GPUCHAN_FMT_CS8)      { re = (int)(signed char)b.x;        im = (int)(signed char)b.y; }
        else if (format == GPUCHAN_FMT_CU8) { re = (int)(signed char)b.x - 127;  im = (int)(signed char)b.y - 127; }
        else                                { re = ((int)b.x - 127) << 7;        im = ((int)b.y - 127) << 7; }
        dst[i] = pack16(re, im);
    }
}

/* The discriminator alone (multifm/fm_demod.c:36-85) over an already filtered int16 IQ stream: output k from
 * samples k and k - 1 (the one before the first is carried state) */
__global__ void fm_only_kernel(const int *__restrict__ iq, size_t n, int last_in, const float2 *__restrict__ tab, AtanParams ap,
                               short *__restrict__ pcm)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int cur = iq[k], prev = k ? iq[k - 1] : last_in;
        pcm[k] = (short)fm_pcm(lo16(cur), hi16(cur), lo16(prev), hi16(prev), tab, ap);
    }
}

template <int CPT, int R>
size_t imad_smem_bytes(int T, int D)
{
    constexpr int CG = 32 * CPT, KT = FIR_WARPS * R, KP = KT - 1, QP = KT + 1;
    const size_t W = (size_t)KP * D + T;
    return W * 8 + (size_t)T * CG * 4 + 2 * (size_t)CG * QP * 4 + 256 * 8;
}

} // namespace

/* ------------------------------------------------------------------------------------------ */
/* host object                                                                                */
/* ------------------------------------------------------------------------------------------ */
struct gpuchan {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_own = nullptr, ev_ext = nullptr;
    int T = 0, D = 0, C = 0, Cpad = 0;
    uint32_t fs = 0, flags = 0;
    int engine = GPUCHAN_ENGINE_IMAD;
    size_t max_batch = 0;
    int smem_max = 0;

    std::vector<int16_t> h_re, h_im;     /* [C][T] */
    std::vector<int> h_incr;             /* packed */

    int *d_taps = nullptr;               /* [T][Cpad] packed */
    int *d_incr = nullptr, *d_rot = nullptr;
    int *d_last[2] = { nullptr, nullptr };
    uint32_t *d_mu = nullptr, *d_lambda = nullptr;
    int *d_cyc = nullptr;
    int *d_carry[2] = { nullptr, nullptr };
    static constexpr int NSLOT = 2;          /* batches in flight: H2D | kernels | D2H overlap */
    int *d_stage[NSLOT] = { nullptr, nullptr };
    uint8_t *d_stage8[NSLOT] = { nullptr, nullptr };    /* 8-bit staging, allocated on first gpuchan_submit_bytes */
    int *d_ckpt = nullptr;
    size_t ckpt_tiles = 0;
    float2 *d_atan = nullptr;
    int16_t *d_pcm[NSLOT] = { nullptr, nullptr };
    int *d_iq[NSLOT] = { nullptr, nullptr };
    size_t slotK[NSLOT] = { 0, 0 };
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_h2d[NSLOT] = { nullptr, nullptr }, ev_done[NSLOT] = { nullptr, nullptr };
    uint64_t submit_seq = 0, collect_seq = 0;
    int last_slot = -1, collected_slot = -1;
    size_t pitch = 0;

    int pp_last = 0, pp_carry = 0;
    long long carry_len = 0;
    long long skip = 0;                  /* input samples still to be dropped (only when D > T) */
    unsigned long long k_total = 0;
    uint64_t launches = 0;
    AtanParams atan{};
    bool timing = false;
    unsigned long long mu_max = 0;          /* all channels are on their limit cycle once k_total >= mu_max */
    bool all_cyclic = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;   /* around each dominant-kernel launch */

    /* kernel variant */
    int cpt = 2, R = 8;

    /* tensor-core engine */
    TcPlan tc;
    uint8_t *d_tap_img = nullptr;
    int nr_sms = 148;
};

static void host_atan_table(float2 *out)
{
    /* The reference table (multifm/fast_atan2f.c:15-81) holds atan(i/255), i = 0..255, printed with 7
     * significant digits ("%.6e"), plus one duplicated guard entry; regenerated here bit-identically. */
    float t[257];
    for (int i = 0; i < 257; i++) {
        char txt[32];
        const int j = i > 255 ? 255 : i;
        snprintf(txt, sizeof(txt), "%.6e", atan((double)j / 255.0));
        t[i] = strtof(txt, nullptr);
    }
    for (int i = 0; i < 256; i++) out[i] = make_float2(t[i], t[i + 1] - t[i]);
}

static float host_z_small_thr()
{
    /* (double)z < 0.003921569 as a float comparison (fast_atan2f.c:123) */
    const double res = 0.003921569;
    float f = (float)res;
    if ((double)f < res) f = nextafterf(f, INFINITY);
    return f;
}

namespace tslb200 {
cudaError_t run_math_selftest(uint32_t what, uint64_t seed_or_first, uint64_t count, bool use_fma, const float2 *h_tab,
                              float z_small_thr, uint64_t out[8]);
}

extern "C" int gpuchan_math_selftest(uint32_t what, uint64_t seed_or_first, uint64_t count, uint32_t use_fma, uint64_t out[8])
{
    if (!out || what > 1) return set_err(GPUCHAN_E_BADARGS, "bad selftest arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(GPUCHAN_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    float2 tab[256];
    host_atan_table(tab);
    CUDA_TRY(tslb200::run_math_selftest(what, seed_or_first, count, use_fma != 0, tab, host_z_small_thr(), out));
    return GPUCHAN_OK;
}

namespace tslb200 {
cudaError_t run_math_eval(const int *h_im, const int *h_re, size_t n, bool use_fma, const float2 *h_tab, float z_small_thr,
                          float *h_phi, short *h_pcm);
}

extern "C" int gpuchan_math_eval(const int32_t *s_im, const int32_t *s_re, size_t n, uint32_t use_fma, float *phi, int16_t *pcm)
{
    if (!s_im || !s_re || !phi || !pcm || !n) return set_err(GPUCHAN_E_BADARGS, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(GPUCHAN_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    float2 tab[256];
    host_atan_table(tab);
    CUDA_TRY(tslb200::run_math_eval(s_im, s_re, n, use_fma != 0, tab, host_z_small_thr(), phi, pcm));
    return GPUCHAN_OK;
}

/* Host-only view of the tensor-core plan (no device needed): lets the CPU test-suite check the int8 limb
 * decomposition, the tap image and the per-tile MMA program by emulating them against the direct integer FIR. */
extern "C" int gpuchan_tc_plan_query(const gpuchan_cfg *cfg, uint32_t smem_max_bytes, uint32_t info[16],
                                     uint8_t *tap_image, size_t tap_image_cap, uint32_t *program, size_t program_cap_entries)
{
    if (!cfg || !info || cfg->struct_size != sizeof(gpuchan_cfg) || !cfg->lpf_taps || !cfg->offset_hz || !cfg->nr_channels ||
        cfg->nr_taps < 2 || !cfg->decimation || !cfg->sample_rate_hz)
        return set_err(GPUCHAN_E_BADARGS, "incomplete configuration");
    const int T = (int)cfg->nr_taps, C = (int)cfg->nr_channels;
    std::vector<int16_t> re((size_t)C * T), im((size_t)C * T);
    for (int c = 0; c < C; c++)
        gpuchan_prepare_taps(cfg->lpf_taps, T, cfg->offset_hz[c], cfg->sample_rate_hz, cfg->gain ? cfg->gain[c] : 1.0,
                             &re[(size_t)c * T], &im[(size_t)c * T]);
    const TcPlan pl = tc_make_plan(T, (int)cfg->decimation, C, re.data(), im.data(), smem_max_bytes ? (int)smem_max_bytes : 232448);
    memset(info, 0, 16 * sizeof(uint32_t));
    info[0] = pl.ok ? 1u : 0u;
    if (!pl.ok) return set_err(GPUCHAN_E_INVAL, "tensor-core engine unavailable: %s", pl.why);
    info[1] = (uint32_t)pl.mode; info[2] = (uint32_t)pl.accs; info[3] = (uint32_t)pl.nb_stages; info[4] = (uint32_t)pl.nt_stages;
    info[5] = (uint32_t)pl.a_chunks; info[6] = (uint32_t)pl.a_group_bytes; info[7] = (uint32_t)pl.b_stage_bytes;
    info[8] = (uint32_t)pl.smem_bytes; info[9] = (uint32_t)pl.atan_copies; info[10] = (uint32_t)pl.prog.size();
    info[11] = (uint32_t)pl.prog_split; info[12] = (uint32_t)pl.Kp; info[13] = (uint32_t)pl.Q; info[14] = (uint32_t)pl.R;
    info[15] = (uint32_t)pl.G | ((uint32_t)pl.gpc << 16);
    if (tap_image) {
        std::vector<uint8_t> img;
        tc_build_tap_image(pl, re.data(), im.data(), img);
        if (img.size() > tap_image_cap) return set_err(GPUCHAN_E_BADARGS, "tap image needs %zu bytes", img.size());
        memcpy(tap_image, img.data(), img.size());
    }
    if (program) {
        if (pl.prog.size() > program_cap_entries) return set_err(GPUCHAN_E_BADARGS, "program has %zu entries", pl.prog.size());
        memcpy(program, pl.prog.data(), pl.prog.size() * sizeof(TcMma));
    }
    return GPUCHAN_OK;
}

static int free_all(gpuchan *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaFree(h->d_taps); cudaFree(h->d_incr); cudaFree(h->d_rot);
    cudaFree(h->d_last[0]); cudaFree(h->d_last[1]);
    cudaFree(h->d_mu); cudaFree(h->d_lambda); cudaFree(h->d_cyc);
    cudaFree(h->d_carry[0]); cudaFree(h->d_carry[1]); cudaFree(h->d_ckpt);
    cudaFree(h->d_atan); cudaFree(h->d_tap_img);
    for (int i = 0; i < gpuchan::NSLOT; i++) {
        cudaFree(h->d_stage[i]); cudaFree(h->d_stage8[i]); cudaFree(h->d_pcm[i]); cudaFree(h->d_iq[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->ev_own) cudaEventDestroy(h->ev_own);
    if (h->ev_ext) cudaEventDestroy(h->ev_ext);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

template <int CPT, int R>
static bool variant_fits(const gpuchan *h)
{
    return imad_smem_bytes<CPT, R>(h->T, h->D) <= (size_t)h->smem_max;
}

extern "C" int gpuchan_create(gpuchan_t **ph, const gpuchan_cfg *cfg)
{
    if (!ph || !cfg) return set_err(GPUCHAN_E_BADARGS, "null argument");
    *ph = nullptr;
    if (cfg->struct_size != sizeof(gpuchan_cfg)) return set_err(GPUCHAN_E_BADARGS, "gpuchan_cfg size mismatch");
    if (!cfg->sample_rate_hz || !cfg->decimation || cfg->nr_taps < 2 || !cfg->nr_channels || !cfg->lpf_taps ||
        !cfg->offset_hz || !cfg->max_batch_samples)
        return set_err(GPUCHAN_E_BADARGS, "incomplete configuration");
    if (cfg->decimation > cfg->nr_taps)
        return set_err(GPUCHAN_E_INVAL, "decimation > nr_taps is undefined in the reference (direct_fir.c:394-401)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(GPUCHAN_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return set_err(GPUCHAN_E_BADARGS, "bad device ordinal %d", cfg->device);
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return set_err(GPUCHAN_E_NODEVICE, "device %d is sm_%d%d; this build is sm_100a only", cfg->device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(cfg->device));

    gpuchan *h = new (std::nothrow) gpuchan();
    if (!h) return set_err(GPUCHAN_E_NOMEM, "out of memory");
    h->device = cfg->device;
    h->T = (int)cfg->nr_taps; h->D = (int)cfg->decimation; h->C = (int)cfg->nr_channels;
    h->Cpad = (h->C + 63) & ~63;
    h->fs = cfg->sample_rate_hz; h->flags = cfg->flags; h->max_batch = cfg->max_batch_samples;
    h->smem_max = (int)prop.sharedMemPerBlockOptin;
    h->nr_sms = prop.multiProcessorCount;
    h->engine = GPUCHAN_ENGINE_IMAD;

#define FAIL_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            set_err(GPUCHAN_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            free_all(h);                                                                        \
            return GPUCHAN_E_CUDA;                                                              \
        }                                                                                       \
    } while (0)

    FAIL_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    FAIL_TRY(cudaEventCreateWithFlags(&h->ev_own, cudaEventDisableTiming));
    FAIL_TRY(cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming));

    /* --- a1: taps and derotator increments on the host --- */
    const int T = h->T, C = h->C, Cpad = h->Cpad;
    h->h_re.resize((size_t)C * T); h->h_im.resize((size_t)C * T); h->h_incr.resize(C);
    std::vector<int> packed((size_t)T * Cpad, 0);
    for (int c = 0; c < C; c++) {
        const double g = cfg->gain ? cfg->gain[c] : 1.0;
        gpuchan_prepare_taps(cfg->lpf_taps, T, cfg->offset_hz[c], h->fs, g, &h->h_re[(size_t)c * T], &h->h_im[(size_t)c * T]);
        int16_t inc[2];
        gpuchan_derot_increment(cfg->offset_hz[c], h->fs, h->D, inc);
        h->h_incr[c] = ((int)inc[0] & 0xffff) | ((int)inc[1] << 16);
        for (int i = 0; i < T; i++)
            packed[(size_t)i * Cpad + c] = ((int)h->h_re[(size_t)c * T + i] & 0xffff) | ((int)h->h_im[(size_t)c * T + i] << 16);
    }

    /* --- engine selection --- */
    if (cfg->engine != GPUCHAN_ENGINE_AUTO && cfg->engine != GPUCHAN_ENGINE_IMAD && cfg->engine != GPUCHAN_ENGINE_TC) {
        free_all(h);
        return set_err(GPUCHAN_E_BADARGS, "unknown engine %u", cfg->engine);
    }
    if (cfg->engine == GPUCHAN_ENGINE_TC || (cfg->engine == GPUCHAN_ENGINE_AUTO)) {
        h->tc = tc_make_plan(T, h->D, C, h->h_re.data(), h->h_im.data(), h->smem_max);
        if (h->tc.ok) h->engine = GPUCHAN_ENGINE_TC;
        else if (cfg->engine == GPUCHAN_ENGINE_TC) {
            const char *why = h->tc.why;
            free_all(h);
            return set_err(GPUCHAN_E_INVAL, "tensor-core engine unavailable: %s", why);
        } else {
            /* never silent: the exact CUDA-core engine is an order of magnitude slower */
            fprintf(stderr, "gpuchan: tensor-core engine unavailable for %d taps, decimation %d, %d channels (%s): using the int32 CUDA-core "
                            "engine (about 15x slower)\n", T, h->D, C, h->tc.why);
        }
    }
    if (getenv("GPUCHAN_VERBOSE") && h->engine == GPUCHAN_ENGINE_TC)
        fprintf(stderr, "gpuchan: tensor-core engine, tap limbs by %s (%d MMAs per 64-output tile, %d channel group%s, %d per CTA, %d sample stages)\n",
                h->tc.mode == TC_MODE_SUM ? "sum of int8 terms" : "radix 256", (int)h->tc.prog.size(), h->tc.G, h->tc.G == 1 ? "" : "s",
                h->tc.gpc, h->tc.nb_stages);

    /* --- pick the IMAD kernel variant that fits shared memory (also the fallback engine) --- */
    if (h->engine == GPUCHAN_ENGINE_IMAD) {
        if      (C > 32 && variant_fits<2, 8>(h)) { h->cpt = 2; h->R = 8; }
        else if (variant_fits<1, 8>(h) && C <= 32) { h->cpt = 1; h->R = 8; }
        else if (C > 32 && variant_fits<2, 4>(h)) { h->cpt = 2; h->R = 4; }
        else if (variant_fits<1, 8>(h))           { h->cpt = 1; h->R = 8; }
        else if (variant_fits<1, 4>(h))           { h->cpt = 1; h->R = 4; }
        else { free_all(h); return set_err(GPUCHAN_E_INVAL, "taps=%d decimation=%d do not fit shared memory", T, h->D); }
    }

    const bool use_tc = h->engine == GPUCHAN_ENGINE_TC;
    const int KP = FIR_WARPS * h->R - 1;
    const int sub = use_tc ? TC_SUB : 1;
    const size_t max_avail = h->max_batch + (size_t)T;
    const size_t max_K = max_avail / h->D + 2;
    h->pitch = (max_K + 63) & ~(size_t)63;
    h->ckpt_tiles = use_tc ? tc_max_ckpt_tiles(h->tc, (long long)max_K, h->nr_sms) : (max_K + KP - 1) / KP + 1;

    if (use_tc) {
        std::vector<uint8_t> img;
        tc_build_tap_image(h->tc, h->h_re.data(), h->h_im.data(), img);
        FAIL_TRY(cudaMalloc(&h->d_tap_img, img.size()));
        FAIL_TRY(cudaMemcpy(h->d_tap_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    }

    FAIL_TRY(cudaMalloc(&h->d_taps, packed.size() * sizeof(int)));
    FAIL_TRY(cudaMemcpy(h->d_taps, packed.data(), packed.size() * sizeof(int), cudaMemcpyHostToDevice));
    FAIL_TRY(cudaMalloc(&h->d_incr, C * sizeof(int)));
    FAIL_TRY(cudaMemcpy(h->d_incr, h->h_incr.data(), C * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<int> rot0(C, 16384);             /* direct_fir.c:78-79: rot = (1<<14, 0) */
    FAIL_TRY(cudaMalloc(&h->d_rot, C * sizeof(int)));
    FAIL_TRY(cudaMemcpy(h->d_rot, rot0.data(), C * sizeof(int), cudaMemcpyHostToDevice));
    for (int i = 0; i < 2; i++) {
        FAIL_TRY(cudaMalloc(&h->d_last[i], C * sizeof(int)));
        FAIL_TRY(cudaMemset(h->d_last[i], 0, C * sizeof(int)));     /* fm_demod.c: last sample starts at 0 */
        FAIL_TRY(cudaMalloc(&h->d_carry[i], (size_t)(T + 1) * sizeof(int)));
    }
    FAIL_TRY(cudaMalloc(&h->d_mu, C * sizeof(uint32_t)));
    FAIL_TRY(cudaMalloc(&h->d_lambda, C * sizeof(uint32_t)));
    FAIL_TRY(cudaMalloc(&h->d_cyc, (size_t)C * ROT_LMAX * sizeof(int)));
    FAIL_TRY(cudaMalloc(&h->d_ckpt, h->ckpt_tiles * sub * C * sizeof(int)));
    FAIL_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    FAIL_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < gpuchan::NSLOT; i++) {
        FAIL_TRY(cudaMalloc(&h->d_stage[i], h->max_batch * sizeof(int)));
        FAIL_TRY(cudaMalloc(&h->d_pcm[i], (size_t)C * h->pitch * sizeof(int16_t)));
        if (h->flags & GPUCHAN_F_KEEP_IQ) FAIL_TRY(cudaMalloc(&h->d_iq[i], (size_t)C * h->pitch * sizeof(int)));
        FAIL_TRY(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
        FAIL_TRY(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    float2 tab[256];
    host_atan_table(tab);
    FAIL_TRY(cudaMalloc(&h->d_atan, sizeof(tab)));
    FAIL_TRY(cudaMemcpy(h->d_atan, tab, sizeof(tab), cudaMemcpyHostToDevice));

    h->atan.z_small_thr = host_z_small_thr();
    h->atan.use_fma = (h->flags & GPUCHAN_F_ATAN_FMA) ? 1 : 0;

    /* --- derotator limit cycles --- */
    rot_cycle_detect_kernel<<<(C + 63) / 64, 64, 0, h->stream>>>(h->d_incr, C, h->d_mu, h->d_lambda, h->d_cyc);
    h->launches++;
    FAIL_TRY(cudaGetLastError());
    FAIL_TRY(cudaStreamSynchronize(h->stream));
    {
        std::vector<uint32_t> mu(C), lam(C);
        FAIL_TRY(cudaMemcpy(mu.data(), h->d_mu, C * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        FAIL_TRY(cudaMemcpy(lam.data(), h->d_lambda, C * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        h->all_cyclic = true; h->mu_max = 0;
        for (int c = 0; c < C; c++) {
            if (lam[c] == 0) h->all_cyclic = false;
            else if (mu[c] > h->mu_max) h->mu_max = mu[c];
        }
    }
#undef FAIL_TRY
    *ph = h;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_sync(gpuchan_t *h);

extern "C" int gpuchan_destroy(gpuchan_t **ph)
{
    if (!ph || !*ph) return set_err(GPUCHAN_E_BADARGS, "null handle");
    cudaSetDevice((*ph)->device);
    gpuchan_sync(*ph);
    cudaDeviceSynchronize();
    free_all(*ph);
    *ph = nullptr;
    return GPUCHAN_OK;
}

template <int CPT, int R>
static cudaError_t launch_imad(gpuchan *h, const FirFmParams &p, int nr_tiles, cudaStream_t st)
{
    const size_t smem = imad_smem_bytes<CPT, R>(h->T, h->D);
    cudaError_t e = cudaFuncSetAttribute(fir_fm_imad_kernel<CPT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(nr_tiles, (h->C + 32 * CPT - 1) / (32 * CPT));
    fir_fm_imad_kernel<CPT, R><<<grid, CTA_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

static int slot_acquire(gpuchan *h)
{
    if (h->submit_seq - h->collect_seq >= (uint64_t)gpuchan::NSLOT)
        return set_err(GPUCHAN_E_BUSY, "%d batches in flight: gpuchan_collect() or gpuchan_discard() first", gpuchan::NSLOT);
    return GPUCHAN_OK;
}

/* enqueue all kernels of one submit; the FIR/FM kernel runs on stream st, outputs land in slot.
 * in_ready: event after which the fresh samples are readable (nullptr: already ordered on st). */
static int run_batch(gpuchan *h, const int *d_fresh, size_t n_complex, cudaStream_t st, int slot, cudaEvent_t in_ready)
{
    const long long avail = h->carry_len + (long long)n_complex;
    const int T = h->T, D = h->D;
    const unsigned long long K = (avail >= T) ? (unsigned long long)((avail - T) / D + 1) : 0;
    const bool use_tc = h->engine == GPUCHAN_ENGINE_TC;

    InWindow in;
    in.carry = h->d_carry[h->pp_carry];
    in.fresh = d_fresh;
    in.carry_len = h->carry_len;
    in.total = avail;

    /* Everything of one submit runs in order on st.  (A second stream running the derotator prepass one batch ahead
     * was measured on B200 and lost: next to the persistent FIR/FM kernel the small kernels get ~2 CTAs per SM and
     * become the critical path, step 0.29 -> 0.38 ms.) */
    if (in_ready) CUDA_TRY(cudaStreamWaitEvent(st, in_ready, 0));
    int *ckpt = h->d_ckpt;

    CkptGeom cg{};
    TcGeom tg;
    int nr_tiles;
    if (use_tc) {
        tg = tc_geometry(h->tc, (long long)K, h->nr_sms);
        cg.mode = 1; cg.sub = TC_SUB;
        nr_tiles = tg.total_tiles;
    } else {
        cg.mode = 0; cg.KP = FIR_WARPS * h->R - 1; cg.sub = 1; cg.step = 16;
        nr_tiles = (int)((K + cg.KP - 1) / cg.KP);
    }
    /* Tensor-core engine in steady state (every channel past its transient, cycle tabulated): the kernel reads the
     * phases from the cycle table itself -- no prepass, no checkpoint traffic. */
    const bool steady = h->all_cyclic && h->k_total >= h->mu_max + 1;
    const bool tc_table = use_tc && steady;
    if (K > 0) {
        if ((size_t)nr_tiles > h->ckpt_tiles || K > h->pitch) return set_err(GPUCHAN_E_INVAL, "internal capacity exceeded");
        if (tc_table) {
            /* nothing to do */
        } else if (steady) {
            dim3 g((h->C + 63) / 64, (nr_tiles + 15) / 16);
            rot_prepass_table_kernel<<<g, 64, 0, st>>>(h->d_rot, h->C, h->d_mu, h->d_lambda, h->d_cyc, h->k_total, K, cg,
                                                        nr_tiles, ckpt, 0);
            h->launches++;
        } else {
            /* some channel is still in its transient: walk that part sequentially (one thread per channel), fill
             * everything on a tabulated cycle in parallel */
            dim3 g((h->C + 63) / 64, (nr_tiles + 15) / 16);
            rot_prepass_table_kernel<<<g, 64, 0, st>>>(h->d_rot, h->C, h->d_mu, h->d_lambda, h->d_cyc, h->k_total, K, cg,
                                                        nr_tiles, ckpt, 1);
            rot_prepass_kernel<<<(h->C + 63) / 64, 64, 0, st>>>(h->d_incr, h->d_rot, h->C, h->d_mu, h->d_lambda, h->d_cyc,
                                                                 h->k_total, K, cg, nr_tiles, ckpt, 1);
            h->launches += 2;
        }
        CUDA_TRY(cudaGetLastError());
    }

    TcBatch tb;
    tb.in = in;

    /* keep what the next submit still needs: samples [K*D, avail).  (IMAD engine: after the FIR kernel, see below) */
    const long long from = (long long)K * D;
    const long long keep = avail - from;     /* < T */
    auto save_carry = [&](cudaStream_t s2) -> int {
        if (keep > 0) {
            carry_save_kernel<<<(unsigned)((keep + 255) / 256), 256, 0, s2>>>(in, from, h->d_carry[h->pp_carry ^ 1], (int)keep);
            h->launches++;
            CUDA_TRY(cudaGetLastError());
        }
        return GPUCHAN_OK;
    };
    const bool carry_in_kernel = use_tc && K > 0;
    if (use_tc && !carry_in_kernel) { if (int rc = save_carry(st)) return rc; }

    if (K > 0) {
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (h->timing) {
            CUDA_TRY(cudaEventCreate(&t0)); CUDA_TRY(cudaEventCreate(&t1));
            CUDA_TRY(cudaEventRecord(t0, st));
        }
        if (use_tc) {
            tb.tap_img = h->d_tap_img; tb.incr = h->d_incr; tb.ckpt = ckpt;
            tb.last_in = h->d_last[h->pp_last]; tb.last_out = h->d_last[h->pp_last ^ 1];
            tb.atan_tab = h->d_atan; tb.pcm = h->d_pcm[slot]; tb.iq_out = h->d_iq[slot]; tb.pitch = (long long)h->pitch;
            if (tc_table) {
                tb.ckpt = nullptr; tb.mu = h->d_mu; tb.lambda = h->d_lambda; tb.cyc = h->d_cyc; tb.cyc_pitch = ROT_LMAX;
                tb.k_base = h->k_total;
            }
            if (keep > 0) { tb.carry_out = h->d_carry[h->pp_carry ^ 1]; tb.carry_from = from; tb.carry_keep = (int)keep; }
            tb.K = K; tb.geom = tg; tb.atan = h->atan;
            CUDA_TRY(tc_launch_fir_fm(h->tc, tb, st));
            h->launches++;
        } else {
            FirFmParams p;
            p.in = in;
            p.taps = h->d_taps; p.incr = h->d_incr; p.ckpt = ckpt;
            p.last_in = h->d_last[h->pp_last]; p.last_out = h->d_last[h->pp_last ^ 1];
            p.atan_tab = h->d_atan;
            p.pcm = h->d_pcm[slot]; p.iq_out = h->d_iq[slot]; p.pitch = (long long)h->pitch;
            p.K = K; p.T = T; p.D = D; p.C = h->C; p.Cpad = h->Cpad;
            p.first_stream = (h->k_total == 0);
            p.atan = h->atan;
            cudaError_t e;
            if      (h->cpt == 2 && h->R == 8) e = launch_imad<2, 8>(h, p, nr_tiles, st);
            else if (h->cpt == 2 && h->R == 4) e = launch_imad<2, 4>(h, p, nr_tiles, st);
            else if (h->cpt == 1 && h->R == 8) e = launch_imad<1, 8>(h, p, nr_tiles, st);
            else                               e = launch_imad<1, 4>(h, p, nr_tiles, st);
            h->launches++;
            CUDA_TRY(e);
        }
        if (h->timing) { CUDA_TRY(cudaEventRecord(t1, st)); h->timed.emplace_back(t0, t1); }
        h->pp_last ^= 1;
        h->k_total += K;
    }
    if (!use_tc) { if (int rc = save_carry(st)) return rc; }
    h->pp_carry ^= 1;
    h->carry_len = keep > 0 ? keep : 0;
    h->slotK[slot] = (size_t)K;
    h->last_slot = slot;
    CUDA_TRY(cudaEventRecord(h->ev_done[slot], st));
    h->submit_seq++;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_submit_device(gpuchan_t *h, const int16_t *d_iq, size_t n_complex, void *cuda_stream)
{
    if (!h || (!d_iq && n_complex)) return set_err(GPUCHAN_E_BADARGS, "null argument");
    if (n_complex > h->max_batch) return set_err(GPUCHAN_E_INVAL, "submit of %zu samples exceeds max_batch_samples %zu", n_complex, h->max_batch);
    if (int rc = slot_acquire(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    /* cuda_stream only says WHEN the samples are readable; the bank always works on its own streams, so a
     * producer stream (NCCL broadcast, H2D copy) can run ahead of the FIR/FM kernel of the previous batch. */
    cudaEvent_t in_ready = nullptr;
    if (cuda_stream) {
        CUDA_TRY(cudaEventRecord(h->ev_ext, (cudaStream_t)cuda_stream));
        in_ready = h->ev_ext;
    }
    const int slot = (int)(h->submit_seq % gpuchan::NSLOT);
    return run_batch(h, reinterpret_cast<const int *>(d_iq), n_complex, h->stream, slot, in_ready);
}

/* Make cuda_stream wait for everything submitted to the bank so far (for callers that time or chain work on
 * their own stream). */
extern "C" int gpuchan_stream_wait(gpuchan_t *h, void *cuda_stream)
{
    if (!h || !cuda_stream) return set_err(GPUCHAN_E_BADARGS, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventRecord(h->ev_own, h->stream));
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)cuda_stream, h->ev_own, 0));
    return GPUCHAN_OK;
}

extern "C" int gpuchan_submit(gpuchan_t *h, const int16_t *iq_host, size_t n_complex)
{
    if (!h || (!iq_host && n_complex)) return set_err(GPUCHAN_E_BADARGS, "null argument");
    if (n_complex > h->max_batch) return set_err(GPUCHAN_E_INVAL, "submit of %zu samples exceeds max_batch_samples %zu", n_complex, h->max_batch);
    if (int rc = slot_acquire(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    const int slot = (int)(h->submit_seq % gpuchan::NSLOT);
    /* copy engine: H2D of batch i overlaps the kernels of batch i-1 and the D2H of batch i-2.
     * The staging slot was last read by the kernels of batch i-2 (ev_done of this slot). */
    CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_done[slot], 0));
    if (n_complex) CUDA_TRY(cudaMemcpyAsync(h->d_stage[slot], iq_host, n_complex * 4, cudaMemcpyHostToDevice, h->s_in));
    CUDA_TRY(cudaEventRecord(h->ev_h2d[slot], h->s_in));
    return run_batch(h, h->d_stage[slot], n_complex, h->stream, slot, h->ev_h2d[slot]);
}

extern "C" int gpuchan_submit_bytes(gpuchan_t *h, const uint8_t *iq8_host, size_t n_complex, uint32_t format)
{
    if (!h || (!iq8_host && n_complex)) return set_err(GPUCHAN_E_BADARGS, "null argument");
    if (format != GPUCHAN_FMT_CS8 && format != GPUCHAN_FMT_CU8 && format != GPUCHAN_FMT_CU8_RTL)
        return set_err(GPUCHAN_E_BADARGS, "unknown sample format %u", format);
    if (n_complex > h->max_batch) return set_err(GPUCHAN_E_INVAL, "submit of %zu samples exceeds max_batch_samples %zu", n_complex, h->max_batch);
    if (int rc = slot_acquire(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    const int slot = (int)(h->submit_seq % gpuchan::NSLOT);
    if (!h->d_stage8[slot]) CUDA_TRY(cudaMalloc(&h->d_stage8[slot], h->max_batch * 2));
    /* same overlap as gpuchan_submit: copy + widen of batch i on the input stream while batch i-1 computes */
    CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_done[slot], 0));
    if (n_complex) {
        CUDA_TRY(cudaMemcpyAsync(h->d_stage8[slot], iq8_host, n_complex * 2, cudaMemcpyHostToDevice, h->s_in));
        const unsigned blocks = (unsigned)std::min<size_t>((n_complex + 255) / 256, (size_t)h->nr_sms * 8);
        widen_bytes_kernel<<<blocks, 256, 0, h->s_in>>>(reinterpret_cast<const uchar2 *>(h->d_stage8[slot]), h->d_stage[slot],
                                                        n_complex, format);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(h->ev_h2d[slot], h->s_in));
    return run_batch(h, h->d_stage[slot], n_complex, h->stream, slot, h->ev_h2d[slot]);
}

extern "C" int gpuchan_sync(gpuchan_t *h)
{
    if (!h) return set_err(GPUCHAN_E_BADARGS, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < gpuchan::NSLOT; i++) CUDA_TRY(cudaEventSynchronize(h->ev_done[i]));
    CUDA_TRY(cudaStreamSynchronize(h->s_in));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    return GPUCHAN_OK;
}

extern "C" int gpuchan_pending(gpuchan_t *h, size_t *n)
{
    if (!h || !n) return set_err(GPUCHAN_E_BADARGS, "null argument");
    *n = (h->collect_seq < h->submit_seq) ? h->slotK[h->collect_seq % gpuchan::NSLOT] : 0;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_in_flight(gpuchan_t *h)
{
    return h ? (int)(h->submit_seq - h->collect_seq) : GPUCHAN_E_BADARGS;
}

extern "C" int gpuchan_discard(gpuchan_t *h)
{
    if (!h) return set_err(GPUCHAN_E_BADARGS, "null handle");
    if (h->collect_seq < h->submit_seq) { h->collected_slot = (int)(h->collect_seq % gpuchan::NSLOT); h->collect_seq++; }
    return GPUCHAN_OK;
}

/* copy-out of the oldest batch, split so that several banks (gpuchan_multi) can run theirs side by side */
extern "C" int gpuchan_collect_begin(gpuchan_t *h, int16_t *pcm_host, size_t cap)
{
    if (!h || !pcm_host) return set_err(GPUCHAN_E_BADARGS, "null argument");
    if (h->collect_seq >= h->submit_seq) return GPUCHAN_OK;          /* nothing in flight */
    const int slot = (int)(h->collect_seq % gpuchan::NSLOT);
    const size_t K = h->slotK[slot];
    if (K > cap) return set_err(GPUCHAN_E_INVAL, "collect capacity %zu < %zu outputs", cap, K);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_done[slot], 0));
    if (K)
        CUDA_TRY(cudaMemcpy2DAsync(pcm_host, cap * sizeof(int16_t), h->d_pcm[slot], h->pitch * sizeof(int16_t),
                                   K * sizeof(int16_t), h->C, cudaMemcpyDeviceToHost, h->s_out));
    return GPUCHAN_OK;
}

extern "C" int gpuchan_collect_end(gpuchan_t *h, size_t *n)
{
    if (!h || !n) return set_err(GPUCHAN_E_BADARGS, "null argument");
    *n = 0;
    if (h->collect_seq >= h->submit_seq) return GPUCHAN_OK;
    const int slot = (int)(h->collect_seq % gpuchan::NSLOT);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    *n = h->slotK[slot];
    h->collected_slot = slot;
    h->collect_seq++;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_collect(gpuchan_t *h, int16_t *pcm_host, size_t cap, size_t *n)
{
    if (!h || !pcm_host || !n) return set_err(GPUCHAN_E_BADARGS, "null argument");
    *n = 0;
    if (int rc = gpuchan_collect_begin(h, pcm_host, cap)) return rc;
    return gpuchan_collect_end(h, n);
}

extern "C" int gpuchan_collect_iq(gpuchan_t *h, int16_t *iq_host, size_t cap, size_t *n)
{
    if (!h || !iq_host || !n) return set_err(GPUCHAN_E_BADARGS, "null argument");
    if (!(h->flags & GPUCHAN_F_KEEP_IQ)) return set_err(GPUCHAN_E_INVAL, "bank was created without GPUCHAN_F_KEEP_IQ");
    *n = 0;
    const int slot = h->collected_slot;
    if (slot < 0) return GPUCHAN_OK;
    const size_t K = h->slotK[slot];
    if (K > cap) return set_err(GPUCHAN_E_INVAL, "collect capacity %zu < %zu outputs", cap, K);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_done[slot], 0));
    if (K)
        CUDA_TRY(cudaMemcpy2DAsync(iq_host, cap * 4, h->d_iq[slot], h->pitch * 4, K * 4, h->C,
                                   cudaMemcpyDeviceToHost, h->s_out));
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    *n = K;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_device_pcm(gpuchan_t *h, const int16_t **d_pcm, size_t *pitch, size_t *n)
{
    if (!h || !d_pcm || !pitch || !n) return set_err(GPUCHAN_E_BADARGS, "null argument");
    if (h->last_slot < 0) { *d_pcm = nullptr; *pitch = h->pitch; *n = 0; return GPUCHAN_OK; }
    *d_pcm = h->d_pcm[h->last_slot]; *pitch = h->pitch; *n = h->slotK[h->last_slot];
    return GPUCHAN_OK;
}

extern "C" int gpuchan_get_taps(gpuchan_t *h, uint32_t channel, int16_t *c_re, int16_t *c_im)
{
    if (!h || !c_re || !c_im || channel >= (uint32_t)h->C) return set_err(GPUCHAN_E_BADARGS, "bad argument");
    memcpy(c_re, &h->h_re[(size_t)channel * h->T], h->T * sizeof(int16_t));
    memcpy(c_im, &h->h_im[(size_t)channel * h->T], h->T * sizeof(int16_t));
    return GPUCHAN_OK;
}

extern "C" int gpuchan_get_rot_state(gpuchan_t *h, uint32_t channel, int16_t rot[2], int16_t incr[2],
                                     uint64_t *outputs_so_far, uint32_t *cycle_mu, uint32_t *cycle_lambda)
{
    if (!h || channel >= (uint32_t)h->C) return set_err(GPUCHAN_E_BADARGS, "bad argument");
    if (int rc = gpuchan_sync(h)) return rc;
    if (h->engine == GPUCHAN_ENGINE_TC && h->all_cyclic && h->k_total >= h->mu_max + 1) {
        rot_state_table_kernel<<<(h->C + 63) / 64, 64, 0, h->stream>>>(h->d_rot, h->C, h->d_mu, h->d_lambda, h->d_cyc, h->k_total);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    int w = 0;
    uint32_t mu = 0, lam = 0;
    CUDA_TRY(cudaMemcpy(&w, h->d_rot + channel, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&mu, h->d_mu + channel, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&lam, h->d_lambda + channel, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (rot) { rot[0] = (int16_t)(w & 0xffff); rot[1] = (int16_t)(w >> 16); }
    if (incr) { incr[0] = (int16_t)(h->h_incr[channel] & 0xffff); incr[1] = (int16_t)(h->h_incr[channel] >> 16); }
    if (outputs_so_far) *outputs_so_far = h->k_total;
    if (cycle_mu) *cycle_mu = mu;
    if (cycle_lambda) *cycle_lambda = lam;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_engine(gpuchan_t *h) { return h ? h->engine : GPUCHAN_E_BADARGS; }
extern "C" uint64_t gpuchan_kernel_launches(gpuchan_t *h) { return h ? h->launches : 0; }

/* Optional instrumentation for bench.py's roofline: CUDA events around every launch of the dominant
 * (FIR+FM) kernel, on the stream it is launched on. */
extern "C" int gpuchan_timing_enable(gpuchan_t *h, int on)
{
    if (!h) return set_err(GPUCHAN_E_BADARGS, "null handle");
    h->timing = on != 0;
    return GPUCHAN_OK;
}

extern "C" int gpuchan_timing_read(gpuchan_t *h, double *total_ms, uint64_t *nr_launches)
{
    if (!h || !total_ms || !nr_launches) return set_err(GPUCHAN_E_BADARGS, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    double sum = 0.0;
    uint64_t n = 0;
    for (auto &pr : h->timed) {
        CUDA_TRY(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        sum += ms; n++;
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    h->timed.clear();
    *total_ms = sum; *nr_launches = n;
    return GPUCHAN_OK;
}

/* What one launch of the dominant kernel issues to the tensor cores (bench.py's MAC/s roofline): out = { tcgen05.mma
 * instructions per tile, N of one instruction (M = 128, K = 32 int8), PCM outputs per channel a tile produces, channel
 * groups (CTAs working on the same samples) }.  All zero on the IMAD engine. */
extern "C" int gpuchan_tc_model(gpuchan_t *h, uint64_t out[4])
{
    if (!h || !out) return set_err(GPUCHAN_E_BADARGS, "null argument");
    out[0] = out[1] = out[2] = out[3] = 0;
    if (h->engine != GPUCHAN_ENGINE_TC) return GPUCHAN_OK;
    out[0] = h->tc.prog.size(); out[1] = TC_N; out[2] = TC_OUT; out[3] = (uint64_t)h->tc.G;
    return GPUCHAN_OK;
}

/* Pinned host memory for callers that stay free of CUDA headers (the C host side). */
extern "C" int gpuchan_host_alloc(void **pp, size_t bytes)
{
    if (!pp || !bytes) return set_err(GPUCHAN_E_BADARGS, "bad argument");
    CUDA_TRY(cudaHostAlloc(pp, bytes, cudaHostAllocDefault));
    return GPUCHAN_OK;
}

extern "C" int gpuchan_host_free(void *p)
{
    if (p) CUDA_TRY(cudaFreeHost(p));
    return GPUCHAN_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* FM discriminator as an object of its own (multifm/fm_demod.h:22-34)                        */
/* ------------------------------------------------------------------------------------------ */
struct gpufm {
    int device = 0;
    cudaStream_t stream = nullptr;
    size_t max_samples = 0;
    int *d_iq = nullptr;
    short *d_pcm = nullptr;
    float2 *d_atan = nullptr;
    AtanParams atan{};
    int last = 0;                       /* packed previous input sample, (0, 0) at start (fm_demod.c: zeroed state) */
    int nr_sms = 148;
};

extern "C" int gpufm_destroy(gpufm_t **ph)
{
    if (!ph || !*ph) return set_err(GPUCHAN_E_BADARGS, "null handle");
    gpufm *h = *ph;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->d_iq); cudaFree(h->d_pcm); cudaFree(h->d_atan);
    delete h;
    *ph = nullptr;
    return GPUCHAN_OK;
}

extern "C" int gpufm_create(gpufm_t **ph, int32_t device, uint32_t max_samples, uint32_t flags)
{
    if (!ph || !max_samples) return set_err(GPUCHAN_E_BADARGS, "bad argument");
    *ph = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(GPUCHAN_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return set_err(GPUCHAN_E_BADARGS, "bad device ordinal %d", device);
    CUDA_TRY(cudaSetDevice(device));
    gpufm *h = new (std::nothrow) gpufm();
    if (!h) return set_err(GPUCHAN_E_NOMEM, "out of memory");
    h->device = device; h->max_samples = max_samples;
    cudaDeviceProp prop{};
    float2 tab[256];
    host_atan_table(tab);
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_iq, (size_t)max_samples * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_pcm, (size_t)max_samples * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_atan, sizeof(tab));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_atan, tab, sizeof(tab), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_err(GPUCHAN_E_CUDA, "gpufm_create: %s", cudaGetErrorString(e));
        gpufm_destroy(&h);
        return GPUCHAN_E_CUDA;
    }
    h->nr_sms = prop.multiProcessorCount;
    h->atan.z_small_thr = host_z_small_thr();
    h->atan.use_fma = (flags & GPUCHAN_F_ATAN_FMA) ? 1 : 0;
    *ph = h;
    return GPUCHAN_OK;
}

extern "C" int gpufm_process(gpufm_t *h, const int16_t *iq_host, size_t n, int16_t *pcm_host)
{
    if (!h || !iq_host || !pcm_host || !n) return set_err(GPUCHAN_E_BADARGS, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    for (size_t done = 0; done < n; ) {
        const size_t m = std::min(n - done, h->max_samples);
        CUDA_TRY(cudaMemcpyAsync(h->d_iq, iq_host + 2 * done, m * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        const unsigned blocks = (unsigned)std::min<size_t>((m + 255) / 256, (size_t)h->nr_sms * 8);
        fm_only_kernel<<<blocks, 256, 0, h->stream>>>(h->d_iq, m, h->last, h->d_atan, h->atan, h->d_pcm);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(pcm_host + done, h->d_pcm, m * sizeof(short), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->last = ((int)iq_host[2 * (done + m) - 2] & 0xffff) | ((int)iq_host[2 * (done + m) - 1] << 16);
        done += m;
    }
    return GPUCHAN_OK;
}
