/*
 * gpupager.cu -- B200 pager bank: rational resampler + POCSAG decoder for every channel of a channel bank.
 * C ABI in include/tslb200_gpupager.h.
 *
 * Kernels:
 *   resample_kernel   one thread per (channel, output), 256 outputs per block over a shared-memory input tile:
 *                     y_m = rq(sum_j h_{p_m}[j] * x[n_m + j]), n_m = floor(m*D/I), p_m = (m*D) mod I
 *                     (filter/polyphase_fir.c:184-227, filter/utils.c:46-116); also keeps the <= M input samples per
 *                     channel the next feed still needs.
 *   pocsag_kernel     one WARP per channel over the resampled 38400 Hz stream: optional DC blocker
 *                     (filter/dc_blocker.h:72-92), slicer, 3-rate eye sync detector with the lanes across the eye-phase
 *                     registers, 16-word batch sampled 32 bits per ballot, BCH(31,21) correction, address / alpha /
 *                     numeric assembly (pager/pager_pocsag.c:82-543, pager/bch_code.c:307-398).  Messages are queued per
 *                     channel for the host callbacks.
 *   flex_kernel       one WARP per channel over the resampled 16000 Hz stream (30 samples of bit-sync search per step,
 *                     lane 0 through the rest of a frame): optional DC blocker, Sync 1
 *                     (10-phase bit-sync search, A / B / inverted A, FIW), slicer training, Sync 2, 4 codings
 *                     (1600/2, 3200/2, 3200/4, 6400/4), block de-interleave into up to 4 phases, BCH + checksum,
 *                     BIW / address / vector walk, alphanumeric / numeric / tone / SIV assembly
 *                     (pager/pager_flex.c, whole file).
 *   pcm_carry_kernel  (negated / channel-gathered) copy of a feed for banks without a resampler.
 *
 * All arithmetic is integer/bitwise and reproduces the reference bit for bit, including its quirks:
 * LSB-first batch words (`bit << bit_count`, count masked to 5 bits as x86 does), the bit-reversed idle
 * word 0x6983915e, bit-reversed address/function reporting, EOT/NUL padding kept in the delivered text.
 */
#include "../../include/tslb200_gpupager.h"
#include "fm_math.cuh"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace tslb200;

static thread_local std::string g_pager_error;

static int perr(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_pager_error = buf;
    return code;
}

#define PCUDA(expr)                                                                                 \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return perr(GPUPAGER_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

extern "C" const char *gpupager_last_error(void) { return g_pager_error.c_str(); }

extern "C" int gpupager_quantize_taps(const double *coeffs, size_t nr, int16_t *taps_q14)
{
    if (!coeffs || !taps_q14) return perr(GPUPAGER_E_BADARGS, "null argument");
    const double q15 = (double)(1 << 14);                       /* decoder/decoder.c:528 */
    for (size_t i = 0; i < nr; i++) taps_q14[i] = (int16_t)(coeffs[i] * q15);   /* decoder.c:532 */
    return GPUPAGER_OK;
}

namespace {

constexpr uint32_t POCSAG_SYNC = 0x7cd215d8u;   /* pager/pager_pocsag_priv.h:40 */
constexpr uint32_t POCSAG_IDLE = 0x6983915eu;   /* pager/pager_pocsag_priv.h:46 */

enum { ST_SEARCH = 0, ST_SYNCHRONIZED = 1, ST_BATCH = 2, ST_SYNCWORD = 3 };
enum { MT_NONE = 0, MT_UNKNOWN = 1 };

/* GF(2^5), x^5 + x^2 + 1 (pager/pager_pocsag.c:150): alpha^i and discrete log */
__constant__ uint8_t c_gf_exp[32];
__constant__ int8_t c_gf_log[32];

struct PocsagState {
    int state;
    uint32_t sample_skip, baud;
    uint32_t b_skip, b_word, b_word_bit, b_bits;
    uint32_t s_skip, s_bits, s_word;
    uint32_t eye_cur[3], eye_matches[3];
    uint32_t n_alpha, n_numeric;
    int score, seen_nonprint;
    uint32_t capcode, w_alpha, w_numeric, vb_alpha, vb_numeric, function;
    int msg_type;
    int dc_x, dc_y, dc_acc;
    uint32_t batch[16];
    uint32_t eye_reg[75 + 32 + 16];
};

struct MsgSink {
    const unsigned *chan_id;    /* [C] channel index reported with a message (nullptr = the decoder channel itself) */
    gpupager_msg *msgs;         /* [C][cap] */
    uint32_t *count;            /* [C] */
    unsigned long long *dropped;
    uint32_t cap;               /* messages per channel between two drains */
};

struct InPcm {
    const short *carry;         /* [C][carry_pitch] */
    const short *fresh;         /* [C][fresh_pitch] */
    long long carry_pitch, fresh_pitch;
    long long carry_len, total; /* window = carry_len + fresh_len samples, starting at global index base */
    int invert;                 /* decoder -i: fresh samples are negated (int16 wrap); carried ones already were */
    const unsigned *map;        /* [C] row of `fresh` each decoder channel reads (nullptr = identity); carry rows are per decoder channel */
};

__device__ __forceinline__ int pcm_at(const InPcm &w, int c, long long i)
{
    if (i < 0 || i >= w.total) return 0;
    if (i < w.carry_len) return (int)w.carry[(size_t)c * w.carry_pitch + i];
    const int v = (int)w.fresh[(size_t)(w.map ? w.map[c] : (unsigned)c) * w.fresh_pitch + (i - w.carry_len)];
    return w.invert ? (int)(short)(-v) : v;
}

/* One block = 256 consecutive outputs of one channel.  The input window they need ([n_first, n_last + M), carry and
 * fresh part resolved once) is staged in shared memory, then every thread runs its own phase filter over it.  The last
 * block column (blockIdx.x == gridDim.x - 1) instead keeps the input samples the next feed still needs. */
constexpr int RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS) resample_kernel(InPcm in, unsigned long long base, const short *__restrict__ phase_filters, int M,
                                unsigned interp, unsigned decim, unsigned long long m0, unsigned nr_out,
                                short *__restrict__ out, long long out_pitch, int span_cap,
                                long long keep_from, short *__restrict__ keep_dst, long long keep_pitch, int keep_n)
{
    extern __shared__ short win[];
    const int c = blockIdx.y;
    if (blockIdx.x == gridDim.x - 1) {              /* the carry for the next feed */
        for (int i = threadIdx.x; i < keep_n; i += RS_THREADS) keep_dst[(size_t)c * keep_pitch + i] = (short)pcm_at(in, c, keep_from + i);
        return;
    }
    const unsigned i0 = blockIdx.x * RS_THREADS;
    const unsigned cnt = min((unsigned)RS_THREADS, nr_out - i0);
    const unsigned long long n_first = (m0 + i0) * decim / interp;
    const unsigned long long n_last = (m0 + i0 + cnt - 1) * decim / interp;
    const int span = (int)(n_last - n_first) + M;
    const long long off0 = (long long)(n_first - base);
    for (int i = threadIdx.x; i < span && i < span_cap; i += RS_THREADS) win[i] = (short)pcm_at(in, c, off0 + i);
    __syncthreads();
    if (threadIdx.x >= cnt) return;
    const unsigned long long adv = (m0 + i0 + threadIdx.x) * decim;
    const unsigned long long n_m = adv / interp;
    const unsigned p = (unsigned)(adv % interp);
    const short *h = phase_filters + (size_t)p * M;
    const short *x = win + (int)(n_m - n_first);
    int acc = 0;
#pragma unroll 4
    for (int j = 0; j < M; j++) acc += (int)x[j] * (int)h[j];
    out[(size_t)c * out_pitch + i0 + threadIdx.x] = (short)rq14(acc);
}

__global__ void pcm_carry_kernel(InPcm in, long long from, short *__restrict__ dst, long long dst_pitch, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i < n) dst[(size_t)c * dst_pitch + i] = (short)pcm_at(in, c, from + i);
}

/* ---- BCH(31,21,t=2), bit (30-j) of the word = coefficient j (pager/bch_code.c:325-394) ---- */
__device__ int bch_decode(uint32_t &word)
{
    uint32_t r = word;
    int s[5];
    bool any = false;
#pragma unroll
    for (int i = 1; i <= 4; i++) {
        int v = 0;
        for (int j = 0; j < 31; j++)
            if ((r >> (30 - j)) & 1) v ^= c_gf_exp[(i * j) % 31];
        any |= (v != 0);
        s[i] = c_gf_log[v];
    }
    if (!any) return 0;
    if (s[1] != -1) {
        const int s3 = (s[1] * 3) % 31;
        if (s[3] == s3) {
            r ^= 1u << (30 - s[1]);
        } else {
            const int aux = (s[3] != -1) ? (c_gf_exp[s3] ^ c_gf_exp[s[3]]) : c_gf_exp[s3];
            int reg1 = (s[2] - c_gf_log[aux] + 31) % 31;
            int reg2 = (s[1] - c_gf_log[aux] + 31) % 31;
            int loc0 = 0, loc1 = 0, count = 0;
            for (int i = 1; i <= 31; i++) {
                reg1 = (reg1 + 1) % 31;
                reg2 = (reg2 + 2) % 31;
                const int q = 1 ^ c_gf_exp[reg1] ^ c_gf_exp[reg2];
                if (!q) { if (count == 0) loc0 = i % 31; else if (count == 1) loc1 = i % 31; count++; }
            }
            if (count != 2) return 1;
            r ^= 1u << (30 - loc0);
            r ^= 1u << (30 - loc1);
        }
    } else if (s[2] != -1) {
        return 1;
    }
    word = r;
    return 0;
}

__device__ void msg_reset(PocsagState &p)
{
    p.n_alpha = p.n_numeric = 0;
    p.w_alpha = p.w_numeric = 0;
    p.vb_alpha = p.vb_numeric = 0;
    p.seen_nonprint = 0; p.score = 0;
    p.msg_type = MT_NONE; p.function = 0;
}

/* pager/pager_pocsag.c:242-297 */
__device__ void deliver(PocsagState &p, char *alpha, char *numeric, const MsgSink &sink, int c)
{
    if (p.msg_type == MT_NONE) return;
    if (p.n_alpha != 0) {
        const char last = alpha[p.n_alpha - 1];
        if (last == 0x4 || last == 0x3 || last == 0x0 || last == 0x17) p.score = 1;
    }
    if (p.n_numeric > 40) p.score = 1;
    const bool is_alpha = p.score > 0;
    const uint32_t slot = sink.count[c];
    if (slot < sink.cap) {
        gpupager_msg *m = sink.msgs + (size_t)c * sink.cap + slot;
        m->channel = sink.chan_id ? sink.chan_id[c] : (unsigned)c; m->kind = is_alpha ? GPUPAGER_MSG_ALPHA : GPUPAGER_MSG_NUMERIC;
        m->baud = p.baud; m->capcode = p.capcode; m->function = p.function;
        m->capcode_hi = 0;
        for (int i = 0; i < 6; i++) m->aux[i] = 0;
        const uint32_t len = is_alpha ? p.n_alpha : p.n_numeric;
        m->len = len;
        const char *src = is_alpha ? alpha : numeric;
        for (uint32_t i = 0; i < len; i++) m->text[i] = src[i];
        if (len < GPUPAGER_MSG_TEXT_MAX) m->text[len] = 0;
        sink.count[c] = slot + 1;
    } else {
        atomicAdd(sink.dropped, 1ull);
    }
    msg_reset(p);
}

/* pager/pager_pocsag.c:320-432 */
__device__ void process_batch(PocsagState &p, char *alpha, char *numeric, const MsgSink &sink, int c)
{
    const char bcd_map[16] = { '0', '1', '2', '3', '4', '5', '6', '7', '8', '9', 'X', 'U', ' ', '-', '[', ']' };
    for (unsigned z = 0; z < 16; z++) {
        uint32_t w = p.batch[z] & 0x7fffffffu;
        if (bch_decode(w)) {
            if (p.msg_type != MT_NONE) deliver(p, alpha, numeric, sink, c);
            return;
        }
        if (w == POCSAG_IDLE) {
            if (p.msg_type != MT_NONE) deliver(p, alpha, numeric, sink, c);
            continue;
        }
        if ((w & 1) == 0) {
            deliver(p, alpha, numeric, sink, c);
            p.msg_type = MT_UNKNOWN;
            p.function = (w >> 19) & 0x3;
            p.capcode = (((w >> 1) & 0x3ffffu) << 3) + ((z >> 1) & 0x7);
        } else if (p.msg_type == MT_UNKNOWN) {
            const uint32_t val = (w >> 1) & 0xfffffu;
            p.w_alpha |= val << p.vb_alpha;
            p.vb_alpha += 20;
            while (p.vb_alpha >= 7) {
                const char ch = (char)(p.w_alpha & 0x7f);
                if (p.n_alpha < 511) alpha[p.n_alpha++] = ch;       /* the reference buffer is 512 bytes, unchecked */
                if ((ch >= 0x20 && ch <= 0x7e) || ch == 0xa || ch == 0xd) {
                    if (!p.seen_nonprint) p.score++;
                } else {
                    p.seen_nonprint = 1;
                    if (ch != 0x03 && ch != 0x04 && ch != 0x17 && ch != 0x0) p.score -= 10;
                }
                p.w_alpha >>= 7;
                p.vb_alpha -= 7;
            }
            if (p.n_numeric < 511) {
                p.w_numeric |= val << p.vb_numeric;
                p.vb_numeric += 20;
                while (p.vb_numeric >= 4 && p.n_numeric < 511) {
                    numeric[p.n_numeric++] = bcd_map[p.w_numeric & 0xf];
                    p.w_numeric >>= 4;
                    p.vb_numeric -= 4;
                }
            }
        }
    }
}

__device__ __forceinline__ bool sync_ok(uint32_t w) { return __popc(w ^ POCSAG_SYNC) <= 4; }

/* ---- POCSAG decoder, one WARP per channel (pager/pager_pocsag.c:434-543 and :82-117) --------------------------------
 * The reference walks the 38400 Hz stream sample by sample through a 4-state machine.  The same machine, event driven:
 *   pass 1  the optional DC blocker (a sequential IIR, every lane computes the identical recurrence so nobody diverges)
 *           and the slicer: sign bits of the whole feed, 32 samples per ballot, into a per-channel bit array;
 *   pass 2  SEARCH: 32 samples per step -- lane l shifts sample l's bit into ITS eye-phase register of each of the three
 *           rates (75 + 32 + 16 registers, eye_cur[] picks the phase exactly like the reference; the 16 registers of the
 *           2400-baud detector are hit twice per step, lanes 0-15 then 16-31) and tests it against the sync word; the
 *           three match masks are then scanned in sample order for the run-length logic (in noise: one iteration);
 *           BATCH / SYNCWORD: the next 32 bits are sampled every sample_skip-th sample -- lane j fetches bit j of the
 *           word straight from the bit array, one ballot assembles it (LSB first for batch words, MSB first for the
 *           sync word, like the reference); BCH, address / message assembly and delivery run on lane 0.
 * Registers updated past a detection inside a SEARCH step are never looked at again: the eye state is only read in
 * SEARCH, and every return to SEARCH goes through eyes_reset (:524-527). */
constexpr int PW_WARPS = 4;                 /* channels (warps) per block */

__device__ __forceinline__ unsigned bit_at(const unsigned *__restrict__ bits, unsigned long long pos)
{
    return (bits[pos >> 5] >> (pos & 31)) & 1u;
}

/* samples until (and including) the next one a bit is taken from: skip counts up modulo 2^16 and fires on == sample_skip */
__device__ __forceinline__ unsigned until_sampling(unsigned sample_skip, unsigned skip)
{
    const unsigned k = (sample_skip - skip) & 0xffffu;
    return k ? k : 65536u;
}

__global__ void __launch_bounds__(32 * PW_WARPS) pocsag_kernel(PocsagState *__restrict__ states, char *__restrict__ text, int nr_channels,
                              short *__restrict__ pcm, long long pitch, unsigned n, unsigned *__restrict__ bits_all, unsigned bits_pitch,
                              MsgSink sink, int use_dc, int dc_p)
{
    __shared__ PocsagState sst[PW_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int c = blockIdx.x * PW_WARPS + wib;
    if (c >= nr_channels) return;                                   /* whole warps leave: no block-wide barrier below */
    PocsagState &S = sst[wib];
    {
        const unsigned *src = reinterpret_cast<const unsigned *>(states + c);
        unsigned *dst = reinterpret_cast<unsigned *>(&S);
        for (int i = lane; i < (int)(sizeof(PocsagState) / 4); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    char *alpha = text + (size_t)c * 1024, *numeric = alpha + 512;   /* message_alpha[512], message_numeric[512] */
    short *x = pcm + (size_t)c * pitch;
    unsigned *bits = bits_all + (size_t)c * bits_pitch;

    /* ---- pass 1: DC blocker (filter/dc_blocker.h:79-88) and slicer ---- */
    {
        int dc_x = S.dc_x, dc_y = S.dc_y, dc_acc = S.dc_acc;
        __syncwarp();                               /* every lane has read the state before lane 0 writes it back below */
        for (unsigned g0 = 0; g0 < n; g0 += 32) {
            const unsigned idx = g0 + lane;
            int v = idx < n ? (int)x[idx] : 0;
            if (use_dc) {
                const unsigned cnt = min(32u, n - g0);
                int mine = v;
                for (unsigned t = 0; t < cnt; t++) {
                    const int sample = __shfl_sync(0xffffffffu, v, t);
                    dc_acc -= dc_x;
                    dc_x = sample << 14;
                    dc_acc += dc_x - dc_p * dc_y;
                    dc_y = dc_acc >> 14;
                    if ((unsigned)lane == t) mine = (int)(short)dc_y;
                }
                v = mine;
                if (idx < n) x[idx] = (short)v;
            }
            const unsigned w = __ballot_sync(0xffffffffu, idx < n && v < 0);
            if (lane == 0) bits[g0 >> 5] = w;
        }
        if (lane == 0) { bits[(n + 31) >> 5] = 0; S.dc_x = dc_x; S.dc_y = dc_y; S.dc_acc = dc_acc; }
        __syncwarp();
    }

    /* ---- pass 2: the state machine ---- */
    const unsigned spbs[3] = { 75, 32, 16 }, bauds[3] = { 512, 1200, 2400 }, bases[3] = { 0, 75, 75 + 32 };
    unsigned i = 0;
    if (S.state == ST_SYNCHRONIZED) { if (lane == 0) S.state = ST_BATCH; __syncwarp(); }
    while (i < n) {
        const int state = S.state;
        __syncwarp();                               /* scalars are read by every lane before lane 0 updates them further down */
        if (state == ST_SEARCH) {
            const unsigned cnt = min(32u, n - i);
            const unsigned word = __funnelshift_r(bits[i >> 5], bits[(i >> 5) + 1], i & 31);
            const unsigned bit = (word >> lane) & 1u;
            unsigned mask[3], matches[3], cur[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const unsigned spb = spbs[b];
                cur[b] = S.eye_cur[b]; matches[b] = S.eye_matches[b];
                bool m = false;
                /* spb = 16: two samples of a step share a register -- lanes 0-15 first, then 16-31 */
                for (int half = 0; half < (spb < 32 ? 2 : 1); half++) {
                    const bool mine = (unsigned)lane < cnt && (spb >= 32 || (lane >> 4) == half);
                    if (mine) {
                        const unsigned ph = (cur[b] + (unsigned)lane) % spb;
                        unsigned r = S.eye_reg[bases[b] + ph];
                        r = (r << 1) | bit;
                        S.eye_reg[bases[b] + ph] = r;
                        m = sync_ok(r);
                    }
                    __syncwarp();
                }
                mask[b] = __ballot_sync(0xffffffffu, m);
            }
            /* run-length logic in sample order; stops as soon as nothing more can happen in this step */
            unsigned consumed = cnt;
            bool trig = false;
            unsigned t_skip = 0, t_baud = 0, t_bskip = 0;
            for (unsigned sidx = 0; sidx < cnt; sidx++) {
                if ((((mask[0] | mask[1] | mask[2]) >> sidx) == 0) && (matches[0] | matches[1] | matches[2]) == 0) break;
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    if ((mask[b] >> sidx) & 1u) matches[b]++;
                    else if (matches[b] > spbs[b] / 2) { trig = true; t_skip = spbs[b]; t_baud = bauds[b]; t_bskip = (matches[b] / 2) & 0xffffu; }
                    else matches[b] = 0;
                }
                if (trig) { consumed = sidx + 1; break; }
            }
            __syncwarp();
            if (trig) {
                if (lane < 16) S.batch[lane] = 0;                   /* batch_reset */
                if (lane == 0) {
                    S.sample_skip = t_skip; S.baud = t_baud;
                    S.b_word = 0; S.b_word_bit = 0; S.b_bits = 0; S.b_skip = t_bskip;
                    S.state = ST_BATCH;                             /* SYNCHRONIZED for the rest of this sample, BATCH from the next */
                }
            }
            if (lane < 3) { S.eye_cur[lane] = (cur[lane] + consumed) % spbs[lane]; S.eye_matches[lane] = matches[lane]; }
            i += consumed;
            __syncwarp();
        } else {
            /* BATCH (LSB-first into batch[]) or SYNCWORD (MSB-first into s_word): up to 32 bits, one every sample_skip samples */
            const bool batch = state == ST_BATCH;
            const unsigned spb = S.sample_skip;
            const unsigned skip = batch ? S.b_skip : S.s_skip;
            const unsigned have = batch ? S.b_word_bit : S.s_bits;
            __syncwarp();
            const unsigned long long first = (unsigned long long)i + until_sampling(spb, skip) - 1;
            const unsigned long long pos = first + (unsigned long long)lane * spb;
            const bool valid = (unsigned)lane < 32 - have && pos < n;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const unsigned w = __ballot_sync(0xffffffffu, valid && bit_at(bits, pos));
            const unsigned nv = __popc(vmask);
            if (nv == 0) {                                          /* the feed ends before the next bit */
                if (lane == 0) { if (batch) S.b_skip = (skip + (n - i)) & 0xffffu; else S.s_skip = (skip + (n - i)) & 0xffffu; }
                i = n;
                __syncwarp();
                continue;
            }
            const unsigned long long last = first + (unsigned long long)(nv - 1) * spb;
            i = (unsigned)(last + 1);
            if (batch) {
                if (lane == 0) {
                    S.batch[S.b_word] |= w << have;
                    S.b_bits = (S.b_bits + nv) & 0xffffu; S.b_skip = 0;
                    unsigned wb = have + nv;
                    if (wb == 32) {
                        wb = 0;
                        if (++S.b_word == 16) {
                            process_batch(S, alpha, numeric, sink, c);
                            S.state = ST_SYNCWORD;
                            S.b_word = 0;
                            S.s_skip = 0; S.s_bits = 0; S.s_word = 0;
                        }
                    }
                    S.b_word_bit = wb;
                }
            } else if (lane == 0) {
                const unsigned msb_first = __brev(w) >> (32 - nv);  /* first sampled bit ends up highest */
                S.s_word = (nv == 32) ? msb_first : ((S.s_word << nv) | msb_first);
                S.s_skip = 0;
                S.s_bits += nv;
            }
            __syncwarp();
            if (!batch && S.s_bits == 32) {
                const bool ok = sync_ok(S.s_word);
                __syncwarp();
                if (!ok) {
                    for (int r = lane; r < 75 + 32 + 16; r += 32) S.eye_reg[r] = 0;     /* eyes_reset */
                    if (lane < 3) { S.eye_cur[lane] = 0; S.eye_matches[lane] = 0; }
                    if (lane == 0) { S.state = ST_SEARCH; S.sample_skip = 0; deliver(S, alpha, numeric, sink, c); }
                } else {
                    if (lane < 16) S.batch[lane] = 0;                                   /* batch_reset */
                    if (lane == 0) { S.state = ST_BATCH; S.b_word = 0; S.b_word_bit = 0; S.b_skip = 0; S.b_bits = 0; }
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    {
        unsigned *dst = reinterpret_cast<unsigned *>(states + c);
        const unsigned *src = reinterpret_cast<const unsigned *>(&S);
        for (int k = lane; k < (int)(sizeof(PocsagState) / 4); k += 32) dst[k] = src[k];
    }
}


/* ============================================================================================== */
/* FLEX (pager/pager_flex.c).  State mirrors struct pager_flex (pager/pager_flex_priv.h:238-326).   */
/* ============================================================================================== */
enum { FX_SYNC_1 = 0, FX_SYNC_2 = 1, FX_BLOCK = 2 };
enum { FS_SEARCH_BS1 = 0, FS_BS1, FS_A, FS_B, FS_INV_A, FS_FIW, FS_SYNCED };
enum { F2_COMMA = 0, F2_C, F2_INV_COMMA, F2_INV_C, F2_SYNCED };

/* pager_flex.c:47-96 */
struct FlexCoding { unsigned short seq_a, baud; unsigned char fsk_levels, sample_skip, sync_2_samples, sym_bits, sample_fudge;
                    unsigned short symbols_per_block; unsigned char nr_phases; };
__constant__ FlexCoding c_flex_codings[4] = {
    { 0x78f3, 1600, 2, 9,  4, 1, 0, 2816, 1 },
    { 0x84e7, 3200, 2, 4, 24, 1, 2, 5632, 2 },
    { 0x4f97, 3200, 4, 9, 12, 2, 0, 2816, 2 },
    { 0x215f, 6400, 4, 4, 32, 2, 2, 5632, 4 },
};
__constant__ char c_flex_num_lut[16] = { '0','1','2','3','4','5','6','7','8','9','X','U',' ','-',']','[' };   /* :686-704 */

/* struct pager_flex_block as the reference lays it out in memory, in 32-bit words: 4 phases x (88 words + one word
 * of cur_bit | cur_word << 8 | base_word << 16), nr_symbols, phase_ff.  The reference walks phase_words[] with
 * unchecked offsets read off the air (start word <= 127, length <= 127), so reads and in-place BCH fix-ups may land
 * in a later phase or in the bookkeeping words; one flat array reproduces that.  Offsets past the block (the
 * reference would read its heap) end the vector instead. */
constexpr int FX_PHASE_STRIDE = 89;
constexpr int FX_BLOCK_WORDS = 4 * FX_PHASE_STRIDE + 2;

struct FlexState {
    int sample_range, sample_delta;             /* int16 values */
    int state, skip, skip_count;
    uint32_t cycle_id, frame_id;
    int sync_state;
    uint32_t sample_counter, bit_counter;       /* uint8 in the reference */
    uint32_t a, b, inv_a, fiw;
    int coding;
    int sum_high, sum_low;
    uint32_t cnt_high, cnt_low;
    int s2_state;
    uint32_t nr_dots, c, inv_c, nr_c;
    int nr_symbols, phase_ff;
    uint32_t msg_len;
    int dc_x, dc_y, dc_acc;
    uint32_t sync_words[10];
    uint8_t cur_bit[4], cur_word[4], base_word[4];
    uint32_t blk[FX_BLOCK_WORDS];
    char msg_buf[256];
};

__device__ __forceinline__ uint32_t fx_cksum(uint32_t w)                       /* :107-119 */
{
    uint32_t s = 0;
    w &= 0x1fffff;
    for (int i = 0; i < 6; i++) { s += w & 0xf; w >>= 4; }
    return s & 0xf;
}

__device__ __forceinline__ int fx_slice(const FlexState &f, int sample)        /* :129-171 */
{
    if (c_flex_codings[f.coding].fsk_levels == 2) return sample >= 0 ? 1 : 0;
    sample = (int)(short)(sample - f.sample_delta);
    if (sample < 0) return (-sample > f.sample_range / 4) ? 0 : 1;
    return (sample > f.sample_range / 4) ? 2 : 3;
}

__device__ void fx_sync_reset(FlexState &f)                                    /* :209-233 */
{
    for (int i = 0; i < 10; i++) f.sync_words[i] = 0;
    f.sync_state = FS_BS1;
    f.sample_counter = 0; f.bit_counter = 0;
    f.a = 0; f.b = 0; f.inv_a = 0; f.fiw = 0; f.coding = -1;
    f.sum_high = f.sum_low = 0; f.cnt_high = f.cnt_low = 0;
}

__device__ void fx_reset(FlexState &f)                                         /* :173-262 */
{
    f.state = FX_SYNC_1;
    f.skip = 0; f.skip_count = 0;
    f.sample_range = 0; f.sample_delta = 0;
    f.frame_id = 0; f.cycle_id = 0;
    fx_sync_reset(f);
    f.s2_state = F2_COMMA; f.nr_dots = 0; f.c = 0; f.inv_c = 0; f.nr_c = 0;
    f.nr_symbols = 0; f.phase_ff = 0;
    for (int i = 0; i < 4; i++) { f.cur_bit[i] = 0; f.cur_word[i] = 0; f.base_word[i] = 0; }
}

__global__ void flex_init_kernel(FlexState *states, int nr_channels)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nr_channels) fx_reset(states[c]);
}

__device__ gpupager_msg *fx_msg(const FlexState &f, const MsgSink &sink, int c, uint32_t kind, uint32_t phase, uint64_t capcode)
{
    const uint32_t slot = sink.count[c];
    if (slot >= sink.cap) { atomicAdd(sink.dropped, 1ull); return nullptr; }
    gpupager_msg *m = sink.msgs + (size_t)c * sink.cap + slot;
    m->channel = sink.chan_id ? sink.chan_id[c] : (unsigned)c; m->kind = kind; m->baud = c_flex_codings[f.coding].baud;
    m->capcode = (uint32_t)capcode; m->capcode_hi = (uint32_t)(capcode >> 32);
    m->function = phase; m->len = 0;
    m->aux[0] = f.cycle_id; m->aux[1] = f.frame_id; m->aux[2] = m->aux[3] = m->aux[4] = m->aux[5] = 0;
    m->text[0] = 0;
    sink.count[c] = slot + 1;
    return m;
}

__device__ void fx_msg_text(gpupager_msg *m, const FlexState &f)
{
    if (!m) return;
    m->len = f.msg_len;
    for (uint32_t i = 0; i < f.msg_len; i++) m->text[i] = f.msg_buf[i];
    m->text[f.msg_len] = 0;
}

__device__ __forceinline__ void fx_range(FlexState &f, int sample)             /* :352-358 and twins */
{
    if (sample > 0) { f.sum_high += sample; f.cnt_high++; }
    else            { f.sum_low += sample;  f.cnt_low++; }
}

__device__ void fx_sync_update(FlexState &f, int sample)                       /* :295-458 */
{
    f.sample_counter = (f.sample_counter + 1) % 10;
    const uint32_t sym = sample >= 0 ? 1u : 0u;
    switch (f.sync_state) {
    case FS_SEARCH_BS1: {
        uint32_t &w = f.sync_words[f.sample_counter];
        w = (w << 1) | sym;
        if (w == 0xaaaaaaaau) { f.bit_counter = 1; f.sync_state = FS_BS1; }
        break;
    }
    case FS_BS1: {
        uint32_t &w = f.sync_words[f.sample_counter];
        w = (w << 1) | sym;
        if (w == 0xaaaaaaaau) {
            f.bit_counter = (f.bit_counter + 1) & 0xff;
        } else {
            if (f.bit_counter < 3) f.sync_state = FS_SEARCH_BS1;
            else { f.sync_state = FS_A; f.sample_counter = f.bit_counter / 2; }
            f.bit_counter = 0;
        }
        break;
    }
    case FS_A:
        if (f.sample_counter == 0) {
            f.a = (f.a << 1) | sym;
            fx_range(f, sample);
            if (++f.bit_counter == 32) { f.sync_state = FS_B; f.bit_counter = 0; }
        }
        break;
    case FS_B:
        if (f.sample_counter == 0) {
            f.b = ((f.b << 1) | sym) & 0xffffu;
            fx_range(f, sample);
            if (++f.bit_counter == 16) { f.sync_state = FS_INV_A; f.bit_counter = 0; }
        }
        break;
    case FS_INV_A:
        if (f.sample_counter == 0) {
            f.inv_a = (f.inv_a << 1) | sym;
            fx_range(f, sample);
            if (++f.bit_counter == 32) {
                /* _pager_flex_sync_check_baud (:264-287): only the A word decides -- the reference's test on the
                 * inverted word is evaluated in int, where ~seq_a has 16 extra one bits, and can never pass. */
                const uint32_t coding_a = (f.a >> 16) & 0xffffu;
                int found = -1;
                for (int i = 0; i < 4 && found < 0; i++)
                    if (__popc((uint32_t)c_flex_codings[i].seq_a ^ coding_a) < 4) found = i;
                if (found >= 0) { f.coding = found; f.sync_state = FS_FIW; }
                else fx_sync_reset(f);
                f.bit_counter = 0;
            }
        }
        break;
    case FS_FIW:
        if (f.sample_counter == 0) {
            f.fiw = (f.fiw >> 1) | (sym << 31);
            fx_range(f, sample);
            if (++f.bit_counter == 32) {
                /* :438-442.  A zero count is a division by zero (SIGFPE) in the reference; the frame is dropped here. */
                if (f.cnt_high == 0 || f.cnt_low == 0) { fx_reset(f); break; }
                const int hi = (int)(short)(f.sum_high / (int)f.cnt_high), lo = (int)(short)(f.sum_low / (int)f.cnt_low);
                f.sample_range = (int)(short)(hi - lo);
                f.sample_delta = (int)(short)(hi - f.sample_range / 2);
                f.sync_state = FS_SYNCED;
            }
        }
        break;
    default:
        break;
    }
}

__device__ void fx_sync2_update(FlexState &f, int sample)                      /* :460-525 */
{
    const FlexCoding &cd = c_flex_codings[f.coding];
    switch (f.s2_state) {
    case F2_COMMA:
        f.nr_dots = (f.nr_dots + 1) & 0xffffu;
        if (cd.sync_2_samples == f.nr_dots) f.s2_state = F2_C;
        break;
    case F2_C:
        f.c = ((f.c << cd.sym_bits) | (uint32_t)fx_slice(f, sample)) & 0xffffu;
        f.nr_c += cd.sym_bits;
        if (f.nr_c == 16) { f.s2_state = F2_INV_COMMA; f.nr_dots = 0; }
        break;
    case F2_INV_COMMA:
        f.nr_dots = (f.nr_dots + 1) & 0xffffu;
        if (cd.sync_2_samples == f.nr_dots) { f.s2_state = F2_INV_C; f.nr_c = 0; }
        break;
    case F2_INV_C:
        f.inv_c = ((f.inv_c << cd.sym_bits) | (uint32_t)fx_slice(f, sample)) & 0xffffu;
        f.nr_c += cd.sym_bits;
        if (f.nr_c == 16) f.s2_state = F2_SYNCED;
        break;
    default:
        break;
    }
}

__device__ __forceinline__ bool fx_in(size_t idx) { return idx < (size_t)FX_BLOCK_WORDS; }

/* :597-681 */
__device__ int fx_alnum(FlexState &f, const MsgSink &sink, int c, uint32_t phase, uint64_t capcode, uint32_t long_word,
                        size_t base, size_t nr_words)
{
    size_t first_char_word = 1;
    int skip_word = 0;
    uint32_t status;
    if (long_word != 0xffffffffu) { first_char_word = 0; status = long_word; }
    else {
        if (!fx_in(base)) return -1;
        status = f.blk[base];
        if (bch_decode(status)) return -1;
    }
    const uint32_t fragment = (status >> 10) & 1;
    const uint32_t seq = (status >> 11) & 3;
    uint32_t maildrop = 0;
    if (seq == 3) { skip_word = 1; maildrop = (status >> 20) & 1; }
    for (size_t i = first_char_word; i < nr_words; i++) {
        if (!fx_in(base + i)) return -1;
        uint32_t cw = f.blk[base + i];
        if (bch_decode(cw)) return -1;
        if (skip_word) cw >>= 7;
        for (int j = skip_word; j < 3; j++) {
            const uint32_t ch = cw & 0x7f;
            if (ch != 0x3) f.msg_buf[f.msg_len++] = (char)ch; else break;
            if (f.msg_len == 255) break;
            cw >>= 7;
        }
        skip_word = 0;
        if (f.msg_len == 255) break;
    }
    gpupager_msg *m = fx_msg(f, sink, c, GPUPAGER_MSG_FLEX_ALNUM, phase, capcode);
    if (m) { m->aux[2] = fragment; m->aux[3] = maildrop; m->aux[4] = seq; }
    fx_msg_text(m, f);
    return 0;
}

/* :709-824 */
__device__ int fx_numeric(FlexState &f, const MsgSink &sink, int c, uint32_t phase, uint64_t capcode, uint32_t long_word,
                          size_t base, size_t nr_words)
{
    uint32_t cur = 0, next = 0;
    unsigned long long nr_bits = (unsigned long long)nr_words * 21, cur_bits = 19, next_offs = 0, next_bits = 21;
    if (long_word != 0xffffffffu) {
        cur = (long_word & 0x1fffff) >> 2;
        nr_bits += 19; cur_bits = 19; next_offs = 0;
    } else {
        if (!fx_in(base)) return -1;
        cur = f.blk[base];
        if (bch_decode(cur)) return -1;
        cur &= 0x1fffff; cur >>= 2;
        cur_bits = 19; nr_bits -= 2; next_offs = 1;
    }
    if (next_offs < nr_words) {
        if (!fx_in(base + next_offs)) return -1;
        next = f.blk[base + next_offs];
        if (bch_decode(next)) return -1;
        next_bits = 21; next &= 0x1fffff;
    }
    nr_bits &= ~3ull;
    do {
        const unsigned long long rem = cur_bits & ~3ull;
        for (unsigned long long i = 0; i < rem; i += 4) {
            f.msg_buf[f.msg_len++] = c_flex_num_lut[cur & 0xf];
            if (f.msg_len == 255) break;
            cur >>= 4; cur_bits -= 4; nr_bits -= 4;
        }
        if (f.msg_len == 255) break;
        if (cur_bits != 0 && nr_bits != 0) {
            switch (cur_bits) {
            case 1: cur |= (next & 0x7) << 1; next >>= 3; next_bits -= 3; break;
            case 2: cur |= (next & 0x3) << 2; next >>= 2; next_bits -= 2; break;
            case 3: cur |= (next & 0x1) << 3; next >>= 1; next_bits -= 1; break;
            }
            cur_bits = 4;
        } else if (cur_bits == 0 && nr_bits != 0) {
            cur = next; cur_bits = next_bits; next_bits = 21; next_offs++;
            if (next_offs < nr_words) {
                if (!fx_in(base + next_offs)) return -1;
                next = f.blk[base + next_offs];
                if (bch_decode(next)) return -1;
                next &= 0x1fffff;
            }
        }
    } while (nr_bits != 0);
    fx_msg_text(fx_msg(f, sink, c, GPUPAGER_MSG_FLEX_NUM, phase, capcode), f);
    return 0;
}

/* :829-883 */
__device__ int fx_tone(FlexState &f, const MsgSink &sink, int c, uint32_t phase, uint64_t capcode, uint32_t first, uint32_t second)
{
    first &= 0x1fffff;
    switch ((first >> 7) & 3) {
    case 0:
        first >>= 9;
        for (int i = 0; i < 3; i++) { f.msg_buf[f.msg_len++] = c_flex_num_lut[first & 0xf]; first >>= 4; }
        if (second != 0xffffffffu) {
            second &= 0x1fffff;
            for (int i = 0; i < 5; i++) { f.msg_buf[f.msg_len++] = c_flex_num_lut[second & 0xf]; second >>= 4; }
        }
        fx_msg_text(fx_msg(f, sink, c, GPUPAGER_MSG_FLEX_NUM, phase, capcode), f);
        return 0;
    case 1: case 2: return 0;           /* the reference only logs these */
    default: return -1;
    }
}

/* :938-1033 */
__device__ int fx_vector(FlexState &f, const MsgSink &sink, int c, uint32_t phase, uint64_t capcode, size_t vi, size_t nr_vec,
                         size_t base)
{
    f.msg_len = 0;
    for (size_t i = 0; i < nr_vec; i++) {
        if (!fx_in(vi + i)) return -1;
        if (bch_decode(f.blk[vi + i])) return -1;
    }
    const uint32_t vec = f.blk[vi];
    if (fx_cksum(vec) != 0xf) return -1;
    const uint32_t type = (vec >> 4) & 7;
    const size_t start = (vec >> 7) & 0x7f;
    const uint32_t long_word = (nr_vec == 2) ? f.blk[vi + 1] : 0xffffffffu;
    size_t len;
    switch (type) {
    case 2: return fx_tone(f, sink, c, phase, capcode, vec, long_word);
    case 3:
        len = ((vec >> 14) & 7) + 1;
        if (nr_vec == 2) len -= 1;
        return fx_numeric(f, sink, c, phase, capcode, long_word, base + start, len);
    case 5:
        len = (vec >> 14) & 0x7f;
        if (nr_vec == 2) len -= 1;      /* wraps for length 0, like the reference's size_t */
        return fx_alnum(f, sink, c, phase, capcode, long_word, base + start, len);
    case 1: {                           /* short instruction vector, :885-933 */
        const uint32_t v = vec & 0x7fffff;
        if (fx_cksum(v) != 0xf) return -1;
        gpupager_msg *m = fx_msg(f, sink, c, GPUPAGER_MSG_FLEX_SIV, phase, capcode);
        if (m) { m->aux[2] = (v >> 7) & 7; m->aux[3] = (v >> 10) & 0x7ff; }
        return 0;
    }
    default: return 0;                  /* unsupported vector types are logged only */
    }
}

/* :1088-1198 with _pager_flex_decode_address :527-573 */
__device__ void fx_phase_process(FlexState &f, const MsgSink &sink, int c, uint32_t ph)
{
    const size_t base = (size_t)ph * FX_PHASE_STRIDE;
    uint32_t biw = f.blk[base] & 0x7fffffffu;
    if (bch_decode(biw)) return;
    if (fx_cksum(biw) != 0xf) return;
    const uint32_t vsw = (biw >> 10) & 0x3f, eob = (biw >> 8) & 3;
    if (eob > vsw) return;
    const size_t addr_start = 1 + (size_t)eob;
    for (size_t i = addr_start; i < vsw; i++) {
        const size_t vec_offs = i + vsw - addr_start;
        const size_t ai = base + i;
        if (!fx_in(ai)) return;
        if (bch_decode(f.blk[ai])) return;
        const uint32_t first = f.blk[ai] &= 0x1fffff;
        uint64_t capcode;
        size_t nr_words = 0;
        if ((first > 0x8000 && first <= 0x1e0000) || (first > 0x1f0000 && first < 0x1f7fff)) {
            capcode = first - 32768;
        } else {
            if (!fx_in(ai + 1)) return;
            if (bch_decode(f.blk[ai + 1])) return;
            const uint32_t second = f.blk[ai + 1] &= 0x1fffff;
            nr_words = 1;
            capcode = (uint32_t)(0x1f9001u + (((0x1fffffu - second) * 32768u) + first - 1u));     /* 32-bit arithmetic */
        }
        (void)fx_vector(f, sink, c, ph, capcode, base + vec_offs, nr_words + 1, base);
        i += nr_words;
    }
}

__device__ __forceinline__ void fx_append_bit(FlexState &f, int ph, bool bit)  /* :1200-1222 */
{
    uint32_t &w = f.blk[ph * FX_PHASE_STRIDE + f.base_word[ph] + f.cur_word[ph]];
    w = (w >> 1) | ((uint32_t)bit << 31);
    f.cur_word[ph] = (uint8_t)((f.cur_word[ph] + 1) % 8);
    if (f.cur_word[ph] == 0) f.cur_bit[ph]++;
    if (f.cur_bit[ph] == 32) { f.base_word[ph] += 8; f.cur_bit[ph] = 0; f.cur_word[ph] = 0; }
}

__device__ void fx_block_update(FlexState &f, const MsgSink &sink, int c, int sample)    /* :1224-1310 */
{
    const FlexCoding &cd = c_flex_codings[f.coding];
    const int sym = fx_slice(f, sample);
    switch (cd.nr_phases) {
    case 1: fx_append_bit(f, 0, sym == 1); break;
    case 2:
        if (cd.fsk_levels == 2) { fx_append_bit(f, f.phase_ff ? 2 : 0, sym == 1); f.phase_ff = !f.phase_ff; }
        else { fx_append_bit(f, 0, (sym & 2) != 0); fx_append_bit(f, 2, (sym & 1) != 0); }
        break;
    default:
        if (!f.phase_ff) { fx_append_bit(f, 0, (sym & 2) != 0); fx_append_bit(f, 1, (sym & 1) != 0); }
        else             { fx_append_bit(f, 2, (sym & 2) != 0); fx_append_bit(f, 3, (sym & 1) != 0); }
        f.phase_ff = !f.phase_ff;
        break;
    }
    if (++f.nr_symbols == cd.symbols_per_block) {
        for (int p = 0; p < 4; p++)
            f.blk[p * FX_PHASE_STRIDE + 88] = f.cur_bit[p] | (uint32_t)f.cur_word[p] << 8 | (uint32_t)f.base_word[p] << 16;
        f.blk[4 * FX_PHASE_STRIDE] = (uint32_t)f.nr_symbols;
        f.blk[4 * FX_PHASE_STRIDE + 1] = (uint32_t)f.phase_ff;
        if (cd.nr_phases == 1) fx_phase_process(f, sink, c, 0);
        else if (cd.nr_phases == 2) { fx_phase_process(f, sink, c, 0); fx_phase_process(f, sink, c, 2); }
        else for (uint32_t p = 0; p < 4; p++) fx_phase_process(f, sink, c, p);
        fx_reset(f);
    }
}

/* pager_flex_on_pcm (:1401-1455), one WARP per channel.
 *   pass 1  the optional DC blocker over the whole feed (every lane computes the identical recurrence);
 *   pass 2  the state machine.  Where a receiver spends nearly all of its time -- searching for bit sync, Sync 1 state
 *           SEARCH_BS1 (:296-345): every sample shifts into one of 10 phase registers, round robin, and is compared with
 *           0xaaaaaaaa -- the warp takes 30 samples per step: lane j handles sample j, in three sub-passes of ten lanes
 *           (ten distinct registers each); on a hit the lanes behind it in the same sub-pass put their register back and
 *           lane 0 carries on sample by sample with the reference's logic (fx_sync_update and friends, unchanged) until
 *           the machine is back in SEARCH_BS1.  Those sequential stretches (one frame: 1.9 s of signal) see every
 *           (skip + 1)-th sample only. */
constexpr int FW_WARPS = 4;

__device__ void flex_step_sequential(FlexState &f, const MsgSink &sink, int c, int sample)
{
    if (f.skip_count != 0) { f.skip_count--; return; }
    f.skip_count = f.skip;
    switch (f.state) {
    case FX_SYNC_1:
        fx_sync_update(f, sample);
        if (f.sync_state == FS_SYNCED) {
            /* _pager_flex_handle_fiw :1312-1345 */
            uint32_t fiw = f.fiw & 0x7fffffffu;
            bool ok = bch_decode(fiw) == 0;
            if (ok) {
                f.cycle_id = (fiw >> 4) & 0xf;
                f.frame_id = (fiw >> 8) & 0x7f;
                ok = fx_cksum(fiw) == 0xf;
            }
            if (ok) {
                f.state = FX_SYNC_2;
                f.skip = c_flex_codings[f.coding].sample_skip;
                f.skip_count = f.skip + c_flex_codings[f.coding].sample_fudge;
            } else {
                fx_reset(f);
            }
        }
        break;
    case FX_SYNC_2:
        fx_sync2_update(f, sample);
        if (f.s2_state == F2_SYNCED) f.state = FX_BLOCK;
        break;
    default:
        fx_block_update(f, sink, c, sample);
        break;
    }
}

__device__ __forceinline__ bool flex_searching(const FlexState &f)
{
    return f.state == FX_SYNC_1 && f.sync_state == FS_SEARCH_BS1 && f.skip == 0 && f.skip_count == 0;
}

__global__ void __launch_bounds__(32 * FW_WARPS) flex_kernel(FlexState *states, int nr_channels, short *pcm,
                                                             long long pitch, unsigned n, MsgSink sink, int use_dc, int dc_p)
{
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * FW_WARPS + (threadIdx.x >> 5);
    if (c >= nr_channels) return;                   /* whole warps leave */
    FlexState &f = states[c];
    short *x = pcm + (size_t)c * pitch;

    if (use_dc) {                                   /* filter/dc_blocker.h:79-88 */
        int dc_x = f.dc_x, dc_y = f.dc_y, dc_acc = f.dc_acc;
        __syncwarp();
        for (unsigned g0 = 0; g0 < n; g0 += 32) {
            const unsigned idx = g0 + lane, cnt = min(32u, n - g0);
            const int v = idx < n ? (int)x[idx] : 0;
            int mine = v;
            for (unsigned t = 0; t < cnt; t++) {
                const int sample = __shfl_sync(0xffffffffu, v, t);
                dc_acc -= dc_x;
                dc_x = sample << 14;
                dc_acc += dc_x - dc_p * dc_y;
                dc_y = dc_acc >> 14;
                if ((unsigned)lane == t) mine = (int)(short)dc_y;
            }
            if (idx < n) x[idx] = (short)mine;
        }
        __syncwarp();
        if (lane == 0) { f.dc_x = dc_x; f.dc_y = dc_y; f.dc_acc = dc_acc; }
        __syncwarp();
    }

    unsigned i = 0;
    while (i < n) {
        if (flex_searching(f)) {
            /* ---- 30 samples of bit-sync search ---- */
            const unsigned cnt = min(30u, n - i);
            const unsigned sc0 = f.sample_counter;
            const bool have = (unsigned)lane < cnt;
            const uint32_t sym = (have && x[i + lane] >= 0) ? 1u : 0u;
            const unsigned ph = (sc0 + 1 + (unsigned)lane) % 10;
            int hit_at = -1;
            __syncwarp();
            for (int sub = 0; sub < 3 && hit_at < 0; sub++) {
                const bool mine = have && lane / 10 == sub;
                uint32_t old = 0, w = 0;
                if (mine) { old = f.sync_words[ph]; w = (old << 1) | sym; f.sync_words[ph] = w; }
                const unsigned hits = __ballot_sync(0xffffffffu, mine && w == 0xaaaaaaaau);
                if (hits) {
                    hit_at = __ffs(hits) - 1;
                    if (mine && lane > hit_at) f.sync_words[ph] = old;      /* samples behind the hit have not happened yet */
                }
                __syncwarp();
            }
            if (hit_at >= 0) {
                if (lane == 0) { f.sample_counter = (sc0 + 1 + (unsigned)hit_at) % 10; f.bit_counter = 1; f.sync_state = FS_BS1; }
                i += (unsigned)hit_at + 1;
            } else {
                if (lane == 0) f.sample_counter = (sc0 + cnt) % 10;
                i += cnt;
            }
            __syncwarp();
        } else {
            /* ---- lane 0 walks the reference's state machine until it is searching again ---- */
            unsigned j = i;
            if (lane == 0) {
                do {
                    flex_step_sequential(f, sink, c, (int)x[j]);
                    j++;
                } while (j < n && !flex_searching(f));
            }
            i = __shfl_sync(0xffffffffu, j, 0);
            __syncwarp();
        }
    }
}

} // namespace

/* ------------------------------------------------------------------------------------------ */
struct gpupager {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_own = nullptr, ev_ext = nullptr;
    int C = 0, nr_src = 0;
    unsigned interp = 1, decim = 1;
    int M = 0;                          /* taps per phase */
    uint32_t flags = 0;
    size_t max_feed = 0;
    int dc_p = 0;

    short *d_phase = nullptr;           /* [interp][M] */
    short *d_carry[2] = { nullptr, nullptr };
    long long carry_pitch = 0, carry_len = 0;
    int pp = 0;
    unsigned long long in_base = 0;     /* global index of window sample 0 */
    unsigned long long total_in = 0;    /* samples fed so far */
    unsigned long long m_next = 0;      /* next resampled output index */
    short *d_stage = nullptr;           /* host feeds: [C][max_feed] */
    short *d_res = nullptr;             /* resampled output of the last feed [C][res_pitch] */
    long long res_pitch = 0;
    size_t last_out = 0;
    int decoder = GPUPAGER_DECODER_POCSAG;
    PocsagState *d_states = nullptr;
    unsigned *d_bits = nullptr;         /* POCSAG: sliced bits of the current feed, [C][bits_pitch] words */
    unsigned bits_pitch = 0;
    int rs_span_cap = 0;                /* shorts of shared memory one resampler block stages */
    FlexState *d_fstates = nullptr;
    char *d_text = nullptr;
    uint32_t msg_cap = 32;
    gpupager_msg *d_msgs = nullptr;
    uint32_t *d_count = nullptr;
    unsigned long long *d_dropped = nullptr;
    unsigned *d_map = nullptr;          /* [C] source row per decoder channel (cfg.channel_map), nullptr = identity */
    std::vector<gpupager_msg> queue;    /* decoded, not yet handed out */
    uint64_t launches = 0, dropped = 0;
};

static void pager_free(gpupager *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_phase); cudaFree(h->d_carry[0]); cudaFree(h->d_carry[1]); cudaFree(h->d_stage); cudaFree(h->d_res);
    cudaFree(h->d_states); cudaFree(h->d_fstates); cudaFree(h->d_text); cudaFree(h->d_msgs); cudaFree(h->d_count); cudaFree(h->d_dropped);
    cudaFree(h->d_map); cudaFree(h->d_bits);
    if (h->ev_own) cudaEventDestroy(h->ev_own);
    if (h->ev_ext) cudaEventDestroy(h->ev_ext);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int gpupager_create(gpupager_t **ph, const gpupager_cfg *cfg)
{
    if (!ph || !cfg) return perr(GPUPAGER_E_BADARGS, "null argument");
    *ph = nullptr;
    if (cfg->struct_size != sizeof(gpupager_cfg)) return perr(GPUPAGER_E_BADARGS, "gpupager_cfg size mismatch");
    const bool bypass = (cfg->flags & GPUPAGER_F_NO_RESAMPLE) != 0;
    if (cfg->decoder != GPUPAGER_DECODER_POCSAG && cfg->decoder != GPUPAGER_DECODER_FLEX)
        return perr(GPUPAGER_E_BADARGS, "unknown decoder %u", cfg->decoder);
    if (!cfg->nr_channels || !cfg->max_feed_samples) return perr(GPUPAGER_E_BADARGS, "incomplete configuration");
    if (!bypass && (!cfg->interpolate || !cfg->decimate || !cfg->nr_taps || !cfg->taps))
        return perr(GPUPAGER_E_BADARGS, "resampler needs interpolate, decimate and taps");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return perr(GPUPAGER_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return perr(GPUPAGER_E_BADARGS, "bad device ordinal");
    PCUDA(cudaSetDevice(cfg->device));

    gpupager *h = new (std::nothrow) gpupager();
    if (!h) return perr(GPUPAGER_E_NOMEM, "out of memory");
    h->device = cfg->device; h->C = (int)cfg->nr_channels; h->flags = cfg->flags; h->max_feed = cfg->max_feed_samples;
    h->decoder = (int)cfg->decoder;
    if (cfg->flags & GPUPAGER_F_DC_BLOCK) {
        const double pole = cfg->dc_pole != 0.0 ? cfg->dc_pole : 0.9999;
        h->dc_p = (int16_t)((1.0 - pole) * (double)(1 << 14));      /* filter/dc_blocker.h:57 */
    }
#define PFAIL(expr)                                                                                 \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            perr(GPUPAGER_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            pager_free(h);                                                                          \
            return GPUPAGER_E_CUDA;                                                                 \
        }                                                                                           \
    } while (0)
    PFAIL(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    PFAIL(cudaEventCreateWithFlags(&h->ev_own, cudaEventDisableTiming));
    PFAIL(cudaEventCreateWithFlags(&h->ev_ext, cudaEventDisableTiming));

    const int C = h->C;
    if (cfg->channel_map) {
        PFAIL(cudaMalloc(&h->d_map, C * sizeof(unsigned)));
        PFAIL(cudaMemcpy(h->d_map, cfg->channel_map, C * sizeof(unsigned), cudaMemcpyHostToDevice));
    }
    size_t max_out = h->max_feed;
    if (!bypass) {
        h->interp = cfg->interpolate; h->decim = cfg->decimate;
        size_t m = (cfg->nr_taps + h->interp - 1) / h->interp;      /* polyphase_fir.c:70 */
        m = (m + 3) & ~(size_t)3;                                   /* polyphase_fir.c:73 */
        h->M = (int)m;
        std::vector<short> phases((size_t)h->interp * m, 0);
        for (size_t i = 0; i < cfg->nr_taps; i++)                   /* polyphase_fir.c:81-83 */
            phases[(i % h->interp) * m + i / h->interp] = cfg->taps[i];
        PFAIL(cudaMalloc(&h->d_phase, phases.size() * sizeof(short)));
        PFAIL(cudaMemcpy(h->d_phase, phases.data(), phases.size() * sizeof(short), cudaMemcpyHostToDevice));
        h->carry_pitch = (long long)((m + 8 + 7) & ~(size_t)7);
        for (int i = 0; i < 2; i++) PFAIL(cudaMalloc(&h->d_carry[i], (size_t)C * h->carry_pitch * sizeof(short)));
        max_out = (h->max_feed + m) * h->interp / h->decim + 8;
        /* input samples 256 consecutive outputs span, plus the filter length */
        h->rs_span_cap = (int)((unsigned long long)RS_THREADS * h->decim / h->interp + m + 4);
        if ((size_t)h->rs_span_cap * sizeof(short) > 200 * 1024) {
            pager_free(h);
            return perr(GPUPAGER_E_INVAL, "decimate / interpolate = %u / %u is too steep for the resampler tile", h->decim, h->interp);
        }
        if ((size_t)h->rs_span_cap * sizeof(short) > 48 * 1024)
            PFAIL(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->rs_span_cap * (int)sizeof(short)));
    }
    h->res_pitch = (long long)((max_out + 63) & ~(size_t)63);
    PFAIL(cudaMalloc(&h->d_res, (size_t)C * h->res_pitch * sizeof(short)));
    h->nr_src = C;                      /* rows a host feed carries: every row the map refers to */
    if (cfg->channel_map) for (int c = 0; c < C; c++) if ((int)cfg->channel_map[c] + 1 > h->nr_src) h->nr_src = (int)cfg->channel_map[c] + 1;
    PFAIL(cudaMalloc(&h->d_stage, (size_t)h->nr_src * h->max_feed * sizeof(short)));
    if (h->decoder == GPUPAGER_DECODER_FLEX) {
        PFAIL(cudaMalloc(&h->d_fstates, (size_t)C * sizeof(FlexState)));
        PFAIL(cudaMemset(h->d_fstates, 0, (size_t)C * sizeof(FlexState)));
        flex_init_kernel<<<(C + 63) / 64, 64, 0, h->stream>>>(h->d_fstates, C);
        PFAIL(cudaGetLastError());
    } else {
        PFAIL(cudaMalloc(&h->d_states, (size_t)C * sizeof(PocsagState)));
        PFAIL(cudaMemset(h->d_states, 0, (size_t)C * sizeof(PocsagState)));
        PFAIL(cudaMalloc(&h->d_text, (size_t)C * 1024));
        PFAIL(cudaMemset(h->d_text, 0, (size_t)C * 1024));
        h->bits_pitch = (unsigned)(h->res_pitch / 32 + 2);
        PFAIL(cudaMalloc(&h->d_bits, (size_t)C * h->bits_pitch * sizeof(unsigned)));
    }
    h->msg_cap = 32 + (uint32_t)(max_out / 1000);       /* shortest message: 2 codewords x 16 samples/bit */
    if (h->decoder == GPUPAGER_DECODER_FLEX)
        h->msg_cap = 384 + (uint32_t)(max_out / 80);    /* a block end can deliver every address of up to 4 phases at once */
    PFAIL(cudaMalloc(&h->d_msgs, (size_t)C * h->msg_cap * sizeof(gpupager_msg)));
    PFAIL(cudaMalloc(&h->d_count, C * sizeof(uint32_t)));
    PFAIL(cudaMemset(h->d_count, 0, C * sizeof(uint32_t)));
    PFAIL(cudaMalloc(&h->d_dropped, sizeof(unsigned long long)));
    PFAIL(cudaMemset(h->d_dropped, 0, sizeof(unsigned long long)));

    /* GF(2^5) tables, primitive polynomial x^5 + x^2 + 1 */
    uint8_t gexp[32]; int8_t glog[32];
    int v = 1;
    for (int i = 0; i < 31; i++) { gexp[i] = (uint8_t)v; glog[v] = (int8_t)i; v <<= 1; if (v & 0x20) v ^= 0x25; }
    gexp[31] = 0; glog[0] = -1;
    PFAIL(cudaMemcpyToSymbol(c_gf_exp, gexp, sizeof(gexp)));
    PFAIL(cudaMemcpyToSymbol(c_gf_log, glog, sizeof(glog)));
#undef PFAIL
    *ph = h;
    return GPUPAGER_OK;
}

extern "C" int gpupager_destroy(gpupager_t **ph)
{
    if (!ph || !*ph) return perr(GPUPAGER_E_BADARGS, "null handle");
    cudaSetDevice((*ph)->device);
    cudaStreamSynchronize((*ph)->stream);
    pager_free(*ph);
    *ph = nullptr;
    return GPUPAGER_OK;
}

static int pager_run(gpupager *h, const short *d_pcm, size_t pitch, size_t n, cudaStream_t st)
{
    const int C = h->C;
    const short *dec_in = d_pcm;
    long long dec_pitch = (long long)pitch;
    unsigned nr_dec = (unsigned)n;

    if (!(h->flags & GPUPAGER_F_NO_RESAMPLE)) {
        InPcm in;
        in.carry = h->d_carry[h->pp]; in.fresh = d_pcm;
        in.carry_pitch = h->carry_pitch; in.fresh_pitch = (long long)pitch;
        in.carry_len = h->carry_len; in.total = h->carry_len + (long long)n;
        in.invert = (h->flags & GPUPAGER_F_INVERT) ? 1 : 0;
        in.map = h->d_map;
        const unsigned long long total = h->total_in + n;
        /* outputs m with n_m + M < total  <=>  m < (total - M) * I / D   (polyphase_fir.c:184, strict) */
        unsigned long long m_end = h->m_next;
        if (total > (unsigned long long)h->M) {
            const unsigned long long lim = (total - h->M) * h->interp;          /* m * D < lim */
            m_end = (lim + h->decim - 1) / h->decim;
            if (m_end < h->m_next) m_end = h->m_next;
        }
        const unsigned long long nr_out = m_end - h->m_next;
        if (nr_out > (unsigned long long)h->res_pitch) return perr(GPUPAGER_E_INVAL, "feed too large for the output buffer");
        /* keep input from n_next on (the resampler kernel's last block column saves it) */
        unsigned long long n_next = m_end * h->decim / h->interp;
        if (n_next > total) n_next = total;
        if (n_next < h->in_base) n_next = h->in_base;
        const long long keep = (long long)(total - n_next);
        if (keep > h->carry_pitch) return perr(GPUPAGER_E_INVAL, "internal: carry %lld exceeds capacity", keep);
        {
            dim3 grid((unsigned)((nr_out + RS_THREADS - 1) / RS_THREADS) + 1, C);
            resample_kernel<<<grid, RS_THREADS, h->rs_span_cap * sizeof(short), st>>>(
                in, h->in_base, h->d_phase, h->M, h->interp, h->decim, h->m_next, (unsigned)nr_out, h->d_res, h->res_pitch,
                h->rs_span_cap, (long long)(n_next - h->in_base), h->d_carry[h->pp ^ 1], h->carry_pitch, (int)keep);
            h->launches++;
            PCUDA(cudaGetLastError());
        }
        h->m_next = m_end;
        h->total_in = total;
        h->pp ^= 1;
        h->carry_len = keep;
        h->in_base = n_next;
        dec_in = h->d_res; dec_pitch = h->res_pitch; nr_dec = (unsigned)nr_out;
    } else if ((h->flags & GPUPAGER_F_INVERT) || h->d_map) {
        /* no resampler to fold the negation / the channel selection into: a (negated, gathered) copy */
        if (n > (size_t)h->res_pitch) return perr(GPUPAGER_E_INVAL, "feed too large");
        InPcm in;
        in.carry = nullptr; in.fresh = d_pcm; in.carry_pitch = 0; in.fresh_pitch = (long long)pitch;
        in.carry_len = 0; in.total = (long long)n; in.invert = (h->flags & GPUPAGER_F_INVERT) ? 1 : 0;
        in.map = h->d_map;
        dim3 grid((unsigned)((n + 255) / 256), C);
        pcm_carry_kernel<<<grid, 256, 0, st>>>(in, 0, h->d_res, h->res_pitch, (int)n);
        h->launches++;
        PCUDA(cudaGetLastError());
        dec_in = h->d_res; dec_pitch = h->res_pitch;
    } else if (h->flags & (GPUPAGER_F_DC_BLOCK | GPUPAGER_F_KEEP_PCM)) {
        /* the decoder may rewrite samples in place (DC blocker) and the -d tap wants them: work on a copy */
        if (n > (size_t)h->res_pitch) return perr(GPUPAGER_E_INVAL, "feed too large");
        PCUDA(cudaMemcpy2DAsync(h->d_res, h->res_pitch * sizeof(short), d_pcm, pitch * sizeof(short), n * sizeof(short), C,
                                cudaMemcpyDeviceToDevice, st));
        dec_in = h->d_res; dec_pitch = h->res_pitch;
    }
    h->last_out = nr_dec;
    if (nr_dec) {
        MsgSink sink{ h->d_map, h->d_msgs, h->d_count, h->d_dropped, h->msg_cap };
        if (h->decoder == GPUPAGER_DECODER_FLEX)
            flex_kernel<<<(C + FW_WARPS - 1) / FW_WARPS, 32 * FW_WARPS, 0, st>>>(h->d_fstates, C, const_cast<short *>(dec_in), dec_pitch, nr_dec, sink,
                                                                                  (h->flags & GPUPAGER_F_DC_BLOCK) ? 1 : 0, h->dc_p);
        else
            pocsag_kernel<<<(C + PW_WARPS - 1) / PW_WARPS, 32 * PW_WARPS, 0, st>>>(h->d_states, h->d_text, C, const_cast<short *>(dec_in), dec_pitch,
                                                                                    nr_dec, h->d_bits, h->bits_pitch, sink,
                                                                                    (h->flags & GPUPAGER_F_DC_BLOCK) ? 1 : 0, h->dc_p);
        h->launches++;
        PCUDA(cudaGetLastError());
    }
    return GPUPAGER_OK;
}

extern "C" int gpupager_feed_device(gpupager_t *h, const int16_t *d_pcm, size_t pitch, size_t n, void *cuda_stream)
{
    if (!h || (!d_pcm && n)) return perr(GPUPAGER_E_BADARGS, "null argument");
    if (n > h->max_feed) return perr(GPUPAGER_E_INVAL, "feed of %zu samples exceeds max_feed_samples %zu", n, h->max_feed);
    if (n == 0) return GPUPAGER_OK;
    PCUDA(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (st != h->stream) {
        PCUDA(cudaEventRecord(h->ev_own, h->stream));
        PCUDA(cudaStreamWaitEvent(st, h->ev_own, 0));
    }
    if (int rc = pager_run(h, d_pcm, pitch, n, st)) return rc;
    if (st != h->stream) {
        PCUDA(cudaEventRecord(h->ev_ext, st));
        PCUDA(cudaStreamWaitEvent(h->stream, h->ev_ext, 0));
    }
    return GPUPAGER_OK;
}

extern "C" int gpupager_feed(gpupager_t *h, const int16_t *pcm_host, size_t pitch, size_t n)
{
    if (!h || (!pcm_host && n)) return perr(GPUPAGER_E_BADARGS, "null argument");
    if (n > h->max_feed) return perr(GPUPAGER_E_INVAL, "feed of %zu samples exceeds max_feed_samples %zu", n, h->max_feed);
    if (n == 0) return GPUPAGER_OK;
    PCUDA(cudaSetDevice(h->device));
    PCUDA(cudaMemcpy2DAsync(h->d_stage, h->max_feed * sizeof(short), pcm_host, pitch * sizeof(short), n * sizeof(short),
                            h->nr_src, cudaMemcpyHostToDevice, h->stream));
    return pager_run(h, h->d_stage, h->max_feed, n, h->stream);
}

static int pager_drain(gpupager *h)
{
    PCUDA(cudaSetDevice(h->device));
    PCUDA(cudaStreamSynchronize(h->stream));
    std::vector<uint32_t> cnt(h->C);
    PCUDA(cudaMemcpy(cnt.data(), h->d_count, h->C * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    bool any = false;
    for (int c = 0; c < h->C; c++) {
        if (!cnt[c]) continue;
        any = true;
        const size_t at = h->queue.size();
        h->queue.resize(at + cnt[c]);
        PCUDA(cudaMemcpy(&h->queue[at], h->d_msgs + (size_t)c * h->msg_cap, cnt[c] * sizeof(gpupager_msg),
                         cudaMemcpyDeviceToHost));
    }
    if (any) PCUDA(cudaMemset(h->d_count, 0, h->C * sizeof(uint32_t)));
    unsigned long long d = 0;
    PCUDA(cudaMemcpy(&d, h->d_dropped, sizeof(d), cudaMemcpyDeviceToHost));
    h->dropped = d;
    return GPUPAGER_OK;
}

extern "C" int gpupager_poll(gpupager_t *h, gpupager_msg *out, size_t cap, size_t *nr_msgs)
{
    if (!h || !nr_msgs || (!out && cap)) return perr(GPUPAGER_E_BADARGS, "null argument");
    if (int rc = pager_drain(h)) return rc;
    const size_t n = h->queue.size() < cap ? h->queue.size() : cap;
    if (n) memcpy(out, h->queue.data(), n * sizeof(gpupager_msg));
    h->queue.erase(h->queue.begin(), h->queue.begin() + n);
    *nr_msgs = n;
    return GPUPAGER_OK;
}

extern "C" int gpupager_dispatch(gpupager_t *h, gpupager_on_msg_func_t on_numeric, gpupager_on_msg_func_t on_alpha, void *user,
                                 size_t *nr_msgs)
{
    if (!h) return perr(GPUPAGER_E_BADARGS, "null handle");
    if (int rc = pager_drain(h)) return rc;
    for (const gpupager_msg &m : h->queue) {
        if (m.kind != GPUPAGER_MSG_ALPHA && m.kind != GPUPAGER_MSG_NUMERIC) continue;
        gpupager_on_msg_func_t cb = (m.kind == GPUPAGER_MSG_ALPHA) ? on_alpha : on_numeric;
        if (cb) cb(user, m.channel, (uint16_t)m.baud, m.capcode, m.text, m.len, (uint8_t)m.function);
    }
    if (nr_msgs) *nr_msgs = h->queue.size();
    h->queue.clear();
    return GPUPAGER_OK;
}

extern "C" int gpupager_dispatch_flex(gpupager_t *h, gpupager_on_flex_alnum_func_t on_alnum, gpupager_on_flex_num_func_t on_num,
                                      gpupager_on_flex_siv_func_t on_siv, void *user, size_t *nr_msgs)
{
    if (!h) return perr(GPUPAGER_E_BADARGS, "null handle");
    if (int rc = pager_drain(h)) return rc;
    for (const gpupager_msg &m : h->queue) {
        const uint64_t cap = (uint64_t)m.capcode | ((uint64_t)m.capcode_hi << 32);
        if (m.kind == GPUPAGER_MSG_FLEX_ALNUM && on_alnum)
            on_alnum(user, m.channel, (uint16_t)m.baud, (uint8_t)m.function, (uint8_t)m.aux[0], (uint8_t)m.aux[1], cap,
                     (int)m.aux[2], (int)m.aux[3], (uint8_t)m.aux[4], m.text, m.len);
        else if (m.kind == GPUPAGER_MSG_FLEX_NUM && on_num)
            on_num(user, m.channel, (uint16_t)m.baud, (uint8_t)m.function, (uint8_t)m.aux[0], (uint8_t)m.aux[1], cap, m.text, m.len);
        else if (m.kind == GPUPAGER_MSG_FLEX_SIV && on_siv)
            on_siv(user, m.channel, (uint16_t)m.baud, (uint8_t)m.function, (uint8_t)m.aux[0], (uint8_t)m.aux[1], cap,
                   (uint8_t)m.aux[2], m.aux[3]);
    }
    if (nr_msgs) *nr_msgs = h->queue.size();
    h->queue.clear();
    return GPUPAGER_OK;
}

extern "C" int gpupager_collect_pcm(gpupager_t *h, int16_t *out, size_t cap, size_t *n)
{
    if (!h || !out || !n) return perr(GPUPAGER_E_BADARGS, "null argument");
    if (!(h->flags & GPUPAGER_F_KEEP_PCM)) return perr(GPUPAGER_E_INVAL, "bank was created without GPUPAGER_F_KEEP_PCM");
    *n = h->last_out;
    if (h->last_out > cap) return perr(GPUPAGER_E_INVAL, "capacity %zu < %zu samples", cap, h->last_out);
    PCUDA(cudaSetDevice(h->device));
    if (h->last_out)
        PCUDA(cudaMemcpy2DAsync(out, cap * sizeof(short), h->d_res, h->res_pitch * sizeof(short), h->last_out * sizeof(short),
                                h->C, cudaMemcpyDeviceToHost, h->stream));
    PCUDA(cudaStreamSynchronize(h->stream));
    return GPUPAGER_OK;
}

extern "C" uint64_t gpupager_kernel_launches(gpupager_t *h) { return h ? h->launches : 0; }
extern "C" uint64_t gpupager_dropped_msgs(gpupager_t *h) { return h ? h->dropped : 0; }

/* ============================================================================================== */
/* Mueller-Muller timing recovery (pager/mueller_muller.c:10-115), one thread per channel.           */
/* ============================================================================================== */
namespace {

struct MmState { float w, m, next_offset, last_sample; };

__device__ __forceinline__ float mm_sign(float v) { return (float)(v > 0.0f) - (float)(v < 0.0f); }   /* :34-38 */

template <bool FMA>
__global__ void mm_kernel(MmState *__restrict__ states, int nr_channels, const short *__restrict__ pcm, long long pitch,
                          unsigned n, float kw, float km, float emin, float emax, short *__restrict__ out, long long cap,
                          unsigned *__restrict__ nr_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nr_channels) return;
    MmState st = states[c];
    const short *x = pcm + (size_t)c * pitch;
    short *d = out + (size_t)c * cap;
    float cur = st.next_offset, w = st.w, m = st.m, last = st.last_sample;
    const float nf = (float)n;
    unsigned nd = 0;
    while (cur < nf && nd < cap) {                                          /* :65 */
        const float sample = (float)x[(size_t)__fadd_rn(cur, 0.5f)];        /* :66 */
        d[nd++] = (short)sample;                                            /* :70 */
        /* :76 -- both products are exact (a sign times an integer-valued float) */
        const float w_error = __fsub_rn(__fmul_rn(mm_sign(last), sample), __fmul_rn(mm_sign(sample), last));
        w = FMA ? __fmaf_rn(w_error, kw, w) : __fadd_rn(w, __fmul_rn(w_error, kw));                 /* :79 */
        if (emin > w) w = emin; else if (emax < w) w = emax;                                        /* :86-90 */
        m = FMA ? __fadd_rn(m, __fmaf_rn(km, sample, w)) : __fadd_rn(m, __fadd_rn(w, __fmul_rn(km, sample)));   /* :92 */
        const float fl = floorf(m);
        cur = __fadd_rn(cur, fl);                                           /* :95 */
        m = __fsub_rn(m, fl);
        last = sample;
    }
    st.next_offset = __fsub_rn(cur, nf);                                    /* :107-109 */
    st.w = w; st.m = m; st.last_sample = last;
    states[c] = st;
    nr_out[c] = nd;
}

} // namespace

struct gpumm {
    int device = 0, C = 0;
    uint32_t flags = 0;
    size_t max_feed = 0;
    float kw = 0, km = 0, emin = 0, emax = 0;
    cudaStream_t stream = nullptr;
    MmState *d_states = nullptr;
    short *d_in = nullptr, *d_out = nullptr;
    unsigned *d_nr = nullptr;
};

extern "C" int gpumm_create(gpumm_t **ph, uint32_t nr_channels, int32_t device, float kw, float km, float samples_per_bit,
                            float error_min, float error_max, uint32_t max_feed_samples, uint32_t flags)
{
    if (!ph || !nr_channels || !max_feed_samples) return perr(GPUPAGER_E_BADARGS, "bad argument");
    *ph = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return perr(GPUPAGER_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return perr(GPUPAGER_E_BADARGS, "bad device ordinal");
    PCUDA(cudaSetDevice(device));
    gpumm *h = new (std::nothrow) gpumm();
    if (!h) return perr(GPUPAGER_E_NOMEM, "out of memory");
    h->device = device; h->C = (int)nr_channels; h->flags = flags; h->max_feed = max_feed_samples;
    h->kw = kw; h->km = km; h->emin = error_min; h->emax = error_max;
    const size_t C = nr_channels;
    std::vector<MmState> init(C);
    for (auto &s : init) { s.w = s.m = samples_per_bit; s.next_offset = 0.0f; s.last_sample = 0.0f; }   /* mueller_muller.c:19-20 */
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_states, C * sizeof(MmState));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_states, init.data(), C * sizeof(MmState), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_in, C * h->max_feed * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_out, C * h->max_feed * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_nr, C * sizeof(unsigned));
    if (e != cudaSuccess) {
        perr(GPUPAGER_E_CUDA, "gpumm_create: %s", cudaGetErrorString(e));
        gpumm_destroy(&h);
        return GPUPAGER_E_CUDA;
    }
    *ph = h;
    return GPUPAGER_OK;
}

extern "C" int gpumm_destroy(gpumm_t **ph)
{
    if (!ph || !*ph) return perr(GPUPAGER_E_BADARGS, "null handle");
    gpumm *h = *ph;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->d_states); cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_nr);
    delete h;
    *ph = nullptr;
    return GPUPAGER_OK;
}

extern "C" int gpumm_process(gpumm_t *h, const int16_t *pcm_host, size_t pitch, size_t n, int16_t *decisions_host, size_t cap,
                             uint32_t *nr_out)
{
    if (!h || !pcm_host || !decisions_host || !nr_out) return perr(GPUPAGER_E_BADARGS, "null argument");
    if (n > h->max_feed) return perr(GPUPAGER_E_INVAL, "feed of %zu samples exceeds max_feed_samples %zu", n, h->max_feed);
    if (n >= (1u << 24)) return perr(GPUPAGER_E_INVAL, "mm_process counts samples in float: feeds must stay below 2^24 samples");
    for (int c = 0; c < h->C; c++) nr_out[c] = 0;
    if (n == 0) return GPUPAGER_OK;
    PCUDA(cudaSetDevice(h->device));
    PCUDA(cudaMemcpy2DAsync(h->d_in, h->max_feed * sizeof(short), pcm_host, pitch * sizeof(short), n * sizeof(short), h->C,
                            cudaMemcpyHostToDevice, h->stream));
    const long long dcap = (long long)(cap < h->max_feed ? cap : h->max_feed);
    if (h->flags & GPUMM_F_FMA)
        mm_kernel<true><<<(h->C + 31) / 32, 32, 0, h->stream>>>(h->d_states, h->C, h->d_in, (long long)h->max_feed, (unsigned)n, h->kw,
                                                               h->km, h->emin, h->emax, h->d_out, dcap, h->d_nr);
    else
        mm_kernel<false><<<(h->C + 31) / 32, 32, 0, h->stream>>>(h->d_states, h->C, h->d_in, (long long)h->max_feed, (unsigned)n, h->kw,
                                                                h->km, h->emin, h->emax, h->d_out, dcap, h->d_nr);
    PCUDA(cudaGetLastError());
    PCUDA(cudaMemcpyAsync(nr_out, h->d_nr, h->C * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    PCUDA(cudaStreamSynchronize(h->stream));
    unsigned mx = 0;
    for (int c = 0; c < h->C; c++) mx = nr_out[c] > mx ? nr_out[c] : mx;
    if (mx) PCUDA(cudaMemcpy2D(decisions_host, cap * sizeof(short), h->d_out, dcap * sizeof(short), mx * sizeof(short), h->C,
                               cudaMemcpyDeviceToHost));
    return GPUPAGER_OK;
}

extern "C" int gpumm_get_state(gpumm_t *h, uint32_t channel, float state[4])
{
    if (!h || !state || channel >= (uint32_t)h->C) return perr(GPUPAGER_E_BADARGS, "bad argument");
    PCUDA(cudaSetDevice(h->device));
    PCUDA(cudaMemcpy(state, h->d_states + channel, sizeof(MmState), cudaMemcpyDeviceToHost));
    return GPUPAGER_OK;
}
