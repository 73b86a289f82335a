/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Stream-form CPU restatement of the reference algorithms.  Each function
 * names the reference lines it follows (paths relative to /root/reference).
 * Compiled with -ffp-contract=off: the single FMA the reference's own Release
 * build fuses (multifm/fast_atan2f.c:131 under gcc -O3 -march=<fma capable>)
 * is selected explicitly through the `fma` argument.
 */
#include "oracle.h"

#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define Q14_SHIFT 14    /* filter/filter.h:16 -- "Q_15_SHIFT" is 14 */

/* filter/complex.h:31-34: (a >> 14) + ((a >> 13) & 1), truncated to int16 by the return type */
static inline int16_t rq(int32_t a)
{
    return (int16_t)((a >> Q14_SHIFT) + ((a >> (Q14_SHIFT - 1)) & 1));
}

/* all accumulations wrap modulo 2^32 like x86 int32 arithmetic does */
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

/* ---------------------------------------------------------------------- */
/* a1  multifm/demod.c:205-261, multifm/receiver.c:220, filter/direct_fir.c:72-83 */
/* ---------------------------------------------------------------------- */
void orc_prepare_taps(const double *lpf, size_t nr_taps, int32_t offset_hz, uint32_t sample_rate,
                      double gain, int16_t *c_re, int16_t *c_im)
{
    const double w = -2.0 * M_PI * (double)offset_hz / (double)sample_rate;     /* demod.c:210 */
    for (size_t i = 0; i < nr_taps; i++) {
        const double complex t = gain * cexp(CMPLX(0, w * (double)i)) * lpf[i];  /* demod.c:234 */
        c_re[i] = (int16_t)(creal(t) * 16384.0);                                 /* demod.c:242 */
        c_im[i] = (int16_t)(cimag(t) * 16384.0);                                 /* demod.c:243 */
    }
}

void orc_derot_incr(int32_t offset_hz, uint32_t sample_rate, unsigned decimation, int16_t incr[2])
{
    const double fwt0 = 2.0 * M_PI * (double)offset_hz / (double)sample_rate;    /* direct_fir.c:73 */
    const double complex d = cexp(CMPLX(0, -fwt0 * (double)decimation));         /* direct_fir.c:75 */
    incr[0] = (int16_t)(int32_t)(creal(d) * 16384.0);                            /* direct_fir.c:76 */
    incr[1] = (int16_t)(int32_t)(cimag(d) * 16384.0);                            /* direct_fir.c:77 */
}

double orc_db_to_gain(double db)
{
    return pow(10.0, db / 10.0);                                                 /* receiver.c:220 */
}

/* ---------------------------------------------------------------------- */
/* a4  multifm/fast_atan2f.c:15-81 (table), :101-174; multifm/fm_demod.c:36-85 */
/* ---------------------------------------------------------------------- */
static float g_atan_tab[257];
static int g_atan_tab_ready;

/* The reference table holds atan(i/255), i = 0..255, written with 7 significant
 * decimal digits ("%.6e") and one duplicated guard entry; regenerating it that
 * way is bit-identical (checked against the reference in tests). */
void orc_atan_table(float tab[257])
{
    for (int i = 0; i < 257; i++) {
        char txt[32];
        int j = i > 255 ? 255 : i;
        snprintf(txt, sizeof(txt), "%.6e", atan((double)j / 255.0));
        tab[i] = strtof(txt, NULL);
    }
}

float orc_fast_atan2f(float y, float x, int fma)
{
    if (!g_atan_tab_ready) { orc_atan_table(g_atan_tab); g_atan_tab_ready = 1; }

    const float ya = fabsf(y), xa = fabsf(x);
    float z, base;

    if (!(ya > 0.0f || xa > 0.0f)) return 0.0f;                 /* :113 */
    z = (ya < xa) ? ya / xa : xa / ya;                          /* :116-119 */

    if ((double)z < 0.003921569) {                              /* :123, double compare */
        base = z;
    } else {
        float alpha = z * 255.0f;                               /* :127 */
        int idx = ((int)alpha) & 0xff;                          /* :128 */
        alpha -= (float)idx;                                    /* :129 */
        float lo = g_atan_tab[idx];
        float d = g_atan_tab[idx + 1] - g_atan_tab[idx];
        base = fma ? fmaf(d, alpha, lo) : (lo + d * alpha);     /* :132-133 */
    }

    const float pi_f = (float)3.14159265358979323846;
    const float hpi_f = (float)1.57079632679489661923;
    float angle;
    if (xa > ya) {                                              /* :136 */
        if (x >= 0.0f) angle = (y >= 0.0f) ? base : -base;
        else           angle = (y >= 0.0f) ? (pi_f - base) : (base - pi_f);
    } else {
        if (y >= 0.0f) angle = (x >= 0.0f) ? (hpi_f - base) : (hpi_f + base);
        else           angle = (x >= 0.0f) ? (-hpi_f + base) : (-hpi_f - base);
    }
    return angle;
}

int16_t orc_fm_pcm(int32_t s_re, int32_t s_im, int fma)
{
    float phi = orc_fast_atan2f((float)s_im, (float)s_re, fma); /* fm_demod.c:68 */
    float scaled = (float)(((double)phi / M_PI) * (double)16384.0f); /* fm_demod.c:71 */
    return (int16_t)scaled;                                     /* fm_demod.c:72 */
}

/* ---------------------------------------------------------------------- */
/* a2-a4  filter/direct_fir.c:329-417 + :152-172, multifm/fm_demod.c:53-77 */
/* ---------------------------------------------------------------------- */
void orc_chan_state_init(orc_chan_state *st, int32_t offset_hz, uint32_t sample_rate, unsigned decimation)
{
    int16_t incr[2];
    memset(st, 0, sizeof(*st));
    orc_derot_incr(offset_hz, sample_rate, decimation, incr);
    st->incr_re = incr[0]; st->incr_im = incr[1];
    st->rot_re = 16384; st->rot_im = 0;                         /* direct_fir.c:78-79 */
}

size_t orc_chan_stream(orc_chan_state *st, size_t nr_taps, const int16_t *c_re, const int16_t *c_im,
                       unsigned decimation, const int16_t *iq, size_t n, int fma,
                       int16_t *out_iq, int16_t *out_pcm)
{
    if (n < nr_taps) return 0;
    const size_t K = (n - nr_taps) / decimation + 1;
    const int derotate = !(st->incr_re == 0 && st->incr_im == 0);   /* direct_fir.c:406 */

    for (size_t k = 0; k < K; k++) {
        const int16_t *x = iq + 2 * k * (size_t)decimation;
        int32_t acc_re = 0, acc_im = 0;
        for (size_t i = 0; i < nr_taps; i++) {                  /* direct_fir.c:366-385 */
            int32_t s_re = x[2 * i], s_im = x[2 * i + 1], t_re = c_re[i], t_im = c_im[i];
            acc_re = wadd(acc_re, wsub(wmul(t_re, s_re), wmul(t_im, s_im)));
            acc_im = wadd(acc_im, wadd(wmul(t_re, s_im), wmul(t_im, s_re)));
        }
        if (derotate) {                                         /* direct_fir.c:406-409, :162-167 */
            int32_t q_re = rq(acc_re), q_im = rq(acc_im);
            int32_t r_re = st->rot_re, r_im = st->rot_im, i_re = st->incr_re, i_im = st->incr_im;
            acc_re = wsub(wmul(q_re, r_re), wmul(q_im, r_im));
            acc_im = wadd(wmul(q_re, r_im), wmul(q_im, r_re));
            st->rot_re = rq(wsub(wmul(r_re, i_re), wmul(r_im, i_im)));
            st->rot_im = rq(wadd(wmul(r_re, i_im), wmul(r_im, i_re)));
            st->rot_counter++;
        }
        const int32_t y_re = rq(acc_re), y_im = rq(acc_im);     /* direct_fir.c:412-413 */
        if (out_iq) { out_iq[2 * k] = (int16_t)y_re; out_iq[2 * k + 1] = (int16_t)y_im; }

        /* fm_demod.c:55-64: s = y * conj(last) */
        const int32_t b_re = st->last_re, b_im = -st->last_im;
        const int32_t s_re = wsub(wmul(y_re, b_re), wmul(y_im, b_im));
        const int32_t s_im = wadd(wmul(y_re, b_im), wmul(y_im, b_re));
        if (out_pcm) out_pcm[k] = orc_fm_pcm(s_re, s_im, fma);
        st->last_re = y_re; st->last_im = y_im;
    }
    return K;
}

/* ---------------------------------------------------------------------- */
/* a5  filter/polyphase_fir.c:47-105,162-233; filter/utils.c:46-116         */
/* ---------------------------------------------------------------------- */
int orc_resamp_init(orc_resamp *r, const int16_t *taps, size_t nr_taps, unsigned interp, unsigned decim)
{
    memset(r, 0, sizeof(*r));
    if (!nr_taps || !interp || !decim) return -1;
    size_t m = (nr_taps + interp - 1) / interp;                 /* polyphase_fir.c:70 */
    m = (m + 3) & ~(size_t)3;                                   /* polyphase_fir.c:73 */
    r->phase_filters = calloc((size_t)interp * m, sizeof(int16_t));
    if (!r->phase_filters) return -1;
    for (size_t i = 0; i < nr_taps; i++)                        /* polyphase_fir.c:81-83 */
        r->phase_filters[(i % interp) * m + i / interp] = taps[i];
    r->interp = interp; r->decim = decim; r->phase_len = m; r->phase = 0;
    return 0;
}

void orc_resamp_free(orc_resamp *r)
{
    free(r->phase_filters);
    r->phase_filters = NULL;
}

size_t orc_resamp_stream(orc_resamp *r, const int16_t *x, size_t n, int16_t *out, size_t cap, size_t *consumed)
{
    size_t off = 0, produced = 0;
    const size_t m = r->phase_len;
    while (produced < cap && off < n && (n - off) > m) {        /* polyphase_fir.c:184, strict '>' */
        const int16_t *h = r->phase_filters + r->phase * m;
        int32_t acc = 0;
        for (size_t j = 0; j < m; j++)                          /* utils.c:88-93 */
            acc = wadd(acc, wmul((int32_t)x[off + j], (int32_t)h[j]));
        out[produced++] = rq(acc);                              /* utils.c:109 */
        size_t p = r->phase + r->decim;                         /* polyphase_fir.c:206-211 */
        off += p / r->interp;
        r->phase = p % r->interp;
    }
    *consumed = off;
    return produced;
}

/* ---------------------------------------------------------------------- */
/* f1  filter/dc_blocker.h:46-92                                            */
/* ---------------------------------------------------------------------- */
void orc_dc_blocker_init(orc_dc_blocker *b, double pole)
{
    memset(b, 0, sizeof(*b));
    b->p = (int16_t)((1.0 - pole) * (double)(1 << 14));         /* dc_blocker.h:57 */
}

void orc_dc_blocker_apply(orc_dc_blocker *b, int16_t *x, size_t n)
{
    for (size_t i = 0; i < n; i++) {                            /* dc_blocker.h:79-88 */
        b->acc = wsub(b->acc, b->x_prev);
        b->x_prev = (int32_t)((uint32_t)(int32_t)x[i] << 14);
        b->acc = wadd(b->acc, wsub(b->x_prev, wmul(b->p, b->y_prev)));
        b->y_prev = b->acc >> 14;
        x[i] = (int16_t)b->y_prev;
    }
}

/* ---------------------------------------------------------------------- */
/* a7  pager/bch_code.c:42-74 (field), :307-398 (decode); poly per          */
/*     pager/pager_pocsag.c:150 : x^5 + x^2 + 1, n=31, k=21, t=2            */
/* ---------------------------------------------------------------------- */
static int g_exp[32], g_log[32], g_gf_ready;

static void gf_init(void)
{
    /* alpha^i in polynomial form; log(0) = -1 */
    int v = 1;
    for (int i = 0; i < 31; i++) {
        g_exp[i] = v; g_log[v] = i;
        v <<= 1;
        if (v & 0x20) v ^= 0x25;        /* x^5 = x^2 + 1 */
    }
    g_exp[31] = 0;                      /* never indexed with 31 after a % 31 */
    g_log[0] = -1;
    g_gf_ready = 1;
}

int orc_bch_decode(uint32_t *word)
{
    if (!g_gf_ready) gf_init();
    uint32_t r = *word;
    int s[5], any = 0, fail = 0;

    for (int i = 1; i <= 4; i++) {                              /* bch_code.c:325-340 */
        int v = 0;
        for (int j = 0; j < 31; j++)
            if ((r >> (30 - j)) & 1) v ^= g_exp[(i * j) % 31];
        if (v) any = 1;
        s[i] = g_log[v];
    }

    if (any) {
        if (s[1] != -1) {
            const int s3 = (s[1] * 3) % 31;
            if (s[3] == s3) {                                   /* single error, :345-347 */
                r ^= 1u << (30 - s[1]);
            } else {                                            /* two errors, :353-388 */
                int aux = (s[3] != -1) ? (g_exp[s3] ^ g_exp[s[3]]) : g_exp[s3];
                int reg1 = (s[2] - g_log[aux] + 31) % 31;
                int reg2 = (s[1] - g_log[aux] + 31) % 31;
                int loc[3], count = 0;
                for (int i = 1; i <= 31; i++) {                 /* Chien search */
                    int q = 1;
                    reg1 = (reg1 + 1) % 31; q ^= g_exp[reg1];
                    reg2 = (reg2 + 2) % 31; q ^= g_exp[reg2];
                    if (!q && count < 3) loc[count++] = i % 31;
                }
                if (count == 2) {
                    r ^= 1u << (30 - loc[0]);
                    r ^= 1u << (30 - loc[1]);
                } else {
                    fail = 1;
                }
            }
        } else if (s[2] != -1) {                                /* :389-391 */
            fail = 1;
        }
    }
    *word = r;
    return fail;
}

/* ---------------------------------------------------------------------- */
/* a6/a7  pager/pager_pocsag.c:82-117 (eye), :242-297 (deliver),            */
/*        :320-432 (batch), :434-543 (FSM)                                  */
/* ---------------------------------------------------------------------- */
#define POCSAG_SYNC 0x7cd215d8u     /* pager_pocsag_priv.h:40 */
#define POCSAG_IDLE 0x6983915eu     /* pager_pocsag_priv.h:46 (bit-reversed standard idle) */

enum { ST_SEARCH = 0, ST_SYNCHRONIZED = 1, ST_BATCH = 2, ST_SYNCWORD = 3 };
enum { MT_NONE = 0, MT_UNKNOWN = 1, MT_ALPHA = 2, MT_NUMERIC = 3 };

struct eye {
    uint32_t spb, baud, cur, matches;
    uint32_t reg[75];
};

struct orc_pocsag {
    int state;
    uint16_t sample_skip, baud;
    /* batch */
    uint16_t b_skip, b_word, b_word_bit, b_bits;
    uint32_t batch[16];
    /* sync search */
    uint16_t s_skip;
    size_t s_bits;
    uint32_t s_word;
    struct eye eyes[3];
    /* message under assembly */
    char alpha[512], numeric[512];
    size_t n_alpha, n_numeric;
    int score;
    int seen_nonprint;
    uint32_t capcode, w_alpha, w_numeric;
    size_t vb_alpha, vb_numeric;
    uint8_t function;
    int msg_type;
    /* sink */
    orc_msg *msgs;
    size_t cap, n;
};

static int sync_ok(uint32_t w) { return __builtin_popcount(w ^ POCSAG_SYNC) <= 4; }

static void eyes_reset(orc_pocsag *p)
{
    static const uint32_t spb[3] = { 75, 32, 16 }, baud[3] = { 512, 1200, 2400 };
    for (int i = 0; i < 3; i++) {
        memset(&p->eyes[i], 0, sizeof(p->eyes[i]));
        p->eyes[i].spb = spb[i]; p->eyes[i].baud = baud[i];
    }
}

static void msg_reset(orc_pocsag *p)
{
    p->n_alpha = p->n_numeric = 0;
    p->w_alpha = p->w_numeric = 0;
    p->vb_alpha = p->vb_numeric = 0;
    p->seen_nonprint = 0; p->score = 0;
    p->msg_type = MT_NONE; p->function = 0;
}

static void batch_reset(orc_pocsag *p)
{
    memset(p->batch, 0, sizeof(p->batch));
    p->b_word = p->b_word_bit = p->b_skip = p->b_bits = 0;
}

orc_pocsag *orc_pocsag_new(size_t max_msgs)
{
    orc_pocsag *p = calloc(1, sizeof(*p));
    p->msgs = calloc(max_msgs ? max_msgs : 1, sizeof(orc_msg));
    p->cap = max_msgs;
    eyes_reset(p);
    msg_reset(p);
    return p;
}

void orc_pocsag_delete(orc_pocsag *p)
{
    if (!p) return;
    free(p->msgs);
    free(p);
}

size_t orc_pocsag_msgs(orc_pocsag *p, orc_msg **msgs) { *msgs = p->msgs; return p->n; }
size_t orc_msg_size(void) { return sizeof(orc_msg); }

static void deliver(orc_pocsag *p)                              /* pager_pocsag.c:242-297 */
{
    if (p->msg_type == MT_NONE) return;
    if (p->n_alpha != 0) {
        char last = p->alpha[p->n_alpha - 1];
        if (last == 0x4 || last == 0x3 || last == 0x0 || last == 0x17) p->score = 1;
    }
    if (p->n_numeric > 40) p->score = 1;
    const int is_alpha = p->score > 0;
    if (p->n < p->cap) {
        orc_msg *m = &p->msgs[p->n++];
        memset(m, 0, sizeof(*m));
        m->kind = (uint32_t)is_alpha; m->baud = p->baud; m->capcode_lo = p->capcode; m->function = p->function;
        if (is_alpha) { m->len = (uint32_t)p->n_alpha; memcpy(m->data, p->alpha, p->n_alpha); }
        else          { m->len = (uint32_t)p->n_numeric; memcpy(m->data, p->numeric, p->n_numeric); }
    }
    msg_reset(p);
}

static void process_batch(orc_pocsag *p)                        /* pager_pocsag.c:320-432 */
{
    static const char bcd_map[16] = { '0','1','2','3','4','5','6','7','8','9','X','U',' ','-','[',']' };
    for (unsigned z = 0; z < 16; z++) {
        uint32_t w = p->batch[z] & 0x7fffffffu;
        if (orc_bch_decode(&w)) {
            if (p->msg_type != MT_NONE) deliver(p);
            return;
        }
        if (w == POCSAG_IDLE) {
            if (p->msg_type != MT_NONE) deliver(p);
            continue;
        }
        if ((w & 1) == 0) {                                     /* address word, :356-363 */
            deliver(p);
            p->msg_type = MT_UNKNOWN;
            p->function = (w >> 19) & 0x3;
            p->capcode = (((w >> 1) & 0x3ffffu) << 3) + ((z >> 1) & 0x7);
        } else if (p->msg_type == MT_UNKNOWN) {                 /* data word, :364-417 */
            const uint32_t val = (w >> 1) & 0xfffffu;
            p->w_alpha |= val << p->vb_alpha;
            p->vb_alpha += 20;
            while (p->vb_alpha >= 7) {
                char c = (char)(p->w_alpha & 0x7f);
                /* the reference writes message_alpha[] unchecked (512 bytes); we stop storing at 511 */
                if (p->n_alpha < 511) p->alpha[p->n_alpha++] = c;
                if ((c >= 0x20 && c <= 0x7e) || c == 0xa || c == 0xd) {     /* isprint() in the C locale */
                    if (!p->seen_nonprint) p->score++;
                } else {
                    p->seen_nonprint = 1;
                    if (c != 0x03 && c != 0x04 && c != 0x17 && c != 0x0) p->score -= 10;
                }
                p->w_alpha >>= 7;
                p->vb_alpha -= 7;
            }
            if (p->n_numeric < 511) {
                p->w_numeric |= val << p->vb_numeric;
                p->vb_numeric += 20;
                while (p->vb_numeric >= 4 && p->n_numeric < 511) {
                    p->numeric[p->n_numeric++] = bcd_map[p->w_numeric & 0xf];
                    p->w_numeric >>= 4;
                    p->vb_numeric -= 4;
                }
            }
        }
    }
}

static void eye_on_sample(orc_pocsag *p, struct eye *e, int16_t sample)  /* pager_pocsag.c:82-117 */
{
    uint32_t *r = &e->reg[e->cur];
    *r = (*r << 1) | (sample < 0 ? 1u : 0u);
    if (sync_ok(*r)) {
        e->matches++;
    } else if (e->matches > e->spb / 2) {
        p->sample_skip = (uint16_t)e->spb;
        p->baud = (uint16_t)e->baud;
        batch_reset(p);
        p->b_skip = (uint16_t)(e->matches / 2);
        p->state = ST_SYNCHRONIZED;
    } else {
        e->matches = 0;
    }
    e->cur = (e->cur + 1) % e->spb;
}

void orc_pocsag_on_pcm(orc_pocsag *p, const int16_t *pcm, size_t n)     /* pager_pocsag.c:434-543 */
{
    size_t i = 0;
    while (i < n) {
        switch (p->state) {
        case ST_SEARCH:
            while (i < n) {
                eye_on_sample(p, &p->eyes[0], pcm[i]);
                eye_on_sample(p, &p->eyes[1], pcm[i]);
                eye_on_sample(p, &p->eyes[2], pcm[i]);
                i++;
                if (p->state == ST_SYNCHRONIZED) break;
            }
            break;
        case ST_SYNCHRONIZED:
            p->state = ST_BATCH;
            /* fall through */
        case ST_BATCH:
            while (i < n) {
                if (++p->b_skip == p->sample_skip) {
                    uint32_t bit = pcm[i] < 0 ? 1u : 0u;
                    /* `bit << bit_count` with bit_count up to 511: x86 masks the count to 5 bits */
                    p->batch[p->b_word] |= bit << (p->b_bits & 31);
                    p->b_word_bit++; p->b_bits++; p->b_skip = 0;
                    if (p->b_word_bit == 32) {
                        p->b_word_bit = 0;
                        if (++p->b_word == 16) {
                            process_batch(p);
                            p->state = ST_SYNCWORD;
                            p->b_word = 0;
                            p->s_skip = 0; p->s_bits = 0; p->s_word = 0;
                            i++;
                            break;
                        }
                    }
                }
                i++;
            }
            break;
        case ST_SYNCWORD:
            while (i < n) {
                if (++p->s_skip == p->sample_skip) {
                    p->s_skip = 0;
                    p->s_word = (p->s_word << 1) | (pcm[i] < 0 ? 1u : 0u);
                    if (++p->s_bits == 32) {
                        if (!sync_ok(p->s_word)) {
                            p->state = ST_SEARCH;
                            p->sample_skip = 0;
                            eyes_reset(p);
                            deliver(p);
                        } else {
                            p->state = ST_BATCH;
                            batch_reset(p);
                        }
                        i++;
                        break;
                    }
                }
                i++;
            }
            break;
        }
    }
}

/* ---------------------------------------------------------------------- */
/* a8  FLEX: pager/pager_flex.c (whole file); state pager/pager_flex_priv.h */
/* ---------------------------------------------------------------------- */
enum { FX_SYNC_1 = 0, FX_SYNC_2 = 1, FX_BLOCK = 2 };                                    /* pager_flex_priv.h:11-31 */
enum { FS_SEARCH_BS1 = 0, FS_BS1, FS_A, FS_B, FS_INV_A, FS_FIW, FS_SYNCED };            /* :33-70 */
enum { F2_COMMA = 0, F2_C, F2_INV_COMMA, F2_INV_C, F2_SYNCED };                         /* :113-138 */

struct flex_coding { uint16_t seq_a, baud; uint8_t fsk_levels, sample_skip, sync_2_samples, sym_bits, sample_fudge;
                     uint16_t symbols_per_block; uint8_t nr_phases; };
static const struct flex_coding flex_codings[4] = {                                     /* pager_flex.c:47-96 */
    { 0x78f3, 1600, 2, 9,  4, 1, 0, 2816, 1 },
    { 0x84e7, 3200, 2, 4, 24, 1, 2, 5632, 2 },
    { 0x4f97, 3200, 4, 9, 12, 2, 0, 2816, 2 },
    { 0x215f, 6400, 4, 4, 32, 2, 2, 5632, 4 },
};

/* struct pager_flex_block laid out as the reference has it in memory (pager_flex_priv.h:175-233), as 32-bit
 * words: 4 phases x (88 words + one word holding cur_bit | cur_word << 8 | base_word << 16), then nr_symbols,
 * then phase_ff.  The reference indexes phase_words[] with unchecked offsets taken from the air (vector start
 * words up to 127 + length 127, pager_flex.c:977,1003), so reads and in-place BCH fix-ups can land in a later
 * phase or in these bookkeeping words; modelling the block as one flat array reproduces that.  Offsets past the
 * block (the reference would read its own heap) abort the vector instead. */
#define FX_PHASE_STRIDE 89
#define FX_BLOCK_WORDS (4 * FX_PHASE_STRIDE + 2)

struct orc_flex {
    int16_t sample_range, sample_delta;
    int state;
    int16_t skip, skip_count;
    uint8_t cycle_id, frame_id;
    /* sync 1 */
    uint32_t sync_words[10];
    int sync_state;
    uint8_t sample_counter, bit_counter;
    uint32_t a; uint16_t b; uint32_t inv_a; uint32_t fiw;
    int coding;                         /* -1 = none */
    int32_t sum_high, sum_low;
    unsigned cnt_high, cnt_low;
    /* sync 2 */
    int s2_state;
    uint16_t nr_dots, c, inv_c;
    uint8_t nr_c;
    /* block */
    uint32_t blk[FX_BLOCK_WORDS];
    uint8_t cur_bit[4], cur_word[4], base_word[4];
    int32_t nr_symbols;
    int phase_ff;
    char msg_buf[256];
    size_t msg_len;
    /* sink */
    orc_msg *msgs;
    size_t cap, n;
};

static uint8_t fx_cksum(uint32_t w)                                 /* pager_flex.c:107-119 */
{
    uint8_t s = 0;
    w &= 0x1fffff;
    for (int i = 0; i < 6; i++) { s += w & 0xf; w >>= 4; }
    return s & 0xf;
}

static int fx_slice2(int16_t sample) { return !((uint16_t)sample >> 15); }             /* :129-138 */
static int fx_slice4(const orc_flex *f, int16_t sample)                                 /* :148-171 */
{
    sample = (int16_t)(sample - f->sample_delta);
    if (sample < 0) return (-sample > f->sample_range / 4) ? 0 : 1;
    return (sample > f->sample_range / 4) ? 2 : 3;
}
static int fx_slice(const orc_flex *f, int16_t sample)
{
    return flex_codings[f->coding].fsk_levels == 2 ? fx_slice2(sample) : fx_slice4(f, sample);
}

static void fx_sync_reset(orc_flex *f)                              /* :209-233 */
{
    memset(f->sync_words, 0, sizeof(f->sync_words));
    f->sync_state = FS_BS1;
    f->sample_counter = 0; f->bit_counter = 0;
    f->a = 0; f->b = 0; f->inv_a = 0; f->fiw = 0; f->coding = -1;
    f->sum_high = f->sum_low = 0; f->cnt_high = f->cnt_low = 0;
}

static void fx_reset(orc_flex *f)                                   /* :238-262, :173-207 */
{
    f->state = FX_SYNC_1;
    f->skip = 0; f->skip_count = 0;
    f->sample_range = 0; f->sample_delta = 0;
    f->frame_id = 0; f->cycle_id = 0;
    fx_sync_reset(f);
    f->s2_state = F2_COMMA; f->nr_dots = 0; f->c = 0; f->inv_c = 0; f->nr_c = 0;
    f->nr_symbols = 0; f->phase_ff = 0;
    for (int i = 0; i < 4; i++) { f->cur_bit[i] = 0; f->cur_word[i] = 0; f->base_word[i] = 0; }
}

orc_flex *orc_flex_new(size_t max_msgs)
{
    orc_flex *f = calloc(1, sizeof(*f));
    f->msgs = calloc(max_msgs ? max_msgs : 1, sizeof(orc_msg));
    f->cap = max_msgs;
    fx_reset(f);
    return f;
}
void orc_flex_delete(orc_flex *f) { if (f) { free(f->msgs); free(f); } }
size_t orc_flex_msgs(orc_flex *f, orc_msg **msgs) { *msgs = f->msgs; return f->n; }

static orc_msg *fx_msg(orc_flex *f, uint32_t kind, uint8_t phase, uint64_t capcode)
{
    if (f->n >= f->cap) return NULL;
    orc_msg *m = &f->msgs[f->n++];
    memset(m, 0, sizeof(*m));
    m->kind = kind; m->baud = flex_codings[f->coding].baud; m->function = phase;
    m->capcode_lo = (uint32_t)capcode; m->capcode_hi = (uint32_t)(capcode >> 32);
    m->aux[0] = f->cycle_id; m->aux[1] = f->frame_id;
    return m;
}

static void fx_range(orc_flex *f, int16_t sample)                   /* :352-358 and twins */
{
    if (sample > 0) { f->sum_high += sample; f->cnt_high++; }
    else            { f->sum_low += sample;  f->cnt_low++; }
}

static void fx_sync_update(orc_flex *f, int16_t sample)             /* :295-458 */
{
    f->sample_counter = (uint8_t)((f->sample_counter + 1) % 10);
    const uint32_t sym = (uint32_t)fx_slice2(sample);
    uint32_t *w = &f->sync_words[f->sample_counter];
    switch (f->sync_state) {
    case FS_SEARCH_BS1:
        *w = (*w << 1) | sym;
        if (*w == 0xaaaaaaaau) { f->bit_counter = 1; f->sync_state = FS_BS1; }
        break;
    case FS_BS1:
        *w = (*w << 1) | sym;
        if (*w == 0xaaaaaaaau) {
            f->bit_counter++;
        } else {
            if (f->bit_counter < 3) f->sync_state = FS_SEARCH_BS1;
            else { f->sync_state = FS_A; f->sample_counter = f->bit_counter / 2; }
            f->bit_counter = 0;
        }
        break;
    case FS_A:
        if (f->sample_counter == 0) {
            f->a = (f->a << 1) | sym;
            fx_range(f, sample);
            if (++f->bit_counter == 32) { f->sync_state = FS_B; f->bit_counter = 0; }
        }
        break;
    case FS_B:
        if (f->sample_counter == 0) {
            f->b = (uint16_t)((f->b << 1) | sym);
            fx_range(f, sample);
            if (++f->bit_counter == 16) { f->sync_state = FS_INV_A; f->bit_counter = 0; }
        }
        break;
    case FS_INV_A:
        if (f->sample_counter == 0) {
            f->inv_a = (f->inv_a << 1) | sym;
            fx_range(f, sample);
            if (++f->bit_counter == 32) {
                /* _pager_flex_sync_check_baud :264-287.  The second test, popcount(~seq_a ^ inv_a>>16) < 4, is
                 * evaluated in int: ~seq_a has its upper 16 bits set, the count is always >= 16 -- dead. */
                const uint16_t coding_a = (f->a >> 16) & 0xffff, inv_coding_a = (f->inv_a >> 16) & 0xffff;
                int found = -1;
                for (int i = 0; i < 4 && found < 0; i++)
                    if (__builtin_popcount(flex_codings[i].seq_a ^ coding_a) < 4 ||
                        __builtin_popcount(~flex_codings[i].seq_a ^ inv_coding_a) < 4) found = i;
                if (found >= 0) { f->coding = found; f->sync_state = FS_FIW; }
                else fx_sync_reset(f);
                f->bit_counter = 0;
            }
        }
        break;
    case FS_FIW:
        if (f->sample_counter == 0) {
            f->fiw = (f->fiw >> 1) | (sym << 31);
            fx_range(f, sample);
            if (++f->bit_counter == 32) {
                /* :438-442.  A zero count divides by zero in the reference (SIGFPE); here the frame is dropped. */
                if (f->cnt_high == 0 || f->cnt_low == 0) { fx_reset(f); break; }
                const int16_t hi = (int16_t)(f->sum_high / (int)f->cnt_high), lo = (int16_t)(f->sum_low / (int)f->cnt_low);
                f->sample_range = (int16_t)(hi - lo);
                f->sample_delta = (int16_t)(hi - (int)f->sample_range / 2);
                f->sync_state = FS_SYNCED;
            }
        }
        break;
    default:
        break;
    }
}

static int fx_handle_fiw(orc_flex *f)                               /* :1312-1345 */
{
    uint32_t fiw = f->fiw & 0x7fffffffu;
    if (orc_bch_decode(&fiw)) return 0;
    f->cycle_id = (fiw >> 4) & 0xf;
    f->frame_id = (fiw >> 8) & 0x7f;
    return fx_cksum(fiw) == 0xf;
}

static void fx_sync2_update(orc_flex *f, int16_t sample)            /* :460-525 */
{
    const struct flex_coding *cd = &flex_codings[f->coding];
    switch (f->s2_state) {
    case F2_COMMA:
        if (cd->sync_2_samples == ++f->nr_dots) f->s2_state = F2_C;
        break;
    case F2_C:
        f->c = (uint16_t)((f->c << cd->sym_bits) | fx_slice(f, sample));
        f->nr_c += cd->sym_bits;
        if (f->nr_c == 16) { f->s2_state = F2_INV_COMMA; f->nr_dots = 0; }
        break;
    case F2_INV_COMMA:
        if (cd->sync_2_samples == ++f->nr_dots) { f->s2_state = F2_INV_C; f->nr_c = 0; }
        break;
    case F2_INV_C:
        f->inv_c = (uint16_t)((f->inv_c << cd->sym_bits) | fx_slice(f, sample));
        f->nr_c += cd->sym_bits;
        if (f->nr_c == 16) f->s2_state = F2_SYNCED;
        break;
    default:
        break;
    }
}

/* --- phase processing over the flat block image --- */
#define FX_OOB 0xffffffffu
static int fx_in(size_t idx) { return idx < FX_BLOCK_WORDS; }

static int fx_decode_address(orc_flex *f, size_t ai, uint64_t *capcode, size_t *nr_words)     /* :527-573 */
{
    *capcode = 0; *nr_words = 0;
    if (!fx_in(ai)) return -1;
    if (orc_bch_decode(&f->blk[ai])) return -1;
    const uint32_t first = f->blk[ai] &= 0x1fffff;
    if ((first > 0x8000 && first <= 0x1e0000) || (first > 0x1f0000 && first < 0x1f7fff)) {
        *capcode = first - 32768;
    } else {
        if (!fx_in(ai + 1)) return -1;
        if (orc_bch_decode(&f->blk[ai + 1])) return -1;
        const uint32_t second = f->blk[ai + 1] &= 0x1fffff;
        *nr_words = 1;
        *capcode = (uint32_t)(0x1f9001u + (((0x1fffffu - second) * 32768u) + first - 1u));    /* 32-bit arithmetic */
    }
    return 0;
}

static int fx_alnum(orc_flex *f, uint8_t phase, uint64_t capcode, uint32_t long_word, size_t base, size_t nr_words)  /* :597-681 */
{
    size_t first_char_word = 1;
    int skip_word = 0;
    uint32_t status;
    if (long_word != 0xffffffffu) { first_char_word = 0; status = long_word; }
    else {
        if (!fx_in(base)) return -1;
        status = f->blk[base];
        if (orc_bch_decode(&status)) return -1;
    }
    const int fragment = (status & (1u << 10)) != 0;
    const uint8_t seq = (status >> 11) & 3;
    int maildrop = 0;
    if (seq == 3) { skip_word = 1; maildrop = (status & (1u << 20)) != 0; }
    for (size_t i = first_char_word; i < nr_words; i++) {
        if (!fx_in(base + i)) return -1;
        uint32_t cw = f->blk[base + i];
        if (orc_bch_decode(&cw)) return -1;
        if (skip_word) cw >>= 7;
        for (size_t j = (size_t)skip_word; j < 3; j++) {
            const uint8_t ch = cw & 0x7f;
            if (ch != 0x3) f->msg_buf[f->msg_len++] = (char)ch; else break;
            if (f->msg_len == 255) break;
            cw >>= 7;
        }
        skip_word = 0;
        if (f->msg_len == 255) break;
    }
    orc_msg *m = fx_msg(f, 2, phase, capcode);
    if (m) { m->aux[2] = (uint32_t)fragment; m->aux[3] = (uint32_t)maildrop; m->aux[4] = seq;
             m->len = (uint32_t)f->msg_len; memcpy(m->data, f->msg_buf, f->msg_len); }
    return 0;
}

static const char fx_num_lut[16] = { '0','1','2','3','4','5','6','7','8','9','X','U',' ','-',']','[' };  /* :686-704 */

static int fx_numeric(orc_flex *f, uint8_t phase, uint64_t capcode, uint32_t long_word, size_t base, size_t nr_words)  /* :709-824 */
{
    uint32_t cur = 0, next = 0;
    size_t nr_bits = nr_words * 21, cur_bits = 19, next_offs = 0, next_bits = 21;
    if (long_word != 0xffffffffu) {
        cur = (long_word & 0x1fffff) >> 2;
        nr_bits += 19; cur_bits = 19; next_offs = 0;
    } else {
        if (!fx_in(base)) return -1;
        cur = f->blk[base];
        if (orc_bch_decode(&cur)) return -1;
        cur &= 0x1fffff; cur >>= 2;
        cur_bits = 19; nr_bits -= 2; next_offs = 1;
    }
    if (next_offs < nr_words) {
        if (!fx_in(base + next_offs)) return -1;
        next = f->blk[base + next_offs];
        if (orc_bch_decode(&next)) return -1;
        next_bits = 21; next &= 0x1fffff;
    }
    nr_bits &= ~(size_t)3;
    do {
        const size_t rem = cur_bits & ~(size_t)3;
        for (size_t i = 0; i < rem; i += 4) {
            f->msg_buf[f->msg_len++] = fx_num_lut[cur & 0xf];
            if (f->msg_len == 255) break;
            cur >>= 4; cur_bits -= 4; nr_bits -= 4;
        }
        if (f->msg_len == 255) break;
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
                next = f->blk[base + next_offs];
                if (orc_bch_decode(&next)) return -1;
                next &= 0x1fffff;
            }
        }
    } while (nr_bits != 0);
    orc_msg *m = fx_msg(f, 3, phase, capcode);
    if (m) { m->len = (uint32_t)f->msg_len; memcpy(m->data, f->msg_buf, f->msg_len); }
    return 0;
}

static int fx_tone(orc_flex *f, uint8_t phase, uint64_t capcode, uint32_t first, uint32_t second)   /* :829-883 */
{
    first &= 0x1fffff;
    switch ((first >> 7) & 3) {
    case 0: {
        first >>= 9;
        for (int i = 0; i < 3; i++) { f->msg_buf[f->msg_len++] = fx_num_lut[first & 0xf]; first >>= 4; }
        if (second != 0xffffffffu) {
            second &= 0x1fffff;
            for (int i = 0; i < 5; i++) { f->msg_buf[f->msg_len++] = fx_num_lut[second & 0xf]; second >>= 4; }
        }
        orc_msg *m = fx_msg(f, 3, phase, capcode);
        if (m) { m->len = (uint32_t)f->msg_len; memcpy(m->data, f->msg_buf, f->msg_len); }
        return 0;
    }
    case 1: case 2: return 0;           /* logged only */
    default: return -1;
    }
}

static int fx_siv(orc_flex *f, uint8_t phase, uint64_t capcode, uint32_t vec)                      /* :885-933 */
{
    vec &= 0x7fffff;
    if (fx_cksum(vec) != 0xf) return -1;
    orc_msg *m = fx_msg(f, 4, phase, capcode);
    if (m) { m->aux[2] = (vec >> 7) & 7; m->aux[3] = (vec >> 10) & 0x7ff; }
    return 0;
}

static int fx_vector(orc_flex *f, uint8_t phase, uint64_t capcode, size_t vi, size_t nr_vec, size_t base)  /* :938-1033 */
{
    f->msg_len = 0;
    for (size_t i = 0; i < nr_vec; i++) {
        if (!fx_in(vi + i)) return -1;
        if (orc_bch_decode(&f->blk[vi + i])) return -1;
    }
    const uint32_t vec = f->blk[vi];
    if (fx_cksum(vec) != 0xf) return -1;
    const uint8_t type = (vec >> 4) & 7;
    const size_t start = (vec >> 7) & 0x7f;
    const uint32_t long_word = (nr_vec == 2) ? f->blk[vi + 1] : 0xffffffffu;
    size_t len;
    switch (type) {
    case 2: return fx_tone(f, phase, capcode, vec, long_word);
    case 3:
        len = ((vec >> 14) & 7) + 1;
        if (nr_vec == 2) len -= 1;
        return fx_numeric(f, phase, capcode, long_word, base + start, len);
    case 5:
        len = (vec >> 14) & 0x7f;
        if (nr_vec == 2) len -= 1;      /* wraps to SIZE_MAX for length 0, like the reference */
        return fx_alnum(f, phase, capcode, long_word, base + start, len);
    case 1: return fx_siv(f, phase, capcode, vec);
    default: return 0;                  /* unsupported types are logged only */
    }
}

static void fx_phase_process(orc_flex *f, unsigned ph)                                              /* :1088-1198 */
{
    const size_t base = (size_t)ph * FX_PHASE_STRIDE;
    uint32_t biw = f->blk[base] & 0x7fffffffu;
    if (orc_bch_decode(&biw)) return;
    if (fx_cksum(biw) != 0xf) return;
    const uint8_t vsw = (biw >> 10) & 0x3f, eob = (biw >> 8) & 3;
    if (eob > vsw) return;
    /* extra BIWs (:1157-1159) only log */
    const size_t addr_start = 1 + (size_t)eob;
    for (size_t i = addr_start; i < vsw; i++) {
        const size_t vec_offs = i + vsw - addr_start;
        uint64_t capcode; size_t nr_words;
        if (fx_decode_address(f, base + i, &capcode, &nr_words)) return;
        (void)fx_vector(f, (uint8_t)ph, capcode, base + vec_offs, nr_words + 1, base);
        i += nr_words;
    }
}

static void fx_append_bit(orc_flex *f, int ph, int bit)                                             /* :1200-1222 */
{
    uint32_t *w = &f->blk[(size_t)ph * FX_PHASE_STRIDE + f->base_word[ph] + f->cur_word[ph]];
    *w = (*w >> 1) | ((uint32_t)(bit != 0) << 31);
    f->cur_word[ph] = (uint8_t)((f->cur_word[ph] + 1) % 8);
    if (f->cur_word[ph] == 0) f->cur_bit[ph]++;
    if (f->cur_bit[ph] == 32) { f->base_word[ph] += 8; f->cur_bit[ph] = 0; f->cur_word[ph] = 0; }
}

static void fx_block_update(orc_flex *f, int16_t sample)                                            /* :1224-1310 */
{
    const struct flex_coding *cd = &flex_codings[f->coding];
    const int sym = fx_slice(f, sample);
    switch (cd->nr_phases) {
    case 1: fx_append_bit(f, 0, sym == 1); break;
    case 2:
        if (cd->fsk_levels == 2) { fx_append_bit(f, f->phase_ff ? 2 : 0, sym == 1); f->phase_ff = !f->phase_ff; }
        else { fx_append_bit(f, 0, sym & 2); fx_append_bit(f, 2, sym & 1); }
        break;
    default:
        if (!f->phase_ff) { fx_append_bit(f, 0, sym & 2); fx_append_bit(f, 1, sym & 1); }
        else              { fx_append_bit(f, 2, sym & 2); fx_append_bit(f, 3, sym & 1); }
        f->phase_ff = !f->phase_ff;
        break;
    }
    if (++f->nr_symbols == cd->symbols_per_block) {
        /* materialise the bookkeeping words of struct pager_flex_block before the unchecked walks */
        for (int p = 0; p < 4; p++)
            f->blk[(size_t)p * FX_PHASE_STRIDE + 88] = f->cur_bit[p] | (uint32_t)f->cur_word[p] << 8 | (uint32_t)f->base_word[p] << 16;
        f->blk[4 * FX_PHASE_STRIDE] = (uint32_t)f->nr_symbols;
        f->blk[4 * FX_PHASE_STRIDE + 1] = (uint32_t)f->phase_ff;
        switch (cd->nr_phases) {
        case 1: fx_phase_process(f, 0); break;
        case 2: fx_phase_process(f, 0); fx_phase_process(f, 2); break;
        default: for (unsigned p = 0; p < 4; p++) fx_phase_process(f, p); break;
        }
        fx_reset(f);
    }
}

void orc_flex_on_pcm(orc_flex *f, const int16_t *pcm, size_t n)                                     /* :1401-1455 */
{
    for (size_t i = 0; i < n; i++) {
        if (f->skip_count != 0) { f->skip_count--; continue; }
        f->skip_count = f->skip;
        switch (f->state) {
        case FX_SYNC_1:
            fx_sync_update(f, pcm[i]);
            if (f->sync_state == FS_SYNCED) {
                if (fx_handle_fiw(f)) {
                    f->state = FX_SYNC_2;
                    f->skip = flex_codings[f->coding].sample_skip;
                    f->skip_count = (int16_t)(f->skip + flex_codings[f->coding].sample_fudge);
                } else {
                    fx_reset(f);
                }
            }
            break;
        case FX_SYNC_2:
            fx_sync2_update(f, pcm[i]);
            if (f->s2_state == F2_SYNCED) f->state = FX_BLOCK;
            break;
        case FX_BLOCK:
            fx_block_update(f, pcm[i]);
            break;
        }
    }
}

/* ---------------------------------------------------------------------- */
/* f4  Mueller-Muller timing recovery: pager/mueller_muller.c:10-115        */
/* (dead code in the reference's pipelines -- SURVEY.md F5 -- restated for  */
/* the standalone differential test).  fma = 1 reproduces what GNU C's      */
/* default -ffp-contract=fast makes of :80 and :95 on an FMA machine.       */
/* ---------------------------------------------------------------------- */
void orc_mm_init(orc_mm *mm, float kw, float km, float samples_per_bit, float error_min, float error_max)
{
    memset(mm, 0, sizeof(*mm));
    mm->w = mm->m = samples_per_bit;                            /* :19-20 */
    mm->kw = kw; mm->km = km; mm->error_min = error_min; mm->error_max = error_max;
}

static float mm_sign(float v) { return (float)(v > 0) - (float)(v < 0); }      /* :34-38 */

size_t orc_mm_process(orc_mm *mm, const int16_t *samples, size_t n, int16_t *decisions, size_t cap, int fma)
{
    float cur = mm->next_offset, w = mm->w, m = mm->m;
    const float nf = (float)n;
    size_t nd = 0;
    while (cur < nf && nd < cap) {                              /* :65 */
        const float sample = samples[(size_t)(cur + 0.5f)];     /* :66 */
        decisions[nd++] = (int16_t)sample;                      /* :70 */
        const float w_error = mm_sign(mm->last_sample) * sample - mm_sign(sample) * mm->last_sample;   /* :76, exact */
        if (fma) w = fmaf(w_error, mm->kw, w); else w += w_error * mm->kw;                             /* :79 */
        if (mm->error_min > w) w = mm->error_min; else if (mm->error_max < w) w = mm->error_max;       /* :86-90 */
        if (fma) m += fmaf(mm->km, sample, w); else m += w + mm->km * sample;                          /* :92 */
        cur += floorf(m);                                       /* :95 */
        m -= floorf(m);
        mm->last_sample = sample;
    }
    mm->next_offset = cur - nf;                                 /* :107-109 */
    mm->w = w; mm->m = m;
    return nd;
}
