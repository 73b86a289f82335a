/*
 * ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver around the reference's own, UNMODIFIED numeric objects
 * (compiled by oracle/Makefile straight from /root/reference into
 * oracle/_ref/libtslref*.so).  The reference's glue (multifm/demod.c,
 * multifm/receiver.c, decoder/decoder.c) needs the un-vendored TSL runtime
 * (work queues, worker threads, config) and cannot be compiled here, so this
 * file re-states ONLY their control flow and calls the real arithmetic:
 *
 *   ref_prepare_taps   <- multifm/demod.c:205-261 (_demod_fir_prepare), gain per
 *                         multifm/receiver.c:220
 *   ref_chan_*         <- multifm/demod.c:49-121 (demod_thread_process):
 *                         push a 4096-sample sample_buf (multifm/file_if.c:18),
 *                         loop { direct_fir_process(<=1024) ; multifm_fm_demod_process }
 *   ref_resamp_*       <- decoder/decoder.c:581-673 (process_samples):
 *                         1024-sample sample_bufs -> polyphase_fir_process(<=1024)
 *   ref_pocsag_*       <- decoder/decoder.c:651 + callbacks decoder.c:264-318
 *   ref_flex_*         <- decoder/decoder.c:649 + callbacks
 *   ref_bench_multifm  <- one pthread per channel (multifm/demod.c:339), all
 *                         channels consuming the same in-memory buffers.
 */
#include <filter/filter.h>
#include <filter/direct_fir.h>
#include <filter/polyphase_fir.h>
#include <filter/sample_buf.h>
#include <filter/complex.h>
#include <filter/dc_blocker.h>
#include <multifm/fm_demod.h>
#include <multifm/demod_base.h>
#include <multifm/fast_atan2f.h>
#include <pager/pager.h>
#include <pager/pager_pocsag.h>
#include <pager/pager_flex.h>
#include <pager/mueller_muller.h>
#include <pager/bch_code.h>

#include <tsl/assert.h>
#include <tsl/safe_alloc.h>

#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#define REF_IQ_BUF_SAMPLES   4096   /* multifm/file_if.c:18 SAMPLES_PER_BUF */
#define REF_LPF_OUTPUT_LEN   1024   /* multifm/demod.h:12 */
#define REF_PCM_BUF_SAMPLES  1024   /* decoder/decoder.c NR_SAMPLES */

/* ------------------------------------------------------------------ */
/* sample_buf helpers                                                   */
/* ------------------------------------------------------------------ */
static aresult_t _ref_release_free(struct sample_buf *buf)
{
    free(buf);
    return A_OK;
}

static aresult_t _ref_release_noop(struct sample_buf *buf)
{
    (void)buf;
    return A_OK;
}

static struct sample_buf *_ref_buf_new(size_t data_bytes, enum sample_type st)
{
    struct sample_buf *b = NULL;
    if (posix_memalign((void **)&b, 64, sizeof(*b) + data_bytes)) abort();
    memset(b, 0, sizeof(*b));
    b->refcount = 1;
    b->sample_type = st;
    b->sample_buf_bytes = (uint32_t)data_bytes;
    b->release = _ref_release_free;
    return b;
}

/* ------------------------------------------------------------------ */
/* a1: tap preparation (multifm/demod.c:205-261)                        */
/* ------------------------------------------------------------------ */
void ref_prepare_taps(const double *lpf_taps, size_t nr_taps, int32_t offset_hz, uint32_t sample_rate,
                      double gain, int16_t *c_re, int16_t *c_im)
{
    double f_offs = -2.0 * M_PI * (double)offset_hz / (double)sample_rate;
    for (size_t i = 0; i < nr_taps; i++) {
        const double complex lpf_tap = gain * cexp(CMPLX(0, f_offs * (double)i)) * lpf_taps[i];
        const double q15 = 1ll << Q_15_SHIFT;
        c_re[i] = (int16_t)(creal(lpf_tap) * q15);
        c_im[i] = (int16_t)(cimag(lpf_tap) * q15);
    }
}

double ref_db_to_gain(double db)
{
    return pow(10.0, db / 10.0); /* multifm/receiver.c:220 */
}

/* ------------------------------------------------------------------ */
/* a2-a4: one channel = direct_fir + fm_demod                           */
/* ------------------------------------------------------------------ */
struct ref_chan {
    struct direct_fir fir;
    struct demod_base *demod;
    int16_t filt_samp_buf[2 * REF_LPF_OUTPUT_LEN];
    int16_t out_buf[REF_LPF_OUTPUT_LEN];
    /* pending partial input buffer */
    struct sample_buf *pend;
    /* output sinks for the current call */
    int16_t *sink_iq;
    int16_t *sink_pcm;
    size_t sink_cap;
    size_t sink_n;
    size_t total_out;
};

void *ref_chan_new_taps(size_t nr_taps, const int16_t *c_re, const int16_t *c_im, unsigned decimation,
                        uint32_t sample_rate, int32_t offset_hz)
{
    struct ref_chan *ch = calloc(1, sizeof(*ch));
    if (FAILED(direct_fir_init(&ch->fir, nr_taps, c_re, c_im, decimation, true, sample_rate, offset_hz))) {
        free(ch);
        return NULL;
    }
    if (FAILED(multifm_fm_demod_init(&ch->demod))) abort();
    return ch;
}

void *ref_chan_new(const double *lpf_taps, size_t nr_taps, int32_t offset_hz, uint32_t sample_rate,
                   unsigned decimation, double gain)
{
    int16_t *c = calloc(2 * nr_taps, sizeof(int16_t));
    ref_prepare_taps(lpf_taps, nr_taps, offset_hz, sample_rate, gain, c, c + nr_taps);
    void *ch = ref_chan_new_taps(nr_taps, c, c + nr_taps, decimation, sample_rate, offset_hz);
    free(c);
    return ch;
}

void ref_chan_delete(void *h)
{
    struct ref_chan *ch = h;
    if (!ch) return;
    direct_fir_cleanup(&ch->fir);
    multifm_fm_demod_cleanup(&ch->demod);
    if (ch->pend) free(ch->pend);
    free(ch);
}

void ref_chan_get_state(void *h, int16_t *rot, int16_t *incr, uint32_t *rot_counter)
{
    struct ref_chan *ch = h;
    rot[0] = ch->fir.rot_phase_re; rot[1] = ch->fir.rot_phase_im;
    incr[0] = ch->fir.rot_phase_incr_re; incr[1] = ch->fir.rot_phase_incr_im;
    *rot_counter = ch->fir.rot_counter;
}

void ref_chan_get_taps(void *h, int16_t *c_re, int16_t *c_im)
{
    struct ref_chan *ch = h;
    memcpy(c_re, ch->fir.fir_real_coeff, ch->fir.nr_coeffs * sizeof(int16_t));
    memcpy(c_im, ch->fir.fir_imag_coeff, ch->fir.nr_coeffs * sizeof(int16_t));
}

/* demod_thread_process (multifm/demod.c:49-121) with FIFO writes replaced by memcpy sinks */
static void _ref_chan_deliver(struct ref_chan *ch, struct sample_buf *sbuf)
{
    bool can_process = false;

    TSL_BUG_IF_FAILED(direct_fir_push_sample_buf(&ch->fir, sbuf));
    TSL_BUG_IF_FAILED(direct_fir_can_process(&ch->fir, &can_process, NULL));

    while (true == can_process) {
        size_t nr_samples = 0, nr_pcm = 0, nr_bytes = 0;

        TSL_BUG_IF_FAILED(direct_fir_process(&ch->fir, ch->filt_samp_buf, REF_LPF_OUTPUT_LEN, &nr_samples));
        if (nr_samples == 0) {
            fprintf(stderr, "ref_chan: FIR made no progress (unequal buffer sizes?)\n");
            abort();
        }
        {
            TSL_BUG_IF_FAILED(multifm_fm_demod_process(ch->demod, ch->filt_samp_buf, nr_samples,
                        ch->out_buf, &nr_pcm, &nr_bytes));
            if (ch->sink_n + nr_pcm > ch->sink_cap) {
                fprintf(stderr, "ref_chan: output sink overflow\n");
                abort();
            }
            if (ch->sink_iq) memcpy(ch->sink_iq + 2 * ch->sink_n, ch->filt_samp_buf, nr_samples * 2 * sizeof(int16_t));
            if (ch->sink_pcm) memcpy(ch->sink_pcm + ch->sink_n, ch->out_buf, nr_pcm * sizeof(int16_t));
            ch->sink_n += nr_pcm;
            ch->total_out += nr_pcm;
        }
        TSL_BUG_IF_FAILED(direct_fir_can_process(&ch->fir, &can_process, NULL));
    }
}

/*
 * Stream n_complex IQ samples through the channel.  Input is cut into
 * 4096-sample sample_bufs exactly like _file_read_cs16 (multifm/file_if.c:47).
 * A partial trailing buffer is held back until completed by a later call, or
 * delivered by ref_chan_flush() (a short final read() at EOF).
 * Returns the number of outputs written to out_iq (2 int16 each) / out_pcm.
 */
size_t ref_chan_run(void *h, const int16_t *iq, size_t n_complex, int16_t *out_iq, int16_t *out_pcm, size_t cap)
{
    struct ref_chan *ch = h;
    ch->sink_iq = out_iq; ch->sink_pcm = out_pcm; ch->sink_cap = cap; ch->sink_n = 0;

    while (n_complex != 0) {
        if (!ch->pend) ch->pend = _ref_buf_new(REF_IQ_BUF_SAMPLES * 4, COMPLEX_INT_16);
        size_t room = REF_IQ_BUF_SAMPLES - ch->pend->nr_samples;
        size_t take = n_complex < room ? n_complex : room;
        memcpy(ch->pend->data_buf + 4 * (size_t)ch->pend->nr_samples, iq, take * 4);
        ch->pend->nr_samples += (uint32_t)take;
        iq += 2 * take; n_complex -= take;
        if (ch->pend->nr_samples == REF_IQ_BUF_SAMPLES) {
            struct sample_buf *b = ch->pend;
            ch->pend = NULL;
            _ref_chan_deliver(ch, b);
        }
    }
    return ch->sink_n;
}

size_t ref_chan_flush(void *h, int16_t *out_iq, int16_t *out_pcm, size_t cap)
{
    /* A short final buffer cannot be replayed: with unequal buffer sizes the x86 path computes the
     * new sample_offset from the NEW buffer's length (filter/direct_fir.c:394-401), after which
     * nr_samples over-counts and demod_thread_process spins on a FIR that returns 0 outputs.
     * The reference itself aborts at EOF (multifm/receiver.c:84).  So the partial tail is dropped:
     * the comparable output prefix is floor(N/4096)*4096 input samples. */
    struct ref_chan *ch = h;
    (void)out_iq; (void)out_pcm; (void)cap;
    if (ch->pend) { free(ch->pend); ch->pend = NULL; }
    return 0;
}

/* a4 in isolation */
void ref_fm_demod(const int16_t *iq, size_t n, int16_t *pcm)
{
    struct demod_base *d = NULL;
    size_t n_out = 0, n_bytes = 0;
    if (n == 0) return;
    multifm_fm_demod_init(&d);
    multifm_fm_demod_process(d, (int16_t *)iq, n, pcm, &n_out, &n_bytes);
    multifm_fm_demod_cleanup(&d);
}

float ref_fast_atan2f(float y, float x)
{
    return fast_atan2f(y, x);
}

/* ------------------------------------------------------------------ */
/* a5: rational resampler driven like decoder.c:process_samples        */
/* ------------------------------------------------------------------ */
struct ref_resamp {
    struct polyphase_fir *pfir;
    struct sample_buf *pend;
    struct dc_blocker blck;
    int use_dc;
    int16_t output_buf[REF_PCM_BUF_SAMPLES];
};

void *ref_resamp_new(const int16_t *taps, size_t nr_taps, unsigned interp, unsigned decim, int use_dc, double dc_pole)
{
    struct ref_resamp *r = calloc(1, sizeof(*r));
    if (FAILED(polyphase_fir_new(&r->pfir, nr_taps, taps, interp, decim))) { free(r); return NULL; }
    r->use_dc = use_dc;
    if (use_dc) dc_blocker_init(&r->blck, dc_pole);
    return r;
}

void ref_resamp_delete(void *h)
{
    struct ref_resamp *r = h;
    if (!r) return;
    /* the reference never releases queued buffers on delete; drop them here */
    polyphase_fir_delete(&r->pfir);
    if (r->pend) free(r->pend);
    free(r);
}

typedef void (*ref_pcm_sink_t)(void *ctx, const int16_t *pcm, size_t n);

static void _ref_resamp_drain(struct ref_resamp *r, ref_pcm_sink_t sink, void *ctx)
{
    for (;;) {
        size_t new_samples = 0;
        TSL_BUG_IF_FAILED(polyphase_fir_process(r->pfir, r->output_buf, REF_PCM_BUF_SAMPLES, &new_samples));
        if (0 == new_samples) break;
        if (r->use_dc) dc_blocker_apply(&r->blck, r->output_buf, new_samples);
        sink(ctx, r->output_buf, new_samples);
    }
}

static void _ref_resamp_feed(struct ref_resamp *r, const int16_t *pcm, size_t n, ref_pcm_sink_t sink, void *ctx)
{
    while (n != 0) {
        bool full = false;
        TSL_BUG_IF_FAILED(polyphase_fir_full(r->pfir, &full));
        if (full) {
            _ref_resamp_drain(r, sink, ctx);
            TSL_BUG_IF_FAILED(polyphase_fir_full(r->pfir, &full));
            if (full) { fprintf(stderr, "ref_resamp: stuck full\n"); abort(); }
        }
        if (!r->pend) r->pend = _ref_buf_new(REF_PCM_BUF_SAMPLES * sizeof(int16_t), REAL_UINT_16);
        size_t room = REF_PCM_BUF_SAMPLES - r->pend->nr_samples;
        size_t take = n < room ? n : room;
        memcpy(r->pend->data_buf + 2 * (size_t)r->pend->nr_samples, pcm, take * 2);
        r->pend->nr_samples += (uint32_t)take;
        pcm += take; n -= take;
        if (r->pend->nr_samples == REF_PCM_BUF_SAMPLES) {
            TSL_BUG_IF_FAILED(polyphase_fir_push_sample_buf(r->pfir, r->pend));
            r->pend = NULL;
        }
        _ref_resamp_drain(r, sink, ctx);
    }
}

struct _ref_memsink { int16_t *out; size_t cap, n; };
static void _ref_memsink_put(void *ctx, const int16_t *pcm, size_t n)
{
    struct _ref_memsink *s = ctx;
    if (s->n + n > s->cap) { fprintf(stderr, "ref_resamp: sink overflow\n"); abort(); }
    memcpy(s->out + s->n, pcm, n * sizeof(int16_t));
    s->n += n;
}

size_t ref_resamp_run(void *h, const int16_t *pcm, size_t n, int16_t *out, size_t cap)
{
    struct _ref_memsink s = { out, cap, 0 };
    _ref_resamp_feed(h, pcm, n, _ref_memsink_put, &s);
    return s.n;
}

/* ------------------------------------------------------------------ */
/* a6/a7: POCSAG, messages captured as records                          */
/* ------------------------------------------------------------------ */
struct ref_msg {
    uint32_t kind;      /* 0 numeric, 1 alpha; FLEX: 2 alnum 3 num 4 siv */
    uint32_t baud;
    uint32_t capcode_lo;
    uint32_t capcode_hi;
    uint32_t function;  /* POCSAG function; FLEX: phase */
    uint32_t len;
    uint32_t aux[6];    /* FLEX: cycle, frame, fragmented, maildrop, seq, 0 */
    char data[520];
};

struct ref_msgbuf {
    struct ref_msg *msgs;
    size_t cap, n, dropped;
};

static struct ref_msg *_ref_msg_next(struct ref_msgbuf *mb)
{
    if (mb->n >= mb->cap) { mb->dropped++; return NULL; }
    struct ref_msg *m = &mb->msgs[mb->n++];
    memset(m, 0, sizeof(*m));
    return m;
}

struct ref_pocsag {
    struct pager_pocsag *pocsag;
    struct ref_msgbuf mb;
};

/* there is no user pointer in the callback, so keep a TLS back-pointer */
static __thread struct ref_pocsag *_ref_cur_pocsag;

static aresult_t _ref_on_pocsag(struct pager_pocsag *p, uint16_t baud, uint32_t capcode, const char *data,
                                size_t len, uint8_t function, uint32_t kind)
{
    (void)p;
    struct ref_msg *m = _ref_msg_next(&_ref_cur_pocsag->mb);
    if (!m) return A_OK;
    m->kind = kind; m->baud = baud; m->capcode_lo = capcode; m->function = function;
    m->len = (uint32_t)len;
    memcpy(m->data, data, len < sizeof(m->data) ? len : sizeof(m->data));
    return A_OK;
}

static aresult_t _ref_on_pocsag_num(struct pager_pocsag *p, uint16_t baud, uint32_t capcode, const char *data,
                                    size_t len, uint8_t function)
{
    return _ref_on_pocsag(p, baud, capcode, data, len, function, 0);
}

static aresult_t _ref_on_pocsag_alpha(struct pager_pocsag *p, uint16_t baud, uint32_t capcode, const char *data,
                                      size_t len, uint8_t function)
{
    return _ref_on_pocsag(p, baud, capcode, data, len, function, 1);
}

void *ref_pocsag_new(uint32_t freq_hz, size_t max_msgs)
{
    struct ref_pocsag *r = calloc(1, sizeof(*r));
    r->mb.msgs = calloc(max_msgs ? max_msgs : 1, sizeof(struct ref_msg));
    r->mb.cap = max_msgs;
    if (FAILED(pager_pocsag_new(&r->pocsag, freq_hz, _ref_on_pocsag_num, _ref_on_pocsag_alpha, false))) abort();
    return r;
}

void ref_pocsag_delete(void *h)
{
    struct ref_pocsag *r = h;
    if (!r) return;
    pager_pocsag_delete(&r->pocsag);
    free(r->mb.msgs);
    free(r);
}

/* feed in `chunk`-sample calls (0 = one call) */
void ref_pocsag_run(void *h, const int16_t *pcm, size_t n, size_t chunk)
{
    struct ref_pocsag *r = h;
    _ref_cur_pocsag = r;
    if (chunk == 0) chunk = n;
    while (n != 0) {
        size_t take = n < chunk ? n : chunk;
        TSL_BUG_IF_FAILED(pager_pocsag_on_pcm(r->pocsag, pcm, take));
        pcm += take; n -= take;
    }
    _ref_cur_pocsag = NULL;
}

size_t ref_pocsag_msgs(void *h, struct ref_msg **pmsgs, size_t *dropped)
{
    struct ref_pocsag *r = h;
    *pmsgs = r->mb.msgs;
    if (dropped) *dropped = r->mb.dropped;
    return r->mb.n;
}

size_t ref_msg_size(void) { return sizeof(struct ref_msg); }

/* BCH(31,21) in isolation (pager/bch_code.c:307), poly per pager_pocsag.c:150 */
int ref_bch_decode(uint32_t *word)
{
    static const int poly[6] = { 1, 0, 1, 0, 0, 1 };
    static __thread struct bch_code *bch;
    if (!bch) TSL_BUG_IF_FAILED(bch_code_new(&bch, poly, 5, 31, 21, 2));
    return bch_code_decode(bch, word);
}

/* resampler -> POCSAG chained exactly like decoder.c:635-651 */
static void _ref_pocsag_sink(void *ctx, const int16_t *pcm, size_t n)
{
    struct ref_pocsag *r = ctx;
    TSL_BUG_IF_FAILED(pager_pocsag_on_pcm(r->pocsag, pcm, n));
}

void ref_decoder_pocsag_run(void *resamp, void *pocsag, const int16_t *pcm, size_t n)
{
    _ref_cur_pocsag = pocsag;
    _ref_resamp_feed(resamp, pcm, n, _ref_pocsag_sink, pocsag);
    _ref_cur_pocsag = NULL;
}

/* ------------------------------------------------------------------ */
/* a8: FLEX                                                             */
/* ------------------------------------------------------------------ */
struct ref_flex {
    struct pager_flex *flex;
    struct ref_msgbuf mb;
};
static __thread struct ref_flex *_ref_cur_flex;

static aresult_t _ref_on_flex_alnum(struct pager_flex *f, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame,
        uint64_t cap_code, bool fragmented, bool maildrop, uint8_t seq, const char *msg, size_t len)
{
    (void)f;
    struct ref_msg *m = _ref_msg_next(&_ref_cur_flex->mb);
    if (!m) return A_OK;
    m->kind = 2; m->baud = baud; m->function = phase;
    m->capcode_lo = (uint32_t)cap_code; m->capcode_hi = (uint32_t)(cap_code >> 32);
    m->aux[0] = cycle; m->aux[1] = frame; m->aux[2] = fragmented; m->aux[3] = maildrop; m->aux[4] = seq;
    m->len = (uint32_t)len;
    memcpy(m->data, msg, len < sizeof(m->data) ? len : sizeof(m->data));
    return A_OK;
}

static aresult_t _ref_on_flex_num(struct pager_flex *f, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame,
        uint64_t cap_code, const char *msg, size_t len)
{
    (void)f;
    struct ref_msg *m = _ref_msg_next(&_ref_cur_flex->mb);
    if (!m) return A_OK;
    m->kind = 3; m->baud = baud; m->function = phase;
    m->capcode_lo = (uint32_t)cap_code; m->capcode_hi = (uint32_t)(cap_code >> 32);
    m->aux[0] = cycle; m->aux[1] = frame;
    m->len = (uint32_t)len;
    memcpy(m->data, msg, len < sizeof(m->data) ? len : sizeof(m->data));
    return A_OK;
}

static aresult_t _ref_on_flex_siv(struct pager_flex *f, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame,
        uint64_t cap_code, uint8_t siv_msg_type, uint32_t data)
{
    (void)f;
    struct ref_msg *m = _ref_msg_next(&_ref_cur_flex->mb);
    if (!m) return A_OK;
    m->kind = 4; m->baud = baud; m->function = phase;
    m->capcode_lo = (uint32_t)cap_code; m->capcode_hi = (uint32_t)(cap_code >> 32);
    m->aux[0] = cycle; m->aux[1] = frame; m->aux[2] = siv_msg_type; m->aux[3] = data;
    return A_OK;
}

void *ref_flex_new(uint32_t freq_hz, size_t max_msgs)
{
    struct ref_flex *r = calloc(1, sizeof(*r));
    r->mb.msgs = calloc(max_msgs ? max_msgs : 1, sizeof(struct ref_msg));
    r->mb.cap = max_msgs;
    if (FAILED(pager_flex_new(&r->flex, freq_hz, _ref_on_flex_alnum, _ref_on_flex_num, _ref_on_flex_siv))) abort();
    return r;
}

void ref_flex_delete(void *h)
{
    struct ref_flex *r = h;
    if (!r) return;
    pager_flex_delete(&r->flex);
    free(r->mb.msgs);
    free(r);
}

void ref_flex_run(void *h, const int16_t *pcm, size_t n, size_t chunk)
{
    struct ref_flex *r = h;
    _ref_cur_flex = r;
    if (chunk == 0) chunk = n;
    while (n != 0) {
        size_t take = n < chunk ? n : chunk;
        TSL_BUG_IF_FAILED(pager_flex_on_pcm(r->flex, pcm, take));
        pcm += take; n -= take;
    }
    _ref_cur_flex = NULL;
}

size_t ref_flex_msgs(void *h, struct ref_msg **pmsgs, size_t *dropped)
{
    struct ref_flex *r = h;
    *pmsgs = r->mb.msgs;
    if (dropped) *dropped = r->mb.dropped;
    return r->mb.n;
}

/* ------------------------------------------------------------------ */
/* f4: Mueller-Muller timing recovery (pager/mueller_muller.c; not wired */
/* into any reference pipeline, tested standalone)                       */
/* ------------------------------------------------------------------ */
size_t ref_mm_run(float kw, float km, float samples_per_bit, float error_min, float error_max,
                  const int16_t *pcm, size_t n, size_t chunk, int16_t *decisions, size_t cap, float state_out[4])
{
    struct mueller_muller mm;
    TSL_BUG_IF_FAILED(mm_init(&mm, kw, km, samples_per_bit, error_min, error_max));
    size_t total = 0;
    if (chunk == 0) chunk = n;
    while (n != 0) {
        size_t take = n < chunk ? n : chunk, got = 0;
        TSL_BUG_IF_FAILED(mm_process(&mm, pcm, take, decisions + total, cap - total, &got));
        total += got; pcm += take; n -= take;
    }
    state_out[0] = mm.w; state_out[1] = mm.m; state_out[2] = mm.next_offset; state_out[3] = mm.last_sample;
    return total;
}

/* ------------------------------------------------------------------ */
/* CPU baseline: the reference's threading model, timed                  */
/* ------------------------------------------------------------------ */
struct _ref_bench_thr {
    pthread_t tid;
    struct ref_chan *ch;
    const int16_t *iq;
    size_t n_bufs;          /* number of 4096-sample buffers in iq */
    size_t reps;            /* how many passes over iq */
    struct sample_buf **bufs; /* shared by all channel threads, like receiver.c:86-95 */
    size_t outputs;
    pthread_barrier_t *bar;
};

static void *_ref_bench_main(void *arg)
{
    struct _ref_bench_thr *t = arg;
    struct ref_chan *ch = t->ch;
    ch->sink_iq = NULL; ch->sink_pcm = NULL; ch->sink_cap = (size_t)-1; ch->sink_n = 0;
    pthread_barrier_wait(t->bar);
    for (size_t r = 0; r < t->reps; r++) {
        for (size_t b = 0; b < t->n_bufs; b++) {
            _ref_chan_deliver(ch, t->bufs[b]);
        }
    }
    t->outputs = ch->total_out;
    pthread_barrier_wait(t->bar);
    return NULL;
}

/*
 * Run nr_channels channel threads (one pthread each, like demod_thread_new)
 * over the same in-memory 4096-sample sample_bufs (shared, refcounted like
 * receiver.c:86-95; the refcount is preset so that no buffer is released
 * before the run ends), `reps` passes.  Buffers are built OUTSIDE the timed
 * region; the timed region is push/FIR/FM only (no file or FIFO I/O).
 * Needs n_complex >= 3*4096.  Returns seconds; total PCM outputs in
 * *total_outputs.
 */
double ref_bench_multifm(size_t nr_channels, const double *lpf_taps, size_t nr_taps, const int32_t *offsets_hz,
                         uint32_t sample_rate, unsigned decimation, const int16_t *iq, size_t n_complex,
                         size_t reps, size_t *total_outputs)
{
    size_t n_bufs = n_complex / REF_IQ_BUF_SAMPLES;
    struct _ref_bench_thr *thr = calloc(nr_channels, sizeof(*thr));
    pthread_barrier_t bar;
    struct timespec t0, t1;

    if (n_bufs < 3) return -1.0;
    struct sample_buf **bufs = calloc(n_bufs, sizeof(struct sample_buf *));
    for (size_t b = 0; b < n_bufs; b++) {
        struct sample_buf *sb = _ref_buf_new(REF_IQ_BUF_SAMPLES * 4, COMPLEX_INT_16);
        sb->release = _ref_release_noop;
        sb->refcount = (uint32_t)(nr_channels * reps + 1);
        sb->nr_samples = REF_IQ_BUF_SAMPLES;
        memcpy(sb->data_buf, iq + 2 * b * REF_IQ_BUF_SAMPLES, REF_IQ_BUF_SAMPLES * 4);
        bufs[b] = sb;
    }
    pthread_barrier_init(&bar, NULL, (unsigned)nr_channels + 1);
    for (size_t c = 0; c < nr_channels; c++) {
        thr[c].ch = ref_chan_new(lpf_taps, nr_taps, offsets_hz[c], sample_rate, decimation, 1.0);
        thr[c].iq = iq; thr[c].n_bufs = n_bufs; thr[c].reps = reps; thr[c].bar = &bar;
        thr[c].bufs = bufs;
        pthread_create(&thr[c].tid, NULL, _ref_bench_main, &thr[c]);
    }

    pthread_barrier_wait(&bar);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&bar);
    clock_gettime(CLOCK_MONOTONIC, &t1);

    size_t total = 0;
    for (size_t c = 0; c < nr_channels; c++) {
        pthread_join(thr[c].tid, NULL);
        total += thr[c].outputs;
        /* detach queued buffers before cleanup so they are not double freed */
        thr[c].ch->fir.sb_active = NULL; thr[c].ch->fir.sb_next = NULL;
        ref_chan_delete(thr[c].ch);
    }
    for (size_t b = 0; b < n_bufs; b++) free(bufs[b]);
    free(bufs);
    pthread_barrier_destroy(&bar);
    free(thr);
    if (total_outputs) *total_outputs = total;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
