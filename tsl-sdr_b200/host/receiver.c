/* receiver.c -- configuration, buffer pool and the single GPU consumer.  See receiver.h. */
#define _GNU_SOURCE
#include "receiver.h"
#include "../../include/tslb200_gpuchan.h"
#include "../../include/tslb200_gpupager.h"

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

/* ---- buffer pool (frame_alloc_new(sizeof(sample_buf)+samples*4, nrSampBufs), receiver.c:154) ---- */
static aresult_t _sample_buf_release(struct sample_buf *buf)
{
    struct receiver *rx = buf->priv;
    pthread_mutex_lock(&rx->pool_mtx);
    rx->pool[rx->pool_free++] = buf;
    pthread_mutex_unlock(&rx->pool_mtx);
    return A_OK;
}

aresult_t receiver_sample_buf_alloc(struct receiver *rx, struct sample_buf **pbuf)
{
    if (!rx || !pbuf) return A_E_BADARGS;
    *pbuf = NULL;
    pthread_mutex_lock(&rx->pool_mtx);
    struct sample_buf *b = rx->pool_free ? rx->pool[--rx->pool_free] : NULL;
    pthread_mutex_unlock(&rx->pool_mtx);
    if (!b) {
        if (0 == rx->nr_samp_buf_alloc_fails)
            B200_MSG("I", "NO-SAMPLE-BUFFER", "There are no available sample buffers, dropping received samples.");
        rx->nr_samp_buf_alloc_fails++;
        return A_E_NOMEM;
    }
    b->release = _sample_buf_release;
    b->priv = rx;
    b->nr_samples = 0;
    atomic_store((_Atomic uint32_t *)&b->refcount, 1);      /* the source owns it until it is delivered */
    *pbuf = b;
    return A_OK;
}

/* ---- delivery: ONE consumer, so refcount = 1 (the reference sets nr_demod_threads, receiver.c:86) ---- */
aresult_t receiver_sample_buf_deliver(struct receiver *rx, struct sample_buf *buf)
{
    if (!rx || !buf) return A_E_BADARGS;
    atomic_store((_Atomic uint32_t *)&buf->refcount, 1);
    if (0 == buf->nr_samples) { sample_buf_decref(buf); return A_E_INVAL; }     /* back to the pool, nothing queued */
    pthread_mutex_lock(&rx->q_mtx);
    while (((rx->q_head + 1) & 127) == rx->q_tail && rx->running)     /* queue full: back-pressure the file source */
        pthread_cond_wait(&rx->q_cv, &rx->q_mtx);
    if (((rx->q_head + 1) & 127) == rx->q_tail) {                     /* still full: the receiver stopped -- drop, never overwrite */
        pthread_mutex_unlock(&rx->q_mtx);
        sample_buf_decref(buf);
        return A_E_BUSY;
    }
    rx->queue[rx->q_head] = buf;
    rx->q_head = (rx->q_head + 1) & 127;
    pthread_mutex_unlock(&rx->q_mtx);
    pthread_cond_broadcast(&rx->q_cv);
    return A_OK;
}

void receiver_end_of_stream(struct receiver *rx)
{
    pthread_mutex_lock(&rx->q_mtx);
    rx->producer_done = true;
    pthread_mutex_unlock(&rx->q_mtx);
    pthread_cond_broadcast(&rx->q_cv);
}

aresult_t receiver_set_mute(struct receiver *rx, bool mute) { if (!rx) return A_E_BADARGS; rx->muted = mute; return A_OK; }
bool receiver_thread_running(struct receiver *rx) { return rx && rx->running; }

/* ---- decoder.c:131-171 character escaping and :264-318 JSON lines ---- */
static void put_alnum_char(FILE *fp, char ch)
{
    switch (ch) {
    case '\n': fputs("\\n", fp); break;
    case '\r': fputs("\\n", fp); break;       /* sic: the reference prints \n for CR as well */
    case '"':  fputs("\\\"", fp); break;
    case '\\': fputs("\\\\", fp); break;
    case '/':  fputs("\\/", fp); break;
    case '\b': fputs("<BKSP>", fp); break;
    case '\f': fputs("<FF>", fp); break;
    case '\t': fputs("\\t", fp); break;
    case 0x03: case 0x04: case 0x17: fputc(' ', fp); break;
    default:
        if (ch >= 0x20 && ch <= 0x7e) fputc(ch, fp);
        else fprintf(fp, "\\u%04x", (unsigned)ch);
    }
}

/* Where a channel's lines go.  With one output for all channels every line carries "channel":N (our extension: the
 * reference runs one decoder process per channel, so its lines need no such key).  With pagerDecode.outFile containing
 * %u there is one file per channel and the lines are byte for byte what decoder.c prints. */
static FILE *line_out(struct receiver *rx, uint32_t channel, bool *with_key)
{
    if (rx->msg_out_ch && channel < rx->nr_demod_threads && rx->msg_out_ch[channel]) { *with_key = false; return rx->msg_out_ch[channel]; }
    *with_key = true;
    return rx->msg_out;
}

static void put_timestamp(FILE *fp)
{
    time_t now = time(NULL);
    struct tm gmt;
    gmtime_r(&now, &gmt);
    fprintf(fp, "\"timestamp\":\"%04i-%02i-%02i %02i:%02i:%02i UTC\",", gmt.tm_year + 1900, gmt.tm_mon + 1, gmt.tm_mday,
            gmt.tm_hour, gmt.tm_min, gmt.tm_sec);
}

/* decoder/decoder.c:264-318 */
static int on_pocsag_msg(struct receiver_pager *pg, const char *type, uint32_t row, uint16_t baud, uint32_t capcode,
                         const char *data, size_t len, uint8_t function)
{
    struct receiver *rx = pg->rx;
    const uint32_t channel = pg->channel_base + row;
    bool key;
    FILE *fp = line_out(rx, channel, &key);
    fprintf(fp, "{\"proto\":\"pocsag\",\"type\":\"%s\",", type);
    put_timestamp(fp);
    fprintf(fp, "\"baud\":%i,\"capCode\":%u,\"function\":%u,", baud, capcode, (unsigned)function);
    if (key) fprintf(fp, "\"channel\":%u,", channel);
    fputs("\"message\":\"", fp);
    for (size_t i = 0; i < len; i++) put_alnum_char(fp, data[i]);
    fputs("\"}\n", fp);
    fflush(fp);
    rx->nr_messages++;
    return 0;
}

static int on_alpha(void *user, uint32_t ch, uint16_t baud, uint32_t cap, const char *d, size_t n, uint8_t fn)
{
    return on_pocsag_msg(user, "alphanumeric", ch, baud, cap, d, n, fn);
}

static int on_numeric(void *user, uint32_t ch, uint16_t baud, uint32_t cap, const char *d, size_t n, uint8_t fn)
{
    return on_pocsag_msg(user, "numeric", ch, baud, cap, d, n, fn);
}

/* ---- FLEX JSON lines: decoder/decoder.c:173-262 (same keys, same order) ---- */
static const char flex_phase_id[4] = { 'A', 'B', 'C', 'D' };

static FILE *flex_head(struct receiver_pager *pg, const char *type, uint32_t row, uint16_t baud, uint8_t phase, uint8_t cycle_no,
                       uint8_t frame_no, uint64_t cap_code)
{
    const uint32_t channel = pg->channel_base + row;
    bool key;
    FILE *fp = line_out(pg->rx, channel, &key);
    fprintf(fp, "{\"proto\":\"flex\",\"type\":\"%s\",", type);
    put_timestamp(fp);
    fprintf(fp, "\"baud\":%i,\"syncLevel\":%i,\"frameNo\":%u,\"cycleNo\":%u,\"phaseNo\":\"%c\",\"capCode\":%llu,",
            baud, 0, frame_no, cycle_no, flex_phase_id[phase & 3], (unsigned long long)cap_code);
    if (key) fprintf(fp, "\"channel\":%u,", channel);
    return fp;
}

static int on_flex_alnum(void *user, uint32_t row, uint16_t baud, uint8_t phase, uint8_t cycle_no, uint8_t frame_no,
                         uint64_t cap_code, int fragmented, int maildrop, uint8_t seq_num, const char *msg, size_t len)
{
    struct receiver_pager *pg = user;
    FILE *fp = flex_head(pg, "alphanumeric", row, baud, phase, cycle_no, frame_no, cap_code);
    fprintf(fp, "\"fragment\":%s,\"maildrop\":%s,\"fragSeq\":%u,\"message\":\"", fragmented ? "true" : "false",
            maildrop ? "true" : "false", seq_num);
    for (size_t i = 0; i < len; i++) put_alnum_char(fp, msg[i]);
    fputs("\"}\n", fp);
    fflush(fp);
    pg->rx->nr_messages++;
    return 0;
}

static int on_flex_num(void *user, uint32_t row, uint16_t baud, uint8_t phase, uint8_t cycle_no, uint8_t frame_no,
                       uint64_t cap_code, const char *msg, size_t len)
{
    struct receiver_pager *pg = user;
    FILE *fp = flex_head(pg, "numeric", row, baud, phase, cycle_no, frame_no, cap_code);
    fputs("\"message\":\"", fp);
    for (size_t i = 0; i < len; i++) put_alnum_char(fp, msg[i]);
    fputs("\"}\n", fp);
    fflush(fp);
    pg->rx->nr_messages++;
    return 0;
}

static int on_flex_siv(void *user, uint32_t row, uint16_t baud, uint8_t phase, uint8_t cycle_no, uint8_t frame_no,
                       uint64_t cap_code, uint8_t siv_msg_type, uint32_t data)
{
    struct receiver_pager *pg = user;
    if (siv_msg_type != 0) return 0;            /* decoder.c:249-257 prints temporary address activations only */
    FILE *fp = flex_head(pg, "tempAddrActivation", row, baud, phase, cycle_no, frame_no, cap_code);
    fprintf(fp, "\"startFrameNo\":%u,\"tempAddressId\":%u}\n", data & 0x7f, (data >> 7) & 0xf);
    fflush(fp);
    pg->rx->nr_messages++;
    return 0;
}

/* ---- consumer: batch sample_bufs into pinned memory, submit, collect the previous batch, write FIFOs ---- */
static void write_outputs(struct receiver *rx, size_t n_out)
{
    for (size_t c = 0; c < rx->nr_demod_threads; c++) {
        struct receiver_channel *ch = &rx->channels[c];
        if (ch->debug_fd >= 0 && rx->iq_host) {
            if (write(ch->debug_fd, rx->iq_host + 2 * c * rx->pcm_cap, n_out * 4) < 0)
                B200_MSG("W", "CANT-WRITE-DEBUG-FILE", "%s", strerror(errno));
        }
        ch->total_nr_demod_samples += n_out;
        if (ch->fifo_fd < 0) continue;
        /* multifm/demod.c:93-110: EPIPE -> count drops until a reader returns; anything else is fatal */
        if (write(ch->fifo_fd, rx->pcm_host + c * rx->pcm_cap, n_out * sizeof(int16_t)) < 0) {
            if (errno == EPIPE) {
                if (0 == ch->nr_dropped_samples)
                    B200_MSG("W", "FIFO-REMOTE-END-DISCONNECTED", "Remote end of FIFO %s disconnected; dropping samples.", ch->out_fifo);
                ch->nr_dropped_samples += n_out;
            } else {
                B200_MSG("F", "FIFO-WRITE", "Failed to write %zu bytes to %s: %s", n_out * 2, ch->out_fifo, strerror(errno));
                abort();
            }
        } else if (ch->nr_dropped_samples) {
            B200_MSG("W", "FIFO-RESUMED", "Remote FIFO end reconnected. Dropped %zu samples in the interim.", ch->nr_dropped_samples);
            ch->nr_dropped_samples = 0;
        }
    }
}

static void collect_one(struct receiver *rx)
{
    size_t n_out = 0;
    if (gpuchan_multi_collect(rx->bank, rx->pcm_host, rx->pcm_cap, &n_out)) {
        B200_MSG("F", "GPU-COLLECT", "%s", gpuchan_last_error());
        abort();
    }
    if (rx->iq_host) {                          /* signalDebugFile taps: per bank, each its own channel range */
        for (uint32_t d = 0; d < gpuchan_multi_devices(rx->bank); d++) {
            gpuchan_t *bank = NULL;
            uint32_t first = 0, cnt = 0;
            size_t n_iq = 0;
            gpuchan_multi_bank(rx->bank, d, &bank, &first, &cnt);
            gpuchan_collect_iq(bank, rx->iq_host + 2 * (size_t)first * rx->pcm_cap, rx->pcm_cap, &n_iq);
        }
    }
    if (n_out) write_outputs(rx, n_out);
    rx->in_flight--;
}

static void submit_batch(struct receiver *rx)
{
    if (0 == rx->batch_fill) return;
    if (gpuchan_multi_submit(rx->bank, rx->batch[rx->batch_cur], rx->batch_fill)) {
        B200_MSG("F", "GPU-SUBMIT", "%s", gpuchan_last_error());
        abort();
    }
    rx->total_iq_samples += rx->batch_fill;
    /* in-process decoders: every pager bank reads its rows of its device's PCM (still on that device); all banks are fed
     * first, the callbacks fire afterwards, so the devices decode side by side */
    for (size_t p = 0; p < rx->nr_pagers; p++) {
        struct receiver_pager *pg = &rx->pagers[p];
        gpuchan_t *bank = NULL;
        const int16_t *d_pcm = NULL;
        size_t pitch = 0, n = 0;
        gpuchan_multi_bank(rx->bank, pg->device_index, &bank, NULL, NULL);
        gpuchan_device_pcm(bank, &d_pcm, &pitch, &n);
        gpuchan_sync(bank);                     /* pager stream is ordered after the bank's work */
        if (n && gpupager_feed_device(pg->bank, d_pcm, pitch, n, NULL)) {
            B200_MSG("F", "GPU-PAGER", "%s", gpupager_last_error());
            abort();
        }
    }
    for (size_t p = 0; p < rx->nr_pagers; p++) {
        struct receiver_pager *pg = &rx->pagers[p];
        size_t nr = 0;
        if (pg->is_flex) gpupager_dispatch_flex(pg->bank, on_flex_alnum, on_flex_num, on_flex_siv, pg, &nr);
        else gpupager_dispatch(pg->bank, on_numeric, on_alpha, pg, &nr);
    }
    rx->in_flight++;
    rx->batch_cur ^= 1;
    rx->batch_fill = 0;
    /* keep one batch in flight: collect the older one while the newer one copies/computes */
    while (rx->in_flight > 1) collect_one(rx);
}

static void *consumer_main(void *arg)
{
    struct receiver *rx = arg;
    const size_t batch_cap = rx->batch_bufs * rx->samples_per_buf;
    for (;;) {
        pthread_mutex_lock(&rx->q_mtx);
        while (rx->q_head == rx->q_tail && !rx->producer_done && rx->running) {
            struct timespec ts;
            clock_gettime(CLOCK_REALTIME, &ts);
            ts.tv_sec += 1;                                   /* 1 s timed wait like multifm/demod.c:146-154 */
            pthread_cond_timedwait(&rx->q_cv, &rx->q_mtx, &ts);
        }
        struct sample_buf *buf = NULL;
        if (rx->q_head != rx->q_tail) {
            buf = rx->queue[rx->q_tail];
            rx->q_tail = (rx->q_tail + 1) & 127;
        }
        const bool done = (buf == NULL) && (rx->producer_done || !rx->running);
        pthread_mutex_unlock(&rx->q_mtx);
        pthread_cond_broadcast(&rx->q_cv);
        if (buf) {
            if (!rx->muted) {
                memcpy(rx->batch[rx->batch_cur] + 2 * rx->batch_fill, buf->data_buf, (size_t)buf->nr_samples * 4);
                rx->batch_fill += buf->nr_samples;
            }
            sample_buf_decref(buf);                           /* the data now lives in pinned staging */
            if (rx->batch_fill + rx->samples_per_buf > batch_cap) submit_batch(rx);
        }
        if (done) break;
    }
    submit_batch(rx);                                         /* ragged tail: any length is a valid submit */
    while (rx->in_flight > 0) collect_one(rx);
    return NULL;
}

/* ---- configuration (multifm/receiver.c:100-256) ---- */
aresult_t receiver_init(struct receiver *rx, const jnode *cfg, receiver_rx_thread_func_t rx_func,
                        receiver_cleanup_func_t cleanup_func, size_t samples_per_buf)
{
    if (!rx || !cfg || !rx_func || !samples_per_buf) return A_E_BADARGS;
    memset(rx, 0, sizeof(*rx));
    rx->thread_func = rx_func; rx->cleanup_func = cleanup_func; rx->samples_per_buf = samples_per_buf;
    rx->muted = true;
    pthread_mutex_init(&rx->pool_mtx, NULL); pthread_mutex_init(&rx->q_mtx, NULL); pthread_cond_init(&rx->q_cv, NULL);

    int v = 0;
    rx->nr_samp_bufs = 64;                                                     /* receiver.c:133 default */
    /* tags and texts of multifm/receiver.c:133-184 */
    if (!json_get_int(cfg, "nrSampBufs", &v)) rx->nr_samp_bufs = v;
    else B200_MSG("I", "DEFAULT-SAMP-BUFS", "Setting sample buffer count to 64");
    if (rx->nr_samp_bufs <= 0) { B200_MSG("E", "BAD-SAMP-BUFS", "nrSampBufs of '%d' is not valid.", rx->nr_samp_bufs); return A_E_INVAL; }
    if (json_get_int(cfg, "sampleRateHz", &v)) { B200_MSG("I", "NO-SAMPLE-RATE", "Need to specify a sample rate, in Hertz."); return A_E_INVAL; }
    if (v <= 0) { B200_MSG("E", "BAD-SAMPLE-RATE", "Sample rate of '%d' is not valid.", v); return A_E_INVAL; }
    rx->sample_rate_hz = (uint32_t)v;
    if (json_get_int(cfg, "centerFreqHz", &v)) { B200_MSG("I", "NO-CENTER-FREQ", "You forgot to specify a center frequency, in Hz."); return A_E_INVAL; }
    rx->center_freq_hz = (uint32_t)v;       /* a 64-bit JSON integer wrapped into an int, like the reference (json_get_int) */
    if (json_get_int(cfg, "decimationFactor", &v)) { B200_MSG("I", "NO-DECIMATION", "Not decimating the output signal: using full bandwidth."); return A_E_INVAL; }
    if (v <= 0) { B200_MSG("E", "BAD-DECIMATION-FACTOR", "Decimation factor of '%d' is not valid.", v); return A_E_INVAL; }
    rx->decimation = (uint32_t)v;

    const jnode *taps = json_get(cfg, "lpfTaps");
    if (!taps || taps->type != J_ARR) { B200_MSG("E", "BAD-FILTER-TAPS", "Need to provide a baseband filter with at least two filter taps as 'lpfTaps'."); return A_E_INVAL; }
    if (taps->len <= 1) { B200_MSG("E", "INSUFF-FILTER-TAPS", "Not enough filter taps for the low-pass filter."); return A_E_INVAL; }
    rx->nr_lpf_taps = taps->len;
    rx->lpf_taps = calloc(taps->len, sizeof(double));
    for (size_t i = 0; i < taps->len; i++) {
        if (taps->items[i]->type != J_NUM) { B200_MSG("E", "BAD-FILTER-TAPS", "lpfTaps[%zu] is not a number", i); return A_E_INVAL; }
        rx->lpf_taps[i] = taps->items[i]->num;
    }

    const jnode *chans = json_get(cfg, "channels");
    if (!chans || chans->type != J_ARR || chans->len == 0) { B200_MSG("E", "MISSING-CHANNELS", "Need to specify at least one channel to demodulate."); return A_E_INVAL; }
    rx->nr_demod_threads = chans->len;
    rx->channels = calloc(chans->len, sizeof(*rx->channels));
    bool any_debug = false;
    for (size_t i = 0; i < chans->len; i++) {
        const jnode *c = chans->items[i];
        struct receiver_channel *ch = &rx->channels[i];
        const char *s = NULL;
        ch->fifo_fd = ch->debug_fd = -1;
        if (json_get_string(c, "outFifo", &s)) { B200_MSG("E", "MISSING-FIFO-ID", "Missing output FIFO filename, aborting."); return A_E_INVAL; }
        ch->out_fifo = strdup(s);
        if (json_get_int(c, "chanCenterFreq", &ch->center_freq_hz)) { B200_MSG("E", "MISSING-CENTER-FREQ", "Missing output channel center frequency."); return A_E_INVAL; }
        if (!json_get_string(c, "signalDebugFile", &s)) { ch->signal_debug = strdup(s); any_debug = true; }
        ch->gain = 1.0;
        if (!json_get_double(c, "dBGain", &ch->gain_db)) ch->gain = gpuchan_db_to_gain(ch->gain_db);   /* key is case-sensitive */
        B200_MSG("I", "CHANNEL", "[%zu]: %4.5f MHz Gain: %f dB -> [%s]", i + 1, (double)ch->center_freq_hz / 1e6, ch->gain_db, ch->out_fifo);
    }

    /* our extensions (all optional) */
    rx->batch_bufs = 64;
    rx->gpu_devices[0] = 0; rx->nr_gpu_devices = 1; rx->gpu_fanout = GPUCHAN_FANOUT_HOST;
    if (!json_get_int(cfg, "gpuDevice", &v)) rx->gpu_devices[0] = v;
    const jnode *gd = json_get(cfg, "gpuDevices");          /* channels shard over these devices (contiguous ranges) */
    if (gd && gd->type == J_ARR && gd->len > 0) {
        if (gd->len > 16) { B200_MSG("E", "BAD-GPU-DEVICES", "at most 16 gpuDevices"); return A_E_INVAL; }
        rx->nr_gpu_devices = (uint32_t)gd->len;
        for (size_t i = 0; i < gd->len; i++) {
            if (gd->items[i]->type != J_NUM || !gd->items[i]->is_int) { B200_MSG("E", "BAD-GPU-DEVICES", "gpuDevices[%zu] is not an integer", i); return A_E_INVAL; }
            rx->gpu_devices[i] = (int)gd->items[i]->inum;
        }
    }
    const char *fo = NULL;
    if (!json_get_string(cfg, "gpuFanout", &fo)) {
        if (!strcmp(fo, "relay")) rx->gpu_fanout = GPUCHAN_FANOUT_RELAY;
        else if (strcmp(fo, "host")) { B200_MSG("E", "BAD-GPU-FANOUT", "gpuFanout must be \"host\" or \"relay\""); return A_E_INVAL; }
    }
    if (!json_get_int(cfg, "gpuBatchBuffers", &v) && v > 0) rx->batch_bufs = (size_t)v;

    /* pool */
    rx->pool = calloc((size_t)rx->nr_samp_bufs, sizeof(*rx->pool));
    for (int i = 0; i < rx->nr_samp_bufs; i++) {
        struct sample_buf *b = NULL;
        if (posix_memalign((void **)&b, 64, sizeof(*b) + samples_per_buf * 4)) return A_E_NOMEM;
        memset(b, 0, sizeof(*b));
        b->sample_type = COMPLEX_INT_16;
        b->sample_buf_bytes = (uint32_t)(samples_per_buf * 4);
        rx->pool[rx->pool_free++] = b;
    }

    /* the channel bank */
    const size_t C = rx->nr_demod_threads;
    int32_t *offs = calloc(C, sizeof(int32_t));
    double *gains = calloc(C, sizeof(double));
    for (size_t i = 0; i < C; i++) {
        /* receiver.c:229: (int32_t)nb_center_freq - center_freq on wrapped ints -- correct modulo 2^32 even above 2^31 Hz */
        offs[i] = (int32_t)((uint32_t)rx->channels[i].center_freq_hz - rx->center_freq_hz);
        gains[i] = rx->channels[i].gain;
    }
    gpuchan_cfg gc;
    memset(&gc, 0, sizeof(gc));
    gc.struct_size = sizeof(gc);
    gc.sample_rate_hz = rx->sample_rate_hz; gc.decimation = rx->decimation;
    gc.nr_taps = (uint32_t)rx->nr_lpf_taps; gc.nr_channels = (uint32_t)C;
    gc.max_batch_samples = (uint32_t)(rx->batch_bufs * samples_per_buf);
    gc.flags = GPUCHAN_F_DEFAULT | (any_debug ? GPUCHAN_F_KEEP_IQ : 0);
    gc.lpf_taps = rx->lpf_taps; gc.offset_hz = offs; gc.gain = gains;
    gpuchan_multi_t *bank = NULL;
    int rc = gpuchan_multi_create(&bank, &gc, rx->gpu_devices, rx->nr_gpu_devices, rx->gpu_fanout);
    free(offs); free(gains);
    if (rc) { B200_MSG("E", "GPU-BANK", "gpuchan_multi_create failed: %s", gpuchan_last_error()); return A_E_INVAL; }
    rx->bank = bank;
    if (gpuchan_multi_devices(bank) > 1)
        B200_MSG("I", "GPU-DEVICES", "%zu channels sharded over %u GPUs, IQ fan-out: %s", C, gpuchan_multi_devices(bank),
                 rx->gpu_fanout == GPUCHAN_FANOUT_RELAY ? "NVLink relay chain" : "one PCIe copy per GPU");
    rx->pcm_cap = gc.max_batch_samples / rx->decimation + 64;
    for (int i = 0; i < 2; i++)
        if (gpuchan_host_alloc((void **)&rx->batch[i], (size_t)gc.max_batch_samples * 4)) return A_E_NOMEM;
    if (gpuchan_host_alloc((void **)&rx->pcm_host, C * rx->pcm_cap * sizeof(int16_t))) return A_E_NOMEM;
    if (any_debug && gpuchan_host_alloc((void **)&rx->iq_host, C * rx->pcm_cap * 4)) return A_E_NOMEM;

    /* optional in-process decoders.  The reference runs one `decoder -m <proto> -I i -D d -F taps [-b] [-i]` process per
     * channel FIFO (decoder/decoder.c:685-697); here pagerDecode is one such specification for all channels, or an array
     * of them, each naming its channels ("channels": [indices into the channels array]) -- a mixed POCSAG / FLEX
     * receiver.  Every specification becomes one pager bank per device that owns some of its channels. */
    const jnode *pd = json_get(cfg, "pagerDecode");
    rx->msg_out = stdout;
    const size_t nr_specs = !pd ? 0 : (pd->type == J_ARR ? pd->len : (pd->type == J_OBJ ? 1 : 0));
    if (nr_specs) rx->pagers = calloc(nr_specs * gpuchan_multi_devices(bank), sizeof(*rx->pagers));
    for (size_t sp = 0; sp < nr_specs; sp++) {
        const jnode *spec = pd->type == J_ARR ? pd->items[sp] : pd;
        if (!spec || spec->type != J_OBJ) { B200_MSG("E", "PAGER-SPEC", "pagerDecode[%zu] is not an object", sp); return A_E_INVAL; }
        int I = 1, Dd = 1;
        const char *s = NULL;
        json_get_int(spec, "interpolate", &I); json_get_int(spec, "decimate", &Dd);
        const jnode *co = json_get(spec, "lpfCoeffs");
        if (!co || co->type != J_ARR || co->len == 0) { B200_MSG("E", "PAGER-TAPS", "pagerDecode.lpfCoeffs missing"); return A_E_INVAL; }
        double *cf = calloc(co->len, sizeof(double));
        int16_t *q = calloc(co->len, sizeof(int16_t));
        for (size_t i = 0; i < co->len; i++) cf[i] = co->items[i]->num;
        gpupager_quantize_taps(cf, co->len, q);
        free(cf);
        /* which channels this specification decodes */
        bool *sel = calloc(C, sizeof(bool));
        const jnode *chs = json_get(spec, "channels");
        if (chs && chs->type == J_ARR) {
            for (size_t i = 0; i < chs->len; i++) {
                const jnode *e = chs->items[i];
                if (e->type != J_NUM || !e->is_int || e->inum < 0 || (size_t)e->inum >= C) {
                    B200_MSG("E", "PAGER-CHANNELS", "pagerDecode channels[%zu] is not a channel index", i);
                    free(sel); free(q);
                    return A_E_INVAL;
                }
                sel[e->inum] = true;
            }
        } else {
            for (size_t i = 0; i < C; i++) sel[i] = true;
        }
        const char *proto = NULL;                           /* decoder -m POCSAG | FLEX */
        const bool is_flex = !json_get_string(spec, "protocol", &proto) && !strncasecmp(proto, "flex", 4);
        for (uint32_t d = 0; d < gpuchan_multi_devices(bank); d++) {
            uint32_t first = 0, cnt = 0, nsel = 0;
            gpuchan_multi_bank(bank, d, NULL, &first, &cnt);
            uint32_t *map = calloc(cnt ? cnt : 1, sizeof(uint32_t));
            for (uint32_t r = 0; r < cnt; r++) if (sel[first + r]) map[nsel++] = r;
            if (nsel) {
                gpupager_cfg pc;
                memset(&pc, 0, sizeof(pc));
                pc.struct_size = sizeof(pc); pc.nr_channels = nsel; pc.device = rx->gpu_devices[d];
                pc.interpolate = (uint32_t)I; pc.decimate = (uint32_t)Dd; pc.nr_taps = (uint32_t)co->len;
                pc.max_feed_samples = (uint32_t)rx->pcm_cap; pc.taps = q;
                pc.channel_map = map;
                double pole = 0.0;
                if (!json_get_double(spec, "dcBlockPole", &pole)) { pc.flags |= GPUPAGER_F_DC_BLOCK; pc.dc_pole = pole; }
                if (is_flex) pc.decoder = GPUPAGER_DECODER_FLEX;
                int inv = 0;                                /* decoder -i */
                if (!json_get_int(spec, "invert", &inv) && inv) pc.flags |= GPUPAGER_F_INVERT;
                struct receiver_pager *pg = &rx->pagers[rx->nr_pagers];
                rc = gpupager_create(&pg->bank, &pc);
                if (rc) { B200_MSG("E", "GPU-PAGER", "gpupager_create failed: %s", gpupager_last_error()); free(map); free(sel); free(q); return A_E_INVAL; }
                pg->rx = rx; pg->is_flex = is_flex; pg->device_index = d; pg->channel_base = first;
                rx->nr_pagers++;
            }
            free(map);
        }
        free(sel); free(q);
        if (!json_get_string(spec, "outFile", &s) && rx->msg_out == stdout && !rx->msg_out_ch) {
            if (strstr(s, "%u")) {                          /* one file per channel, decoder.c's lines byte for byte */
                rx->msg_out_ch = calloc(C, sizeof(FILE *));
                for (size_t i = 0; i < C; i++) {
                    char path[4096];
                    snprintf(path, sizeof(path), s, (unsigned)i);
                    rx->msg_out_ch[i] = fopen(path, "w");
                    if (!rx->msg_out_ch[i]) { B200_MSG("E", "PAGER-OUT", "cannot open %s", path); return A_E_INVAL; }
                }
            } else {
                rx->msg_out = fopen(s, "w");
                if (!rx->msg_out) { B200_MSG("E", "PAGER-OUT", "cannot open %s", s); return A_E_INVAL; }
            }
        }
    }

    /* open outputs last: a FIFO open blocks until a reader appears.  multifm/demod.c:323,331: O_WRONLY only -- the
     * FIFO (or file) must exist, a mistyped path is an error, nothing is created.  A regular file is emptied first so
     * that a rerun never leaves stale PCM behind the new data. */
    signal(SIGPIPE, SIG_IGN);
    for (size_t i = 0; i < C; i++) {
        struct receiver_channel *ch = &rx->channels[i];
        struct stat sb;
        if (ch->signal_debug && *ch->signal_debug) {
            ch->debug_fd = open(ch->signal_debug, O_WRONLY);
            if (ch->debug_fd < 0) { B200_MSG("F", "CANT-OPEN-SIGNAL-DEBUG", "Unable to open signal debug dump file '%s'", ch->signal_debug); return A_E_INVAL; }
            if (0 == fstat(ch->debug_fd, &sb) && S_ISREG(sb.st_mode) && ftruncate(ch->debug_fd, 0)) { /* keep going */ }
        }
        if (strcmp(ch->out_fifo, "/dev/null") == 0 && rx->nr_pagers) continue;  /* decode-only channel */
        ch->fifo_fd = open(ch->out_fifo, O_WRONLY);
        if (ch->fifo_fd < 0) { B200_MSG("F", "CANT-OPEN-FIFO", "Unable to open output fifo '%s'", ch->out_fifo); return A_E_INVAL; }
        if (0 == fstat(ch->fifo_fd, &sb) && S_ISREG(sb.st_mode) && ftruncate(ch->fifo_fd, 0)) { /* keep going */ }
    }
    return A_OK;
}

static void *rx_main(void *arg)
{
    struct receiver *rx = arg;
    rx->thread_func(rx);
    receiver_end_of_stream(rx);
    return NULL;
}

aresult_t receiver_start(struct receiver *rx)
{
    if (!rx) return A_E_BADARGS;
    rx->running = true;
    if (pthread_create(&rx->consumer_thread, NULL, consumer_main, rx)) return A_E_INVAL;
    if (pthread_create(&rx->rx_thread, NULL, rx_main, rx)) return A_E_INVAL;
    return A_OK;
}

aresult_t receiver_drain(struct receiver *rx)
{
    if (!rx) return A_E_BADARGS;
    pthread_join(rx->rx_thread, NULL);
    pthread_join(rx->consumer_thread, NULL);
    rx->running = false;
    return A_OK;
}

aresult_t receiver_cleanup(struct receiver **prx)
{
    if (!prx || !*prx) return A_E_BADARGS;
    struct receiver *rx = *prx;
    if (rx->running) {
        rx->running = false;
        receiver_end_of_stream(rx);
        pthread_join(rx->rx_thread, NULL);
        pthread_join(rx->consumer_thread, NULL);
    }
    for (size_t i = 0; i < rx->nr_demod_threads; i++) {
        struct receiver_channel *ch = &rx->channels[i];
        if (ch->fifo_fd >= 0) close(ch->fifo_fd);
        if (ch->debug_fd >= 0) close(ch->debug_fd);
        free(ch->out_fifo); free(ch->signal_debug);
    }
    for (size_t p = 0; p < rx->nr_pagers; p++) gpupager_destroy(&rx->pagers[p].bank);
    free(rx->pagers);
    if (rx->bank) gpuchan_multi_destroy(&rx->bank);
    gpuchan_host_free(rx->batch[0]); gpuchan_host_free(rx->batch[1]);
    gpuchan_host_free(rx->pcm_host); gpuchan_host_free(rx->iq_host);
    if (rx->msg_out && rx->msg_out != stdout) fclose(rx->msg_out);
    if (rx->msg_out_ch) {
        for (size_t i = 0; i < rx->nr_demod_threads; i++) if (rx->msg_out_ch[i]) fclose(rx->msg_out_ch[i]);
        free(rx->msg_out_ch);
    }
    for (int i = 0; i < rx->nr_samp_bufs; i++) free(rx->pool[i]);   /* all buffers are back in the pool by now */
    free(rx->pool); free(rx->channels); free(rx->lpf_taps);
    pthread_mutex_destroy(&rx->pool_mtx); pthread_mutex_destroy(&rx->q_mtx); pthread_cond_destroy(&rx->q_cv);
    /* the source goes last: its cleanup releases the object that embeds this receiver (file_worker_thread) */
    if (rx->cleanup_func) rx->cleanup_func(rx);
    *prx = NULL;
    return A_OK;
}
