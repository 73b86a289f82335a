/*
 * tslb200_gpupager.h -- C ABI of the B200 pager bank: rational resampler + POCSAG decode for every channel
 * of a channel bank, fed with that bank's int16 PCM.
 *
 * It replaces one `decoder` process per channel (decoder/decoder.c) reading a channel's FIFO:
 *
 *   reference seam                                          replaced by
 *   ------------------------------------------------------  ------------------------------------
 *   decoder/decoder.c:527-533  taps = (int16)(coef*16384)    gpupager_quantize_taps
 *   decoder/decoder.c:685      polyphase_fir_new(I, D)       gpupager_create
 *   decoder/decoder.c:693      pager_pocsag_new(...)         gpupager_create
 *   decoder/decoder.c:596-632  read(fifo) -> 1024-sample buf gpupager_feed / gpupager_feed_device
 *   filter/polyphase_fir.c:162 polyphase_fir_process         resample kernel
 *   filter/dc_blocker.h:72     dc_blocker_apply (-b)         GPUPAGER_F_DC_BLOCK (inside the decode kernel)
 *   pager/pager_pocsag.c:434   pager_pocsag_on_pcm           pocsag kernel (eye sync, slicer, BCH, message assembly)
 *   pager/pager_pocsag.h:8-22  on_numeric / on_alpha         gpupager_dispatch (same argument meaning, fired on the host)
 *   decoder/decoder.c:689      pager_flex_new(...)           gpupager_create with decoder = GPUPAGER_DECODER_FLEX
 *   pager/pager_flex.c:1401    pager_flex_on_pcm             flex kernel (Sync 1 / Sync 2 / block / BCH / vectors)
 *   pager/pager_flex.h:16-83   on_alnum / on_num / on_siv    gpupager_dispatch_flex (same argument meaning)
 *   decoder/decoder.c:621-626  -i (invert the input)         GPUPAGER_F_INVERT
 *
 * Stream contract: resampled sample m of a channel exists as soon as input samples up to n_m + M are
 * available (strict, polyphase_fir.c:184); the decoder consumes every resampled sample in order. Feeds
 * may have any length. aresult_t-compatible returns (0 == A_OK).
 */
#ifndef TSLB200_GPUPAGER_H
#define TSLB200_GPUPAGER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPUPAGER_OK           0
#define GPUPAGER_E_NOMEM    (-1)
#define GPUPAGER_E_BADARGS  (-2)
#define GPUPAGER_E_INVAL    (-5)
#define GPUPAGER_E_CUDA     (-64)
#define GPUPAGER_E_NODEVICE (-65)

#define GPUPAGER_F_DC_BLOCK   0x1u   /* decoder -b (filter/dc_blocker.h), pole from cfg.dc_pole */
#define GPUPAGER_F_KEEP_PCM   0x2u   /* keep the resampled PCM of the last feed (decoder -d tap) */
#define GPUPAGER_F_NO_RESAMPLE 0x4u  /* input is already at the decoder's rate (38400 / 16000 Hz): bypass the resampler */
#define GPUPAGER_F_INVERT     0x8u   /* decoder -i: negate every input sample (int16 wrap) before the resampler */

#define GPUPAGER_DECODER_POCSAG 0    /* decoder -m POCSAG (default) */
#define GPUPAGER_DECODER_FLEX   1    /* decoder -m FLEX */

#define GPUPAGER_MSG_NUMERIC    0    /* POCSAG numeric */
#define GPUPAGER_MSG_ALPHA      1    /* POCSAG alphanumeric */
#define GPUPAGER_MSG_FLEX_ALNUM 2    /* FLEX alphanumeric: function = phase, aux = {cycle, frame, fragmented, maildrop, seq} */
#define GPUPAGER_MSG_FLEX_NUM   3    /* FLEX numeric / tone-with-digits: function = phase, aux = {cycle, frame} */
#define GPUPAGER_MSG_FLEX_SIV   4    /* FLEX short instruction vector: aux = {cycle, frame, siv type, siv data} */
#define GPUPAGER_MSG_TEXT_MAX 512

typedef struct gpupager gpupager_t;

typedef struct gpupager_cfg {
    uint32_t struct_size;
    uint32_t nr_channels;
    int32_t  device;
    uint32_t interpolate;         /* decoder -I */
    uint32_t decimate;            /* decoder -D */
    uint32_t nr_taps;             /* resampler prototype length (lpfCoeffs) */
    uint32_t max_feed_samples;    /* largest per-channel sample count one feed may carry */
    uint32_t flags;
    double   dc_pole;             /* decoder -p (default 0.9999) */
    const int16_t *taps;          /* [nr_taps] Q.14 int16 taps, see gpupager_quantize_taps */
    uint32_t decoder;             /* GPUPAGER_DECODER_* */
    uint32_t reserved;
    /* Optional channel selection: decoder channel c reads row channel_map[c] of the PCM array it is fed (what
     * gpuchan_device_pcm returns) and reports that index in its messages.  The reference starts one `decoder` process
     * per channel and picks the protocol per process (decoder/decoder.c:688-697): a receiver carrying POCSAG and FLEX
     * channels side by side is two banks with different maps over the same channel-bank PCM.  NULL = rows 0..n-1. */
    const uint32_t *channel_map;  /* [nr_channels] */
} gpupager_cfg;

typedef struct gpupager_msg {
    uint32_t channel;
    uint32_t kind;                /* GPUPAGER_MSG_* */
    uint32_t baud;                /* POCSAG 512 / 1200 / 2400; FLEX 1600 / 3200 / 6400 */
    uint32_t capcode;             /* as the reference reports it (pager_pocsag.c:362, pager_flex.c:555,567); low 32 bits */
    uint32_t function;            /* POCSAG function bits; FLEX phase (0 = A .. 3 = D) */
    uint32_t len;                 /* bytes in text */
    uint32_t capcode_hi;          /* FLEX capcodes are uint64_t in the reference's callbacks */
    uint32_t aux[6];              /* see GPUPAGER_MSG_FLEX_* */
    char     text[GPUPAGER_MSG_TEXT_MAX];
} gpupager_msg;

/* Same argument meaning as pager/pager_pocsag.h:8-22, plus the channel index and a user pointer. */
typedef int (*gpupager_on_msg_func_t)(void *user, uint32_t channel, uint16_t baud_rate, uint32_t capcode,
                                      const char *data, size_t data_len, uint8_t function);

/* decoder/decoder.c:530-533: (int16_t)(coef * (1 << 14)) */
/* Same argument meaning as pager/pager_flex.h:16-83, plus the channel index and a user pointer. */
typedef int (*gpupager_on_flex_alnum_func_t)(void *user, uint32_t channel, uint16_t baud, uint8_t phase, uint8_t cycle_no,
                                             uint8_t frame_no, uint64_t cap_code, int fragmented, int maildrop,
                                             uint8_t seq_num, const char *message_bytes, size_t message_len);
typedef int (*gpupager_on_flex_num_func_t)(void *user, uint32_t channel, uint16_t baud, uint8_t phase, uint8_t cycle_no,
                                           uint8_t frame_no, uint64_t cap_code, const char *message_bytes, size_t message_len);
typedef int (*gpupager_on_flex_siv_func_t)(void *user, uint32_t channel, uint16_t baud, uint8_t phase, uint8_t cycle_no,
                                           uint8_t frame_no, uint64_t cap_code, uint8_t siv_msg_type, uint32_t data);

int gpupager_quantize_taps(const double *coeffs, size_t nr, int16_t *taps_q14);

int gpupager_create(gpupager_t **ph, const gpupager_cfg *cfg);
int gpupager_destroy(gpupager_t **ph);

/* n samples per channel from device memory: channel c at d_pcm + c*pitch_samples (what
 * gpuchan_device_pcm returns). Enqueued on cuda_stream (NULL = own stream, ordered after it otherwise). */
int gpupager_feed_device(gpupager_t *h, const int16_t *d_pcm, size_t pitch_samples, size_t n, void *cuda_stream);
/* Same from host memory ([nr_channels][pitch_samples] int16). */
int gpupager_feed(gpupager_t *h, const int16_t *pcm_host, size_t pitch_samples, size_t n);

/* Wait for the kernels, then hand every message decoded since the last call to the callbacks, channel by
 * channel in decode order.  Either callback may be NULL.  Returns the number of messages in *nr_msgs. */
int gpupager_dispatch(gpupager_t *h, gpupager_on_msg_func_t on_numeric, gpupager_on_msg_func_t on_alpha, void *user,
                      size_t *nr_msgs);
/* FLEX banks: the three callbacks of pager/pager_flex.h.  Any of them may be NULL. */
int gpupager_dispatch_flex(gpupager_t *h, gpupager_on_flex_alnum_func_t on_alnum, gpupager_on_flex_num_func_t on_num,
                           gpupager_on_flex_siv_func_t on_siv, void *user, size_t *nr_msgs);
/* Same, copying the records instead (at most cap; the rest stay queued). */
int gpupager_poll(gpupager_t *h, gpupager_msg *out, size_t cap, size_t *nr_msgs);

/* decoder -d tap: resampled PCM of the last feed, channel c at out[c*cap ...]; needs GPUPAGER_F_KEEP_PCM. */
int gpupager_collect_pcm(gpupager_t *h, int16_t *out, size_t cap_per_channel, size_t *n_per_channel);

/* ---- Mueller-Muller soft-decision timing recovery (pager/mueller_muller.c:10-115) for every channel ----
 * The reference ships this block but wires it into no pipeline (mm_init/mm_process have no callers); it is offered
 * here as an optional pre-slicer with the same parameters, state and outputs.  One state per channel; every call
 * consumes n samples per channel (channel c at pcm_host + c*pitch) and appends the decisions (the samples picked
 * at the recovered symbol instants) to decisions_host + c*cap, reporting their number in nr_out[c].
 * GPUMM_F_FMA selects the contraction GNU C applies to mueller_muller.c:79,92 on FMA machines (the reference's
 * Release build); without it every multiply and add rounds separately. */
#define GPUMM_F_FMA 0x1u
typedef struct gpumm gpumm_t;
int gpumm_create(gpumm_t **ph, uint32_t nr_channels, int32_t device, float kw, float km, float samples_per_bit,
                 float error_min, float error_max, uint32_t max_feed_samples, uint32_t flags);
int gpumm_process(gpumm_t *h, const int16_t *pcm_host, size_t pitch_samples, size_t n, int16_t *decisions_host,
                  size_t cap_per_channel, uint32_t *nr_out);
/* mm_process state of one channel after the last call: {w, m, next_offset, last_sample} */
int gpumm_get_state(gpumm_t *h, uint32_t channel, float state[4]);
int gpumm_destroy(gpumm_t **ph);

uint64_t gpupager_kernel_launches(gpupager_t *h);
uint64_t gpupager_dropped_msgs(gpupager_t *h);
const char *gpupager_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
