/* Receiver base "class": the interface every IQ source uses (multifm/receiver.h:77-124 in the reference).
 * Same entry points, same argument meaning; what changed is behind receiver_sample_buf_deliver: instead of
 * fanning the buffer out to N demod threads (multifm/receiver.c:78-98) it goes to ONE consumer that batches
 * buffers into pinned memory and drives the GPU channel bank (include/tslb200_gpuchan.h). */
#ifndef B200_RECEIVER_H
#define B200_RECEIVER_H
#include "b200_result.h"
#include "json_min.h"
#include "sample_buf.h"

#include <pthread.h>

struct receiver;
typedef aresult_t (*receiver_cleanup_func_t)(struct receiver *rx);
typedef aresult_t (*receiver_rx_thread_func_t)(struct receiver *rx);

struct gpuchan_multi;
struct gpupager;

/* one in-process decoder bank: a protocol + resampler for a set of channels living on one device (the reference runs one
 * `decoder` process per channel and picks protocol and resampler per process, decoder/decoder.c:685-697) */
struct receiver_pager {
    struct gpupager *bank;
    struct receiver *rx;
    bool is_flex;
    uint32_t device_index;              /* which bank of the multi-device channel bank feeds it */
    uint32_t channel_base;              /* first channel of that bank: reported channel = base + row */
};

struct receiver_channel {
    char *out_fifo;             /* channels[].outFifo */
    char *signal_debug;         /* channels[].signalDebugFile */
    int center_freq_hz;         /* channels[].chanCenterFreq */
    double gain_db, gain;       /* channels[].dBGain -> 10^(dB/10) (multifm/receiver.c:220) */
    int fifo_fd, debug_fd;
    size_t nr_dropped_samples, total_nr_demod_samples;
};

struct receiver {
    bool muted;
    size_t nr_demod_threads;            /* == number of channels (kept under the reference's name) */
    size_t nr_samp_buf_alloc_fails;
    receiver_cleanup_func_t cleanup_func;
    receiver_rx_thread_func_t thread_func;

    /* configuration (multifm/receiver.c:133-218) */
    uint32_t sample_rate_hz, center_freq_hz, decimation;
    int nr_samp_bufs;
    double *lpf_taps;
    size_t nr_lpf_taps;
    struct receiver_channel *channels;
    size_t samples_per_buf;

    /* sample buffer pool (frame_alloc in the reference) */
    struct sample_buf **pool;
    size_t pool_free;
    pthread_mutex_t pool_mtx;

    /* producer -> consumer queue (work_queue depth 128 + mutex + condvar in the reference) */
    struct sample_buf *queue[128];
    size_t q_head, q_tail;
    pthread_mutex_t q_mtx;
    pthread_cond_t q_cv;

    /* GPU consumer */
    struct gpuchan_multi *bank;         /* all channels, sharded over gpuDevices (one device by default) */
    struct receiver_pager *pagers;      /* pagerDecode: one entry per (decoder spec, device) */
    size_t nr_pagers;
    int gpu_devices[16];
    uint32_t nr_gpu_devices;
    uint32_t gpu_fanout;                /* gpuFanout: "host" (default) | "relay" */
    size_t batch_bufs;                  /* sample_bufs per GPU submit (gpuBatchBuffers, default 64) */
    int16_t *batch[2];                  /* pinned staging, double buffered */
    size_t batch_fill;
    int batch_cur;
    int in_flight;
    int16_t *pcm_host;                  /* pinned: [channels][pcm_cap] */
    int16_t *iq_host;
    size_t pcm_cap;
    FILE *msg_out;                      /* pagerDecode.outFile or stdout */
    FILE **msg_out_ch;                  /* outFile with a %u: one file per channel, lines exactly as decoder.c prints them */
    size_t nr_messages;
    uint64_t total_iq_samples;

    pthread_t rx_thread, consumer_thread;
    volatile bool running, producer_done;
};

aresult_t receiver_init(struct receiver *rx, const jnode *cfg, receiver_rx_thread_func_t rx_func,
                        receiver_cleanup_func_t cleanup_func, size_t samples_per_buf);
aresult_t receiver_start(struct receiver *rx);
/* wait until the source signalled end of stream and everything queued has been processed */
aresult_t receiver_drain(struct receiver *rx);
aresult_t receiver_cleanup(struct receiver **prx);
aresult_t receiver_sample_buf_alloc(struct receiver *rx, struct sample_buf **pbuf);
aresult_t receiver_set_mute(struct receiver *rx, bool mute);
aresult_t receiver_sample_buf_deliver(struct receiver *rx, struct sample_buf *buf);
bool receiver_thread_running(struct receiver *rx);
/* sources call this at end of stream instead of delivering an empty buffer (which aborts in the reference,
 * multifm/receiver.c:84) */
void receiver_end_of_stream(struct receiver *rx);
#endif
