/* file_if.c -- replay a capture file as the receiver's IQ source.
 * Mirrors multifm/file_if.c: 4096-sample sample_bufs (:18), formats cs16 (:47-64, straight read),
 * cs8 (:67-110, plain widening cast) and cu8 (:112-157: the reference reads the bytes through an int8_t
 * pointer before subtracting 127 -- reproduced as is). */
#define _GNU_SOURCE
#include "file_if.h"

#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define SAMPLES_PER_BUF 4096

enum file_fmt { FMT_CS16, FMT_CS8, FMT_CU8 };

struct file_worker_thread {
    struct receiver rx;     /* must be first: the receiver is the base "class" */
    int fd;
    enum file_fmt fmt;
    int8_t *bounce;
};

static ssize_t read_full(int fd, void *dst, size_t bytes)
{
    size_t got = 0;
    while (got < bytes) {
        ssize_t r = read(fd, (uint8_t *)dst + got, bytes - got);
        if (r < 0) { if (errno == EINTR) continue; return -1; }
        if (r == 0) break;
        got += (size_t)r;
    }
    return (ssize_t)got;
}

static aresult_t file_worker_thread_work(struct receiver *rx)
{
    struct file_worker_thread *thr = (struct file_worker_thread *)rx;
    while (receiver_thread_running(rx)) {
        struct sample_buf *sbuf = NULL;
        if (FAILED(receiver_sample_buf_alloc(rx, &sbuf))) {
            usleep(2000);                   /* pool empty: the reference sleeps 500 ms and drops (file_if.c:176-179) */
            continue;
        }
        int16_t *out = (int16_t *)sbuf->data_buf;
        ssize_t nr = 0;
        if (thr->fmt == FMT_CS16) {
            nr = read_full(thr->fd, out, SAMPLES_PER_BUF * 4);
            if (nr > 0) sbuf->nr_samples = (uint32_t)(nr / 4);
        } else {
            nr = read_full(thr->fd, thr->bounce, SAMPLES_PER_BUF * 2);
            for (ssize_t i = 0; i < nr; i++)
                out[i] = (thr->fmt == FMT_CS8) ? (int16_t)thr->bounce[i] : (int16_t)((int16_t)thr->bounce[i] - 127);
            /* file_if.c:147-151: a read that is not a multiple of 4 bytes (only at the end of a file) leaves its last
             * nr % 4 cu8 values without the -127 offset */
            if (thr->fmt == FMT_CU8)
                for (ssize_t i = nr - nr % 4; i < nr; i++) out[i] = (int16_t)thr->bounce[i];
            if (nr > 0) sbuf->nr_samples = (uint32_t)(nr / 2);
        }
        if (nr <= 0 || sbuf->nr_samples == 0) {     /* EOF: the reference aborts here (receiver.c:84) */
            sample_buf_decref(sbuf);                /* never delivered: our reference (set by alloc) returns it to the pool */
            break;
        }
        receiver_sample_buf_deliver(rx, sbuf);
    }
    return A_OK;
}

static aresult_t file_cleanup(struct receiver *rx)
{
    struct file_worker_thread *thr = (struct file_worker_thread *)rx;
    if (thr->fd >= 0) close(thr->fd);
    free(thr->bounce);
    free(thr);              /* receiver_cleanup calls this last: the receiver lives inside thr */
    return A_OK;
}

aresult_t file_worker_thread_new(struct receiver **pthr, const jnode *cfg)
{
    if (!pthr || !cfg) return A_E_BADARGS;
    *pthr = NULL;
    const jnode *dev = json_get(cfg, "device");
    const char *fname = NULL, *fmt = NULL;
    if (!dev || json_get_string(dev, "filename", &fname)) { B200_MSG("E", "MISSING-FILENAME", "device.filename is required"); return A_E_INVAL; }
    if (json_get_string(dev, "fileFormat", &fmt)) fmt = "cs16";
    struct file_worker_thread *thr = calloc(1, sizeof(*thr));
    if (!thr) return A_E_NOMEM;
    if (!strcmp(fmt, "cs16")) thr->fmt = FMT_CS16;
    else if (!strcmp(fmt, "cs8")) thr->fmt = FMT_CS8;
    else if (!strcmp(fmt, "cu8")) thr->fmt = FMT_CU8;
    else { B200_MSG("E", "BAD-FILE-FORMAT", "unknown fileFormat '%s'", fmt); free(thr); return A_E_INVAL; }
    thr->fd = open(fname, O_RDONLY);
    if (thr->fd < 0) { B200_MSG("E", "BAD-FILE", "cannot open %s: %s", fname, strerror(errno)); free(thr); return A_E_INVAL; }
    thr->bounce = malloc(SAMPLES_PER_BUF * 2);
    if (!thr->bounce) { close(thr->fd); free(thr); return A_E_NOMEM; }
    aresult_t ret = receiver_init(&thr->rx, cfg, file_worker_thread_work, file_cleanup, SAMPLES_PER_BUF);
    if (FAILED(ret)) { close(thr->fd); free(thr->bounce); free(thr); return ret; }
    *pthr = &thr->rx;
    return A_OK;
}
