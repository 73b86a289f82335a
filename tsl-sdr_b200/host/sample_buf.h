/* Sample buffer ABI shared by every source (file, and untouched rtl/airspy/uhd sources) and the consumer.
 * Mirrors the reference's filter/sample_buf.h:59-104 field for field so that a source written against the
 * reference keeps working: refcount is set by receiver_sample_buf_deliver, every consumer calls
 * sample_buf_decref exactly once, and the release callback returns the buffer to its pool. */
#ifndef B200_SAMPLE_BUF_H
#define B200_SAMPLE_BUF_H
#include "b200_result.h"

struct sample_buf;

enum sample_type {
    UNKNOWN = 0, REAL_UINT_16 = 1, COMPLEX_UINT_16 = 2, COMPLEX_INT_16 = 3, REAL_UINT_32 = 4, COMPLEX_UINT_32 = 5,
};

typedef aresult_t (*sample_buf_release_func_t)(struct sample_buf *buf);

struct sample_buf {
    uint32_t refcount __attribute__((aligned(16)));
    enum sample_type sample_type;
    uint32_t nr_samples;
    uint32_t sample_buf_bytes;
    uint64_t start_time_ns;
    sample_buf_release_func_t release;
    void *priv;
    uint8_t data_buf[];
};

aresult_t sample_buf_decref(struct sample_buf *buf);
#endif
