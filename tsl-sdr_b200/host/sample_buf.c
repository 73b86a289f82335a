#include "sample_buf.h"
#include <stdatomic.h>
#include <stdlib.h>

/* filter/sample_buf.c:31-43: atomic decrement; the last reference releases the buffer */
aresult_t sample_buf_decref(struct sample_buf *buf)
{
    if (NULL == buf) return A_E_BADARGS;
    if (1 == atomic_fetch_sub((_Atomic uint32_t *)&buf->refcount, 1)) {
        if (NULL == buf->release) abort();
        return buf->release(buf);
    }
    return A_OK;
}
