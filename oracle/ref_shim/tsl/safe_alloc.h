#pragma once
#include <stdlib.h>
#include <string.h>
#include <tsl/result.h>
#include <tsl/cal.h>
/* all allocations are free()-compatible (the reference mixes TFREE and free) */
static inline aresult_t __shim_aalloc(void **pp, size_t bytes, size_t align)
{
    void *p = NULL;
    if (align < sizeof(void *)) align = sizeof(void *);
    if (bytes == 0) bytes = align;
    if (posix_memalign(&p, align, bytes)) { *pp = NULL; return A_E_NOMEM; }
    memset(p, 0, bytes);
    *pp = p;
    return A_OK;
}
#define TACALLOC(pp, n, sz, align)  __shim_aalloc((void **)(pp), (size_t)(n) * (size_t)(sz), (align))
#define TCALLOC(pp, n, sz)          __shim_aalloc((void **)(pp), (size_t)(n) * (size_t)(sz), 16)
#define TZAALLOC(p, align)          __shim_aalloc((void **)&(p), sizeof(*(p)), (align))
#define TZALLOC(p)                  __shim_aalloc((void **)&(p), sizeof(*(p)), 16)
#define TFREE(p)                    do { free(p); (p) = NULL; } while (0)
