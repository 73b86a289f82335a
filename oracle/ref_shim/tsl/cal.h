#pragma once
#include <stddef.h>
#define CAL_ALIGN(n)        __attribute__((aligned(n)))
#define CAL_CACHE_ALIGNED   __attribute__((aligned(64)))
#define CAL_CLEANUP(fn)     __attribute__((cleanup(fn)))
#define CAL_UNUSED          __attribute__((unused))
#define CAL_PACKED          __attribute__((packed))
#define BL_MIN2(a, b)       (((a) < (b)) ? (a) : (b))
#define BL_MAX2(a, b)       (((a) > (b)) ? (a) : (b))
#define BL_ARRAY_ENTRIES(x) (sizeof(x) / sizeof((x)[0]))
#define BL_CONTAINER_OF(p, type, member) ((type *)((char *)(p) - offsetof(type, member)))
