#pragma once
#include <tsl/result.h>
#include <tsl/panic.h>
#include <tsl/cal.h>
#define TSL_ASSERT_ARG(x)          do { if (!(x)) return A_E_BADARGS; } while (0)
#define TSL_ASSERT_ARG_DEBUG(x)    do { } while (0)
#define TSL_ASSERT_PTR_BY_REF(x)   do { if (!(x) || !*(x)) return A_E_BADARGS; } while (0)
#define TSL_BUG_ON(x)              do { if (__builtin_expect(!!(x), 0)) PANIC("BUG: " #x); } while (0)
#define TSL_BUG_IF_FAILED(x)       do { if (__builtin_expect(FAILED(x), 0)) PANIC("BUG (failed): " #x); } while (0)
