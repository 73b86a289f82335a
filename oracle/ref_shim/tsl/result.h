/* TEST INFRASTRUCTURE ONLY. Minimal stand-in for the un-vendored TSL library
 * (github.com/pvachon/tsl) so that the reference's numeric .c files compile
 * unmodified from /root/reference into oracle/_ref/. No arithmetic lives here. */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
typedef int aresult_t;
#define A_OK            0
#define A_E_NOMEM      (-1)
#define A_E_BADARGS    (-2)
#define A_E_NOTFOUND   (-3)
#define A_E_BUSY       (-4)
#define A_E_INVAL      (-5)
#define A_E_EMPTY      (-8)
#define A_E_DONE       (-12)
#define FAILED(x)            ((x) != A_OK)
#define FAILED_UNLIKELY(x)   __builtin_expect(((x) != A_OK), 0)
