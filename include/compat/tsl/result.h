/* Minimal stand-in for the TSL library's <tsl/result.h> (github.com/pvachon/tsl, not vendored by the reference): just
 * what the compat headers in this directory need.  A build that has the real TSL puts it ahead of this directory on the
 * include path; the values below match oracle/ref_shim/tsl/result.h, which the reference's own objects compile against. */
#ifndef TSLB200_COMPAT_TSL_RESULT_H
#define TSLB200_COMPAT_TSL_RESULT_H
#include <stddef.h>
#include <stdint.h>

typedef int aresult_t;
#ifndef A_OK
#define A_OK          0
#define A_E_NOMEM    (-1)
#define A_E_BADARGS  (-2)
#define A_E_NOTFOUND (-3)
#define A_E_BUSY     (-4)
#define A_E_INVAL    (-5)
#define A_E_EMPTY    (-8)
#define A_E_DONE     (-12)
#define FAILED(x)    ((x) != A_OK)
#endif
#endif
