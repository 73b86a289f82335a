/* aresult_t-compatible result codes for the host side (the reference takes these from the un-vendored
 * TSL library: tsl/result.h).  0 == A_OK, negative == error, FAILED(x) == (x != 0). */
#ifndef B200_RESULT_H
#define B200_RESULT_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef int aresult_t;
#define A_OK          0
#define A_E_NOMEM    (-1)
#define A_E_BADARGS  (-2)
#define A_E_NOTFOUND (-3)
#define A_E_BUSY     (-4)
#define A_E_INVAL    (-5)
#define A_E_EMPTY    (-8)
#define A_E_DONE     (-12)
#define FAILED(x)    ((x) != A_OK)

#include <stdio.h>
#define B200_MSG(sev, ident, fmt, ...) fprintf(stderr, "multifm:%s:%s: " fmt "\n", sev, ident, ##__VA_ARGS__)
#endif
