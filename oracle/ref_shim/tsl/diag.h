#pragma once
#include <stdio.h>
#define SEV_INFO    "I"
#define SEV_WARNING "W"
#define SEV_ERROR   "E"
#define SEV_FATAL   "F"
#define DIAG(...) do { } while (0)
#define MESSAGE(subsys, sev, ident, msg, ...) \
    fprintf(stderr, "%s:%s:%s: " msg "\n", subsys, sev, ident, ##__VA_ARGS__)
