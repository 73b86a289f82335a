#pragma once
#include <stdio.h>
#include <stdlib.h>
#define PANIC(msg, ...) do { fprintf(stderr, "PANIC %s:%d: " msg "\n", __FILE__, __LINE__, ##__VA_ARGS__); abort(); } while (0)
