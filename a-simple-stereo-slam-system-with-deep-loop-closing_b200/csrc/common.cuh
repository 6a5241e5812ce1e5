// common.cuh — shared host-side helpers of libslamb200 (status codes, error text, checks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/slamb200.h"

void sb_set_error(const char *fmt, ...);

#define SB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            sb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return SB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define SB_REQUIRE(cond, msg)                                             \
    do {                                                                  \
        if (!(cond)) {                                                    \
            sb_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return SB_ERR_INVALID;                                        \
        }                                                                 \
    } while (0)

static inline int sb_div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t sb_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
