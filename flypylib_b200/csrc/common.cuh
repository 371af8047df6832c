// common.cuh -- shared host/device plumbing for libfplb200 (context, error handling, arena).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include "../../include/fpl_b200.h"

namespace fpl {

void set_error(const char *fmt, ...);

#define FPL_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            fpl::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,             \
                           cudaGetErrorString(_e));                                       \
            return FPL_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define FPL_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            fpl::set_error(__VA_ARGS__);                                                  \
            return FPL_EINVAL;                                                            \
        }                                                                                 \
    } while (0)

#define FPL_TRY(expr)                                                                     \
    do {                                                                                  \
        int _rc = (expr);                                                                 \
        if (_rc != FPL_OK) return _rc;                                                    \
    } while (0)

// Grow-only device arena: one cudaMalloc'd block, bump allocation, reset per call.
struct Arena {
    char  *base = nullptr;
    size_t cap = 0;
    size_t used = 0;
    int reserve(size_t bytes);            // make sure cap >= bytes (frees + reallocs when growing)
    void reset() { used = 0; }
    void *take(size_t bytes) {            // 256-byte aligned bump allocation; nullptr if exhausted
        size_t off = (used + 255) & ~size_t(255);
        if (off + bytes > cap) return nullptr;
        used = off + bytes;
        return base + off;
    }
    void release();
};

}  // namespace fpl

struct fpl_ctx {
    int device = 0;
    int sm_count = 0;
    fpl::Arena arena;
    int64_t launches = 0;
    void *h_pinned = nullptr;             // small pinned staging buffer (counters, thresholds)
    size_t h_pinned_bytes = 0;
};

#define FPL_LAUNCH_CHECK(ctx)                                                             \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            fpl::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,         \
                           cudaGetErrorString(_e));                                       \
            return FPL_ECUDA;                                                             \
        }                                                                                 \
    } while (0)
