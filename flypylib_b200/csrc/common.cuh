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

namespace fpl {
// per-kernel-family device timing (CUDA events on the launching stream), off by default
enum ProfTag { PROF_CONV3 = 0, PROF_CONV1 = 1, PROF_FIRST = 2, PROF_NETAUX = 3, PROF_GAUSS = 4,
               PROF_SELECT = 5, PROF_NMS = 6, PROF_TILER = 7, PROF_NTAGS = 8 };
struct ProfRec { cudaEvent_t a, b; int tag; double work; };
// one activation buffer of the tcgen05 forward pass (conv_umma.cu: pool_take)
struct PoolBuf { void *p; size_t cap; bool busy; };
}  // namespace fpl

struct fpl_ctx {
    int device = 0;
    int sm_count = 0;
    fpl::Arena arena;
    int64_t launches = 0;
    void *h_pinned = nullptr;             // small pinned staging buffer (counters, thresholds)
    size_t h_pinned_bytes = 0;
    bool profiling = false;
    std::vector<fpl::ProfRec> prof;
    std::vector<cudaEvent_t> prof_free;
    std::vector<fpl::PoolBuf> act_pool;   // activation buffers of the networks of this context (this device)
    int v2o_decline = 0;                  // why the last fpl_voxel2obj left the two-tier path (0 = it did not)
    long long v2o_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // counters behind the decline (diagnosis)
    int v2o_skip = 0, v2o_fail_streak = 0; // adaptive: calls that go straight to the exact path after declines
    long long v2o_key[5] = {-1, -1, -1, -1, -1};       // (Z, Y, X, r, lw) of the calls the back-off state belongs to
    std::vector<fpl::PoolBuf> slab_cache; // workspace blocks of finished voxel2obj slab sessions, reused by the next
};

namespace fpl {
// RAII scope: records an event pair around the launches issued inside it when profiling is on.
struct ProfScope {
    fpl_ctx *ctx; cudaStream_t st; ProfRec rec; bool on;
    ProfScope(fpl_ctx *c, cudaStream_t s, int tag, double work) : ctx(c), st(s), on(c->profiling) {
        if (!on) return;
        auto get = [&]() { cudaEvent_t e; if (!ctx->prof_free.empty()) { e = ctx->prof_free.back(); ctx->prof_free.pop_back(); }
                           else cudaEventCreate(&e); return e; };
        rec.a = get(); rec.b = get(); rec.tag = tag; rec.work = work;
        cudaEventRecord(rec.a, st);
    }
    ~ProfScope() { if (on) { cudaEventRecord(rec.b, st); ctx->prof.push_back(rec); } }
};
}  // namespace fpl

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
