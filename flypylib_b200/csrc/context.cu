// context.cu -- fpl_ctx lifetime, error string, arena.
#include "common.cuh"

namespace fpl {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int Arena::reserve(size_t bytes) {
    if (bytes <= cap) return FPL_OK;
    if (base) { cudaFree(base); base = nullptr; cap = 0; }
    size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    cudaError_t e = cudaMalloc((void **)&base, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        base = nullptr;
        return FPL_ENOMEM;
    }
    cap = want;
    used = 0;
    return FPL_OK;
}

void Arena::release() {
    if (base) cudaFree(base);
    base = nullptr; cap = 0; used = 0;
}

}  // namespace fpl

extern "C" {

int fpl_version(void) { return 1; }

const char *fpl_last_error(void) { return fpl::g_err; }

int fpl_device_count(int *count) {
    FPL_REQUIRE(count != nullptr, "fpl_device_count: count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    *count = ok;
    return FPL_OK;
}

int fpl_ctx_create(int device, fpl_ctx **out) {
    FPL_REQUIRE(out != nullptr, "fpl_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        fpl::set_error("fpl_ctx_create: no CUDA device visible (%s); this library has no CPU path",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return FPL_ENODEV;
    }
    FPL_REQUIRE(device >= 0 && device < n, "fpl_ctx_create: device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    FPL_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        fpl::set_error("fpl_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
        return FPL_ENODEV;
    }
    FPL_CUDA_CHECK(cudaSetDevice(device));
    fpl_ctx *c = new fpl_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->h_pinned_bytes = 4096;
    cudaError_t e2 = cudaMallocHost(&c->h_pinned, c->h_pinned_bytes);
    if (e2 != cudaSuccess) {
        delete c;
        fpl::set_error("fpl_ctx_create: cudaMallocHost failed: %s", cudaGetErrorString(e2));
        return FPL_ECUDA;
    }
    *out = c;
    return FPL_OK;
}

int fpl_ctx_destroy(fpl_ctx *ctx) {
    if (!ctx) return FPL_OK;
    cudaSetDevice(ctx->device);
    ctx->arena.release();
    for (auto &b : ctx->act_pool) cudaFree(b.p);
    ctx->act_pool.clear();
    for (auto &b : ctx->slab_cache) cudaFree(b.p);
    ctx->slab_cache.clear();
    for (auto &r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx;
    return FPL_OK;
}

int fpl_ctx_workspace_bytes(fpl_ctx *ctx, int64_t *bytes) {
    FPL_REQUIRE(ctx && bytes, "fpl_ctx_workspace_bytes: NULL argument");
    *bytes = (int64_t)ctx->arena.cap;
    return FPL_OK;
}

int fpl_ctx_release_workspace(fpl_ctx *ctx) {
    FPL_REQUIRE(ctx, "fpl_ctx_release_workspace: NULL ctx");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    FPL_CUDA_CHECK(cudaDeviceSynchronize());
    ctx->arena.release();
    for (auto &b : ctx->act_pool) cudaFree(b.p);
    ctx->act_pool.clear();
    for (auto &b : ctx->slab_cache) cudaFree(b.p);
    ctx->slab_cache.clear();
    return FPL_OK;
}

int fpl_ctx_profile_begin(fpl_ctx *ctx) {
    FPL_REQUIRE(ctx, "fpl_ctx_profile_begin: NULL ctx");
    for (auto &r : ctx->prof) { ctx->prof_free.push_back(r.a); ctx->prof_free.push_back(r.b); }
    ctx->prof.clear();
    // pre-create the event pool so that no cudaEventCreate happens inside a timed region
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    while (ctx->prof_free.size() < 8192) {
        cudaEvent_t e;
        FPL_CUDA_CHECK(cudaEventCreate(&e));
        ctx->prof_free.push_back(e);
    }
    ctx->profiling = true;
    return FPL_OK;
}

int fpl_ctx_profile_end(fpl_ctx *ctx, double *ms_by_tag, double *work_by_tag, int64_t *count_by_tag) {
    FPL_REQUIRE(ctx && ms_by_tag && work_by_tag && count_by_tag, "fpl_ctx_profile_end: NULL argument");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    FPL_CUDA_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < fpl::PROF_NTAGS; ++i) { ms_by_tag[i] = 0; work_by_tag[i] = 0; count_by_tag[i] = 0; }
    for (auto &r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        ms_by_tag[r.tag] += ms; work_by_tag[r.tag] += r.work; count_by_tag[r.tag] += 1;
        ctx->prof_free.push_back(r.a); ctx->prof_free.push_back(r.b);
    }
    ctx->prof.clear();
    ctx->profiling = false;
    return FPL_OK;
}

int fpl_ctx_launch_count(fpl_ctx *ctx, int64_t *launches) {
    FPL_REQUIRE(ctx && launches, "fpl_ctx_launch_count: NULL argument");
    *launches = ctx->launches;
    return FPL_OK;
}

}  // extern "C"
