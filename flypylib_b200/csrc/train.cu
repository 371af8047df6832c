// train.cu -- one data-parallel training step of the VGG T-bar classifiers (BASELINE config 5).
//
// Replaces, for `FplNetwork.train` on the VGG builders (compile_args None -> binary_crossentropy + adam,
// flypylib/fplnetwork.py:74-79, :112-128) what Keras/TensorFlow do inside fit_generator for one batch:
//   forward in training mode  : Conv3D(valid, no bias) -> BatchNormalization with BATCH statistics
//                               (biased variance, eps 1e-3; per tower, flypylib/multi_gpu.py:32-46)
//                               -> ReLU [-> Dropout(0.5)] ... -> Conv3D(1,(1,1,1)) + bias -> sigmoid
//   loss                      : binary_crossentropy, mean over the GLOBAL batch
//   backward                  : sigmoid/BCE, dropout, ReLU, BN (batch statistics), conv wgrad + dgrad,
//                               max-pool routing
//   update                    : Adam (Keras defaults), moving averages of the BN statistics (momentum 0.99)
// The gradients live in one flat buffer (Keras get_weights() order) owned by the caller, who
// all-reduces it across ranks (NCCL, torch.distributed) between fpl_train_forward_backward and
// fpl_train_apply.  Patches are tiny (64 x 24^3 per GPU, 1.1 GFLOP forward each): everything here is
// plain fp32 CUDA-core code, correctness first; activations are (N,z,y,x,C) float32.
#include "net.cuh"
#include <math.h>

namespace fpl {
namespace train {

using net::Op;
using net::OP_CONV;
using net::OP_FINAL;
using net::OP_POOL;

struct Layer {                 // one Conv3D + BN + ReLU (+ Dropout) block
    int k, cin, cout, din, dout;
    bool dropout, pool_after;
    size_t off_kernel, off_gamma, off_beta, off_mean, off_var;   // offsets into the flat parameter vector
    size_t off_bn;                                               // offset into the batch-statistics vector
    float *x = nullptr;        // conv output (pre BN)            (N, dout^3, cout)
    float *y = nullptr;        // after BN + ReLU (pre dropout)
    float *yd = nullptr;       // after dropout (== y when no dropout)
    float *yp = nullptr;       // after pooling (when pool_after)
    double *stat = nullptr;    // [2*cout] sum, sumsq  then  [2*cout] dgamma, dbeta  (double accumulators)
    float *mean = nullptr, *invstd = nullptr;
};

}  // namespace train
}  // namespace fpl

struct fpl_trainer {
    fpl_ctx *ctx = nullptr;
    int arch = 0, patch = 0, batch = 0;
    std::vector<fpl::train::Layer> layers;
    size_t off_final_kernel = 0, off_final_bias = 0, n_params = 0, n_bn = 0;
    int final_cin = 0;
    float *g0 = nullptr, *g1 = nullptr;       // gradient ping-pong (max activation size)
    float *logit = nullptr, *dlogit = nullptr;
    double *loss_acc = nullptr;               // [0] sum of per-sample BCE, [1] correct count
    std::vector<void *> allocs;
};

namespace fpl {
namespace train {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// raw convolution, one thread = one output voxel x 8 output channels
template <int K>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const float *__restrict__ in, const float *__restrict__ w, float *__restrict__ out, int n,
                int din, int cin, int cout) {
    const int dout = din - (K - 1);
    const int cg = cout / 8;
    const long long total = (long long)n * dout * dout * dout * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg); long long v = i / cg;
        const int x = (int)(v % dout); v /= dout;
        const int y = (int)(v % dout); v /= dout;
        const int z = (int)(v % dout); const int t = (int)(v / dout);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int kd = 0; kd < K; ++kd)
            for (int kh = 0; kh < K; ++kh)
                for (int kw = 0; kw < K; ++kw) {
                    const float *ip = in + ((((size_t)t * din + z + kd) * din + y + kh) * din + x + kw) * cin;
                    const float *wp = w + (size_t)((kd * K + kh) * K + kw) * cin * cout + g * 8;
                    for (int ci = 0; ci < cin; ++ci) {
                        const float a = __ldg(ip + ci);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, __ldg(wp + (size_t)ci * cout + j), acc[j]);
                    }
                }
        float *op = out + (((size_t)t * dout + z) * dout + y) * dout * cout + (size_t)x * cout + g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = acc[j];
    }
}

// per-channel sum and sum of squares over `rows` rows of C channels (double accumulation)
__global__ void __launch_bounds__(256)
bn_sums_kernel(const float *__restrict__ x, long long rows, int c, double *__restrict__ stat) {
    const int lanes = 256 / c;                 // row lanes per block
    const int ch = threadIdx.x % c, rl = threadIdx.x / c;
    double s = 0.0, ss = 0.0;
    if (rl < lanes)
        for (long long r = (long long)blockIdx.x * lanes + rl; r < rows; r += (long long)gridDim.x * lanes) {
            const double v = (double)x[r * c + ch];
            s += v; ss += v * v;
        }
    if (rl < lanes) { atomicAdd(&stat[ch], s); atomicAdd(&stat[c + ch], ss); }
}

__global__ void bn_finalize_kernel(const double *__restrict__ stat, long long rows, int c, float eps,
                                   float *__restrict__ mean, float *__restrict__ invstd,
                                   float *__restrict__ bn_batch) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const double m = stat[ch] / (double)rows;
    double var = stat[c + ch] / (double)rows - m * m;       // biased variance (Keras non-fused BN)
    if (var < 0) var = 0;
    mean[ch] = (float)m;
    invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
    bn_batch[ch] = (float)m;
    bn_batch[c + ch] = (float)var;
}

// y = relu(gamma * (x - mean) * invstd + beta); optional dropout(0.5) into yd
__device__ __forceinline__ unsigned long long splitmix(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ bool keep_bit(unsigned long long seed, unsigned long long layer, unsigned long long idx) {
    return splitmix(idx + (seed + layer * 0x632BE59BD9B4E019ULL) * 0x9E3779B97F4A7C15ULL) & 1ULL;
}

__global__ void __launch_bounds__(256)
bn_relu_dropout_kernel(const float *__restrict__ x, long long total, int c, const float *__restrict__ gamma,
                       const float *__restrict__ beta, const float *__restrict__ mean,
                       const float *__restrict__ invstd, float *__restrict__ y, float *__restrict__ yd,
                       unsigned long long seed, unsigned long long layer) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        float v = (x[i] - mean[ch]) * invstd[ch] * gamma[ch] + beta[ch];
        v = fmaxf(v, 0.f);
        y[i] = v;
        if (yd) yd[i] = keep_bit(seed, layer, (unsigned long long)i) ? v * 2.f : 0.f;
    }
}

__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, int n, int din, int c) {
    const int dout = din / 2;
    const long long total = (long long)n * dout * dout * dout * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c); long long v = i / c;
        const int x = (int)(v % dout); v /= dout;
        const int y = (int)(v % dout); v /= dout;
        const int z = (int)(v % dout); const int t = (int)(v / dout);
        float m = -INFINITY;
        for (int dz = 0; dz < 2; ++dz)
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx)
                    m = fmaxf(m, in[((((size_t)t * din + 2 * z + dz) * din + 2 * y + dy) * din + 2 * x + dx) * c + ch]);
        out[i] = m;
    }
}

// gradient of max-pool: the first position (dz,dy,dx order) that attains the max receives the gradient
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float *__restrict__ in, const float *__restrict__ pooled, const float *__restrict__ dpooled,
                float *__restrict__ din_, int n, int din, int c) {
    const int dout = din / 2;
    const long long total = (long long)n * dout * dout * dout * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c); long long v = i / c;
        const int x = (int)(v % dout); v /= dout;
        const int y = (int)(v % dout); v /= dout;
        const int z = (int)(v % dout); const int t = (int)(v / dout);
        const float m = pooled[i], g = dpooled[i];
        bool given = false;
        for (int dz = 0; dz < 2; ++dz)
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const size_t o = ((((size_t)t * din + 2 * z + dz) * din + 2 * y + dy) * din + 2 * x + dx) * c + ch;
                    const bool hit = !given && in[o] == m;
                    din_[o] = hit ? g : 0.f;
                    given = given || hit;
                }
    }
}

// final Conv3D(1,(1,1,1)) + bias + sigmoid + BCE on 1x1x1 outputs; one block
__global__ void __launch_bounds__(256)
final_fwd_bwd_kernel(const float *__restrict__ yin, int n, int c, const float *__restrict__ w,
                     const float *__restrict__ bias, const unsigned char *__restrict__ labels, float loss_scale,
                     float *__restrict__ gw, float *__restrict__ gb, float *__restrict__ dy,
                     double *__restrict__ loss_acc, float *__restrict__ logit_out) {
    __shared__ float s_dl[1024];
    __shared__ double s_loss[256];
    __shared__ int s_ok[256];
    double my_loss = 0.0; int my_ok = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float l = bias[0];
        for (int ch = 0; ch < c; ++ch) l = fmaf(yin[(size_t)i * c + ch], w[ch], l);
        const float t = labels[i] ? 1.f : 0.f;
        const float p = 1.f / (1.f + expf(-l));
        // BCE from the logit (softplus form); Keras additionally clips p to [1e-7, 1-1e-7]
        const float bce = fmaxf(l, 0.f) - l * t + log1pf(expf(-fabsf(l)));
        my_loss += (double)bce;
        my_ok += ((p > 0.5f) == (t > 0.5f)) ? 1 : 0;
        s_dl[i] = (p - t) * loss_scale;
        logit_out[i] = l;
    }
    s_loss[threadIdx.x] = my_loss; s_ok[threadIdx.x] = my_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        double L = 0; int K = 0;
        for (int i = 0; i < blockDim.x; ++i) { L += s_loss[i]; K += s_ok[i]; }
        loss_acc[0] = L; loss_acc[1] = (double)K;
        float g = 0.f;
        for (int i = 0; i < n; ++i) g += s_dl[i];
        gb[0] = g;
    }
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float g = 0.f;
        for (int i = 0; i < n; ++i) g = fmaf(s_dl[i], yin[(size_t)i * c + ch], g);
        gw[ch] = g;
    }
    for (int i = threadIdx.x; i < n * c; i += blockDim.x) dy[i] = s_dl[i / c] * w[i % c];
}

// dz = dOut (through dropout mask) * (y > 0); accumulate dgamma = sum dz*xhat, dbeta = sum dz
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float *__restrict__ dout, const float *__restrict__ x, const float *__restrict__ y,
                     long long rows, int c, const float *__restrict__ mean, const float *__restrict__ invstd,
                     int dropout, unsigned long long seed, unsigned long long layer, float *__restrict__ dz,
                     double *__restrict__ acc) {
    const int lanes = 256 / c;
    const int ch = threadIdx.x % c, rl = threadIdx.x / c;
    double dg = 0.0, db = 0.0;
    if (rl < lanes)
        for (long long r = (long long)blockIdx.x * lanes + rl; r < rows; r += (long long)gridDim.x * lanes) {
            const long long i = r * c + ch;
            float g = dout[i];
            if (dropout) g = keep_bit(seed, layer, (unsigned long long)i) ? g * 2.f : 0.f;
            g = y[i] > 0.f ? g : 0.f;
            dz[i] = g;
            const float xhat = (x[i] - mean[ch]) * invstd[ch];
            dg += (double)g * (double)xhat; db += (double)g;
        }
    if (rl < lanes) { atomicAdd(&acc[ch], dg); atomicAdd(&acc[c + ch], db); }
}

// dx = gamma*invstd * (dz - dbeta/M - xhat*dgamma/M); also emits dgamma/dbeta as float gradients
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float *__restrict__ dz_dx, const float *__restrict__ x, long long total, int c, long long rows,
                    const float *__restrict__ gamma, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const double *__restrict__ acc,
                    float *__restrict__ g_gamma, float *__restrict__ g_beta) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const float xhat = (x[i] - mean[ch]) * invstd[ch];
        const float dg = (float)(acc[ch] / (double)rows), db = (float)(acc[c + ch] / (double)rows);
        dz_dx[i] = gamma[ch] * invstd[ch] * (dz_dx[i] - db - xhat * dg);
        if (i < c) { g_gamma[i] = (float)acc[i]; g_beta[i] = (float)acc[c + i]; }
    }
}

// weight gradient: dW[tap][ci][co] += sum over a chunk of output voxels of in[.. + tap][ci] * dx[..][co]
constexpr int kWgRows = 8, kWgAcc = 36;
template <int K>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float *__restrict__ in, const float *__restrict__ dx, float *__restrict__ dw, int n,
                  int din, int cin, int cout, long long rows_per_block) {
    __shared__ float s_in[kWgRows][96];
    __shared__ float s_dx[kWgRows][96];
    const int dout = din - (K - 1);
    const int tap = blockIdx.x;
    const int kd = tap / (K * K), kh = (tap / K) % K, kw = tap % K;
    const long long rows = (long long)n * dout * dout * dout;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    const int pairs = cin * cout;
    float acc[kWgAcc];
#pragma unroll
    for (int j = 0; j < kWgAcc; ++j) acc[j] = 0.f;
    for (long long rb = r0; rb < r1; rb += kWgRows) {
        for (int i = threadIdx.x; i < kWgRows * (cin + cout); i += blockDim.x) {
            const int rr = i / (cin + cout), e = i % (cin + cout);
            const long long r = rb + rr;
            float v = 0.f;
            if (r < r1) {
                if (e < cin) {
                    long long q = r;
                    const int x = (int)(q % dout); q /= dout;
                    const int y = (int)(q % dout); q /= dout;
                    const int z = (int)(q % dout); const int t = (int)(q / dout);
                    v = in[((((size_t)t * din + z + kd) * din + y + kh) * din + x + kw) * cin + e];
                } else {
                    v = dx[(size_t)r * cout + (e - cin)];
                }
            }
            if (e < cin) s_in[rr][e] = v; else s_dx[rr][e - cin] = v;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kWgAcc; ++j) {
            const int p = threadIdx.x + j * 256;
            if (p < pairs) {
                const int ci = p / cout, co = p % cout;
#pragma unroll
                for (int rr = 0; rr < kWgRows; ++rr) acc[j] = fmaf(s_in[rr][ci], s_dx[rr][co], acc[j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kWgAcc; ++j) {
        const int p = threadIdx.x + j * 256;
        if (p < pairs) atomicAdd(&dw[(size_t)tap * pairs + p], acc[j]);
    }
}

// input gradient: din[n,z,y,x,ci] = sum_{tap,co} dx[n,z-kd,y-kh,x-kw,co] * W[tap][ci][co]
template <int K>
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(const float *__restrict__ dx, const float *__restrict__ w, float *__restrict__ din_, int n,
                  int din, int cin, int cout) {
    const int dout = din - (K - 1);
    const int cg = cin / 8;
    const long long total = (long long)n * din * din * din * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg); long long v = i / cg;
        const int x = (int)(v % din); v /= din;
        const int y = (int)(v % din); v /= din;
        const int z = (int)(v % din); const int t = (int)(v / din);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int kd = 0; kd < K; ++kd) {
            const int oz = z - kd; if (oz < 0 || oz >= dout) continue;
            for (int kh = 0; kh < K; ++kh) {
                const int oy = y - kh; if (oy < 0 || oy >= dout) continue;
                for (int kw = 0; kw < K; ++kw) {
                    const int ox = x - kw; if (ox < 0 || ox >= dout) continue;
                    const float *dp = dx + ((((size_t)t * dout + oz) * dout + oy) * dout + ox) * cout;
                    const float *wp = w + ((size_t)((kd * K + kh) * K + kw) * cin + g * 8) * cout;
                    for (int co = 0; co < cout; ++co) {
                        const float d = __ldg(dp + co);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = fmaf(d, __ldg(wp + (size_t)j * cout + co), acc[j]);
                    }
                }
            }
        }
        float *op = din_ + ((((size_t)t * din + z) * din + y) * din + x) * cin + g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = acc[j];
    }
}

// Adam (Keras: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)) on the trainable ranges and
// moving-average update of the BN statistics
struct Range { long long begin, end; };
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
            long long begin, long long end, float lr_t, float b1, float b2, float eps) {
    for (long long i = begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < end;
         i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
}
__global__ void moving_kernel(float *__restrict__ p, long long off_mean, long long off_var,
                              const float *__restrict__ bn_batch, int c, float momentum) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    p[off_mean + ch] = p[off_mean + ch] * momentum + bn_batch[ch] * (1.f - momentum);
    p[off_var + ch] = p[off_var + ch] * momentum + bn_batch[c + ch] * (1.f - momentum);
}

static int blocks_for(fpl_ctx *ctx, long long total) {
    long long b = (total + 255) / 256, cap = (long long)ctx->sm_count * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

static int row_blocks(fpl_ctx *ctx, long long rows, int c) {     // grid for the (rows x C) reductions
    const long long lanes = 256 / c;
    long long b = (rows + lanes - 1) / lanes, cap = (long long)ctx->sm_count * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace train
}  // namespace fpl

using namespace fpl::train;

extern "C" {

int fpl_train_destroy(fpl_trainer *t) {
    if (!t) return FPL_OK;
    cudaSetDevice(t->ctx->device);
    for (void *p : t->allocs) cudaFree(p);
    delete t;
    return FPL_OK;
}

int fpl_train_create(fpl_ctx *ctx, int arch, int patch_sz, int batch, fpl_trainer **out) {
    FPL_REQUIRE(ctx && out, "fpl_train_create: NULL argument");
    FPL_REQUIRE(arch == FPL_ARCH_VGG_LIKE || arch == FPL_ARCH_VGG_LIKE2,
                "fpl_train_create: the training step covers the VGG builders (binary_crossentropy/adam defaults)");
    FPL_REQUIRE(batch > 0 && batch <= 1024, "fpl_train_create: batch must be in 1..1024");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    fpl_net *gnet = nullptr;
    FPL_TRY(fpl_net_create(ctx, arch, &gnet));
    fpl_trainer *t = new fpl_trainer();
    t->ctx = ctx; t->arch = arch; t->patch = patch_sz; t->batch = batch;
    int d = patch_sz, c = 1, li = 0;
    size_t off = 0, off_bn = 0, max_elems = (size_t)batch * d * d * d;
    bool ok = true;
    for (size_t i = 0; i < gnet->ops.size() && ok; ++i) {
        const Op &o = gnet->ops[i];
        if (o.kind == OP_CONV) {
            Layer L;
            L.k = o.k; L.cin = o.cin; L.cout = o.cout; L.din = d; L.dout = d - (o.k - 1);
            if (L.dout <= 0) { ok = false; break; }
            L.dropout = (li == 5 || li == 6);       // Dropout(0.5) after full1 and full2 (fplmodels.py:127,131,163,167)
            L.pool_after = i + 1 < gnet->ops.size() && gnet->ops[i + 1].kind == OP_POOL;
            L.off_kernel = off; off += (size_t)o.k * o.k * o.k * o.cin * o.cout;
            L.off_gamma = off; off += o.cout;
            L.off_beta = off; off += o.cout;
            L.off_mean = off; off += o.cout;
            L.off_var = off; off += o.cout;
            L.off_bn = off_bn; off_bn += 2 * (size_t)o.cout;
            d = L.dout; c = o.cout;
            size_t e = (size_t)batch * d * d * d * c;
            if (e > max_elems) max_elems = e;
            if (L.pool_after) { if (d % 2) { ok = false; break; } d /= 2; }
            t->layers.push_back(L);
            ++li;
        } else if (o.kind == OP_FINAL) {
            t->final_cin = o.cin;
            t->off_final_kernel = off; off += o.cin;
            t->off_final_bias = off; off += 1;
        }
    }
    fpl_net_destroy(gnet);
    if (!ok || d != 1) {
        delete t;
        fpl::set_error("fpl_train_create: patch edge %d does not reduce to a 1x1x1 output for this architecture", patch_sz);
        return FPL_EINVAL;
    }
    t->n_params = off; t->n_bn = off_bn;
    auto alloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        t->allocs.push_back(p);
        return p;
    };
    bool mem_ok = true;
    for (Layer &L : t->layers) {
        const size_t e = (size_t)batch * L.dout * L.dout * L.dout * L.cout;
        L.x = (float *)alloc(e * 4); L.y = (float *)alloc(e * 4);
        L.yd = L.dropout ? (float *)alloc(e * 4) : L.y;
        if (L.pool_after) L.yp = (float *)alloc(e / 8 * 4);
        L.stat = (double *)alloc(4 * L.cout * sizeof(double));
        L.mean = (float *)alloc(L.cout * 4); L.invstd = (float *)alloc(L.cout * 4);
        mem_ok = mem_ok && L.x && L.y && L.yd && (!L.pool_after || L.yp) && L.stat && L.mean && L.invstd;
    }
    t->g0 = (float *)alloc(max_elems * 4); t->g1 = (float *)alloc(max_elems * 4);
    t->logit = (float *)alloc(batch * 4); t->dlogit = (float *)alloc(batch * 4);
    t->loss_acc = (double *)alloc(2 * sizeof(double));
    if (!mem_ok || !t->g0 || !t->g1 || !t->logit || !t->dlogit || !t->loss_acc) {
        fpl_train_destroy(t);
        fpl::set_error("fpl_train_create: device allocation failed");
        return FPL_ENOMEM;
    }
    *out = t;
    return FPL_OK;
}

int fpl_train_sizes(const fpl_trainer *t, int64_t *n_params, int64_t *n_bn) {
    FPL_REQUIRE(t, "fpl_train_sizes: NULL trainer");
    if (n_params) *n_params = (int64_t)t->n_params;
    if (n_bn) *n_bn = (int64_t)t->n_bn;
    return FPL_OK;
}

int fpl_train_forward_backward(fpl_trainer *t, const float *d_x, const uint8_t *d_labels, const float *d_params,
                               float *d_grads, float *d_bn_batch, float loss_scale, uint64_t dropout_seed,
                               double *h_loss_sum, int64_t *h_correct, void *stream) {
    FPL_REQUIRE(t && d_x && d_labels && d_params && d_grads && d_bn_batch, "fpl_train_forward_backward: NULL argument");
    fpl_ctx *ctx = t->ctx;
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = t->batch;
    FPL_CUDA_CHECK(cudaMemsetAsync(d_grads, 0, t->n_params * sizeof(float), st));
    // ---------------- forward (training mode)
    const float *cur = d_x;
    for (size_t li = 0; li < t->layers.size(); ++li) {
        Layer &L = t->layers[li];
        const long long rows = (long long)n * L.dout * L.dout * L.dout, total = rows * L.cout;
        if (L.k == 3) conv_fwd_kernel<3><<<blocks_for(ctx, total / 8), 256, 0, st>>>(cur, d_params + L.off_kernel, L.x, n, L.din, L.cin, L.cout);
        else conv_fwd_kernel<1><<<blocks_for(ctx, total / 8), 256, 0, st>>>(cur, d_params + L.off_kernel, L.x, n, L.din, L.cin, L.cout);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaMemsetAsync(L.stat, 0, 4 * L.cout * sizeof(double), st));
        bn_sums_kernel<<<row_blocks(ctx, rows, L.cout), 256, 0, st>>>(L.x, rows, L.cout, L.stat);
        FPL_LAUNCH_CHECK(ctx);
        bn_finalize_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(L.stat, rows, L.cout, 1e-3f, L.mean, L.invstd, d_bn_batch + L.off_bn);
        FPL_LAUNCH_CHECK(ctx);
        bn_relu_dropout_kernel<<<blocks_for(ctx, total), 256, 0, st>>>(L.x, total, L.cout, d_params + L.off_gamma,
                                                                      d_params + L.off_beta, L.mean, L.invstd, L.y,
                                                                      L.dropout ? L.yd : nullptr, dropout_seed, li);
        FPL_LAUNCH_CHECK(ctx);
        cur = L.yd;
        if (L.pool_after) {
            pool_fwd_kernel<<<blocks_for(ctx, total / 8), 256, 0, st>>>(L.yd, L.yp, n, L.dout, L.cout);
            FPL_LAUNCH_CHECK(ctx);
            cur = L.yp;
        }
    }
    // ---------------- loss + backward
    float *ga = t->g0, *gb = t->g1;
    final_fwd_bwd_kernel<<<1, 256, 0, st>>>(cur, n, t->final_cin, d_params + t->off_final_kernel,
                                            d_params + t->off_final_bias, d_labels, loss_scale,
                                            d_grads + t->off_final_kernel, d_grads + t->off_final_bias, ga,
                                            t->loss_acc, t->logit);
    FPL_LAUNCH_CHECK(ctx);
    for (int li = (int)t->layers.size() - 1; li >= 0; --li) {
        Layer &L = t->layers[li];
        const long long rows = (long long)n * L.dout * L.dout * L.dout, total = rows * L.cout;
        if (L.pool_after) {      // `ga` holds the gradient w.r.t. the pooled tensor
            pool_bwd_kernel<<<blocks_for(ctx, total / 8), 256, 0, st>>>(L.yd, L.yp, ga, gb, n, L.dout, L.cout);
            FPL_LAUNCH_CHECK(ctx);
            float *tmp = ga; ga = gb; gb = tmp;
        }
        FPL_CUDA_CHECK(cudaMemsetAsync(L.stat + 2 * L.cout, 0, 2 * L.cout * sizeof(double), st));
        bn_bwd_reduce_kernel<<<row_blocks(ctx, rows, L.cout), 256, 0, st>>>(ga, L.x, L.y, rows, L.cout, L.mean, L.invstd,
                                                                   L.dropout ? 1 : 0, dropout_seed, (unsigned long long)li,
                                                                   gb, L.stat + 2 * L.cout);
        FPL_LAUNCH_CHECK(ctx);
        bn_bwd_apply_kernel<<<blocks_for(ctx, total), 256, 0, st>>>(gb, L.x, total, L.cout, rows, d_params + L.off_gamma,
                                                                   L.mean, L.invstd, L.stat + 2 * L.cout,
                                                                   d_grads + L.off_gamma, d_grads + L.off_beta);
        FPL_LAUNCH_CHECK(ctx);
        // gb now holds dx (gradient w.r.t. the conv output)
        const float *lin = li == 0 ? d_x : (t->layers[li - 1].pool_after ? t->layers[li - 1].yp : t->layers[li - 1].yd);
        FPL_REQUIRE(L.cin * L.cout <= 256 * kWgAcc && L.cin <= 96 && L.cout <= 96, "wgrad: layer too wide");
        const long long rpb = 2048;
        dim3 wg(L.k * L.k * L.k, (unsigned)((rows + rpb - 1) / rpb));
        if (L.k == 3) conv_wgrad_kernel<3><<<wg, 256, 0, st>>>(lin, gb, d_grads + L.off_kernel, n, L.din, L.cin, L.cout, rpb);
        else conv_wgrad_kernel<1><<<wg, 256, 0, st>>>(lin, gb, d_grads + L.off_kernel, n, L.din, L.cin, L.cout, rpb);
        FPL_LAUNCH_CHECK(ctx);
        if (li > 0) {
            const long long tin = (long long)n * L.din * L.din * L.din * L.cin;
            if (L.k == 3) conv_dgrad_kernel<3><<<blocks_for(ctx, tin / 8), 256, 0, st>>>(gb, d_params + L.off_kernel, ga, n, L.din, L.cin, L.cout);
            else conv_dgrad_kernel<1><<<blocks_for(ctx, tin / 8), 256, 0, st>>>(gb, d_params + L.off_kernel, ga, n, L.din, L.cin, L.cout);
            FPL_LAUNCH_CHECK(ctx);
            // ga = gradient w.r.t. this layer's input = previous block's (pooled / dropped-out) output
        }
    }
    double *h = (double *)ctx->h_pinned;
    FPL_CUDA_CHECK(cudaMemcpyAsync(h, t->loss_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_loss_sum) *h_loss_sum = h[0];
    if (h_correct) *h_correct = (int64_t)h[1];
    return FPL_OK;
}

int fpl_train_apply(fpl_trainer *t, float *d_params, const float *d_grads, float *d_m, float *d_v,
                    const float *d_bn_batch, int64_t step, float lr, float beta1, float beta2, float eps,
                    float bn_momentum, void *stream) {
    FPL_REQUIRE(t && d_params && d_grads && d_m && d_v && d_bn_batch && step >= 1, "fpl_train_apply: bad argument");
    fpl_ctx *ctx = t->ctx;
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    auto adam = [&](size_t b, size_t e) {
        adam_kernel<<<blocks_for(ctx, (long long)(e - b)), 256, 0, st>>>(d_params, d_grads, d_m, d_v, (long long)b,
                                                                        (long long)e, (float)lr_t, beta1, beta2, eps);
        ctx->launches++;
    };
    for (Layer &L : t->layers) {
        adam(L.off_kernel, L.off_mean);                       // kernel, gamma, beta are contiguous
        moving_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(d_params, (long long)L.off_mean, (long long)L.off_var,
                                                            d_bn_batch + L.off_bn, L.cout, bn_momentum);
        ctx->launches++;
    }
    adam(t->off_final_kernel, t->off_final_bias + 1);
    FPL_CUDA_CHECK(cudaGetLastError());
    return FPL_OK;
}

}  // extern "C"
