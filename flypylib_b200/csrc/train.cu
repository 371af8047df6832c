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
// fpl_train_apply.  Activations are (N,z,y,x,C) float32.  The three contractions of every convolution (forward,
// dgrad, wgrad) run on the tensor cores (train_tc.cuh: tcgen05 / TMEM, bf16 hi/lo split operands = fp32-class
// results by default); the fp32 CUDA-core kernels below remain as the validation path (FPL_PREC_FP32).
#include "net.cuh"
#include "train_tc.cuh"
#include <math.h>
#include <stdlib.h>

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
    float *gpad = nullptr;                    // zero-padded dx of the layer whose dgrad runs (tensor-core path)
    float *gemb = nullptr;                    // dx on the input grid of the layer whose wgrad runs (slab wgrad)
    void *wimg = nullptr;                     // packed bf16 hi/lo weight image of the convolution that runs (slab conv)
    int precision = FPL_PREC_TF32;            // FPL_PREC_TF32: bf16 hi/lo x3 on tcgen05; FPL_PREC_BF16; FPL_PREC_FP32: CUDA cores
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

// per-channel sum and sum of squares over `rows` rows of C channels (double accumulation).  Each thread owns one
// channel of `lanes` interleaved rows, keeps 8 independent loads in flight, the block reduces its row lanes in shared
// memory and issues ONE atomic per channel and statistic.
constexpr int kBnUnroll = 8;
__global__ void __launch_bounds__(256)
bn_sums_kernel(const float *__restrict__ x, long long rows, int c, double *__restrict__ stat) {
    __shared__ double s_red[2][256];
    const int lanes = 256 / c;                 // row lanes per block
    const int ch = threadIdx.x % c, rl = threadIdx.x / c;
    double s = 0.0, ss = 0.0;
    if (rl < lanes) {
        const long long stride = (long long)gridDim.x * lanes;
        for (long long r0 = (long long)blockIdx.x * lanes + rl; r0 < rows; r0 += stride * kBnUnroll) {
            float v[kBnUnroll];
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) {
                const long long r = r0 + u * stride;
                v[u] = r < rows ? __ldg(x + r * c + ch) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kBnUnroll; ++u) { const double d = (double)v[u]; s += d; ss += d * d; }
        }
    }
    s_red[0][threadIdx.x] = s; s_red[1][threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.x < c) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < lanes; ++l) { a += s_red[0][l * c + threadIdx.x]; b += s_red[1][l * c + threadIdx.x]; }
        atomicAdd(&stat[threadIdx.x], a); atomicAdd(&stat[c + threadIdx.x], b);
    }
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
    __shared__ float s_m[96], s_k[96], s_g[96], s_b[96];
    for (int i = threadIdx.x; i < c; i += blockDim.x) { s_m[i] = mean[i]; s_k[i] = invstd[i]; s_g[i] = gamma[i]; s_b[i] = beta[i]; }
    __syncthreads();
    const int c4 = c >> 2;
    const long long total4 = total >> 2;            // C is a multiple of 4
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c4) * 4;
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + i);
        const float in[4] = {xv.x, xv.y, xv.z, xv.w};
        float o[4], od[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // same operation order as the scalar form: ((x - mean) * invstd) * gamma + beta
            float v = (in[e] - s_m[ch + e]) * s_k[ch + e] * s_g[ch + e] + s_b[ch + e];
            v = fmaxf(v, 0.f);
            o[e] = v;
            od[e] = (yd && !keep_bit(seed, layer, (unsigned long long)(i * 4 + e))) ? 0.f : v * 2.f;
        }
        reinterpret_cast<float4 *>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (yd) reinterpret_cast<float4 *>(yd)[i] = make_float4(od[0], od[1], od[2], od[3]);
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
    __shared__ double s_red[2][256];
    constexpr int U = 4;
    const int lanes = 256 / c;
    const int ch = threadIdx.x % c, rl = threadIdx.x / c;
    double dg = 0.0, db = 0.0;
    if (rl < lanes) {
        const float m = mean[ch], is = invstd[ch];
        const long long stride = (long long)gridDim.x * lanes;
        for (long long r0 = (long long)blockIdx.x * lanes + rl; r0 < rows; r0 += stride * U) {
            float gv[U], xv[U], yv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * stride;
                const bool ok = r < rows;
                const long long i = r * c + ch;
                gv[u] = ok ? __ldg(dout + i) : 0.f; xv[u] = ok ? __ldg(x + i) : 0.f; yv[u] = ok ? __ldg(y + i) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * stride;
                if (r >= rows) continue;
                const long long i = r * c + ch;
                float g = gv[u];
                if (dropout) g = keep_bit(seed, layer, (unsigned long long)i) ? g * 2.f : 0.f;
                g = yv[u] > 0.f ? g : 0.f;
                dz[i] = g;
                const float xhat = (xv[u] - m) * is;
                dg += (double)g * (double)xhat; db += (double)g;
            }
        }
    }
    s_red[0][threadIdx.x] = dg; s_red[1][threadIdx.x] = db;
    __syncthreads();
    if (threadIdx.x < c) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < lanes; ++l) { a += s_red[0][l * c + threadIdx.x]; b += s_red[1][l * c + threadIdx.x]; }
        atomicAdd(&acc[threadIdx.x], a); atomicAdd(&acc[c + threadIdx.x], b);
    }
}

// dx = gamma*invstd * (dz - dbeta/M - xhat*dgamma/M); also emits dgamma/dbeta as float gradients
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float *__restrict__ dz_dx, const float *__restrict__ x, long long total, int c, long long rows,
                    const float *__restrict__ gamma, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const double *__restrict__ acc,
                    float *__restrict__ g_gamma, float *__restrict__ g_beta) {
    __shared__ float s_m[96], s_is[96], s_k[96], s_dg[96], s_db[96];
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        s_m[i] = mean[i]; s_is[i] = invstd[i]; s_k[i] = gamma[i] * invstd[i];
        s_dg[i] = (float)(acc[i] / (double)rows); s_db[i] = (float)(acc[c + i] / (double)rows);
        if (blockIdx.x == 0) { g_gamma[i] = (float)acc[i]; g_beta[i] = (float)acc[c + i]; }
    }
    __syncthreads();
    const int c4 = c >> 2;
    const long long total4 = total >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c4) * 4;
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + i);
        const float4 dv = reinterpret_cast<const float4 *>(dz_dx)[i];
        const float xi[4] = {xv.x, xv.y, xv.z, xv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float xhat = (xi[e] - s_m[ch + e]) * s_is[ch + e];
            o[e] = s_k[ch + e] * (di[e] - s_db[ch + e] - xhat * s_dg[ch + e]);
        }
        reinterpret_cast<float4 *>(dz_dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
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
    long long b = (rows + lanes - 1) / lanes, cap = (long long)ctx->sm_count * 8;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}


// ------------------------------------------------------------------------------------------------
// tensor-core launches (train_tc.cuh)
// ------------------------------------------------------------------------------------------------
static bool tc_layer_ok(int k, int cin, int cout) {
    const bool cin_ok = (cin == 1 && k == 3) || cin % 48 == 0;
    return cin_ok && cout % 48 == 0 && cout <= kTcMaxN && cin <= kTcMaxN && (k == 1 || k == 3);
}

template <int NS>
static int tc_set_smem() {        // per launch: the attribute belongs to the current device
    FPL_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FPL_CUDA_CHECK(cudaFuncSetAttribute(tc_wgrad_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return FPL_OK;
}

// valid convolution of `act` (n, din^3, ca) with the k^3 kernel `w`; flip = 0: forward (nn = cout, w read as
// (tap, ca, nn)), flip = 1: dgrad on the zero-padded dx (nn = cin, w read as (T-1-tap, nn, ca)); out (n*dout^3, nn)
static int launch_tc_conv(fpl_ctx *ctx, int prec, const float *act, const float *w, float *out, int n, int din, int k,
                          int ca, int nn, int flip, cudaStream_t st) {
    TcArgs a{};
    a.act = act; a.mat = w; a.out = out;
    a.n_img = n; a.din = din; a.dout = din - (k - 1); a.k = k; a.ca = ca; a.nn = nn; a.flip = flip;
    a.K = k * k * k * ca;
    a.rows = (long long)n * a.dout * a.dout * a.dout;
    const int apc = ca == 1 ? 4 : 6;
    a.a_bytes = (uint32_t)apc * 2048u; a.b_bytes = (uint32_t)apc * (uint32_t)nn * 16u;
    const int np = prec == FPL_PREC_BF16 ? 1 : 2;
    const size_t smem = 2 * (size_t)np * (a.a_bytes + a.b_bytes);
    const long long tiles = (a.rows + 127) / 128;
    const long long cap = (long long)ctx->sm_count * 3;
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (prec == FPL_PREC_BF16) { FPL_TRY(tc_set_smem<1>()); tc_conv_kernel<1><<<grid, kTcThreads, smem, st>>>(a); }
    else { FPL_TRY(tc_set_smem<3>()); tc_conv_kernel<3><<<grid, kTcThreads, smem, st>>>(a); }
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

// dW (k^3*cin, cout) += sum over the n*dout^3 output voxels of in[v + tap, ci] * dx[v, co]
static int launch_tc_wgrad(fpl_ctx *ctx, int prec, const float *in, const float *dx, float *dw, int n, int din, int k,
                           int cin, int cout, cudaStream_t st) {
    TcArgs a{};
    a.act = in; a.mat = dx; a.out = dw;
    a.n_img = n; a.din = din; a.dout = din - (k - 1); a.k = k; a.ca = cin; a.nn = cout;
    a.Mtot = k * k * k * cin;
    a.rows = (long long)n * a.dout * a.dout * a.dout;
    FPL_REQUIRE((long long)n * din * din * din * cin < (1ll << 31), "wgrad: activation tensor too large for 32-bit offsets");
    a.a_bytes = 8u * 2048u; a.b_bytes = 8u * (uint32_t)cout * 16u;
    const int np = prec == FPL_PREC_BF16 ? 1 : 2;
    const size_t smem = 2 * (size_t)np * (a.a_bytes + a.b_bytes);
    const int mtiles = (a.Mtot + 127) / 128;
    const long long nchunks = (a.rows + 63) / 64;
    long long ksplit = ((long long)ctx->sm_count * 3 + mtiles - 1) / mtiles;
    if (ksplit > nchunks) ksplit = nchunks;
    if (ksplit < 1) ksplit = 1;
    a.chunks_per_cta = (int)((nchunks + ksplit - 1) / ksplit);
    ksplit = (nchunks + a.chunks_per_cta - 1) / a.chunks_per_cta;
    dim3 grid((unsigned)mtiles, (unsigned)ksplit);
    if (prec == FPL_PREC_BF16) { FPL_TRY(tc_set_smem<1>()); tc_wgrad_kernel<1><<<grid, kTcThreads, smem, st>>>(a); }
    else { FPL_TRY(tc_set_smem<3>()); tc_wgrad_kernel<3><<<grid, kTcThreads, smem, st>>>(a); }
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

// ---- slab kernels (48 | Cin): see train_tc.cuh
static int g_tc_gather = -1;     // FPL_TC_GATHER=1: A/B switch back to the per-tap gather kernels
static bool slab_layer_ok(int k, int cin, int cout) {
    if (g_tc_gather < 0) { const char *e = getenv("FPL_TC_GATHER"); g_tc_gather = (e && e[0] == '1') ? 1 : 0; }
    return !g_tc_gather && cin % 48 == 0 && cout % 48 == 0 && cin <= 96 && cout <= 96 && (k == 1 || k == 3);
}
static size_t slab_wimg_bytes(int k, int ca, int nn) { return (size_t)k * k * k * (ca / 48) * 2 * 6 * nn * 16; }

template <int NS>
static int slab_set_smem() {      // per launch: the attribute belongs to the current device
    FPL_CUDA_CHECK(cudaFuncSetAttribute(tc_slab_conv_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    FPL_CUDA_CHECK(cudaFuncSetAttribute(tc_slab_wgrad_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    return FPL_OK;
}

// valid convolution (forward: flip 0, dgrad on the zero-padded dx: flip 1); wimg: scratch for the packed weights
static int launch_slab_conv(fpl_ctx *ctx, int prec, const float *act, const float *w, void *wimg, float *out, int n,
                            int din, int k, int ca, int nn, int flip, cudaStream_t st) {
    const int ns = prec == FPL_PREC_BF16 ? 1 : 3, np = ns == 3 ? 2 : 1;
    const int ntaps = k * k * k;
    const int total = ntaps * (ca / 48) * 6 * nn;
    if (ns == 1) tc_pack_w_kernel<1><<<(total + 255) / 256, 256, 0, st>>>(w, (uint8_t *)wimg, ntaps, ca, nn, flip);
    else tc_pack_w_kernel<3><<<(total + 255) / 256, 256, 0, st>>>(w, (uint8_t *)wimg, ntaps, ca, nn, flip);
    FPL_LAUNCH_CHECK(ctx);
    SlabConvArgs a{};
    a.act = act; a.wimg = (const __nv_bfloat16 *)wimg; a.out = out;
    a.k = k; a.ca = ca; a.nn = nn;
    const int dout = din - (k - 1);
    if (k == 1) { a.flat = 1; a.n_img = 1; a.din = 1 << 14; a.dout = a.din; a.u_max = n * din * din * din; a.img_vox = a.u_max; }
    else { a.flat = 0; a.n_img = n; a.din = din; a.dout = dout; a.u_max = ((dout - 1) * din + (dout - 1)) * din + dout;
           a.img_vox = (long long)din * din * din; }
    if (k == 1) FPL_REQUIRE((long long)n * din * din * din < (1ll << 30), "slab conv: too many rows");
    const int halo = k == 1 ? 0 : (k - 1) * (din + 1);
    const size_t b_ring = (size_t)kSlabRing * np * 6 * nn * 16;
    size_t smem = 0;
    for (int mt = 2; mt >= 1; --mt) {
        int sp = mt * 128 + halo;
        sp += (9 - sp % 8) % 8;                                   // == 1 (mod 8): staging stores spread over the banks
        smem = (size_t)np * k * (ca / 8) * sp * 16 + b_ring;
        a.mt = mt; a.s_pad = sp;
        if (smem <= 227 * 1024 - 1024 && 2 * mt * np * nn <= 512) break;
    }
    FPL_REQUIRE(smem <= 227 * 1024 - 1024 && 2 * a.mt * np * nn <= 512, "slab conv: layer does not fit shared memory / TMEM");
    const int rows_pass = a.mt * 128;
    a.pairs_per_img = (a.u_max + rows_pass - 1) / rows_pass;
    const long long passes = (long long)a.n_img * a.pairs_per_img;
    const int grid = (int)(passes < ctx->sm_count ? passes : ctx->sm_count);
    if (ns == 1) { FPL_TRY(slab_set_smem<1>()); tc_slab_conv_kernel<1><<<grid, kSlabThreads, smem, st>>>(a); }
    else { FPL_TRY(slab_set_smem<3>()); tc_slab_conv_kernel<3><<<grid, kSlabThreads, smem, st>>>(a); }
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

// dW += in^T * dx; dxe: dx on the input grid (k == 3: zero-embedded copy; k == 1: dx itself)
static int launch_slab_wgrad(fpl_ctx *ctx, int prec, const float *in, const float *dxe, float *dw, int n, int din, int k,
                             int ci, int co, cudaStream_t st) {
    const int ns = prec == FPL_PREC_BF16 ? 1 : 3, np = ns == 3 ? 2 : 1;
    SlabWgradArgs a{};
    a.act = in; a.dxe = dxe; a.dw = dw; a.k = k; a.ci = ci; a.co = co;
    const int dout = din - (k - 1);
    long long u_max;
    if (k == 1) { a.n_img = 1; a.din = 1 << 14; a.dout = a.din; u_max = (long long)n * din * din * din; a.img_vox = u_max; }
    else { a.n_img = n; a.din = din; a.dout = dout; u_max = ((long long)(dout - 1) * din + (dout - 1)) * din + dout;
           a.img_vox = (long long)din * din * din; }
    FPL_REQUIRE(u_max < (1ll << 30), "slab wgrad: too many voxels");
    const int halo = k == 1 ? 0 : (k - 1) * (din + 1);
    int sp = kWgKc + halo;
    sp += (9 - sp % 8) % 8;
    a.s_pad = sp;
    a.chunks_per_img = (int)((u_max + kWgKc - 1) / kWgKc);
    const size_t pa = (size_t)kWgAPad * 16, a_img = (size_t)(co / 8) * pa, x_img = (size_t)(ci / 8) * sp * 16;
    const size_t stage = np * (a_img + x_img);
    const size_t over = (np - 1) * a_img + 16 * pa;               // reach of the M = 128 read from the last A image
    size_t smem = 2 * stage + (over > stage ? over - stage : 0);
    smem = (smem + 15) & ~(size_t)15;
    FPL_REQUIRE(smem <= 227 * 1024 - 1024, "slab wgrad: layer does not fit shared memory");
    FPL_REQUIRE(k * k * ci <= 512, "slab wgrad: accumulators do not fit TMEM");
    a.smem_bytes = (uint32_t)smem;
    const long long n_units = (long long)a.n_img * a.chunks_per_img;
    long long per_kd = ctx->sm_count / k;
    if (per_kd < 1) per_kd = 1;
    if (per_kd > n_units) per_kd = n_units;
    a.units_per_cta = (int)((n_units + per_kd - 1) / per_kd);
    per_kd = (n_units + a.units_per_cta - 1) / a.units_per_cta;
    dim3 grid((unsigned)k, (unsigned)per_kd);
    if (ns == 1) { FPL_TRY(slab_set_smem<1>()); tc_slab_wgrad_kernel<1><<<grid, kWgThreads, smem, st>>>(a); }
    else { FPL_TRY(slab_set_smem<3>()); tc_slab_wgrad_kernel<3><<<grid, kWgThreads, smem, st>>>(a); }
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
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
    size_t pad_elems = 16;
    for (size_t li = 1; li < t->layers.size(); ++li) {
        const Layer &L = t->layers[li];
        const size_t dp = (size_t)L.dout + 2 * (L.k - 1);
        if (L.k > 1 && (size_t)batch * dp * dp * dp * L.cout > pad_elems) pad_elems = (size_t)batch * dp * dp * dp * L.cout;
    }
    t->gpad = (float *)alloc(pad_elems * 4);
    size_t emb_elems = 16, wimg_bytes = 16;
    for (size_t li = 0; li < t->layers.size(); ++li) {
        const Layer &L = t->layers[li];
        if (L.cin % 48 || L.cout % 48) continue;
        if (L.k > 1 && (size_t)batch * L.din * L.din * L.din * L.cout > emb_elems) emb_elems = (size_t)batch * L.din * L.din * L.din * L.cout;
        if (slab_wimg_bytes(L.k, L.cin, L.cout) > wimg_bytes) wimg_bytes = slab_wimg_bytes(L.k, L.cin, L.cout);
        if (slab_wimg_bytes(L.k, L.cout, L.cin) > wimg_bytes) wimg_bytes = slab_wimg_bytes(L.k, L.cout, L.cin);
    }
    t->gemb = (float *)alloc(emb_elems * 4);
    t->wimg = alloc(wimg_bytes);
    mem_ok = mem_ok && t->gpad && t->gemb && t->wimg;
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

int fpl_train_set_precision(fpl_trainer *t, int precision) {
    FPL_REQUIRE(t, "fpl_train_set_precision: NULL trainer");
    FPL_REQUIRE(precision == FPL_PREC_FP32 || precision == FPL_PREC_BF16 || precision == FPL_PREC_TF32,
                "fpl_train_set_precision: unknown precision %d", precision);
    t->precision = precision;
    return FPL_OK;
}

// test hooks (not part of include/fpl_b200.h): the three tensor-core contractions on caller-provided device tensors
// what = 0: forward conv (act (n,din^3,cin), w (k^3,cin,cout) -> out (n,dout^3,cout))
// what = 1: dgrad       (act = dx (n,dout^3,cout), w -> out (n,din^3,cin)); scratch >= n*(dout+2(k-1))^3*cout floats
// what = 2: wgrad       (act = layer input (n,din^3,cin), w = dx (n,dout^3,cout) -> out (k^3,cin,cout), overwritten)
int fpl_debug_train_tc(fpl_ctx *ctx, int what, int precision, const float *act, const float *w, float *out,
                       float *scratch, int n, int din, int k, int cin, int cout, void *stream) {
    FPL_REQUIRE(ctx && act && w && out, "fpl_debug_train_tc: NULL argument");
    FPL_REQUIRE(tc_layer_ok(k, cin, cout), "fpl_debug_train_tc: layer shape not covered by the tensor-core kernels");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int dout = din - (k - 1);
    const bool slab = slab_layer_ok(k, cin, cout);
    const int p = k - 1, dp = dout + 2 * p;
    FPL_REQUIRE(what != 1 || k == 1 || scratch, "fpl_debug_train_tc: dgrad needs the scratch buffer");
    void *wimg = nullptr; float *emb = nullptr;
    if (slab) {
        const size_t wb = slab_wimg_bytes(k, cin, cout) > slab_wimg_bytes(k, cout, cin) ? slab_wimg_bytes(k, cin, cout) : slab_wimg_bytes(k, cout, cin);
        FPL_CUDA_CHECK(cudaMalloc(&wimg, wb));
        FPL_CUDA_CHECK(cudaMalloc((void **)&emb, (size_t)n * din * din * din * cout * sizeof(float)));
    }
    int rc = FPL_OK;
    if (what == 0) {
        rc = slab ? launch_slab_conv(ctx, precision, act, w, wimg, out, n, din, k, cin, cout, 0, st)
                  : launch_tc_conv(ctx, precision, act, w, out, n, din, k, cin, cout, 0, st);
    } else if (what == 1) {
        const float *dxp = act;
        if (k > 1) {
            tc_embed_kernel<<<blocks_for(ctx, (long long)n * dp * dp * dp * (cout / 4)), 256, 0, st>>>(
                (const float4 *)act, (float4 *)scratch, n, dout, p, dp, cout / 4);
            ctx->launches++;
            dxp = scratch;
        }
        rc = slab ? launch_slab_conv(ctx, precision, dxp, w, wimg, out, n, dp, k, cout, cin, 1, st)
                  : launch_tc_conv(ctx, precision, dxp, w, out, n, dp, k, cout, cin, 1, st);
    } else {
        cudaMemsetAsync(out, 0, (size_t)k * k * k * cin * cout * sizeof(float), st);
        if (slab) {
            const float *dxe = w;
            if (k > 1) {
                tc_embed_kernel<<<blocks_for(ctx, (long long)n * din * din * din * (cout / 4)), 256, 0, st>>>(
                    (const float4 *)w, (float4 *)emb, n, dout, 0, din, cout / 4);
                ctx->launches++;
                dxe = emb;
            }
            rc = launch_slab_wgrad(ctx, precision, act, dxe, out, n, din, k, cin, cout, st);
        } else {
            rc = launch_tc_wgrad(ctx, precision, act, w, out, n, din, k, cin, cout, st);
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (wimg) cudaFree(wimg);
    if (emb) cudaFree(emb);
    if (rc == FPL_OK && e != cudaSuccess) { fpl::set_error("fpl_debug_train_tc: %s", cudaGetErrorString(e)); rc = FPL_ECUDA; }
    return rc;
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
        const bool tc = t->precision != FPL_PREC_FP32 && tc_layer_ok(L.k, L.cin, L.cout);
        if (tc && slab_layer_ok(L.k, L.cin, L.cout))
            FPL_TRY(launch_slab_conv(ctx, t->precision, cur, d_params + L.off_kernel, t->wimg, L.x, n, L.din, L.k, L.cin, L.cout, 0, st));
        else if (tc) FPL_TRY(launch_tc_conv(ctx, t->precision, cur, d_params + L.off_kernel, L.x, n, L.din, L.k, L.cin, L.cout, 0, st));
        else if (L.k == 3) conv_fwd_kernel<3><<<blocks_for(ctx, total / 8), 256, 0, st>>>(cur, d_params + L.off_kernel, L.x, n, L.din, L.cin, L.cout);
        else conv_fwd_kernel<1><<<blocks_for(ctx, total / 8), 256, 0, st>>>(cur, d_params + L.off_kernel, L.x, n, L.din, L.cin, L.cout);
        if (!tc) FPL_LAUNCH_CHECK(ctx);
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
        const bool tc = t->precision != FPL_PREC_FP32 && tc_layer_ok(L.k, L.cin, L.cout);
        if (tc) {
            const bool slab = slab_layer_ok(L.k, L.cin, L.cout);
            const int p = L.k - 1;
            if (slab) {
                const float *dxe = gb;                   // dx on the input grid (zeros where there is no output)
                if (L.k > 1) {
                    const long long te = (long long)n * L.din * L.din * L.din * (L.cout / 4);
                    tc_embed_kernel<<<blocks_for(ctx, te), 256, 0, st>>>((const float4 *)gb, (float4 *)t->gemb, n, L.dout, 0, L.din, L.cout / 4);
                    FPL_LAUNCH_CHECK(ctx);
                    dxe = t->gemb;
                }
                FPL_TRY(launch_slab_wgrad(ctx, t->precision, lin, dxe, d_grads + L.off_kernel, n, L.din, L.k, L.cin, L.cout, st));
            } else {
                FPL_TRY(launch_tc_wgrad(ctx, t->precision, lin, gb, d_grads + L.off_kernel, n, L.din, L.k, L.cin, L.cout, st));
            }
            if (li > 0) {
                // dgrad = valid convolution of the zero-padded dx with the flipped kernel (channels swapped)
                const float *dxp = gb;
                if (L.k > 1) {
                    const int dp = L.dout + 2 * p;
                    const long long tp = (long long)n * dp * dp * dp * (L.cout / 4);
                    tc_embed_kernel<<<blocks_for(ctx, tp), 256, 0, st>>>((const float4 *)gb, (float4 *)t->gpad, n, L.dout, p, dp, L.cout / 4);
                    FPL_LAUNCH_CHECK(ctx);
                    dxp = t->gpad;
                }
                if (slab) FPL_TRY(launch_slab_conv(ctx, t->precision, dxp, d_params + L.off_kernel, t->wimg, ga, n, L.dout + 2 * p, L.k,
                                                   L.cout, L.cin, 1, st));
                else FPL_TRY(launch_tc_conv(ctx, t->precision, dxp, d_params + L.off_kernel, ga, n, L.dout + 2 * p, L.k, L.cout, L.cin, 1, st));
            }
        } else {
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
            }
        }
        // ga = gradient w.r.t. this layer's input = previous block's (pooled / dropped-out) output
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
