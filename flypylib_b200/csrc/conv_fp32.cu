// conv_fp32.cu -- fp32 CUDA-core execution of the network graph (FPL_PREC_FP32).
//
// This is the high-precision validation path (prob-map parity <= 2e-3 against the float64
// restatement of the Keras graph); the throughput path is the tcgen05 implicit GEMM in conv_umma.cu.
// Layout: activations (tile, z, y, x, C) float32, channels innermost (Keras channels_last).
#include "net.cuh"

namespace fpl {
namespace net {

// Direct convolution, K in {1,3}, 'valid'.  One thread = one output voxel x CPT output channels.
// blockDim = (32 voxels along x, Cout/CPT channel groups).
template <int K, int CPT>
__global__ void __launch_bounds__(256)
conv_fp32_kernel(const float *__restrict__ in, const float *__restrict__ w,
                 const float *__restrict__ scale, const float *__restrict__ bias,
                 float *__restrict__ out, int n_tiles, int din, int cin, int cout, int relu) {
    const int dout = din - (K - 1);
    const int xb = (dout + 31) / 32;
    long long bid = blockIdx.x;
    const int x = (int)(bid % xb) * 32 + threadIdx.x; bid /= xb;
    const int y = (int)(bid % dout); bid /= dout;
    const int z = (int)(bid % dout); bid /= dout;
    const int t = (int)bid;
    const int co0 = threadIdx.y * CPT;
    if (x >= dout) return;
    float acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
    const float *tin = in + (size_t)t * din * din * din * cin;
    for (int kd = 0; kd < K; ++kd)
        for (int kh = 0; kh < K; ++kh)
            for (int kw = 0; kw < K; ++kw) {
                const float *ip = tin + ((size_t)((z + kd) * din + (y + kh)) * din + (x + kw)) * cin;
                const float *wp = w + (size_t)((kd * K + kh) * K + kw) * cin * cout + co0;
                for (int ci = 0; ci < cin; ++ci) {
                    const float a = __ldg(ip + ci);
#pragma unroll
                    for (int j = 0; j < CPT; ++j) acc[j] = fmaf(a, __ldg(wp + (size_t)ci * cout + j), acc[j]);
                }
            }
    float *op = out + ((size_t)t * dout * dout * dout + ((size_t)z * dout + y) * dout + x) * cout + co0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        float v = fmaf(acc[j], scale[co0 + j], bias[co0 + j]);
        if (relu) v = fmaxf(v, 0.f);
        op[j] = v;
    }
}

// MaxPooling3D((2,2,2)), floor.
__global__ void __launch_bounds__(256)
pool_fp32_kernel(const float *__restrict__ in, float *__restrict__ out, int n_tiles, int din, int c) {
    const int dout = din / 2;
    const long long total = (long long)n_tiles * dout * dout * dout * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % c); long long v = i / c;
        int x = (int)(v % dout); v /= dout;
        int y = (int)(v % dout); v /= dout;
        int z = (int)(v % dout); int t = (int)(v / dout);
        const float *ip = in + (size_t)t * din * din * din * c;
        float m = -INFINITY;
        for (int dz = 0; dz < 2; ++dz)
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx)
                    m = fmaxf(m, ip[((size_t)((2 * z + dz) * din + (2 * y + dy)) * din + (2 * x + dx)) * c + ch]);
        out[i] = m;
    }
}

// concatenate([UpSampling3D(2)(a), Cropping3D(crop)(skip)], axis=-1)
__global__ void __launch_bounds__(256)
upcat_fp32_kernel(const float *__restrict__ a, int da, int ca, const float *__restrict__ skip, int ds,
                  int cs, int crop, float *__restrict__ out, int n_tiles) {
    const int dout = 2 * da, c = ca + cs;
    const long long total = (long long)n_tiles * dout * dout * dout * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % c); long long v = i / c;
        int x = (int)(v % dout); v /= dout;
        int y = (int)(v % dout); v /= dout;
        int z = (int)(v % dout); int t = (int)(v / dout);
        float r;
        if (ch < ca)
            r = a[((size_t)t * da * da * da + ((size_t)(z / 2) * da + (y / 2)) * da + (x / 2)) * ca + ch];
        else
            r = skip[((size_t)t * ds * ds * ds + ((size_t)(z + crop) * ds + (y + crop)) * ds + (x + crop)) * cs +
                     (ch - ca)];
        out[i] = r;
    }
}

// final Conv3D(1,(1,1,1),activation='sigmoid') + UpSampling3D(stride) of fplnetwork.py:99-105
__global__ void __launch_bounds__(256)
final_fp32_kernel(const float *__restrict__ in, const float *__restrict__ w, float bias,
                  float *__restrict__ out, int n_tiles, int d, int c, int stride) {
    const long long total = (long long)n_tiles * d * d * d;
    const int dout = d * stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const float *ip = in + (size_t)i * c;
        float acc = 0.f;
        for (int ch = 0; ch < c; ++ch) acc = fmaf(ip[ch], __ldg(w + ch), acc);
        acc += bias;
        float p = 1.f / (1.f + expf(-acc));
        long long v = i;
        int x = (int)(v % d); v /= d;
        int y = (int)(v % d); v /= d;
        int z = (int)(v % d); int t = (int)(v / d);
        float *op = out + (size_t)t * dout * dout * dout;
        for (int dz = 0; dz < stride; ++dz)
            for (int dy = 0; dy < stride; ++dy)
                for (int dx = 0; dx < stride; ++dx)
                    op[((size_t)(z * stride + dz) * dout + (y * stride + dy)) * dout + (x * stride + dx)] = p;
    }
}

template <int K>
static int launch_conv(fpl_ctx *ctx, const float *in, const ConvParams &c, float *out, int n_tiles, int din,
                       int relu, cudaStream_t st) {
    const int dout = din - (K - 1);
    const int xb = (dout + 31) / 32;
    const long long blocks = (long long)n_tiles * dout * dout * xb;
    FPL_REQUIRE(blocks < 2147483647LL, "conv_fp32: grid too large");
    if (c.cout % 8 == 0 && c.cout / 8 <= 8) {
        dim3 block(32, c.cout / 8);
        conv_fp32_kernel<K, 8><<<(unsigned)blocks, block, 0, st>>>(in, c.d_kernel, c.d_scale, c.d_bias, out,
                                                                  n_tiles, din, c.cin, c.cout, relu);
    } else if (c.cout % 16 == 0 && c.cout / 16 <= 8) {
        dim3 block(32, c.cout / 16);
        conv_fp32_kernel<K, 16><<<(unsigned)blocks, block, 0, st>>>(in, c.d_kernel, c.d_scale, c.d_bias, out,
                                                                   n_tiles, din, c.cin, c.cout, relu);
    } else {
        set_error("conv_fp32: unsupported Cout %d", c.cout);
        return FPL_EINVAL;
    }
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

int forward_fp32(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out, cudaStream_t st) {
    fpl_ctx *ctx = net->ctx;
    // activations of one tile at a time (the fp32 tensors are large: 98^3 x 48 x 4 B = 181 MB)
    size_t max_elems = 0;
    {
        int d = in_sz, c = 1;
        int sc[4] = {0, 0, 0, 0};
        for (const Op &o : net->ops) {
            if (o.kind == OP_CONV) { d -= o.k - 1; c = o.cout; }
            else if (o.kind == OP_POOL) d /= 2;
            else if (o.kind == OP_SAVE) sc[o.slot] = c;
            else if (o.kind == OP_UPCAT) { d *= 2; c += sc[o.slot]; }
            size_t e = (size_t)d * d * d * c;
            if (e > max_elems) max_elems = e;
        }
    }
    // buffers: ping, pong, two skip slots
    float *bufs[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 4; ++i) {
        cudaError_t e = cudaMalloc((void **)&bufs[i], max_elems * sizeof(float));
        if (e != cudaSuccess) {
            for (int j = 0; j < i; ++j) cudaFree(bufs[j]);
            set_error("forward_fp32: activation buffer allocation failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return FPL_ENOMEM;
        }
    }
    const int stream_blocks = ctx->sm_count * 8;
    int rc = FPL_OK;
    const int out_edge = out_size(net, in_sz);
    for (int t = 0; t < n_tiles && rc == FPL_OK; ++t) {
        const float *cur = d_tiles + (size_t)t * in_sz * in_sz * in_sz;
        int d = in_sz, c = 1;
        int which = 0;                         // next ping/pong target
        const float *skip_ptr[2] = {nullptr, nullptr};
        int skip_d[2] = {0, 0}, skip_c[2] = {0, 0};
        int skip_buf_used = 0;
        for (const Op &o : net->ops) {
            if (o.kind == OP_CONV) {
                const ConvParams &cp = net->convs[o.conv_index];
                float *dst = bufs[which];
                if (cur == dst) { which ^= 1; dst = bufs[which]; }
                rc = (o.k == 3) ? launch_conv<3>(ctx, cur, cp, dst, 1, d, 1, st)
                                : launch_conv<1>(ctx, cur, cp, dst, 1, d, 1, st);
                if (rc != FPL_OK) break;
                d -= o.k - 1; c = o.cout; cur = dst; which ^= 1;
            } else if (o.kind == OP_POOL) {
                float *dst = bufs[which];
                if (cur == dst) { which ^= 1; dst = bufs[which]; }
                pool_fp32_kernel<<<stream_blocks, 256, 0, st>>>(cur, dst, 1, d, c);
                ctx->launches++;
                d /= 2; cur = dst; which ^= 1;
            } else if (o.kind == OP_SAVE) {
                float *dst = bufs[2 + skip_buf_used++];
                cudaMemcpyAsync(dst, cur, (size_t)d * d * d * c * sizeof(float), cudaMemcpyDeviceToDevice, st);
                skip_ptr[o.slot] = dst; skip_d[o.slot] = d; skip_c[o.slot] = c;
            } else if (o.kind == OP_UPCAT) {
                float *dst = bufs[which];
                if (cur == dst) { which ^= 1; dst = bufs[which]; }
                upcat_fp32_kernel<<<stream_blocks, 256, 0, st>>>(cur, d, c, skip_ptr[o.slot], skip_d[o.slot],
                                                                skip_c[o.slot], o.crop, dst, 1);
                ctx->launches++;
                d *= 2; c += skip_c[o.slot]; cur = dst; which ^= 1;
            } else if (o.kind == OP_FINAL) {
                const ConvParams &cp = net->convs[o.conv_index];
                final_fp32_kernel<<<stream_blocks, 256, 0, st>>>(
                    cur, cp.d_kernel, cp.bias[0], d_out + (size_t)t * out_edge * out_edge * out_edge, 1, d, c,
                    net->info.rf_stride);
                ctx->launches++;
            }
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    for (int i = 0; i < 4; ++i) cudaFree(bufs[i]);
    if (rc != FPL_OK) return rc;
    FPL_CUDA_CHECK(e);
    FPL_CUDA_CHECK(cudaGetLastError());
    return FPL_OK;
}

}  // namespace net
}  // namespace fpl
