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

// Register/shared-memory tiled variant for Cin % 16 == 0 (every layer but the first): a block computes
// 64 x-voxels x R rows x all Cout; a thread owns 2 voxels (x, x+32) x 16 output channels.  The weights of one
// (tap, Cin chunk) are staged in shared memory once per block and read as broadcast float4; the activations are
// read as float4 over 4 input channels.  Same FMA order per output as conv_fp32_kernel (taps outer, channels
// inner) -> bit-identical results, ~2 loads per 16 FMAs instead of 17.
template <int K>
__global__ void __launch_bounds__(512)
conv_fp32_tiled_kernel(const float *__restrict__ in, const float *__restrict__ w,
                       const float *__restrict__ scale, const float *__restrict__ bias,
                       float *__restrict__ out, int n_tiles, int din, int cin, int cout, int relu, int chunk) {
    __shared__ __align__(16) float sw[32 * 128];
    const int dout = din - (K - 1);
    const int R = blockDim.z;
    const int xb = (dout + 63) / 64, yb = (dout + R - 1) / R;
    long long bid = blockIdx.x;
    const int x0 = (int)(bid % xb) * 64 + threadIdx.x; bid /= xb;
    const int y = (int)(bid % yb) * R + threadIdx.z; bid /= yb;
    const int z = (int)(bid % dout); bid /= dout;
    const int t = (int)bid;
    const int co0 = threadIdx.y * 16;
    const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * 32 + threadIdx.x, nthr = blockDim.x * blockDim.y * blockDim.z;
    const bool yok = y < dout;
    const bool ok0 = yok && x0 < dout, ok1 = yok && x0 + 32 < dout;
    // clamped coordinates keep every load in bounds; results of invalid voxels are simply not stored
    const int yc = yok ? y : dout - 1, xa = x0 < dout ? x0 : dout - 1, xb1 = x0 + 32 < dout ? x0 + 32 : dout - 1;
    float acc0[16], acc1[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
    const float *tin = in + (size_t)t * din * din * din * cin;
    for (int kd = 0; kd < K; ++kd)
        for (int kh = 0; kh < K; ++kh)
            for (int kw = 0; kw < K; ++kw) {
                const size_t rowbase = ((size_t)(z + kd) * din + (yc + kh)) * din;
                const float *ip0 = tin + (rowbase + xa + kw) * cin;
                const float *ip1 = tin + (rowbase + xb1 + kw) * cin;
                const float *wt = w + (size_t)((kd * K + kh) * K + kw) * cin * cout;
                for (int c0 = 0; c0 < cin; c0 += chunk) {
                    __syncthreads();
                    for (int i = tid * 4; i < chunk * cout; i += nthr * 4)
                        *reinterpret_cast<float4 *>(&sw[i]) = __ldg(reinterpret_cast<const float4 *>(wt + (size_t)c0 * cout + i));
                    __syncthreads();
#pragma unroll 2
                    for (int ci = 0; ci < chunk; ci += 4) {
                        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(ip0 + c0 + ci));
                        const float4 a1 = __ldg(reinterpret_cast<const float4 *>(ip1 + c0 + ci));
                        const float av0[4] = {a0.x, a0.y, a0.z, a0.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 *wr = reinterpret_cast<const float4 *>(&sw[(ci + q) * cout + co0]);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 wv = wr[j4];
                                acc0[4 * j4 + 0] = fmaf(av0[q], wv.x, acc0[4 * j4 + 0]); acc1[4 * j4 + 0] = fmaf(av1[q], wv.x, acc1[4 * j4 + 0]);
                                acc0[4 * j4 + 1] = fmaf(av0[q], wv.y, acc0[4 * j4 + 1]); acc1[4 * j4 + 1] = fmaf(av1[q], wv.y, acc1[4 * j4 + 1]);
                                acc0[4 * j4 + 2] = fmaf(av0[q], wv.z, acc0[4 * j4 + 2]); acc1[4 * j4 + 2] = fmaf(av1[q], wv.z, acc1[4 * j4 + 2]);
                                acc0[4 * j4 + 3] = fmaf(av0[q], wv.w, acc0[4 * j4 + 3]); acc1[4 * j4 + 3] = fmaf(av1[q], wv.w, acc1[4 * j4 + 3]);
                            }
                        }
                    }
                }
            }
    const size_t obase = (size_t)t * dout * dout * dout + ((size_t)z * dout + yc) * dout;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float sc = scale[co0 + j], bi = bias[co0 + j];
        float v0 = fmaf(acc0[j], sc, bi), v1 = fmaf(acc1[j], sc, bi);
        if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        if (ok0) out[(obase + x0) * cout + co0 + j] = v0;
        if (ok1) out[(obase + x0 + 32) * cout + co0 + j] = v1;
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
    if (c.cin % 16 == 0 && c.cout % 16 == 0 && c.cout <= 128 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(c.d_kernel) & 15) == 0) {
        const int g = c.cout / 16;
        int rows = 512 / (32 * g); if (rows > 4) rows = 4; if (rows < 1) rows = 1;
        const long long tb = (long long)n_tiles * dout * ((dout + rows - 1) / rows) * ((dout + 63) / 64);
        FPL_REQUIRE(tb < 2147483647LL, "conv_fp32: grid too large");
        conv_fp32_tiled_kernel<K><<<(unsigned)tb, dim3(32, g, rows), 0, st>>>(in, c.d_kernel, c.d_scale, c.d_bias, out, n_tiles,
                                                                           din, c.cin, c.cout, relu, c.cin % 32 == 0 ? 32 : 16);
    } else if (c.cout % 8 == 0 && c.cout / 8 <= 8) {
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

// out[i] = relu(cur[i] + skip[cropped index]) in place on cur: (d^3, c) channels-last float32, skip (ds^3, c)
__global__ void __launch_bounds__(256)
add_fp32_kernel(float *__restrict__ cur, int d, int c, const float *__restrict__ skip, int ds, int crop) {
    const long long total = (long long)d * d * d * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c); long long v = i / c;
        const int x = (int)(v % d); v /= d;
        const int y = (int)(v % d); const int z = (int)(v / d);
        const float s = skip[((((size_t)(z + crop) * ds + y + crop) * ds) + x + crop) * c + ch];
        cur[i] = fmaxf(cur[i] + s, 0.f);
    }
}

int forward_fp32(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out, cudaStream_t st) {
    fpl_ctx *ctx = net->ctx;
    // activations of one tile at a time (the fp32 tensors are large: 98^3 x 48 x 4 B = 181 MB)
    size_t max_elems = 0;
    {
        int d = in_sz, c = 1;
        int sc[4] = {0, 0, 0, 0};
        for (const Op &o : net->ops) {
            if (o.kind == OP_CONV && o.src_slot >= 0) { sc[o.src_slot] = o.cout; continue; }
            if (o.kind == OP_CONV) { d -= o.k - 1; c = o.cout; }
            else if (o.kind == OP_POOL) d /= 2;
            else if (o.kind == OP_SAVE) sc[o.slot] = c;
            else if (o.kind == OP_UPCAT) { d *= 2; c += sc[o.slot]; }
            size_t e = (size_t)d * d * d * c;
            if (e > max_elems) max_elems = e;
        }
    }
    // buffers: ping, pong, two skip slots, one for a convolved skip (resnet_like shortcut)
    constexpr int kBufs = 5;
    float *bufs[kBufs] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < kBufs; ++i) {
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
            if (o.kind == OP_CONV && o.src_slot >= 0) {
                const ConvParams &cp = net->convs[o.conv_index];
                rc = launch_conv<1>(ctx, skip_ptr[o.src_slot], cp, bufs[4], 1, skip_d[o.src_slot], o.relu ? 1 : 0, st);
                if (rc != FPL_OK) break;
                skip_ptr[o.src_slot] = bufs[4]; skip_c[o.src_slot] = o.cout;
            } else if (o.kind == OP_ADD) {
                add_fp32_kernel<<<stream_blocks, 256, 0, st>>>(const_cast<float *>(cur), d, c, skip_ptr[o.slot], skip_d[o.slot], o.crop);
                ctx->launches++;
            } else if (o.kind == OP_CONV) {
                const ConvParams &cp = net->convs[o.conv_index];
                float *dst = bufs[which];
                if (cur == dst) { which ^= 1; dst = bufs[which]; }
                rc = (o.k == 3) ? launch_conv<3>(ctx, cur, cp, dst, 1, d, o.relu ? 1 : 0, st)
                                : launch_conv<1>(ctx, cur, cp, dst, 1, d, o.relu ? 1 : 0, st);
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
    for (int i = 0; i < kBufs; ++i) cudaFree(bufs[i]);
    if (rc != FPL_OK) return rc;
    FPL_CUDA_CHECK(e);
    FPL_CUDA_CHECK(cudaGetLastError());
    return FPL_OK;
}

}  // namespace net
}  // namespace fpl
