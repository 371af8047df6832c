// net.cu -- graph construction, weight handling, volume tiler and the fpl_net_* C ABI.
//
// Replaces, for the inference hot path:
//   flypylib/fplmodels.py:102-136,138-172,258-304   (graph definitions)
//   flypylib/fplnetwork.py:99-110                    (_set_infer: up-sampling for strided nets, weights)
//   flypylib/fplnetwork.py:136-189                   (infer: tile grid, zero padding, predict, scatter)
#include "net.cuh"
#include <math.h>

namespace fpl {
namespace net {

static Op conv(int k, int cin, int cout) { Op o; o.kind = OP_CONV; o.k = k; o.cin = cin; o.cout = cout; return o; }
static Op pool() { Op o; o.kind = OP_POOL; return o; }
static Op save(int slot) { Op o; o.kind = OP_SAVE; o.slot = slot; return o; }
static Op upcat(int slot, int crop) { Op o; o.kind = OP_UPCAT; o.slot = slot; o.crop = crop; return o; }
static Op conv_no_bn(int k, int cin, int cout) { Op o = conv(k, cin, cout); o.bn = false; return o; }
static Op conv_bn_only(int k, int cin, int cout) { Op o = conv(k, cin, cout); o.relu = false; return o; }
static Op conv_on_skip(int slot, int cin, int cout) { Op o = conv(1, cin, cout); o.relu = false; o.bn = false; o.src_slot = slot; return o; }
static Op add(int slot, int crop) { Op o; o.kind = OP_ADD; o.slot = slot; o.crop = crop; return o; }
static Op final_(int cin) { Op o; o.kind = OP_FINAL; o.k = 1; o.cin = cin; o.cout = 1; return o; }

static int build_graph(fpl_net *n) {
    std::vector<Op> &g = n->ops;
    switch (n->arch) {
        case FPL_ARCH_VGG_LIKE:      // fplmodels.py:102-136
            g = {conv(3, 1, 48), conv(1, 48, 48), pool(), conv(3, 48, 48), conv(1, 48, 48), pool(),
                 conv(3, 48, 48), conv(1, 48, 96), conv(1, 96, 96), final_(96)};
            n->info = {18, 7, 4, 102, true};
            break;
        case FPL_ARCH_VGG_LIKE2:     // fplmodels.py:138-172
            g = {conv(3, 1, 48), conv(3, 48, 48), pool(), conv(3, 48, 48), conv(3, 48, 48), pool(),
                 conv(3, 48, 48), conv(1, 48, 96), conv(1, 96, 96), final_(96)};
            n->info = {24, 10, 4, 100, true};
            break;
        case FPL_ARCH_UNET_LIKE2:    // fplmodels.py:258-304  (concat order: [up-sampled, skip])
            g = {conv(3, 1, 32), conv(3, 32, 32), save(0), pool(), conv(3, 32, 64), conv(3, 64, 64), save(1),
                 pool(), conv(1, 64, 128), upcat(1, 0), conv(3, 192, 64), conv(1, 64, 64), upcat(0, 6),
                 conv(3, 96, 32), conv(1, 32, 32), final_(32)};
            n->info = {24, 9, 1, 100, false};
            break;
        case FPL_ARCH_BASELINE:      // fplmodels.py:73-100
            g = {conv(3, 1, 32), pool(), conv(3, 32, 32), pool(), conv(3, 32, 32), conv(1, 32, 64), final_(64)};
            n->info = {18, 7, 4, 102, true};
            break;
        case FPL_ARCH_UNET_LIKE:     // fplmodels.py:206-256
            g = {conv(3, 1, 32), conv(1, 32, 32), save(0), pool(), conv(3, 32, 64), conv(1, 64, 64), save(1),
                 pool(), conv(1, 64, 128), upcat(1, 0), conv(3, 192, 64), conv(1, 64, 64), upcat(0, 4),
                 conv(3, 96, 32), conv(1, 32, 32), final_(32)};
            n->info = {18, 6, 1, 102, false};
            break;
        case FPL_ARCH_UNET_LIKE3:    // fplmodels.py:306-357
            g = {conv(3, 1, 32), conv(3, 32, 32), save(0), pool(), conv(3, 32, 64), conv(3, 64, 64), save(1),
                 pool(), conv(3, 64, 128), conv(1, 128, 128), upcat(1, 2), conv(3, 192, 64), conv(1, 64, 64),
                 upcat(0, 10), conv(3, 96, 32), conv(1, 32, 32), final_(32)};
            n->info = {32, 13, 1, 100, false};
            break;
        case FPL_ARCH_UNET_LIKE4:    // fplmodels.py:359-410
            g = {conv(3, 1, 32), conv(3, 32, 32), save(0), pool(), conv(3, 32, 64), conv(3, 64, 64), save(1),
                 pool(), conv(3, 64, 128), conv(3, 128, 128), upcat(1, 4), conv(3, 192, 64), conv(1, 64, 64),
                 upcat(0, 14), conv(3, 96, 32), conv(1, 32, 32), final_(32)};
            n->info = {40, 17, 1, 100, false};
            break;
        case FPL_ARCH_UNET_LIKE4B:   // fplmodels.py:412-467
            g = {conv(3, 1, 32), conv(3, 32, 32), save(0), pool(), conv(3, 32, 64), conv(1, 64, 32), conv(3, 32, 64),
                 save(1), pool(), conv(1, 64, 48), conv(3, 48, 128), conv(1, 128, 48), conv(3, 48, 128),
                 conv(1, 128, 48), upcat(1, 4), conv(3, 112, 64), conv(1, 64, 64), upcat(0, 14), conv(3, 96, 32),
                 conv(1, 32, 32), final_(32)};
            n->info = {40, 17, 1, 100, false};
            break;
        case FPL_ARCH_UNET_LIKE_VOL: // fplmodels.py:470-526: Conv3D(activation='relu', use_bias=False), no BatchNormalization
            g = {conv_no_bn(3, 1, 16), conv_no_bn(1, 16, 16), save(0), pool(), conv_no_bn(3, 16, 32), conv_no_bn(1, 32, 32),
                 save(1), pool(), conv_no_bn(1, 32, 64), upcat(1, 0), conv_no_bn(3, 96, 64), conv_no_bn(1, 64, 64),
                 upcat(0, 4), conv_no_bn(3, 80, 32), conv_no_bn(1, 32, 32), final_(32)};
            n->info = {62, 6, 1, 102, false};
            break;
        case FPL_ARCH_RESNET_LIKE:   // fplmodels.py:174-208 (weight order = Keras model.layers order: the shortcut convolution precedes conv3b)
            g = {conv(3, 1, 32), pool(), save(0), conv(3, 32, 32), conv_bn_only(1, 32, 32), add(0, 1), pool(), save(1),
                 conv(3, 32, 64), conv_on_skip(1, 32, 64), conv_bn_only(1, 64, 64), add(1, 1), final_(64)};
            n->info = {18, 7, 4, 102, true};
            break;
        default:
            set_error("fpl_net_create: unknown architecture %d", n->arch);
            return FPL_EINVAL;
    }
    int ci = 0;
    for (Op &o : g)
        if (o.kind == OP_CONV || o.kind == OP_FINAL) {
            o.conv_index = ci++;
            ConvParams p; p.k = o.k; p.cin = o.cin; p.cout = o.cout;
            n->convs.push_back(p);
        }
    return FPL_OK;
}

// Walk the graph on edge lengths only.  Returns the final edge (before up-sampling) or -1.
static int walk_sizes(const fpl_net *n, int in_sz, std::vector<int> *sizes_out = nullptr) {
    int d = in_sz;
    int skip_d[4] = {0, 0, 0, 0};
    for (const Op &o : n->ops) {
        switch (o.kind) {
            case OP_CONV: case OP_FINAL:
                if (o.kind == OP_CONV && o.src_slot >= 0) break;         // 1x1x1 convolution of a stored tensor
                d -= (o.k - 1); if (d <= 0) return -1; break;
            case OP_ADD: if (skip_d[o.slot] - 2 * o.crop != d) return -1; break;
            case OP_POOL: if (d % 2) return -1; d /= 2; if (d <= 0) return -1; break;
            case OP_SAVE: skip_d[o.slot] = d; break;
            case OP_UPCAT: d *= 2; if (skip_d[o.slot] - 2 * o.crop != d) return -1; break;
        }
        if (sizes_out) sizes_out->push_back(d);
    }
    return d;
}

int out_size(const fpl_net *n, int in_sz) {
    int d = walk_sizes(n, in_sz);
    return d < 0 ? -1 : d * n->info.rf_stride;
}

static void free_device_weights(fpl_net *n) {
    for (ConvParams &c : n->convs) {
        if (c.d_kernel) cudaFree(c.d_kernel);
        if (c.d_scale) cudaFree(c.d_scale);
        if (c.d_bias) cudaFree(c.d_bias);
        c.d_kernel = c.d_scale = c.d_bias = nullptr;
    }
    free_packed_umma(n);
}

// ------------------------------------------------------------------------------------------------
// tiler kernels
// ------------------------------------------------------------------------------------------------
// tile t of the batch <- image[start : start+in_sz] (zero beyond the far edge), start = k*out_sz.
// image either float32 or uint8 + (x-mean)/std in float32 (fplobjdetect.py:1106-1107).
// One block per (tile, z, y) row, threads along x: no per-element integer division.
template <typename T>
__global__ void __launch_bounds__(128)
gather_tiles_kernel(const T *__restrict__ img, float *__restrict__ tiles, TileGrid g, int tile0,
                    int n_tiles, float mean, float stdv, const int *__restrict__ tile_ids) {
    const long long rows = (long long)n_tiles * g.in_z * g.in_sz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int y = (int)(row % g.in_sz);
        const int z = (int)((row / g.in_sz) % g.in_z);
        const int t = (int)(row / ((long long)g.in_sz * g.in_z));
        const int tt = tile_ids ? tile_ids[tile0 + t] : tile0 + t;
        const int kx = tt % g.nx, ky = (tt / g.nx) % g.ny, kz = tt / (g.nx * g.ny);
        const long long gz = g.z_base + (long long)kz * g.out_z + z, gy = (long long)ky * g.out_sz + y;
        const long long gx0 = (long long)kx * g.out_sz;
        float *dst = tiles + row * g.in_sz;
        const bool row_ok = gz < g.Z && gy < g.Y;
        const T *src = img + (gz * g.Y + gy) * g.X + gx0;
        for (int x = threadIdx.x; x < g.in_sz; x += blockDim.x) {
            float v = 0.f;
            if (row_ok && gx0 + x < g.X) {
                T raw = src[x];
                if (sizeof(T) == 1) v = __fdiv_rn(__fsub_rn((float)raw, mean), stdv);
                else v = (float)raw;
            }
            dst[x] = v;
        }
    }
}

// pred[off + k*out_sz + (0..ext)] <- tile output[0..ext), ext = min(out_sz, size - off - origin)
__global__ void __launch_bounds__(128)
scatter_tiles_kernel(const float *__restrict__ outs, float *__restrict__ pred, TileGrid g, int tile0,
                     int n_tiles, const int *__restrict__ tile_ids) {
    const long long rows = (long long)n_tiles * g.out_z * g.out_sz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int y = (int)(row % g.out_sz);
        const int z = (int)((row / g.out_sz) % g.out_z);
        const int t = (int)(row / ((long long)g.out_sz * g.out_z));
        const int tt = tile_ids ? tile_ids[tile0 + t] : tile0 + t;
        const int kx = tt % g.nx, ky = (tt / g.nx) % g.ny, kz = tt / (g.nx * g.ny);
        const long long gz = g.z_base + (long long)kz * g.out_z + g.off + z, gy = (long long)ky * g.out_sz + g.off + y;
        const long long gx0 = (long long)kx * g.out_sz + g.off;
        if (gz >= g.Z - g.off || gy >= g.Y - g.off) continue;
        const float *src = outs + row * g.out_sz;
        float *dst = pred + (gz * g.Y + gy) * g.X + gx0;
        for (int x = threadIdx.x; x < g.out_sz; x += blockDim.x)
            if (gx0 + x < g.X - g.off) dst[x] = src[x];
    }
}

static int tiles_along(long long size, int off, int out_sz) {
    // np.mgrid[off : size-off : out_sz]  -> ceil((size - 2 off) / out_sz) origins, 0 if empty
    long long span = size - 2LL * off;
    if (span <= 0) return 0;
    return (int)((span + out_sz - 1) / out_sz);
}

}  // namespace net
}  // namespace fpl

using namespace fpl::net;

static int g_no_slab_mode = 0;   // test hook: keep cubic super-tiles on the tcgen05 path
static int g_no_direct_io = 0;   // test hook: stage tiles through gather/scatter kernels on the tcgen05 path too

extern "C" {

int fpl_debug_no_direct_io(int on) { g_no_direct_io = on; return FPL_OK; }
int fpl_debug_no_slab_mode(int on) { g_no_slab_mode = on; return FPL_OK; }


int fpl_net_create(fpl_ctx *ctx, int arch, fpl_net **out) {
    FPL_REQUIRE(ctx && out, "fpl_net_create: NULL argument");
    fpl_net *n = new fpl_net();
    n->ctx = ctx;
    n->arch = arch;
    int rc = build_graph(n);
    if (rc != FPL_OK) { delete n; return rc; }
    *out = n;
    return FPL_OK;
}

int fpl_net_destroy(fpl_net *net) {
    if (!net) return FPL_OK;
    cudaSetDevice(net->ctx->device);
    free_device_weights(net);
    if (net->d_stage_in) cudaFree(net->d_stage_in);
    if (net->d_stage_out) cudaFree(net->d_stage_out);
    if (net->d_tile_ids) cudaFree(net->d_tile_ids);
    delete net;
    return FPL_OK;
}

int fpl_net_info(const fpl_net *net, int32_t *rf_size, int32_t *rf_offset, int32_t *rf_stride,
                 int32_t *infer_sz) {
    FPL_REQUIRE(net, "fpl_net_info: NULL net");
    if (rf_size) *rf_size = net->info.rf_size;
    if (rf_offset) *rf_offset = net->info.rf_offset;
    if (rf_stride) *rf_stride = net->info.rf_stride;
    if (infer_sz) *infer_sz = net->info.infer_sz;
    return FPL_OK;
}

int fpl_net_num_weights(const fpl_net *net, int32_t *n) {
    FPL_REQUIRE(net && n, "fpl_net_num_weights: NULL argument");
    int c = 0;
    for (const Op &o : net->ops) {
        if (o.kind == OP_CONV) c += o.bn ? 5 : 1;
        if (o.kind == OP_FINAL) c += net->info.final_bias ? 2 : 1;
    }
    *n = c;
    return FPL_OK;
}

int fpl_net_weight_size(const fpl_net *net, int32_t index, int64_t *elems) {
    FPL_REQUIRE(net && elems, "fpl_net_weight_size: NULL argument");
    int c = 0;
    for (const Op &o : net->ops) {
        if (o.kind == OP_CONV) {
            if (index == c) { *elems = (int64_t)o.k * o.k * o.k * o.cin * o.cout; return FPL_OK; }
            if (o.bn && index > c && index < c + 5) { *elems = o.cout; return FPL_OK; }
            c += o.bn ? 5 : 1;
        } else if (o.kind == OP_FINAL) {
            if (index == c) { *elems = o.cin; return FPL_OK; }
            if (net->info.final_bias && index == c + 1) { *elems = 1; return FPL_OK; }
            c += net->info.final_bias ? 2 : 1;
        }
    }
    fpl::set_error("fpl_net_weight_size: index %d out of range", index);
    return FPL_EINVAL;
}

int fpl_net_set_tile_multiplier(fpl_net *net, int32_t m) {
    FPL_REQUIRE(net && m >= 1, "fpl_net_set_tile_multiplier: bad argument");
    FPL_REQUIRE(net->info.rf_stride != 1 || m == 1,
                "tile multiplier > 1 is only valid for the VGG nets (U-Net output depends on the "
                "reference tile grid, SURVEY 5.7)");
    net->tile_mult = m;
    return FPL_OK;
}

int fpl_net_set_weights(fpl_net *net, const float *const *h_arrays, int32_t n, int precision) {
    FPL_REQUIRE(net && h_arrays, "fpl_net_set_weights: NULL argument");
    FPL_REQUIRE(precision == FPL_PREC_FP32 || precision == FPL_PREC_BF16 || precision == FPL_PREC_TF32,
                "fpl_net_set_weights: unknown precision %d", precision);
    int32_t want = 0;
    FPL_TRY(fpl_net_num_weights(net, &want));
    FPL_REQUIRE(n == want, "fpl_net_set_weights: expected %d arrays (Keras get_weights order), got %d", want, n);
    for (int i = 0; i < n; ++i) FPL_REQUIRE(h_arrays[i] != nullptr, "fpl_net_set_weights: array %d is NULL", i);
    FPL_CUDA_CHECK(cudaSetDevice(net->ctx->device));
    FPL_CUDA_CHECK(cudaDeviceSynchronize());
    free_device_weights(net);
    int w = 0;
    for (const Op &o : net->ops) {
        if (o.kind != OP_CONV && o.kind != OP_FINAL) continue;
        ConvParams &c = net->convs[o.conv_index];
        size_t ke = (size_t)o.k * o.k * o.k * o.cin * o.cout;
        c.kernel.assign(h_arrays[w], h_arrays[w] + ke);
        c.scale.assign(o.cout, 1.f);
        c.bias.assign(o.cout, 0.f);
        if (o.kind == OP_CONV && !o.bn) {
            w += 1;
        } else if (o.kind == OP_CONV) {
            const float *gamma = h_arrays[w + 1], *beta = h_arrays[w + 2], *mean = h_arrays[w + 3],
                        *var = h_arrays[w + 4];
            for (int j = 0; j < o.cout; ++j) {   // BatchNormalization(eps=1e-3) inference, folded in double
                double s = (double)gamma[j] / sqrt((double)var[j] + 1e-3);
                c.scale[j] = (float)s;
                c.bias[j] = (float)((double)beta[j] - (double)mean[j] * s);
            }
            w += 5;
        } else {
            if (net->info.final_bias) { c.bias[0] = h_arrays[w + 1][0]; w += 2; }
            else w += 1;
        }
        FPL_CUDA_CHECK(cudaMalloc((void **)&c.d_kernel, ke * sizeof(float)));
        FPL_CUDA_CHECK(cudaMalloc((void **)&c.d_scale, o.cout * sizeof(float)));
        FPL_CUDA_CHECK(cudaMalloc((void **)&c.d_bias, o.cout * sizeof(float)));
        FPL_CUDA_CHECK(cudaMemcpy(c.d_kernel, c.kernel.data(), ke * sizeof(float), cudaMemcpyHostToDevice));
        FPL_CUDA_CHECK(cudaMemcpy(c.d_scale, c.scale.data(), o.cout * sizeof(float), cudaMemcpyHostToDevice));
        FPL_CUDA_CHECK(cudaMemcpy(c.d_bias, c.bias.data(), o.cout * sizeof(float), cudaMemcpyHostToDevice));
    }
    net->precision = precision;
    if (precision != FPL_PREC_FP32) FPL_TRY(pack_weights_umma(net));
    return FPL_OK;
}

int fpl_net_out_size(const fpl_net *net, int32_t in_sz, int32_t *out_sz) {
    FPL_REQUIRE(net && out_sz, "fpl_net_out_size: NULL argument");
    int o = out_size(net, in_sz);
    FPL_REQUIRE(o > 0, "fpl_net_out_size: input edge %d is not valid for this architecture", in_sz);
    *out_sz = o;
    return FPL_OK;
}

int fpl_net_forward_tiles(fpl_net *net, const float *d_tiles, int32_t n_tiles, int32_t in_sz,
                          float *d_out, void *stream) {
    FPL_REQUIRE(net && d_tiles && d_out, "fpl_net_forward_tiles: NULL argument");
    if (net->precision < 0) { fpl::set_error("fpl_net_forward_tiles: network has no weights"); return FPL_ESTATE; }
    FPL_REQUIRE(n_tiles > 0, "fpl_net_forward_tiles: n_tiles must be > 0");
    FPL_REQUIRE(out_size(net, in_sz) > 0, "fpl_net_forward_tiles: input edge %d invalid", in_sz);
    FPL_CUDA_CHECK(cudaSetDevice(net->ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (net->precision == FPL_PREC_FP32) return forward_fp32(net, d_tiles, n_tiles, in_sz, d_out, st);
    return forward_umma(net, d_tiles, n_tiles, in_sz, d_out, st);
}

// clear_lo / clear_hi: whether the rf_offset-wide border planes at the low / high z end of d_pred belong to this
// call (whole volume: both; a z-slab of a larger volume: only where the slab touches the volume's end)
static int infer_volume_impl(fpl_net *net, const void *d_image, int image_is_u8, float norm_mean,
                             float norm_std, int64_t Z, int64_t Y, int64_t X, int32_t z_tile_begin,
                             int32_t z_tile_end, float *d_pred, bool clear_lo, bool clear_hi, void *stream) {
    FPL_REQUIRE(net && d_image && d_pred, "fpl_net_infer_volume: NULL argument");
    if (net->precision < 0) { fpl::set_error("network has not been trained"); return FPL_ESTATE; }
    FPL_REQUIRE(Z > 0 && Y > 0 && X > 0, "fpl_net_infer_volume: empty image");
    FPL_CUDA_CHECK(cudaSetDevice(net->ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    fpl_ctx *ctx = net->ctx;
    const int off = net->info.rf_offset;
    const int ref_out = net->info.infer_sz - 2 * off;
    TileGrid g;
    g.off = off;
    g.out_sz = ref_out;                     // reference grid; super-tiles are scheduled below
    g.in_sz = g.out_sz + 2 * off;
    g.in_z = g.in_sz; g.out_z = g.out_sz; g.z_base = 0;
    g.Z = Z; g.Y = Y; g.X = X;
    g.nz = tiles_along(Z, off, g.out_sz);
    g.ny = tiles_along(Y, off, g.out_sz);
    g.nx = tiles_along(X, off, g.out_sz);
    FPL_REQUIRE(out_size(net, g.in_sz) == g.out_sz, "internal: tile edge %d does not map to %d", g.in_sz, g.out_sz);
    int zb = z_tile_begin < 0 ? 0 : z_tile_begin;
    int ze = (z_tile_end < 0 || z_tile_end > g.nz) ? g.nz : z_tile_end;
    // only this rank's rows of pred are cleared/written when a z-range is given
    {
        long long z0 = (zb == 0) ? (clear_lo ? 0 : off) : (long long)zb * g.out_sz + off;
        long long z1 = (ze == g.nz) ? (clear_hi ? Z : Z - off) : (long long)ze * g.out_sz + off;
        if (z1 > Z) z1 = Z;
        if (z0 > Z) z0 = Z;
        if (z1 > z0)
            FPL_CUDA_CHECK(cudaMemsetAsync(d_pred + z0 * Y * X, 0, sizeof(float) * (size_t)(z1 - z0) * Y * X, st));
    }
    // ---- tile schedule -----------------------------------------------------------------------
    // Phase 1 (VGG, tile_mult m > 1): super-tiles of m x m x m reference tiles.  Their origins stay
    // on the reference grid (multiples of out_sz, itself a multiple of the net stride), the nets are
    // shift-equivariant by rf_stride and the zero padding beyond the image is reproduced by the
    // gather, so every voxel gets exactly the value the reference tile would give it -- with less
    // halo recomputation.  Phase 2: the reference tiles not covered by super-tiles.
    const int m = net->tile_mult;
    struct Phase { TileGrid grid; long long n; int batch; const int *ids; int first; };
    std::vector<Phase> phases;
    // z-slab mode (tcgen05 path, VGG, m > 1): one tile spans the full x/y extent of the volume (no
    // halo recomputation in x and y at all) and m reference tile layers in z; the remaining layers form
    // one thinner slab.  Used when the x/y tile counts are similar and the largest activation of a slab
    // (first conv output, 96 B per voxel) fits comfortably in HBM.
    bool slab_mode = false;
    if (m > 1 && net->precision == FPL_PREC_BF16 && zb == 0 && ze == g.nz && !g_no_slab_mode) {
        const int nxy = g.ny > g.nx ? g.ny : g.nx, nmin = g.ny < g.nx ? g.ny : g.nx;
        int mz = m < g.nz ? m : g.nz;
        const double xy_bytes = 96.0 * ((double)nxy * g.out_sz + 2 * off) * ((double)nxy * g.out_sz + 2 * off);
        while (mz > 1 && xy_bytes * (mz * g.out_sz + 2 * off) > 48e9) --mz;
        if (nxy > 0 && nmin * 5 >= nxy * 4 && xy_bytes * (mz * g.out_sz + 2 * off) <= 48e9) {
            slab_mode = true;
            TileGrid gs = g;
            gs.out_sz = nxy * g.out_sz; gs.in_sz = gs.out_sz + 2 * off;
            gs.ny = gs.nx = 1;
            // full slabs of mz reference layers, then ONE tail slab that holds exactly the planes that are left, rounded
            // up to the net stride (shift-equivariance: any thickness that is a multiple of rf_stride reproduces the
            // reference values) -- a tail of whole layers would compute up to out_sz - 1 planes nobody reads, which is
            // a quarter of the work of a rank that owns 128 planes of a sharded volume
            const long long need = Z - 2LL * off;                        // prediction planes to produce
            const long long slab_out = (long long)mz * g.out_sz;
            const int n_full = (int)(need / slab_out);
            const long long tail = need - (long long)n_full * slab_out;
            const int stride = net->info.rf_stride;
            if (n_full > 0) {
                gs.out_z = (int)slab_out; gs.in_z = gs.out_z + 2 * off; gs.nz = n_full; gs.z_base = 0;
                phases.push_back({gs, (long long)n_full, 1, nullptr, 0});
            }
            if (tail > 0) {
                gs.out_z = (int)((tail + stride - 1) / stride * stride); gs.in_z = gs.out_z + 2 * off; gs.nz = 1;
                gs.z_base = (long long)n_full * slab_out;
                phases.push_back({gs, 1, 1, nullptr, 0});
            }
        }
    }
    const int nsz = (m > 1 && !slab_mode) ? (ze - zb) / m : 0, nsy = (m > 1 && !slab_mode) ? g.ny / m : 0,
              nsx = (m > 1 && !slab_mode) ? g.nx / m : 0;
    const bool have_super = nsz > 0 && nsy > 0 && nsx > 0;
    std::vector<int> rest;
    if (!slab_mode)
    for (int kz = zb; kz < ze; ++kz)
        for (int ky = 0; ky < g.ny; ++ky)
            for (int kx = 0; kx < g.nx; ++kx) {
                const bool in_super = have_super && (kz - zb) < nsz * m && ky < nsy * m && kx < nsx * m;
                if (!in_super) rest.push_back((kz * g.ny + ky) * g.nx + kx);
            }
    int *d_ids = nullptr;
    if (!rest.empty()) {
        if (net->tile_ids_cap < rest.size()) {
            FPL_CUDA_CHECK(cudaStreamSynchronize(st));
            if (net->d_tile_ids) cudaFree(net->d_tile_ids);
            net->d_tile_ids = nullptr; net->tile_ids_cap = 0;
            FPL_CUDA_CHECK(cudaMalloc((void **)&net->d_tile_ids, rest.size() * sizeof(int)));
            net->tile_ids_cap = rest.size();
        }
        d_ids = net->d_tile_ids;
    }
    if (have_super) {
        TileGrid gs = g;
        gs.out_sz = g.out_sz * m; gs.in_sz = gs.out_sz + 2 * off;
        gs.out_z = gs.out_sz; gs.in_z = gs.in_sz;
        gs.nz = nsz; gs.ny = nsy; gs.nx = nsx;
        FPL_REQUIRE(zb == 0, "tile multiplier > 1 cannot be combined with a z tile range");
        FPL_REQUIRE(out_size(net, gs.in_sz) == gs.out_sz, "internal: super-tile edge %d invalid", gs.in_sz);
        const double per_tile_bytes = 96.0 * gs.in_sz * (double)gs.in_sz * gs.in_sz;
        int bs = (int)(3.0e9 / per_tile_bytes); if (bs < 1) bs = 1; if (bs > 8) bs = 8;
        phases.push_back({gs, (long long)nsz * nsy * nsx, bs, nullptr, 0});
    }
    if (!rest.empty()) {
        // the previous call's list may still be in use by queued kernels of the same stream: ordered copy
        FPL_CUDA_CHECK(cudaMemcpyAsync(d_ids, rest.data(), rest.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));     // `rest` is pageable host memory
        phases.push_back({g, (long long)rest.size(), net->precision == FPL_PREC_FP32 ? 4 : 32, d_ids, 0});
    }
    size_t need_in = 0, need_out = 0;
    for (Phase &ph : phases) {
        if (ph.batch > ph.n) ph.batch = (int)ph.n;
        size_t ie = (size_t)ph.grid.in_z * ph.grid.in_sz * ph.grid.in_sz * ph.batch;
        size_t oe = (size_t)ph.grid.out_z * ph.grid.out_sz * ph.grid.out_sz * ph.batch;
        if (net->precision == FPL_PREC_BF16 && !g_no_direct_io) oe = 1;      // the final layer scatters directly
        if (net->precision == FPL_PREC_BF16 && !g_no_direct_io && umma_reads_volume(net)) ie = 1;
        if (ie > need_in) need_in = ie;
        if (oe > need_out) need_out = oe;
    }
    if ((net->stage_in_cap < need_in || net->stage_out_cap < need_out)) {
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        if (net->d_stage_in) cudaFree(net->d_stage_in);
        if (net->d_stage_out) cudaFree(net->d_stage_out);
        net->d_stage_in = net->d_stage_out = nullptr; net->stage_in_cap = net->stage_out_cap = 0;
        FPL_CUDA_CHECK(cudaMalloc((void **)&net->d_stage_in, sizeof(float) * need_in));
        FPL_CUDA_CHECK(cudaMalloc((void **)&net->d_stage_out, sizeof(float) * need_out));
        net->stage_in_cap = need_in; net->stage_out_cap = need_out;
    }
    float *d_in = net->d_stage_in, *d_out = net->d_stage_out;
    int rc = FPL_OK;
    const int blocks = ctx->sm_count * 16;
    for (const Phase &ph : phases) {
        const TileGrid &tg = ph.grid;
        const size_t in_elems = (size_t)tg.in_z * tg.in_sz * tg.in_sz, out_elems = (size_t)tg.out_z * tg.out_sz * tg.out_sz;
        for (long long t0 = 0; t0 < ph.n && rc == FPL_OK; t0 += ph.batch) {
            const int nb = (int)((ph.n - t0) < ph.batch ? (ph.n - t0) : ph.batch);
            // tcgen05 path: the final layer scatters straight into the prediction volume; the input tile is
            // staged by the gather kernel (float32, L2 resident) unless the fused first+second convolution
            // kernel runs, whose builders read the volume directly in the shadow of the MMAs
            const bool fused_scatter = net->precision == FPL_PREC_BF16 && !g_no_direct_io;
            // ... and when the first two convolutions run fused, its builders gather straight from the volume
            const bool direct_in = fused_scatter && umma_reads_volume(net);
            if (!direct_in) {
                fpl::ProfScope prof(ctx, st, fpl::PROF_TILER, (double)nb * in_elems * (image_is_u8 ? 5.0 : 8.0));
                if (image_is_u8)
                    gather_tiles_kernel<uint8_t><<<blocks, 128, 0, st>>>((const uint8_t *)d_image, d_in, tg, (int)t0, nb,
                                                                         norm_mean, norm_std, ph.ids);
                else
                    gather_tiles_kernel<float><<<blocks, 128, 0, st>>>((const float *)d_image, d_in, tg, (int)t0, nb,
                                                                       0.f, 1.f, ph.ids);
                ctx->launches++;
            }
            if (net->precision == FPL_PREC_FP32) rc = forward_fp32(net, d_in, nb, tg.in_sz, d_out, st);
            else if (fused_scatter) {
                VolumeIO vio;
                vio.g = tg; vio.tile0 = (int)t0; vio.ids = ph.ids; vio.pred = d_pred;
                if (direct_in) { vio.img = d_image; vio.is_u8 = image_is_u8; vio.mean = norm_mean; vio.stdv = norm_std; }
                rc = forward_umma(net, d_in, nb, tg.in_sz, nullptr, st, &vio, tg.in_z);
            } else rc = forward_umma(net, d_in, nb, tg.in_sz, d_out, st, nullptr, tg.in_z);
            if (rc != FPL_OK) break;
            if (!fused_scatter) {
                fpl::ProfScope prof(ctx, st, fpl::PROF_TILER, (double)nb * out_elems * 8.0);
                scatter_tiles_kernel<<<blocks, 128, 0, st>>>(d_out, d_pred, tg, (int)t0, nb, ph.ids);
                ctx->launches++;
            }
        }
        if (rc != FPL_OK) break;
    }
    if (rc != FPL_OK) return rc;
    FPL_CUDA_CHECK(cudaGetLastError());
    return FPL_OK;
}

int fpl_net_infer_volume(fpl_net *net, const void *d_image, int image_is_u8, float norm_mean,
                         float norm_std, int64_t Z, int64_t Y, int64_t X, int32_t z_tile_begin,
                         int32_t z_tile_end, float *d_pred, void *stream) {
    return infer_volume_impl(net, d_image, image_is_u8, norm_mean, norm_std, Z, Y, X, z_tile_begin, z_tile_end,
                             d_pred, true, true, stream);
}

int fpl_net_infer_slab(fpl_net *net, const void *d_image_slab, int image_is_u8, float norm_mean, float norm_std,
                       int64_t Z, int64_t z0, int64_t z1, int64_t Y, int64_t X, float *d_pred_slab, void *stream) {
    FPL_REQUIRE(net && d_image_slab && d_pred_slab, "fpl_net_infer_slab: NULL argument");
    FPL_REQUIRE(z0 >= 0 && z1 > z0 && z1 <= Z, "fpl_net_infer_slab: bad plane range [%lld,%lld) of %lld",
                (long long)z0, (long long)z1, (long long)Z);
    const int off = net->info.rf_offset, out = net->info.infer_sz - 2 * off;
    // cut granularity: the VGG nets are shift-equivariant by rf_stride, so any multiple of it reproduces the
    // whole-volume values; the U-Net output depends on the tile phase, so cuts must lie on the reference grid
    const int gran = net->info.rf_stride != 1 ? net->info.rf_stride : out;
    FPL_REQUIRE(z0 % gran == 0, "fpl_net_infer_slab: first plane %lld is not a multiple of %d", (long long)z0, gran);
    FPL_REQUIRE(z1 == Z || (z1 - z0 > 2 * off && (z1 - z0 - 2 * off) % gran == 0),
                "fpl_net_infer_slab: an inner slab must hold 2*rf_offset + k*%d planes (got %lld)", gran,
                (long long)(z1 - z0));
    return infer_volume_impl(net, d_image_slab, image_is_u8, norm_mean, norm_std, z1 - z0, Y, X, 0, -1, d_pred_slab,
                             z0 == 0, z1 == Z, stream);
}

}  // extern "C"
