// net.cuh -- network description shared by the fp32 validation path and the tcgen05 path.
#pragma once
#include "common.cuh"
#include <string>
#include <vector>

namespace fpl {
namespace net {

// One step of the (static) graph.  Mirrors the Keras functional graphs of
// flypylib/fplmodels.py:102-136 / :138-172 / :258-304.
enum OpKind { OP_CONV = 0, OP_POOL = 1, OP_SAVE = 2, OP_UPCAT = 3, OP_FINAL = 4, OP_ADD = 5 };

struct Op {
    OpKind kind;
    int k = 0, cin = 0, cout = 0;   // OP_CONV / OP_FINAL (cout == 1)
    int slot = -1;                  // OP_SAVE: skip slot to store; OP_UPCAT: skip slot to read
    int crop = 0;                   // OP_UPCAT / OP_ADD: symmetric crop of the skip tensor
    bool relu = true;               // OP_CONV: ReLU after the (folded) BatchNormalization (false: resnet_like's BN-only convs)
    bool bn = true;                 // OP_CONV: followed by BatchNormalization (false: plain convolution, one weight array)
    int src_slot = -1;              // OP_CONV: >= 0 convolves the stored skip tensor in place of the current one (1x1x1 only)
    int conv_index = -1;            // index into ConvParams
};

struct ConvParams {                 // host copies (Keras order) + folded device copies
    int k, cin, cout;
    std::vector<float> kernel;      // (kd,kh,kw,cin,cout)
    std::vector<float> scale, bias; // BN folded: y = conv*scale + bias   (final: scale=1, bias)
    float *d_kernel = nullptr;      // fp32, same layout
    float *d_scale = nullptr, *d_bias = nullptr;
    // tcgen05 path
    void *d_packed = nullptr;       // operand-B image for the UMMA kernel (BN scale folded in)
    size_t packed_bytes = 0;
    bool no_rot = false;            // second conv of the fused first+second kernel: split-window weight layout
    bool no_ksplit = false;         // planner: resident-weight plans only (chunk convolutions of the hi/lo path)
    // hi/lo (bf16x3) path: per input-channel chunk one image of bf16(w) and one of bf16(w - bf16(w))
    void *d_packed_lo = nullptr;
    int hilo_chunk = 0;             // input channels per chunk
};

// Tile grid of FplNetwork.infer (fplnetwork.py:146-160) and the optional direct volume I/O of a
// forward pass: the first layer reads its input tile straight from the volume (zero beyond the far
// edge, (x-mean)/std for uint8) and the final layer scatters straight into the prediction volume.
struct TileGrid {
    int nz, ny, nx;          // tiles per axis
    int in_sz, out_sz, off;  // x/y: tile input edge, useful output edge (= stride of origins), rf_offset
    int in_z, out_z;         // z: the same for the z axis (z-slab tiles have a different z extent)
    long long z_base;        // z origin of tile layer 0 of this grid
    long long Z, Y, X;
};
struct VolumeIO {
    const void *img = nullptr;   // (Z,Y,X) uint8 or float32
    int is_u8 = 0;
    float mean = 0.f, stdv = 1.f;
    TileGrid g{};
    int tile0 = 0;               // first tile of the batch (index into ids, or linear tile id)
    const int *ids = nullptr;    // optional linear tile ids
    float *pred = nullptr;       // (Z,Y,X) float32
};

struct ArchInfo {
    int rf_size, rf_offset, rf_stride, infer_sz;
    bool final_bias;
};

}  // namespace net
}  // namespace fpl

struct fpl_net {
    fpl_ctx *ctx = nullptr;
    int arch = 0;
    fpl::net::ArchInfo info{};
    std::vector<fpl::net::Op> ops;
    std::vector<fpl::net::ConvParams> convs;   // every OP_CONV and the OP_FINAL, in graph order
    int precision = -1;                        // -1: no weights yet
    int tile_mult = 1;                         // VGG only: tile edge = tile_mult*out_sz + 2*off
    float *d_stage_in = nullptr, *d_stage_out = nullptr;   // tile staging of fpl_net_infer_volume (grow-only)
    size_t stage_in_cap = 0, stage_out_cap = 0;
    int *d_tile_ids = nullptr; size_t tile_ids_cap = 0;    // reference-tile id list of the tiler
};

namespace fpl {
namespace net {
// fp32 validation path (conv_fp32.cu)
int forward_fp32(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out,
                 cudaStream_t st);
// tcgen05 path (conv_umma.cu)
int forward_umma(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out,
                 cudaStream_t st, const VolumeIO *vio = nullptr, int in_z = 0);   // in_z = 0: cubic tiles
int pack_weights_umma(fpl_net *net);
bool umma_reads_volume(const fpl_net *net);
void free_packed_umma(fpl_net *net);
// output edge of a tile for input edge in_sz (after the x rf_stride up-sampling); -1 if invalid
int out_size(const fpl_net *net, int in_sz);
}  // namespace net
}  // namespace fpl
