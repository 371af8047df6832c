// conv_umma.cu -- tcgen05 / TMEM / TMA implicit-GEMM execution of the network graph (FPL_PREC_BF16).
//
// Replaces infer_network.predict (flypylib/fplnetwork.py:175-176) for the graphs of
// flypylib/fplmodels.py:102-136,138-172,258-304.
//
// Data layout in HBM ("C8-blocked"): activations are (tile, C/8, z, y, x, 8) bf16 -- one 16-byte
// channel atom per voxel per channel group.  A K=16 MMA step consumes two atoms; the UMMA
// K-major no-swizzle canonical layout (8 rows x 16 B core matrices) is then exactly "8 x-consecutive
// voxels of one atom", so every one of the 27 taps of a 3x3x3 convolution is just a different start
// address into ONE halo'd input plane staged in shared memory by a single TMA box load
// (no im2col materialisation, no re-fetch per tap).
//
// conv kernel (one CTA per SM, persistent over work items = (tile, 16x16 xy patch, z chunk)):
//   warp 0      TMA producer : weights once (cp.async.bulk, resident for the CTA's lifetime), then one
//                              halo'd (18 x 18 x Cin) input plane per z step into a 3-plane ring
//   warp 1      MMA issuer   : per output plane, 27 taps x Cin/16 K-steps x 2 M-tiles of
//                              tcgen05.mma.cta_group::1.kind::f16 (M=128 = 8x16 voxels, N=Cout),
//                              fp32 accumulators in TMEM (double buffered across planes)
//   warps 2..5  epilogue     : tcgen05.ld -> + folded-BN bias -> ReLU -> bf16 -> C8-blocked store
// Roofline: tensor pipe (dense contraction); see DESIGN.md for the FLOP count per tile.
#include "net.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace fpl {
namespace net {


// ------------------------------------------------------------------------------------------------
// the implicit-GEMM convolution kernel
// ------------------------------------------------------------------------------------------------
constexpr int kTX = 16, kTY = 16;            // output patch per CTA and z step (two M=128 tiles: x 0..7, 8..15)
constexpr int kAccStages = 2;                // TMEM accumulator double buffering
constexpr int kThreads = 320;             // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct ConvArgs {
    const __nv_bfloat16 *w_packed;   // operand-B image of this Cout split (see pack_weights_umma), BN scale folded
    const float *bias;               // folded BN bias of this split [cout]
    __nv_bfloat16 *out;              // (tile, cout_total/8, Dout, Dout, Dout, 8)  (pooled edge when pool)
    uint32_t w_bytes;
    int n_tiles, din, dout;          // x/y extent of the input / output tensors
    int din_z, dout_z;               // z extent (tiles may be z-slabs: full xy extent, limited z)
    int cin_atoms_total;             // Cin / 8 of the input tensor
    int nsub;                        // input channels are consumed in nsub sub-planes of 2*KSTEPS atoms each
    int cout;                        // output channels of this launch (N of the MMA)
    int cout_total, cout_off;        // position of this split inside the output tensor
    int ring;                        // (sub-)plane ring depth, 2 or 3
    int n_xt, n_yt, n_zc, zc_len;
    int relu;
    int pool;                        // fuse MaxPooling3D(2): `out` is the pooled tensor (edge dout/2)
    int max_blk;                     // cap on accumulator blocks per M-tile region (0: 256/cout)
    uint32_t tmem_cols;
    // K-split: the input channels are contracted in several launches; fp32 partial sums live in HBM between them
    int cin_atom_off;                // first channel atom of this launch's input-channel chunk
    int acc_mode;                    // 0 normal, 1 first chunk (write partial), 2 middle (partial += acc), 3 last (finish)
    float *partial;                  // (tile, cout/8, Dout_z, Dout, Dout, 8) fp32
    // hi/lo path: the nsub sub-planes are the hi and the lo part of the same channels and share ONE weight image
    int shared_w;                    // 1: weights are indexed without the sub-plane offset
    int sub_stride;                  // channel atoms between consecutive sub-planes (0: contiguous, 2*KSTEPS)
};

// Epilogue of one M=128 accumulator tile (N = cout fp32 columns in TMEM):
// TMEM -> registers -> + folded-BN bias (shared memory) -> ReLU -> bf16 -> C8-blocked global store.
// Warp quadrant `q` owns TMEM lanes [32q, 32q+32); lane = accumulator row = voxel (y = row/8, x = row%8).
// The next 16-column TMEM load is in flight while the previous one is converted and stored.
__device__ __forceinline__ void epilogue_store16(const uint32_t (&r)[16], int c0, const float *s_bias, int relu,
                                                 __nv_bfloat16 *__restrict__ out, size_t vox, size_t cg_stride,
                                                 bool ok) {
    if (!ok) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v0 = __uint_as_float(r[h * 8 + 2 * j]) + s_bias[c0 + h * 8 + 2 * j];
            float v1 = __uint_as_float(r[h * 8 + 2 * j + 1]) + s_bias[c0 + h * 8 + 2 * j + 1];
            if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
            pk[j] = *reinterpret_cast<uint32_t *>(&b2);
        }
        *reinterpret_cast<uint4 *>(out + (vox + (size_t)((c0 >> 3) + h) * cg_stride) * 8) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, int q, int lane, int cout, const float *s_bias,
                                              int relu, __nv_bfloat16 *__restrict__ out, int tile, int dout, int z,
                                              int y0, int x0, int cout_total, int cout_off, int dout_z) {
    const int row = q * 32 + lane;
    const int y = y0 + (row >> 3), x = x0 + (row & 7);
    const bool ok = (x < dout) && (y < dout);
    const size_t cg_stride = (size_t)dout_z * dout * dout;
    const size_t vox = ((size_t)tile * (cout_total >> 3) + (cout_off >> 3)) * cg_stride + ((size_t)z * dout + y) * dout + x;
    const uint32_t tcol = tmem_acc + ((uint32_t)(q * 32) << 16);
    uint32_t ra[16], rb[16];
    tmem_ld16(tcol, ra);
#pragma unroll 1
    for (int c0 = 0; c0 < cout; c0 += 32) {
        tmem_ld_wait();
        const bool more = c0 + 16 < cout;
        if (more) tmem_ld16(tcol + (uint32_t)(c0 + 16), rb);
        epilogue_store16(ra, c0, s_bias, relu, out, vox, cg_stride, ok);
        if (more) {
            tmem_ld_wait();
            if (c0 + 32 < cout) tmem_ld16(tcol + (uint32_t)(c0 + 32), ra);
            epilogue_store16(rb, c0 + 16, s_bias, relu, out, vox, cg_stride, ok);
        }
    }
}

// Epilogue with MaxPooling3D((2,2,2)) fused in (Keras floor semantics; the caller guarantees an even
// conv output edge).  Max commutes with the per-channel bias add, ReLU and the bf16 rounding (all
// monotone), so pooling is done on the raw fp32 accumulators: x pairs are adjacent lanes, y pairs are
// lanes 8 apart, z pairs are consecutive output planes of the same warp (even plane parked in `hold`).
// A "transpose-reduce" exchange leaves every lane of a 2x2 group with 4 of the 16 columns of a chunk:
// column offset = 8*(x parity) + 4*(y parity); each lane then stores its 4 channels (8 bytes).
template <int NCH>
__device__ __forceinline__ void epilogue_tile_pool(uint32_t tmem_acc, int q, int lane, const float *s_bias, int relu,
                                                   __nv_bfloat16 *__restrict__ out, int tile, int dpool, int z,
                                                   int y0, int x0, float (&hold)[NCH * 4], int dpool_z) {
    const int row = q * 32 + lane;
    const int y = y0 + (row >> 3), x = x0 + (row & 7);
    const bool ok = (x < 2 * dpool) && (y < 2 * dpool);
    const int xpar = lane & 1, ypar = (lane >> 3) & 1;
    const uint32_t tcol = tmem_acc + ((uint32_t)(q * 32) << 16);
    uint32_t r[NCH][16];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tmem_ld16(tcol + (uint32_t)(c * 16), r[c]);
    tmem_ld_wait();
    const size_t cg_stride = (size_t)dpool_z * dpool * dpool;
    const size_t vox = (size_t)tile * (NCH * 2) * cg_stride + ((size_t)(z >> 1) * dpool + (y >> 1)) * dpool + (x >> 1);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {       // x pair: keep columns [8*xpar, 8*xpar+8), trade the other half
            const float keep = __uint_as_float(xpar ? r[c][8 + j] : r[c][j]);
            const float give = __uint_as_float(xpar ? r[c][j] : r[c][8 + j]);
            v8[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, give, 1));
        }
        float v4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {       // y pair: keep [4*ypar, 4*ypar+4) of those
            const float keep = ypar ? v8[4 + j] : v8[j];
            const float give = ypar ? v8[j] : v8[4 + j];
            v4[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, give, 8));
        }
        if ((z & 1) == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) hold[c * 4 + j] = v4[j];
        } else if (ok) {
            const int ch = c * 16 + xpar * 8 + ypar * 4;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = fmaxf(v4[j], hold[c * 4 + j]) + s_bias[ch + j];
                if (relu) o[j] = fmaxf(o[j], 0.f);
            }
            __nv_bfloat162 b0 = __floats2bfloat162_rn(o[0], o[1]), b1 = __floats2bfloat162_rn(o[2], o[3]);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&b0), *reinterpret_cast<uint32_t *>(&b1));
            *reinterpret_cast<uint2 *>(out + (vox + (size_t)(ch >> 3) * cg_stride) * 8 + (ch & 7)) = pk;
        }
    }
}

// K-split epilogue of one M=128 accumulator tile: partial sums of the input-channel chunks are kept as fp32
// C8-blocked atoms (32 B) in HBM.  mode 1: partial = acc; 2: partial += acc; 3: out = bf16(relu(partial + acc + bias)).
// mode 4 (hi/lo path): as 3, but the fp32 result v is stored as two bf16 tensors hi = bf16(v), lo = bf16(v - hi)
// in channel groups [0, cout/8) and [cout/8, cout/4) of a tensor with 2*cout channels.
__device__ __forceinline__ void epilogue_tile_ksplit(uint32_t tmem_acc, int q, int lane, int cout, const float *s_bias,
                                                     int relu, __nv_bfloat16 *__restrict__ out, float *__restrict__ partial,
                                                     int mode, int tile, int dout, int z, int y0, int x0, int dout_z) {
    const int row = q * 32 + lane;
    const int y = y0 + (row >> 3), x = x0 + (row & 7);
    const bool ok = (x < dout) && (y < dout);
    const size_t cg_stride = (size_t)dout_z * dout * dout;
    const size_t vox = (size_t)tile * (cout >> 3) * cg_stride + ((size_t)z * dout + y) * dout + x;
    const uint32_t tcol = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < cout; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tcol + (uint32_t)c0, r);
        float4 pv[4];
        if (ok && mode != 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 *pp = reinterpret_cast<const float4 *>(partial + (vox + (size_t)((c0 >> 3) + h) * cg_stride) * 8);
                pv[2 * h] = pp[0]; pv[2 * h + 1] = pp[1];
            }
        }
        tmem_ld_wait();
        if (!ok) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (mode != 1) {
#pragma unroll
            for (int h = 0; h < 4; ++h) { v[4 * h] += pv[h].x; v[4 * h + 1] += pv[h].y; v[4 * h + 2] += pv[h].z; v[4 * h + 3] += pv[h].w; }
        }
        if (mode == 4) {
            const size_t ovox = (size_t)tile * (cout >> 2) * cg_stride + ((size_t)z * dout + y) * dout + x;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v0 = v[h * 8 + 2 * j] + s_bias[c0 + h * 8 + 2 * j];
                    float v1 = v[h * 8 + 2 * j + 1] + s_bias[c0 + h * 8 + 2 * j + 1];
                    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v0 - __low2float(hi), v1 - __high2float(hi));
                    ph[j] = *reinterpret_cast<const uint32_t *>(&hi); pl[j] = *reinterpret_cast<const uint32_t *>(&lo);
                }
                *reinterpret_cast<uint4 *>(out + (ovox + (size_t)((c0 >> 3) + h) * cg_stride) * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4 *>(out + (ovox + (size_t)((cout >> 3) + (c0 >> 3) + h) * cg_stride) * 8) =
                    make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
        } else if (mode != 3) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 *pp = reinterpret_cast<float4 *>(partial + (vox + (size_t)((c0 >> 3) + h) * cg_stride) * 8);
                pp[0] = make_float4(v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
                pp[1] = make_float4(v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]);
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v0 = v[h * 8 + 2 * j] + s_bias[c0 + h * 8 + 2 * j];
                    float v1 = v[h * 8 + 2 * j + 1] + s_bias[c0 + h * 8 + 2 * j + 1];
                    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                    pk[j] = *reinterpret_cast<uint32_t *>(&b2);
                }
                *reinterpret_cast<uint4 *>(out + (vox + (size_t)((c0 >> 3) + h) * cg_stride) * 8) =
                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

// Accumulator placement ("kd fusion").  The three kd taps of a 3x3x3 kernel read the SAME shifted A
// tile of input plane p and feed three DIFFERENT output planes z = p, p-1, p-2.  Their accumulators
// are laid out in TMEM so that they sit in adjacent N-column blocks in kd order; one
// tcgen05.mma with N = 3*Cout then updates all three at once and the A tile is read from shared
// memory once instead of three times (the kernel is bound by that read, see profiles/).
// Each M-tile owns a 256-column region of kBlocks = 256/Cout blocks; output number A (monotone
// counter) lives in block kBlocks-1 - (A mod kBlocks), so consecutive outputs occupy descending
// blocks and (p, p-1, p-2) are contiguous except where the ring wraps (then two MMAs are issued).
//
// ROT variant ("rotating window", KS = 3, two M-tiles): three accumulator blocks per M-tile form a ring
// (output A in block 2 - A mod 3), so the blocks fed by one input plane are ALWAYS the three adjacent
// N-column blocks 0,1,2 -- only the assignment kd -> block rotates with A mod 3.  The packed weights
// carry the kd row blocks as [kd0 kd1 kd2 kd0 kd1]; the three cyclic orders are the three windows of
// that sequence, selected by the start address of the B descriptor.  Every (tap, K-step) is then ONE
// tcgen05.mma with N = 3*Cout on every plane (no split where a ring wraps), i.e. the A tile is read from
// shared memory once.  The two M-tiles are issued one after the other (own full/empty barriers), so the
// epilogue of an M-tile drains its finished block while the other M-tile's MMAs run.
template <int KS, int KSTEPS, int TX, bool ROT>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const ConvArgs a) {
    constexpr int SX = TX + KS - 1, SY = kTY + KS - 1;
    constexpr int MT = TX / 8;                      // M=128 tiles per plane patch (x halves)
    static_assert(!ROT || (KS == 3 && MT == 2), "ROT needs a 3x3x3 kernel and two M-tiles");
    constexpr int SUB_ATOMS = 2 * KSTEPS;           // channel atoms per (sub-)plane
    constexpr int kMaxBlocks = 8, kMaxRing = 3;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr uint32_t atom_stride = SY * SX * 16;                  // bytes between channel atoms of a plane
    constexpr uint32_t plane_bytes = SUB_ATOMS * atom_stride;
    constexpr uint32_t plane_pitch = (plane_bytes + 127u) & ~127u;
    const uint32_t w_region = (a.w_bytes + 127u) & ~127u;
    const uint32_t ring = (uint32_t)a.ring;
    uint8_t *s_w = smem_raw;
    uint8_t *s_planes = smem_raw + w_region;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_planes + ring * plane_pitch);
    uint64_t *plane_full = bars;                    // [kMaxRing]
    uint64_t *plane_empty = bars + kMaxRing;        // [kMaxRing]
    uint64_t *acc_full = bars + 2 * kMaxRing;       // [kMaxBlocks]
    uint64_t *acc_empty = acc_full + kMaxBlocks;    // [kMaxBlocks]
    uint64_t *w_full = acc_empty + kMaxBlocks;      // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_full + 1);
    float *s_bias = reinterpret_cast<float *>(bars + 32);          // [cout]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.n_tiles * a.n_yt * a.n_xt * a.n_zc;
    const uint32_t N = (uint32_t)a.cout;
    uint32_t nblk_ = 256u / N > (uint32_t)kMaxBlocks ? (uint32_t)kMaxBlocks : 256u / N;
    if (a.max_blk > 0 && (uint32_t)a.max_blk < nblk_) nblk_ = (uint32_t)a.max_blk;
    const uint32_t nblk = ROT ? 3u : nblk_;
    const int nsub = a.nsub;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxRing; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], 1); }
        for (int i = 0; i < kMaxBlocks; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], ROT ? 4 : 4 * MT); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < a.cout; i += blockDim.x) s_bias[i] = a.bias[i];
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Roles run warp-uniformly (all 32 lanes walk the same loops on uniform values); only the
    // asynchronous instructions themselves are issued by one elected lane.  This keeps descriptors in
    // uniform registers and avoids per-instruction waterfall loops in SASS.
    if (warp == 0) {
        // ===================================== TMA producer =====================================
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(w_full, a.w_bytes);
            for (uint32_t off = 0; off < a.w_bytes; off += 32768u) {
                uint32_t n = a.w_bytes - off < 32768u ? a.w_bytes - off : 32768u;
                bulk_load_1d(s_w + off, reinterpret_cast<const uint8_t *>(a.w_packed) + off, n, w_full);
            }
        }
        uint32_t pc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int t = item;
            // x patch fastest: CTAs running side by side sweep z in step and touch adjacent rows of HBM
            const int xt = t % a.n_xt; t /= a.n_xt;
            const int yt = t % a.n_yt; t /= a.n_yt;
            const int zc = t % a.n_zc; t /= a.n_zc;
            const int tile = t;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            const int np = nz + KS - 1;
            for (int p = 0; p < np; ++p)
                for (int sub = 0; sub < nsub; ++sub, ++pc) {
                    const uint32_t slot = pc % ring, ph = (pc / ring) & 1u;
                    mbar_wait(&plane_empty[slot], ph ^ 1u);
                    if (leader) {
                        mbar_expect_tx(&plane_full[slot], plane_bytes);
                        tma_load_4d(s_planes + slot * plane_pitch, &tmap_in, &plane_full[slot], xt * TX * 8, yt * kTY,
                                    z0 + p, tile * a.cin_atoms_total + a.cin_atom_off + sub * (a.sub_stride ? a.sub_stride : SUB_ATOMS));
                    }
                }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        const bool leader = elect_one();
        const uint32_t idesc_1 = make_idesc_bf16(128, (int)N);         // N field grows linearly with the run length
        // descriptor words in 16-byte units: lo = start | LBO << 16, hi = SBO | version 1 << 14
        const uint32_t a_hi = (SX * 16u >> 4) | (1u << 14);
        const uint32_t b_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = (smem_u32(s_planes) >> 4) | ((atom_stride >> 4) << 16);
        const uint32_t b_lo0 = (smem_u32(s_w) >> 4) | ((N * KS) << 16);   // LBO = KS*Cout*16 bytes
        const uint32_t b_step16 = N * KS * 2u;                            // one (kh,kw,kstep) chunk = KS*Cout*32 bytes
        const uint32_t kt = (uint32_t)((a.shared_w ? 1 : nsub) * KSTEPS);   // K steps per tap in the weight image
        const uint32_t wsub = a.shared_w ? 0u : (uint32_t)KSTEPS;           // weight K-step offset per sub-plane
        mbar_wait(w_full, 0);
        uint32_t pc = 0, ac0 = 0;
        if constexpr (ROT) {
            const uint32_t idesc_3 = make_idesc_bf16(128, (int)(3u * N));
            const uint32_t b_lo0r = (smem_u32(s_w) >> 4) | ((N * 5u) << 16);   // LBO = 5*Cout*16 bytes
            const uint32_t b_step16r = N * 5u * 2u;                            // chunk = 5*Cout*32 bytes
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int zc = (item / (a.n_xt * a.n_yt)) % a.n_zc;
                const int z0 = zc * a.zc_len;
                const int nz = min(a.zc_len, a.dout_z - z0);
                const int np = nz + KS - 1;
#pragma unroll 1
                for (int ip = 0; ip < np; ++ip) {
                    const int kd_lo = ip - (nz - 1) > 0 ? ip - (nz - 1) : 0;
                    const int kd_hi = ip < KS - 1 ? ip : KS - 1;
                    const bool full = kd_lo == 0 && kd_hi == KS - 1;
                    const uint32_t A0 = ac0 + (uint32_t)ip;            // output fed through kd = 0
                    const uint32_t r = A0 % 3u;
                    const uint32_t win = ((r + 1u) % 3u) * N;          // window start (rows) inside [kd0 kd1 kd2 kd0 kd1]
#pragma unroll 1
                    for (int sub = 0; sub < nsub; ++sub, ++pc) {
                        const uint32_t slot = pc % ring, ph = (pc / ring) & 1u;
                        mbar_wait(&plane_full[slot], ph);
                        tc_fence_after();
                        const uint32_t a_pl = a_lo0 + slot * (plane_pitch >> 4);
#pragma unroll 1
                        for (int m = 0; m < 2; ++m) {
                            if (sub == 0 && kd_lo == 0)             // the block of the new output must have been drained
                                mbar_wait(&acc_empty[m * 3 + (2u - r)], ((A0 / 3u) & 1u) ^ 1u);
                            tc_fence_after();
                            if (leader) {
                                const uint32_t d_reg = tmem_base + (uint32_t)m * 256u;
#pragma unroll 1
                                for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                                    for (int kw = 0; kw < KS; ++kw) {
                                        uint32_t b_lo = b_lo0r + ((uint32_t)(kh * KS + kw) * kt + (uint32_t)sub * wsub) * b_step16r;
#pragma unroll
                                        for (int s = 0; s < KSTEPS; ++s) {
                                            const uint32_t a_lo = a_pl + (uint32_t)(kh * SX + kw) + (uint32_t)(2 * s) * (atom_stride >> 4) +
                                                                  (uint32_t)m * 8u;
                                            const uint64_t ad = desc64(a_lo, a_hi);
                                            const bool first = sub == 0 && kh == 0 && kw == 0 && s == 0;
                                            if (full && !first) {
                                                umma_bf16(d_reg, ad, desc64(b_lo + win, b_hi), idesc_3, 1u);
                                            } else {
                                                // plane at a z-chunk edge, or the first MMA of a plane (the kd = 0
                                                // accumulator is overwritten, the others accumulate): one MMA per kd
#pragma unroll
                                                for (int kd = 0; kd < KS; ++kd)
                                                    if (kd >= kd_lo && kd <= kd_hi) {
                                                        const uint32_t bl = 2u - (r + 3u - (uint32_t)kd) % 3u;
                                                        umma_bf16(d_reg + bl * N, ad, desc64(b_lo + (uint32_t)kd * N, b_hi), idesc_1,
                                                                  (first && kd == 0) ? 0u : 1u);
                                                    }
                                            }
                                            b_lo += b_step16r;
                                        }
                                    }
                                if (sub == nsub - 1 && ip >= KS - 1) {   // output ip-(KS-1) has received its last plane
                                    const uint32_t Ad = ac0 + (uint32_t)(ip - (KS - 1));
                                    umma_commit(&acc_full[m * 3 + (2u - Ad % 3u)]);
                                }
                            }
                            __syncwarp();
                        }
                        if (leader) umma_commit(&plane_empty[slot]);
                        __syncwarp();
                    }
                }
                ac0 += (uint32_t)nz;
            }
        } else
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int zc = (item / (a.n_xt * a.n_yt)) % a.n_zc;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            const int np = nz + KS - 1;
#pragma unroll 1
            for (int ip = 0; ip < np; ++ip) {
                const int kd_lo = ip - (nz - 1) > 0 ? ip - (nz - 1) : 0;
                const int kd_hi = ip < KS - 1 ? ip : KS - 1;
                // accumulator blocks of the outputs fed by this plane: kd = kd_lo.. map to ascending
                // blocks starting at blk0; the run breaks once where the block ring wraps to 0
                const uint32_t L = (uint32_t)(kd_hi - kd_lo + 1);
                const uint32_t r = (ac0 + (uint32_t)(ip - kd_lo)) % nblk;
                const uint32_t len0 = L < r + 1u ? L : r + 1u, len1 = L - len0;
                const uint32_t blk0 = nblk - 1u - r;
                const uint32_t kd1 = (uint32_t)kd_lo + len0;
                const uint32_t d_seg0 = tmem_base + blk0 * N, b_seg0 = (uint32_t)kd_lo * N;
                const uint32_t d_seg1 = tmem_base, b_seg1 = kd1 * N;
                const uint32_t i_seg0 = idesc_1 + ((((len0 - 1u) * N) >> 3) << 17);
                const uint32_t i_seg1 = idesc_1 + (((((len1 ? len1 : 1u) - 1u) * N) >> 3) << 17);
                if (kd_lo == 0) {               // a new output starts: its block must have been drained
                    const uint32_t A = ac0 + (uint32_t)ip;
                    mbar_wait(&acc_empty[blk0], ((A / nblk) & 1u) ^ 1u);
                }
#pragma unroll 1
                for (int sub = 0; sub < nsub; ++sub, ++pc) {
                    const uint32_t slot = pc % ring, ph = (pc / ring) & 1u;
                    mbar_wait(&plane_full[slot], ph);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t a_pl = a_lo0 + slot * (plane_pitch >> 4);
#pragma unroll 1
                        for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                            for (int kw = 0; kw < KS; ++kw) {
                                uint32_t b_lo = b_lo0 + ((uint32_t)(kh * KS + kw) * kt + (uint32_t)sub * wsub) * b_step16;
#pragma unroll
                                for (int s = 0; s < KSTEPS; ++s) {
                                    const uint32_t a_lo = a_pl + (uint32_t)(kh * SX + kw) + (uint32_t)(2 * s) * (atom_stride >> 4);
                                    const uint64_t ad0 = desc64(a_lo, a_hi), ad1 = desc64(a_lo + 8u, a_hi);
                                    if (sub == 0 && kh == 0 && kw == 0 && s == 0) {
                                        // first MMA of the plane: the kd = 0 accumulator is overwritten, the
                                        // others accumulate -> one MMA per kd
#pragma unroll
                                        for (int kd = 0; kd < KS; ++kd)
                                            if (kd >= kd_lo && kd <= kd_hi) {
                                                const uint32_t bl = (uint32_t)kd < kd1 ? blk0 + (uint32_t)(kd - kd_lo) : (uint32_t)kd - kd1;
                                                const uint64_t bd = desc64(b_lo + (uint32_t)kd * N, b_hi);
                                                const uint32_t dcol = tmem_base + bl * N;
                                                umma_bf16(dcol, ad0, bd, idesc_1, kd ? 1u : 0u);
                                                if (MT == 2) umma_bf16(dcol + 256u, ad1, bd, idesc_1, kd ? 1u : 0u);
                                            }
                                    } else {
                                        const uint64_t bd0 = desc64(b_lo + b_seg0, b_hi);
                                        umma_bf16(d_seg0, ad0, bd0, i_seg0, 1u);
                                        if (MT == 2) umma_bf16(d_seg0 + 256u, ad1, bd0, i_seg0, 1u);
                                        if (len1) {
                                            const uint64_t bd1 = desc64(b_lo + b_seg1, b_hi);
                                            umma_bf16(d_seg1, ad0, bd1, i_seg1, 1u);
                                            if (MT == 2) umma_bf16(d_seg1 + 256u, ad1, bd1, i_seg1, 1u);
                                        }
                                    }
                                    b_lo += b_step16;
                                }
                            }
                        umma_commit(&plane_empty[slot]);
                        if (sub == nsub - 1 && ip >= KS - 1) {   // output ip-(KS-1) has received its last plane
                            const uint32_t A = ac0 + (uint32_t)(ip - (KS - 1));
                            umma_commit(&acc_full[nblk - 1u - (A % nblk)]);
                        }
                    }
                    __syncwarp();
                }
            }
            ac0 += (uint32_t)nz;
        }
    } else if ((warp - 2) >> 2 < MT) {
        // ===================================== epilogue =========================================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int m = (warp - 2) >> 2;                // which of the M-tiles (x half) this warp drains
        uint32_t A = 0;
        float hold[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hold[j] = 0.f;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int t = item;
            // x patch fastest: CTAs running side by side sweep z in step and touch adjacent rows of HBM
            const int xt = t % a.n_xt; t /= a.n_xt;
            const int yt = t % a.n_yt; t /= a.n_yt;
            const int zc = t % a.n_zc; t /= a.n_zc;
            const int tile = t;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            for (int zo = 0; zo < nz; ++zo, ++A) {
                const uint32_t bl = nblk - 1u - (A % nblk), ph = (A / nblk) & 1u;
                const uint32_t bar_i = ROT ? (uint32_t)m * 3u + bl : bl;       // ROT: barriers per M-tile
                mbar_wait(&acc_full[bar_i], ph);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)m * 256u + bl * N;
                if (a.pool) {
                    if (a.cout == 48)
                        epilogue_tile_pool<3>(tacc, q, lane, s_bias, a.relu, a.out, tile, a.dout >> 1, z0 + zo, yt * kTY,
                                              xt * TX + m * 8, reinterpret_cast<float(&)[12]>(hold), a.dout_z >> 1);
                    else if (a.cout == 32)
                        epilogue_tile_pool<2>(tacc, q, lane, s_bias, a.relu, a.out, tile, a.dout >> 1, z0 + zo, yt * kTY,
                                              xt * TX + m * 8, reinterpret_cast<float(&)[8]>(hold), a.dout_z >> 1);
                    else
                        epilogue_tile_pool<4>(tacc, q, lane, s_bias, a.relu, a.out, tile, a.dout >> 1, z0 + zo, yt * kTY,
                                              xt * TX + m * 8, hold, a.dout_z >> 1);
                } else if (a.acc_mode)
                epilogue_tile_ksplit(tacc, q, lane, a.cout, s_bias, a.relu, a.out, a.partial, a.acc_mode, tile,
                                     a.dout, z0 + zo, yt * kTY, xt * TX + m * 8, a.dout_z);
                else
                epilogue_tile(tacc, q, lane, a.cout, s_bias, a.relu, a.out, tile,
                              a.dout, z0 + zo, yt * kTY, xt * TX + m * 8, a.cout_total, a.cout_off, a.dout_z);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[bar_i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// first layer on the tensor pipe: Conv3D(Cout,(3,3,3)) of the single-channel float32 tile.
// K = 27 taps padded to 32: builder warps write the im2col rows (bf16) of a 16x16 output patch straight
// into shared memory in the UMMA canonical layout, two K=16 MMAs per M-tile do the contraction, the
// common epilogue writes C8-blocked bf16.  The kernel is bound by its bf16 output stream (HBM/L2 write).
//   warp 0      MMA issuer (+ TMEM alloc)      warps 1..4  epilogue      warps 5..8  im2col builders
// ------------------------------------------------------------------------------------------------
constexpr int kFirstThreads = 416;           // warp 0 MMA, warps 1..8 epilogue, warps 9..12 builders
struct FirstArgs {
    const float *in;                 // (tile, din, din, din) float32
    const __nv_bfloat16 *w_packed;   // [2 ksteps][2][Cout][8]; taps 0..26 = kernel*BN scale, 27/28 = bias hi/lo
    const float *bias;               // unused by the kernel (bias rides in K slots 27/28)
    __nv_bfloat16 *out;
    int n_tiles, din, dout, cout;    // din/dout: x/y extent
    int din_z, dout_z;               // z extent
    int n_xt, n_yt, n_zc, zc_len;
    uint32_t tmem_cols;
    VolumeIO vio;                    // vio.img != nullptr: read the tile from the volume instead of `in`
    int dbg;                         // experiments: 1 = skip global fetch, 2 = skip global store, 4 = skip build
};

// first-layer epilogue: the folded-BN bias is already inside the accumulator (two constant-one K
// slots carry bias = hi + lo in bf16), so only ReLU (on packed bf16 pairs) and the store remain.
__device__ __forceinline__ void first_store16(const uint32_t (&r)[16], int c0, __nv_bfloat16 *__restrict__ out,
                                              size_t vox, size_t cg_stride, bool ok) {
    if (!ok) return;
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 b2 = __hmax2(__floats2bfloat162_rn(__uint_as_float(r[h * 8 + 2 * j]),
                                                              __uint_as_float(r[h * 8 + 2 * j + 1])), zero);
            pk[j] = *reinterpret_cast<uint32_t *>(&b2);
        }
        *reinterpret_cast<uint4 *>(out + (vox + (size_t)((c0 >> 3) + h) * cg_stride) * 8) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

__global__ void __launch_bounds__(kFirstThreads, 1)
conv_first_umma_kernel(const FirstArgs a) {
    constexpr int SX = kTX + 2, SY = kTY + 2;
    constexpr int PX = 20;                                     // ring row pitch (floats): 16-byte aligned rows
    constexpr int kRing = 4;
    constexpr int kBuilders = 128;                             // one thread per 2 x-adjacent accumulator rows
    constexpr int kPer = (SY * SX + kBuilders - 1) / kBuilders; // input elements per builder thread and plane
    constexpr uint32_t kAStage = 2 * 4 * 128 * 16;            // 2 M-tiles x 4 atoms x 128 rows x 16 B
    constexpr int kSt = 4;                                     // pipeline depth (A stages and TMEM stages)
    extern __shared__ __align__(128) uint8_t s_a[];            // kSt * kAStage bytes (dynamic)
    __shared__ __align__(128) uint8_t s_w[2 * 2 * 64 * 16];    // Cout <= 64
    __shared__ __align__(16) float s_in[kRing][SY][PX];
    __shared__ float s_lut[256];          // uint8 -> (x-mean)/std, exact IEEE sub/div evaluated once per value
    __shared__ uint64_t bars[4 * kSt];
    __shared__ uint32_t tmem_slot;
    uint64_t *a_full = bars, *a_empty = bars + kSt, *acc_full = bars + 2 * kSt, *acc_empty = bars + 3 * kSt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.n_tiles * a.n_yt * a.n_xt * a.n_zc;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSt; ++i) {
            mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8);
        }
        fence_barrier_init();
    }
    // weights -> smem (generic proxy), made visible to the tensor core with a proxy fence
    for (int i = threadIdx.x; i < a.cout * 4; i += blockDim.x)
        reinterpret_cast<uint4 *>(s_w)[i] = __ldg(reinterpret_cast<const uint4 *>(a.w_packed) + i);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = __fdiv_rn(__fsub_rn((float)i, a.vio.mean), a.vio.stdv);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================================== MMA issuer =======================================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_bf16(128, a.cout);
        const uint32_t a_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B (rows contiguous)
        const uint32_t a_lo0 = (smem_u32(s_a) >> 4) | ((2048u >> 4) << 16);    // LBO = 2048 B between atoms
        const uint32_t b_hi = (128u >> 4) | (1u << 14);
        const uint32_t b_lo0 = (smem_u32(s_w) >> 4) | ((uint32_t)a.cout << 16);
        const uint32_t b_step16 = (uint32_t)a.cout * 2u;
        uint32_t ac = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int zc = (item / (a.n_xt * a.n_yt)) % a.n_zc;
            const int nz = min(a.zc_len, a.dout_z - zc * a.zc_len);
            for (int zo = 0; zo < nz; ++zo, ++ac) {
                const uint32_t st = ac % kSt, ph = (ac / kSt) & 1u;
                mbar_wait(&acc_empty[st], ph ^ 1u);
                mbar_wait(&a_full[st], ph);
                tc_fence_after();
                if (leader) {
                    const uint32_t d0 = tmem_base + st * (2u * (uint32_t)a.cout);
                    const uint32_t a_st = a_lo0 + st * (kAStage >> 4);
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const uint64_t bdesc = desc64(b_lo0 + s * b_step16, b_hi);
                        umma_bf16(d0, desc64(a_st + s * (4096u >> 4), a_hi), bdesc, idesc, s ? 1u : 0u);
                        umma_bf16(d0 + (uint32_t)a.cout, desc64(a_st + (8192u >> 4) + s * (4096u >> 4), a_hi), bdesc,
                                  idesc, s ? 1u : 0u);
                    }
                    umma_commit(&a_empty[st]);
                    umma_commit(&acc_full[st]);
                }
                __syncwarp();
            }
        }
    } else if (warp <= 8) {
        // ===================================== epilogue =========================================
        const int q = warp & 3;
        const int m = (warp - 1) >> 2;
        const int row = q * 32 + lane;
        const size_t cg_stride = (size_t)a.dout_z * a.dout * a.dout;
        uint32_t ac = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int t = item;
            // x patch fastest: CTAs running side by side sweep z in step and touch adjacent rows of HBM
            const int xt = t % a.n_xt; t /= a.n_xt;
            const int yt = t % a.n_yt; t /= a.n_yt;
            const int zc = t % a.n_zc; t /= a.n_zc;
            const int tile = t;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            const int y = yt * kTY + (row >> 3), x = xt * kTX + m * 8 + (row & 7);
            const bool ok = (x < a.dout) && (y < a.dout) && !(a.dbg & 2);
            const size_t vox0 = (size_t)tile * (a.cout >> 3) * cg_stride + ((size_t)z0 * a.dout + y) * a.dout + x;
            for (int zo = 0; zo < nz; ++zo, ++ac) {
                const uint32_t st = ac % kSt, ph = (ac / kSt) & 1u;
                mbar_wait(&acc_full[st], ph);
                tc_fence_after();
                const uint32_t tcol = tmem_base + st * (2u * (uint32_t)a.cout) + (uint32_t)m * (uint32_t)a.cout +
                                      ((uint32_t)(q * 32) << 16);
                const size_t vox = vox0 + (size_t)zo * a.dout * a.dout;
                uint32_t ra[16], rb[16];
                tmem_ld16(tcol, ra);
#pragma unroll 1
                for (int c0 = 0; c0 < a.cout; c0 += 32) {
                    tmem_ld_wait();
                    const bool more = c0 + 16 < a.cout;
                    if (more) tmem_ld16(tcol + (uint32_t)(c0 + 16), rb);
                    first_store16(ra, c0, a.out, vox, cg_stride, ok);
                    if (more) {
                        tmem_ld_wait();
                        if (c0 + 32 < a.cout) tmem_ld16(tcol + (uint32_t)(c0 + 32), ra);
                        first_store16(rb, c0 + 16, a.out, vox, cg_stride, ok);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[st]);
            }
        }
    } else {
        // ===================================== im2col builders ==================================
        // thread -> M-tile bm, accumulator rows (ly, xq), (ly, xq+1): the 27 taps of 2 x-adjacent outputs
        // need 3 planes x 3 rows x 4 floats, fetched as two 8-byte shared loads each.
        const int bt = threadIdx.x - 288;            // 0..127
        const int bm = bt >> 6, ly = (bt & 63) >> 2, xq = (bt & 3) * 2;
        uint32_t ac = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int t = item;
            // x patch fastest: CTAs running side by side sweep z in step and touch adjacent rows of HBM
            const int xt = t % a.n_xt; t /= a.n_xt;
            const int yt = t % a.n_yt; t /= a.n_yt;
            const int zc = t % a.n_zc; t /= a.n_zc;
            const int tile = t;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            // source of this thread's (up to kPer) elements of every input plane: either the float32
            // tile batch or, with direct volume input, the volume itself at the tile origin of the
            // reference grid (fplnetwork.py:149-157; zero beyond the far edge; (x-mean)/std for uint8)
            long long plane_stride, z_lim, z_org = 0;
            long long eoff[kPer];
            bool eok[kPer];
            int edst[kPer];
            const uint8_t *src8 = nullptr;
            const float *src32 = nullptr;
            long long ox = 0, oy = 0, lim_y, lim_x, row_stride;
            if (a.vio.img) {
                const int tt = a.vio.ids ? a.vio.ids[a.vio.tile0 + tile] : a.vio.tile0 + tile;
                ox = (long long)(tt % a.vio.g.nx) * a.vio.g.out_sz;
                oy = (long long)((tt / a.vio.g.nx) % a.vio.g.ny) * a.vio.g.out_sz;
                z_org = a.vio.g.z_base + (long long)(tt / (a.vio.g.nx * a.vio.g.ny)) * a.vio.g.out_z;
                plane_stride = a.vio.g.Y * a.vio.g.X; row_stride = a.vio.g.X;
                z_lim = a.vio.g.Z - z_org < a.din_z ? a.vio.g.Z - z_org : a.din_z;     // planes available
                lim_y = a.vio.g.Y; lim_x = a.vio.g.X;
                src8 = (const uint8_t *)a.vio.img; src32 = (const float *)a.vio.img;
            } else {
                plane_stride = (long long)a.din * a.din; row_stride = a.din;
                z_lim = a.din_z; lim_y = a.din; lim_x = a.din;
                src32 = a.in + (size_t)tile * a.din_z * a.din * a.din;
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = bt + k * kBuilders;
                const int yy = i / SX, xx = i - yy * SX;
                const int gy = yt * kTY + yy, gx = xt * kTX + xx;
                eok[k] = i < SY * SX && gy < a.din && gx < a.din && oy + gy < lim_y && ox + gx < lim_x;
                eoff[k] = (oy + gy) * row_stride + ox + gx;
                edst[k] = yy * PX + xx;
            }
            const bool is_u8 = a.vio.img && a.vio.is_u8;
            float preA[kPer], preB[kPer], preC[kPer];
            auto fetch = [&](int zin, float (&pre)[kPer]) {   // this thread's share of input plane zin
#pragma unroll
                for (int k = 0; k < kPer; ++k) {
                    float v = 0.f;
                    if (eok[k] && zin < z_lim && !(a.dbg & 1)) {
                        const long long o = (z_org + zin) * plane_stride + eoff[k];
                        if (is_u8) v = s_lut[__ldg(src8 + o)];
                        else v = __ldg(src32 + o);
                    }
                    pre[k] = v;
                }
            };
            auto stash = [&](int zin, const float (&pre)[kPer]) {   // registers -> ring slot zin % kRing
                float *dst = &s_in[zin % kRing][0][0];
#pragma unroll
                for (int k = 0; k < kPer; ++k)
                    if (bt + k * kBuilders < SY * SX) dst[edst[k]] = pre[k];
            };
            // all builders are past their last read of the ring (barrier at the end of the previous plane)
            fetch(z0, preA); stash(z0, preA);
            fetch(z0 + 1, preA); stash(z0 + 1, preA);
            fetch(z0 + 2, preA);
            fetch(z0 + 3, preB);
            fetch(z0 + 4, preC);
            for (int zo = 0; zo < nz; ++zo, ++ac) {
                stash(z0 + zo + 2, preA);
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int k = 0; k < kPer; ++k) { preA[k] = preB[k]; preB[k] = preC[k]; }
                if (zo + 3 < nz) fetch(z0 + zo + 5, preC);    // three planes ahead: covers L2/DRAM latency
                const uint32_t st = ac % kSt, ph = (ac / kSt) & 1u;
                mbar_wait(&a_empty[st], ph ^ 1u);
                uint8_t *abase = s_a + st * kAStage + bm * 8192 + (ly * 8 + xq) * 16;
                float f[3][3][4];
                if (!(a.dbg & 4))
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const float *pl = &s_in[(z0 + zo + kd) % kRing][ly][bm * 8 + xq];
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const float2 va = *reinterpret_cast<const float2 *>(pl + kh * PX);
                        const float2 vb = *reinterpret_cast<const float2 *>(pl + kh * PX + 2);
                        f[kd][kh][0] = va.x; f[kd][kh][1] = va.y; f[kd][kh][2] = vb.x; f[kd][kh][3] = vb.y;
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {            // output row x = xq + j
#pragma unroll
                    for (int k = 0; k < 4; ++k) {        // K atom = taps 8k .. 8k+7
                        uint32_t pk[4];
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            float pv[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int tp = 8 * k + 2 * e2 + u;
                                pv[u] = tp < 27 ? f[tp / 9][(tp / 3) % 3][j + tp % 3] : (tp < 29 ? 1.f : 0.f);
                            }
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(pv[0], pv[1]);
                            pk[e2] = *reinterpret_cast<uint32_t *>(&b2);
                        }
                        *reinterpret_cast<uint4 *>(abase + k * 2048 + j * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[st]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");       // ring is free for the next item
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core kernels around the GEMMs (all bandwidth-bound, C8-blocked bf16)
// ------------------------------------------------------------------------------------------------
// first layer: Conv3D(Cout,(3,3,3)) on the single-channel float32 tile + folded BN + ReLU.
// K = 27 is too thin for the tensor pipe; one thread = one output voxel, all Cout channels.
template <int COUT>
__global__ void __launch_bounds__(128)
conv_first_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ scale,
                  const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out, int n_tiles, int din) {
    __shared__ float sw[27 * COUT];
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i] * scale[i % COUT];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int dout = din - 2;
    const int xb = (dout + 127) / 128;
    long long bid = blockIdx.x;
    const int x = (int)(bid % xb) * 128 + threadIdx.x; bid /= xb;
    const int y = (int)(bid % dout); bid /= dout;
    const int z = (int)(bid % dout); bid /= dout;
    const int t = (int)bid;
    if (x >= dout) return;
    const float *tin = in + (size_t)t * din * din * din;
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = sb[j];
#pragma unroll
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float v = __ldg(tin + ((size_t)(z + kd) * din + (y + kh)) * din + (x + kw));
                const float *wp = sw + ((kd * 3 + kh) * 3 + kw) * COUT;
#pragma unroll
                for (int j = 0; j < COUT; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
            }
#pragma unroll
    for (int cg = 0; cg < COUT / 8; ++cg) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(fmaxf(acc[cg * 8 + 2 * j], 0.f), fmaxf(acc[cg * 8 + 2 * j + 1], 0.f));
            pk[j] = *reinterpret_cast<uint32_t *>(&b2);
        }
        size_t o = ((((size_t)t * (COUT / 8) + cg) * dout + z) * dout + y) * dout + x;
        *reinterpret_cast<uint4 *>(out + o * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&a);
    const __nv_bfloat162 *pb = reinterpret_cast<const __nv_bfloat162 *>(&b);
    __nv_bfloat162 *pr = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

// MaxPooling3D((2,2,2)); one thread = one output atom (8 channels of one voxel)
__global__ void __launch_bounds__(256)
pool_blocked_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, long long n_cg_total, int din, int din_z) {
    // one (tile * channel group, z) output plane per block iteration; 32-bit index math inside the plane
    const int dout = din / 2, dout_z = din_z / 2, plane = dout * dout;
    const long long n_planes = n_cg_total * dout_z;
    for (long long p = blockIdx.x; p < n_planes; p += gridDim.x) {
        const int z = (int)(p % dout_z);
        const long long v = p / dout_z;                 // tile*CG + cg
        const uint4 *ip0 = in + ((size_t)v * din_z + 2 * z) * din * din;
        const uint4 *ip1 = ip0 + (size_t)din * din;
        uint4 *dst = out + (size_t)p * plane;
        for (int i = threadIdx.x; i < plane; i += (int)blockDim.x) {
            const int y = i / dout, x = i - y * dout;
            const int o0 = (2 * y) * din + 2 * x, o1 = o0 + din;
            uint4 m = __ldg(ip0 + o0);
            m = bf16x8_max(m, __ldg(ip0 + o0 + 1));
            m = bf16x8_max(m, __ldg(ip0 + o1));
            m = bf16x8_max(m, __ldg(ip0 + o1 + 1));
            m = bf16x8_max(m, __ldg(ip1 + o0));
            m = bf16x8_max(m, __ldg(ip1 + o0 + 1));
            m = bf16x8_max(m, __ldg(ip1 + o1));
            m = bf16x8_max(m, __ldg(ip1 + o1 + 1));
            dst[i] = m;
        }
    }
}

// concatenate([UpSampling3D(2)(a), Cropping3D(crop)(skip)]): channel groups [0,cga) from a, rest from skip
__global__ void __launch_bounds__(256)
upcat_blocked_kernel(const uint4 *__restrict__ a, int da, int cga, const uint4 *__restrict__ skip, int ds, int cgs,
                     int crop, uint4 *__restrict__ out, int n_tiles) {
    // One (tile, channel group, z) output plane per block iteration: the plane is decoded once (the flat form paid five
    // 64-bit divisions per 16-byte element and ran at a sixth of the HBM rate), the voxels of the plane use 32-bit math.
    const int dout = 2 * da, cg_out = cga + cgs, plane = dout * dout;
    const long long n_planes = (long long)n_tiles * cg_out * dout;
    for (long long p = blockIdx.x; p < n_planes; p += gridDim.x) {
        const int z = (int)(p % dout);
        const long long tc = p / dout;
        const int cg = (int)(tc % cg_out), t = (int)(tc / cg_out);
        const bool up = cg < cga;
        const uint4 *src = up ? a + (((size_t)t * cga + cg) * da + (z >> 1)) * da * da
                              : skip + ((((size_t)t * cgs + (cg - cga)) * ds + z + crop) * ds + crop) * ds + crop;
        const int pitch = up ? da : ds;
        uint4 *dst = out + (size_t)p * plane;
        int i = threadIdx.x;
        for (; i + 3 * (int)blockDim.x < plane; i += 4 * (int)blockDim.x) {      // four 16-byte loads in flight per thread
            uint4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = i + j * (int)blockDim.x, y = e / dout, x = e - y * dout;
                r[j] = __ldg(src + (up ? (y >> 1) * pitch + (x >> 1) : y * pitch + x));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[i + j * (int)blockDim.x] = r[j];
        }
        for (; i < plane; i += (int)blockDim.x) {
            const int y = i / dout, x = i - y * dout;
            dst[i] = __ldg(src + (up ? (y >> 1) * pitch + (x >> 1) : y * pitch + x));
        }
    }
}

// final Conv3D(1,(1,1,1)) + sigmoid (+ nearest up-sampling by `stride`, fplnetwork.py:99-105) -> float32 tile
__global__ void __launch_bounds__(256)
final_blocked_kernel(const uint4 *__restrict__ in, const float *__restrict__ w, float bias, float *__restrict__ out,
                     int n_tiles, int d, int cg_in, int stride, const VolumeIO vio, int d_z) {
    const long long vox = (long long)d_z * d * d;
    const long long total = (long long)n_tiles * vox;
    const int dout = d * stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / vox);
        const int v = (int)(i - (long long)t * vox);            // voxel inside the tile: 32-bit math from here on
        float acc = bias;
        for (int cg = 0; cg < cg_in; ++cg) {
            uint4 q = __ldg(in + ((size_t)t * cg_in + cg) * vox + v);
            const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __bfloat1622float2(p[j]);
                acc = fmaf(f.x, __ldg(w + cg * 8 + 2 * j), acc);
                acc = fmaf(f.y, __ldg(w + cg * 8 + 2 * j + 1), acc);
            }
        }
        const float pr = 1.f / (1.f + expf(-acc));
        const int z = v / (d * d), rem = v - z * d * d, y = rem / d, x = rem - y * d;
        if (vio.pred) {
            // scatter of fplnetwork.py:180-187: pred[off + origin + (0..ext)) <- tile output
            const TileGrid &g = vio.g;
            const int tt = vio.ids ? vio.ids[vio.tile0 + t] : vio.tile0 + t;
            const long long bx = (long long)(tt % g.nx) * g.out_sz + g.off + (long long)x * stride;
            const long long by = (long long)((tt / g.nx) % g.ny) * g.out_sz + g.off + (long long)y * stride;
            const long long bz = g.z_base + (long long)(tt / (g.nx * g.ny)) * g.out_z + g.off + (long long)z * stride;
            for (int ez = 0; ez < stride; ++ez)
                for (int ey = 0; ey < stride; ++ey)
                    for (int ex = 0; ex < stride; ++ex)
                        if (bz + ez < g.Z - g.off && by + ey < g.Y - g.off && bx + ex < g.X - g.off)
                            vio.pred[((bz + ez) * g.Y + (by + ey)) * g.X + bx + ex] = pr;
            continue;
        }
        float *op = out + (size_t)t * ((size_t)d_z * stride) * dout * dout;
        for (int ez = 0; ez < stride; ++ez)
            for (int ey = 0; ey < stride; ++ey)
                for (int ex = 0; ex < stride; ++ex)
                    op[((size_t)(z * stride + ez) * dout + (y * stride + ey)) * dout + (x * stride + ex)] = pr;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused first + second convolution (VGG: Conv3D(1->C1,3^3)+BN+ReLU -> Conv3D(C1->C2,3^3)+BN+ReLU
// [+ MaxPooling3D]).  The first layer's output (96 B per voxel) is the largest tensor of the network;
// here it never leaves the SM: for every z step the CTA computes the halo'd 18x18xC1 plane the second
// convolution needs straight into the shared-memory plane ring (in the UMMA layout TMA would have
// produced), on the tensor pipe, in the shadow of the second convolution's MMAs.
//   warp 0        loads the packed second-layer weights (cp.async.bulk), resident for the CTA's lifetime
//   warp 1        MMA issuer: first-layer MMAs (3 M-tiles x 2 K-steps per plane, K = 27 taps + bias slots)
//                 two planes ahead of the kd-fused second-layer MMAs (as in conv_umma_kernel)
//   warps 2..9    second-layer epilogue (bias, ReLU, optional fused max-pool, bf16 store)
//   warps 10..13  builders: raw input planes -> ring; im2col rows of the 324 plane voxels -> A1 tiles
//   warps 14..17  first-layer epilogue: TMEM -> ReLU -> bf16 -> plane ring slot, then signal plane_full
// Arithmetic is identical to the two separate kernels (same MMA shapes, same K order, same rounding),
// so the result is bit-identical to the unfused path.
// TMEM: second-layer accumulators 2 M-tile regions x 4 blocks x C2 columns (384); first-layer tiles rotate
// through 2 slots at 384 + 48 s.
// ------------------------------------------------------------------------------------------------
// resnet_like: cur = relu(cur + crop(skip)) on C8-blocked bf16 tensors (tile, C/8, z, y, x, 8), in place on cur
__global__ void __launch_bounds__(256)
add_blocked_kernel(uint4 *__restrict__ cur, int d, int dz, const uint4 *__restrict__ skip, int ds, int dsz, int crop,
                   long long n_groups) {
    const long long per = (long long)dz * d * d, total = n_groups * per;
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / per; long long v = i - g * per;
        const int x = (int)(v % d); v /= d;
        const int y = (int)(v % d); const int z = (int)(v / d);
        uint4 a = cur[i];
        const uint4 b = skip[((g * dsz + z + crop) * ds + y + crop) * (long long)ds + x + crop];
        __nv_bfloat162 *pa = reinterpret_cast<__nv_bfloat162 *>(&a);
        const __nv_bfloat162 *pb = reinterpret_cast<const __nv_bfloat162 *>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // sum in fp32, one rounding (the operands are bf16 values)
            const float2 fa = __bfloat1622float2(pa[j]), fb = __bfloat1622float2(pb[j]);
            pa[j] = __hmax2(__floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y), zero);
        }
        cur[i] = a;
    }
}

constexpr int kFusedThreads = 576;
struct FusedArgs {
    const float *in;                 // (tile, din_z, din, din) float32 raw tiles
    const __nv_bfloat16 *w1_packed;  // first layer  [2][2][C1][8] (bias in K slots 27/28)
    const __nv_bfloat16 *w2_packed;  // second layer operand-B image (kd-major chunks)
    const float *bias2;
    __nv_bfloat16 *out;              // second layer output (pooled when pool)
    uint32_t w2_bytes;
    int n_tiles, din, din_z;         // raw tile x/y and z extent
    int dout, dout_z;                // second-layer output extents (din - 4)
    int n_xt, n_yt, n_zc, zc_len;
    int pool;
    VolumeIO vio;
};

// ROT: the second convolution uses the rotating-window scheme of conv_umma_kernel<.., ROT> (3 blocks per M-tile,
// weights packed with 5 kd row blocks, barriers per M-tile, M-tiles issued one after the other); used where the
// 5/3 larger weight image fits beside the rings (32/32 channels), not for 48/48.
template <int C1, int C2, bool ROT>
__global__ void __launch_bounds__(kFusedThreads, 1)
conv_fused12_kernel(const FusedArgs a) {
    constexpr int KS = 3, KSTEPS = C1 / 16, TX = 16;
    constexpr int SX = TX + 2, SY = kTY + 2;                       // first-layer plane the second conv reads
    constexpr int NV = SY * SX;                                    // 324 plane voxels
    constexpr int NT1 = (NV + 127) / 128;                          // 3 first-layer M-tiles
    constexpr int RX = SX + 2, RY = SY + 2, RP = 20;               // raw plane 20 x 20, pitch 20 floats
    constexpr int kRing = 4;
    constexpr uint32_t atom_stride = NV * 16;
    constexpr uint32_t plane_bytes = (C1 / 8) * atom_stride;
    constexpr uint32_t plane_pitch = (plane_bytes + 127u) & ~127u;
    constexpr uint32_t kA1Bytes = NT1 * 4 * 128 * 16;
    // second-layer accumulators: 4 blocks per M-tile, so the block a new output starts in was drained a whole
    // plane earlier (with 3 the MMA stream stalled on the epilogue once per plane); the first layer's three
    // M-tiles per plane rotate through 2 accumulator slots
    constexpr uint32_t N = C2, NBLK = ROT ? 3 : 4, kRegion = NBLK * N, kL1Col = 2 * kRegion, kL1Slots = 2;
    static_assert(kL1Col + kL1Slots * C1 <= 512, "TMEM budget");
    static_assert(RX <= RP && C1 % 16 == 0 && C2 % 16 == 0, "shape");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t w_region = (a.w2_bytes + 127u) & ~127u;
    uint8_t *s_w2 = smem_raw;
    uint8_t *s_planes = s_w2 + w_region;
    uint8_t *s_a1 = s_planes + 2 * plane_pitch;
    uint8_t *s_w1 = s_a1 + kA1Bytes;
    float *s_rawp = reinterpret_cast<float *>(s_w1 + 2 * 2 * C1 * 16);
    float *s_lut = s_rawp + kRing * RY * RP;
    float *s_bias = s_lut + 256;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_bias + 128);
    uint64_t *w_full = bars;                 // [1]
    uint64_t *plane_full = bars + 1;         // [2]  first-layer epilogue -> second-layer MMAs
    uint64_t *plane_empty = bars + 3;        // [2]  second-layer MMAs done with the slot
    uint64_t *a1_full = bars + 5;            // [1]  builders -> first-layer MMAs
    uint64_t *a1_empty = bars + 6;           // [1]
    uint64_t *l1_full = bars + 7;            // [3]  first-layer accumulator tiles
    uint64_t *l1_empty = bars + 10;          // [3]
    uint64_t *acc_full = bars + 13;          // [8]  second-layer accumulator blocks (ROT: [2 M-tiles][3])
    uint64_t *acc_empty = bars + 21;         // [8]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.n_tiles * a.n_yt * a.n_xt * a.n_zc;

    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&plane_full[i], 4); mbar_init(&plane_empty[i], 1); }
        mbar_init(a1_full, 4); mbar_init(a1_empty, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(&l1_full[i], 1); mbar_init(&l1_empty[i], 4); }
        for (int i = 0; i < 8; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], ROT ? 4 : 8); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < C1 * 4; i += blockDim.x)
        reinterpret_cast<uint4 *>(s_w1)[i] = __ldg(reinterpret_cast<const uint4 *>(a.w1_packed) + i);
    for (int i = threadIdx.x; i < (int)(kA1Bytes / 16); i += blockDim.x)
        reinterpret_cast<uint4 *>(s_a1)[i] = make_uint4(0, 0, 0, 0);          // rows >= 324 stay zero
    for (int i = threadIdx.x; i < C2; i += blockDim.x) s_bias[i] = a.bias2[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = __fdiv_rn(__fsub_rn((float)i, a.vio.mean), a.vio.stdv);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== weight loader =====================================
        if (elect_one()) {
            mbar_expect_tx(w_full, a.w2_bytes);
            for (uint32_t off = 0; off < a.w2_bytes; off += 32768u) {
                uint32_t n = a.w2_bytes - off < 32768u ? a.w2_bytes - off : 32768u;
                bulk_load_1d(s_w2 + off, reinterpret_cast<const uint8_t *>(a.w2_packed) + off, n, w_full);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        const bool leader = elect_one();
        const uint32_t idesc_1 = make_idesc_bf16(128, (int)N);
        const uint32_t idesc_l1 = make_idesc_bf16(128, C1);
        const uint32_t a_hi = (SX * 16u >> 4) | (1u << 14);
        const uint32_t b_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = (smem_u32(s_planes) >> 4) | ((atom_stride >> 4) << 16);
        const uint32_t idesc_3 = make_idesc_bf16(128, (int)(3u * N));
        constexpr uint32_t kKB = ROT ? 5u : (uint32_t)KS;           // kd row blocks per weight chunk
        const uint32_t b_lo0 = (smem_u32(s_w2) >> 4) | ((N * kKB) << 16);
        const uint32_t b_step16 = N * kKB * 2u;
        const uint32_t a1_hi = (128u >> 4) | (1u << 14);
        const uint32_t a1_lo0 = (smem_u32(s_a1) >> 4) | ((2048u >> 4) << 16);
        const uint32_t b1_lo0 = (smem_u32(s_w1) >> 4) | ((uint32_t)C1 << 16);
        mbar_wait(w_full, 0);
        uint32_t pc1 = 0;          // first-layer planes issued
        uint32_t pc2 = 0;          // second-layer planes issued
        uint32_t ac0 = 0;
        auto issue_l1 = [&]() {    // first-layer MMAs of plane pc1 (all three M-tiles)
            mbar_wait(a1_full, pc1 & 1u);
#pragma unroll 1
            for (int t = 0; t < NT1; ++t) {
                const uint32_t g = pc1 * NT1 + (uint32_t)t, sl = g % kL1Slots, use = g / kL1Slots;
                mbar_wait(&l1_empty[sl], (use & 1u) ^ 1u);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int s = 0; s < 2; ++s)
                        umma_bf16(tmem_base + kL1Col + sl * C1,
                                  desc64(a1_lo0 + (uint32_t)t * (8192u >> 4) + s * (4096u >> 4), a1_hi),
                                  desc64(b1_lo0 + s * (uint32_t)(C1 * 2), b_hi), idesc_l1, s ? 1u : 0u);
                    umma_commit(&l1_full[sl]);
                }
                __syncwarp();
            }
            if (leader) umma_commit(a1_empty);
            __syncwarp();
            ++pc1;
        };
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int zc = (item / (a.n_xt * a.n_yt)) % a.n_zc;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            const int np = nz + KS - 1;
            issue_l1();
            issue_l1();                              // np >= 3 always
            if constexpr (ROT) {
#pragma unroll 1
            for (int ip = 0; ip < np; ++ip, ++pc2) {
                const uint32_t slot = pc2 & 1u, ph = (pc2 >> 1) & 1u;
                const int kd_lo = ip - (nz - 1) > 0 ? ip - (nz - 1) : 0;
                const int kd_hi = ip < KS - 1 ? ip : KS - 1;
                const bool full = kd_lo == 0 && kd_hi == KS - 1;
                const uint32_t A0 = ac0 + (uint32_t)ip;            // output fed through kd = 0
                const uint32_t r = A0 % NBLK;
                const uint32_t win = ((r + 1u) % 3u) * N;          // rotating window, see conv_umma_kernel<.., ROT>
                mbar_wait(&plane_full[slot], ph);
                tc_fence_after();
                const uint32_t a_pl = a_lo0 + slot * (plane_pitch >> 4);
#pragma unroll 1
                for (int m = 0; m < 2; ++m) {
                    if (kd_lo == 0) mbar_wait(&acc_empty[m * 3 + (2u - r)], ((A0 / NBLK) & 1u) ^ 1u);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t d_reg = tmem_base + (uint32_t)m * kRegion;
#pragma unroll 1
                        for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                            for (int kw = 0; kw < KS; ++kw) {
                                uint32_t b_lo = b_lo0 + (uint32_t)((kh * KS + kw) * KSTEPS) * b_step16;
#pragma unroll
                                for (int s = 0; s < KSTEPS; ++s) {
                                    const uint32_t a_lo = a_pl + (uint32_t)(kh * SX + kw) + (uint32_t)(2 * s) * (atom_stride >> 4) +
                                                          (uint32_t)m * 8u;
                                    const uint64_t ad = desc64(a_lo, a_hi);
                                    const bool first = kh == 0 && kw == 0 && s == 0;
                                    if (full && !first) {
                                        umma_bf16(d_reg, ad, desc64(b_lo + win, b_hi), idesc_3, 1u);
                                    } else {
#pragma unroll
                                        for (int kd = 0; kd < KS; ++kd)
                                            if (kd >= kd_lo && kd <= kd_hi) {
                                                const uint32_t bl = 2u - (r + 3u - (uint32_t)kd) % 3u;
                                                umma_bf16(d_reg + bl * N, ad, desc64(b_lo + (uint32_t)kd * N, b_hi), idesc_1,
                                                          (first && kd == 0) ? 0u : 1u);
                                            }
                                    }
                                    b_lo += b_step16;
                                }
                            }
                        if (ip >= KS - 1) {
                            const uint32_t Ad = ac0 + (uint32_t)(ip - (KS - 1));
                            umma_commit(&acc_full[m * 3 + (2u - Ad % NBLK)]);
                        }
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(&plane_empty[slot]);
                __syncwarp();
                if (ip + 2 < np) issue_l1();          // first-layer plane ip+2 behind this plane's MMAs
            }
            } else {
#pragma unroll 1
            for (int ip = 0; ip < np; ++ip, ++pc2) {
                const uint32_t slot = pc2 & 1u, ph = (pc2 >> 1) & 1u;
                const int kd_lo = ip - (nz - 1) > 0 ? ip - (nz - 1) : 0;
                const int kd_hi = ip < KS - 1 ? ip : KS - 1;
                const uint32_t L = (uint32_t)(kd_hi - kd_lo + 1);
                const uint32_t r = (ac0 + (uint32_t)(ip - kd_lo)) % NBLK;
                const uint32_t len0 = L < r + 1u ? L : r + 1u, len1 = L - len0;
                const uint32_t blk0 = NBLK - 1u - r;
                const uint32_t kd1 = (uint32_t)kd_lo + len0;
                const uint32_t d_seg0 = tmem_base + blk0 * N, b_seg0 = (uint32_t)kd_lo * N;
                const uint32_t d_seg1 = tmem_base, b_seg1 = kd1 * N;
                const uint32_t i_seg0 = idesc_1 + ((((len0 - 1u) * N) >> 3) << 17);
                const uint32_t i_seg1 = idesc_1 + (((((len1 ? len1 : 1u) - 1u) * N) >> 3) << 17);
                if (kd_lo == 0) {
                    const uint32_t A = ac0 + (uint32_t)ip;
                    mbar_wait(&acc_empty[blk0], ((A / NBLK) & 1u) ^ 1u);
                }
                mbar_wait(&plane_full[slot], ph);
                tc_fence_after();
                if (leader) {
                    const uint32_t a_pl = a_lo0 + slot * (plane_pitch >> 4);
#pragma unroll 1
                    for (int kh = 0; kh < KS; ++kh)
#pragma unroll
                        for (int kw = 0; kw < KS; ++kw) {
                            uint32_t b_lo = b_lo0 + (uint32_t)((kh * KS + kw) * KSTEPS) * b_step16;
#pragma unroll
                            for (int s = 0; s < KSTEPS; ++s) {
                                const uint32_t a_lo = a_pl + (uint32_t)(kh * SX + kw) + (uint32_t)(2 * s) * (atom_stride >> 4);
                                const uint64_t ad0 = desc64(a_lo, a_hi), ad1 = desc64(a_lo + 8u, a_hi);
                                if (kh == 0 && kw == 0 && s == 0) {
#pragma unroll
                                    for (int kd = 0; kd < KS; ++kd)
                                        if (kd >= kd_lo && kd <= kd_hi) {
                                            const uint32_t bl = (uint32_t)kd < kd1 ? blk0 + (uint32_t)(kd - kd_lo) : (uint32_t)kd - kd1;
                                            const uint64_t bd = desc64(b_lo + (uint32_t)kd * N, b_hi);
                                            const uint32_t dcol = tmem_base + bl * N;
                                            umma_bf16(dcol, ad0, bd, idesc_1, kd ? 1u : 0u);
                                            umma_bf16(dcol + kRegion, ad1, bd, idesc_1, kd ? 1u : 0u);
                                        }
                                } else {
                                    const uint64_t bd0 = desc64(b_lo + b_seg0, b_hi);
                                    umma_bf16(d_seg0, ad0, bd0, i_seg0, 1u);
                                    umma_bf16(d_seg0 + kRegion, ad1, bd0, i_seg0, 1u);
                                    if (len1) {
                                        const uint64_t bd1 = desc64(b_lo + b_seg1, b_hi);
                                        umma_bf16(d_seg1, ad0, bd1, i_seg1, 1u);
                                        umma_bf16(d_seg1 + kRegion, ad1, bd1, i_seg1, 1u);
                                    }
                                }
                                b_lo += b_step16;
                            }
                        }
                    umma_commit(&plane_empty[slot]);
                    if (ip >= KS - 1) {
                        const uint32_t A = ac0 + (uint32_t)(ip - (KS - 1));
                        umma_commit(&acc_full[NBLK - 1u - (A % NBLK)]);
                    }
                }
                __syncwarp();
                if (ip + 2 < np) issue_l1();          // first-layer plane ip+2 behind this plane's MMAs
            }
            }
            ac0 += (uint32_t)nz;
        }
    } else if (warp < 10) {
        // ===================================== second-layer epilogue =============================
        const int q = warp & 3;
        const int m = (warp - 2) >> 2;
        uint32_t A = 0;
        float hold[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hold[j] = 0.f;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int t = item;
            const int xt = t % a.n_xt; t /= a.n_xt;
            const int yt = t % a.n_yt; t /= a.n_yt;
            const int zc = t % a.n_zc; t /= a.n_zc;
            const int tile = t;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            for (int zo = 0; zo < nz; ++zo, ++A) {
                const uint32_t bl = NBLK - 1u - (A % NBLK), ph = (A / NBLK) & 1u;
                const uint32_t bar_i = ROT ? (uint32_t)m * 3u + bl : bl;
                mbar_wait(&acc_full[bar_i], ph);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)m * kRegion + bl * N;
                if (a.pool)
                    epilogue_tile_pool<C2 / 16>(tacc, q, lane, s_bias, 1, a.out, tile, a.dout >> 1, z0 + zo, yt * kTY,
                                                xt * TX + m * 8, reinterpret_cast<float(&)[C2 / 4]>(hold), a.dout_z >> 1);
                else
                    epilogue_tile(tacc, q, lane, C2, s_bias, 1, a.out, tile, a.dout, z0 + zo, yt * kTY, xt * TX + m * 8,
                                  C2, 0, a.dout_z);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[bar_i]);
            }
        }
    } else if (warp < 14) {
        // ===================================== builders ==========================================
        // raw planes: (20 x 20) floats per z, ring of 4; per first-layer plane p the 27 taps of the 324 voxels
        // (flat v = y1*18 + x1) are written as im2col rows of A1; a task = two x-adjacent voxels.
        const int bt = threadIdx.x - 320;             // 0..127
        constexpr int kPer = (RY * RX + 127) / 128;   // raw elements per thread and plane (4)
        constexpr int kTasks = (SY / 2) * SX;         // 162 y-pairs
        uint32_t pc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int t = item;
            const int xt = t % a.n_xt; t /= a.n_xt;
            const int yt = t % a.n_yt; t /= a.n_yt;
            const int zc = t % a.n_zc; t /= a.n_zc;
            const int tile = t;
            const int z0 = zc * a.zc_len;
            const int nz = min(a.zc_len, a.dout_z - z0);
            const int np = nz + KS - 1;                // first-layer planes of this item
            const int nraw = np + 2;                   // raw planes z0 .. z0+np+1
            long long plane_stride, z_lim, z_org = 0, ox = 0, oy = 0, lim_y, lim_x, row_stride;
            const uint8_t *src8 = nullptr; const float *src32 = nullptr;
            if (a.vio.img) {
                const int tt = a.vio.ids ? a.vio.ids[a.vio.tile0 + tile] : a.vio.tile0 + tile;
                ox = (long long)(tt % a.vio.g.nx) * a.vio.g.out_sz;
                oy = (long long)((tt / a.vio.g.nx) % a.vio.g.ny) * a.vio.g.out_sz;
                z_org = a.vio.g.z_base + (long long)(tt / (a.vio.g.nx * a.vio.g.ny)) * a.vio.g.out_z;
                plane_stride = a.vio.g.Y * a.vio.g.X; row_stride = a.vio.g.X;
                z_lim = a.vio.g.Z - z_org < a.din_z ? a.vio.g.Z - z_org : a.din_z;
                lim_y = a.vio.g.Y; lim_x = a.vio.g.X;
                src8 = (const uint8_t *)a.vio.img; src32 = (const float *)a.vio.img;
            } else {
                plane_stride = (long long)a.din * a.din; row_stride = a.din;
                z_lim = a.din_z; lim_y = a.din; lim_x = a.din;
                src32 = a.in + (size_t)tile * a.din_z * a.din * a.din;
            }
            long long eoff[kPer]; bool eok[kPer]; int edst[kPer];
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = bt + k * 128;
                const int yy = i / RX, xx = i - yy * RX;
                const int gy = yt * kTY + yy, gx = xt * TX + xx;
                eok[k] = i < RY * RX && gy < a.din && gx < a.din && oy + gy < lim_y && ox + gx < lim_x;
                eoff[k] = (oy + gy) * row_stride + ox + gx;
                edst[k] = yy * RP + xx;
            }
            const bool is_u8 = a.vio.img && a.vio.is_u8;
            float preA[kPer], preB[kPer], preC[kPer];
            auto fetch = [&](int zin, float (&pre)[kPer]) {
#pragma unroll
                for (int k = 0; k < kPer; ++k) {
                    float v = 0.f;
                    if (eok[k] && zin < z_lim && zin < z0 + nraw) {
                        const long long o = (z_org + zin) * plane_stride + eoff[k];
                        if (is_u8) v = s_lut[__ldg(src8 + o)];
                        else v = __ldg(src32 + o);
                    }
                    pre[k] = v;
                }
            };
            auto stash = [&](int zin, const float (&pre)[kPer]) {
                float *dst = s_rawp + (zin % kRing) * RY * RP;
#pragma unroll
                for (int k = 0; k < kPer; ++k)
                    if (bt + k * 128 < RY * RX) dst[edst[k]] = pre[k];
            };
            fetch(z0, preA); stash(z0, preA);
            fetch(z0 + 1, preA); stash(z0 + 1, preA);
            fetch(z0 + 2, preA);
            fetch(z0 + 3, preB);
            fetch(z0 + 4, preC);
            for (int p = 0; p < np; ++p, ++pc) {
                stash(z0 + p + 2, preA);
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int k = 0; k < kPer; ++k) { preA[k] = preB[k]; preB[k] = preC[k]; }
                fetch(z0 + p + 5, preC);
                mbar_wait(a1_empty, (pc & 1u) ^ 1u);
#pragma unroll 1
                for (int task = bt; task < kTasks; task += 128) {
                    // a task = two y-adjacent voxels (y1, x1), (y1+1, x1); consecutive lanes take consecutive x, so
                    // the raw-plane loads (4-byte stride) and the 16-byte im2col stores (16-byte stride) of a warp
                    // are bank-conflict free (x-adjacent pairs cost 2x the shared-memory wavefronts, profiles/)
                    const int yp = task / SX, x1 = task - yp * SX, y1 = 2 * yp;
                    float f[3][4][3];
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        const float *pl = s_rawp + ((z0 + p + kd) % kRing) * RY * RP + y1 * RP + x1;
#pragma unroll
                        for (int r4 = 0; r4 < 4; ++r4) {
                            f[kd][r4][0] = pl[r4 * RP]; f[kd][r4][1] = pl[r4 * RP + 1]; f[kd][r4][2] = pl[r4 * RP + 2];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int v = (y1 + j) * SX + x1;
                        uint8_t *arow = s_a1 + (v >> 7) * 8192 + (v & 127) * 16;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                float pv[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int tp = 8 * k + 2 * e2 + u;
                                    pv[u] = tp < 27 ? f[tp / 9][j + (tp / 3) % 3][tp % 3] : (tp < 29 ? 1.f : 0.f);
                                }
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(pv[0], pv[1]);
                                pk[e2] = *reinterpret_cast<uint32_t *>(&b2);
                            }
                            *reinterpret_cast<uint4 *>(arow + k * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(a1_full);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");       // raw ring is free for the next item
        }
    } else {
        // ===================================== first-layer epilogue ==============================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
        uint32_t pc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int zc = (item / (a.n_xt * a.n_yt)) % a.n_zc;
            const int nz = min(a.zc_len, a.dout_z - zc * a.zc_len);
            const int np = nz + KS - 1;
            for (int p = 0; p < np; ++p, ++pc) {
                const uint32_t slot = pc & 1u, ph = (pc >> 1) & 1u;
                mbar_wait(&plane_empty[slot], ph ^ 1u);          // second-layer MMAs are done with this slot
                uint8_t *pl = s_planes + slot * plane_pitch;
#pragma unroll 1
                for (int t = 0; t < NT1; ++t) {
                    const uint32_t g = pc * NT1 + (uint32_t)t, sl = g % kL1Slots, use = g / kL1Slots;
                    mbar_wait(&l1_full[sl], use & 1u);
                    tc_fence_after();
                    const int v = t * 128 + row;
                    const uint32_t tcol = tmem_base + kL1Col + sl * C1 + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                    for (int c0 = 0; c0 < C1; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(tcol + (uint32_t)c0, r);
                        tmem_ld_wait();
                        if (v < NV) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                uint32_t pk[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    __nv_bfloat162 b2 = __hmax2(__floats2bfloat162_rn(__uint_as_float(r[h * 8 + 2 * j]),
                                                                                      __uint_as_float(r[h * 8 + 2 * j + 1])), zero);
                                    pk[j] = *reinterpret_cast<uint32_t *>(&b2);
                                }
                                *reinterpret_cast<uint4 *>(pl + (size_t)((c0 >> 3) + h) * atom_stride + (size_t)v * 16) =
                                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&l1_empty[sl]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&plane_full[slot]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Generic CUDA-core convolution on the blocked layout.  Used (a) for layers whose packed weights do
// not fit next to the input ring in shared memory (until their streaming variant lands) and (b) by
// the tests as an independent check of the tcgen05 kernel on identical bf16 inputs.
template <int KS>
__global__ void __launch_bounds__(512)
conv_blocked_direct_kernel(const __nv_bfloat16 *__restrict__ in, const float *__restrict__ w,
                           const float *__restrict__ scale, const float *__restrict__ bias,
                           __nv_bfloat16 *__restrict__ out, int n_tiles, int din, int cin, int cout, int relu) {
    const int dout = din - (KS - 1);
    const int xb = (dout + 31) / 32;
    long long bid = blockIdx.x;
    const int x = (int)(bid % xb) * 32 + threadIdx.x; bid /= xb;
    const int y = (int)(bid % dout); bid /= dout;
    const int z = (int)(bid % dout); bid /= dout;
    const int t = (int)bid;
    const int cg = threadIdx.y;                  // output channel group
    if (x >= dout) return;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int cgi = cin >> 3;
    for (int kd = 0; kd < KS; ++kd)
        for (int kh = 0; kh < KS; ++kh)
            for (int kw = 0; kw < KS; ++kw) {
                const float *wt = w + (size_t)((kd * KS + kh) * KS + kw) * cin * cout + cg * 8;
                for (int ci = 0; ci < cgi; ++ci) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in) +
                                          ((((size_t)t * cgi + ci) * din + z + kd) * din + y + kh) * din + x + kw);
                    const __nv_bfloat16 *e = reinterpret_cast<const __nv_bfloat16 *>(&q);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float av = __bfloat162float(e[k]);
                        const float *wr = wt + (size_t)(ci * 8 + k) * cout;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            acc[j] = fmaf(av, __bfloat162float(__float2bfloat16_rn(__ldg(wr + j) * __ldg(scale + cg * 8 + j))), acc[j]);
                    }
                }
            }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v0 = acc[2 * j] + bias[cg * 8 + 2 * j], v1 = acc[2 * j + 1] + bias[cg * 8 + 2 * j + 1];
        if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
        pk[j] = *reinterpret_cast<uint32_t *>(&b2);
    }
    size_t o = ((((size_t)t * (cout >> 3) + cg) * dout + z) * dout + y) * dout + x;
    *reinterpret_cast<uint4 *>(out + o * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static constexpr size_t kMaxDynSmem = 232448;   // 227 KB

// How one GEMM-shaped convolution is mapped onto conv_umma_kernel so that the packed weights of a
// launch stay resident in shared memory beside the input ring:
//   n_split  : the output channels are computed in n_split launches of n = cout/n_split columns
//   nsub     : the input channels are streamed as nsub sub-planes of 16*ksteps channels
//   tx       : patch width (16: two M-tiles per plane, 8: one)        ring: (sub-)plane ring depth
struct ConvPlan { int n_split, n, nsub, ksteps, tx, ring; size_t smem; bool ok; bool rot; int k_split; };

// rot: the packed weights hold 5 instead of 3 kd row blocks (see the ROT variant of conv_umma_kernel)
static size_t plan_smem(int ks, int cin, int n, int nsub, int tx, int ring, bool rot) {
    const int sx = tx + ks - 1, sy = kTY + ks - 1;
    const size_t kd_blocks = rot ? 5 : ks;
    size_t w = ((size_t)ks * ks * kd_blocks * cin * n * 2 + 127) & ~size_t(127);
    size_t plane = (((size_t)(cin / nsub / 8) * sy * sx * 16) + 127) & ~size_t(127);
    return w + ring * plane + 384 + 512;        // + barriers + bias
}

static bool have_instance(int ks, int ksteps, int tx) {
    if (ks == 3 && tx == 16) return ksteps == 1 || ksteps == 2 || ksteps == 3 || ksteps == 4 || ksteps == 6;   // 1: ROT only
    if (ks == 3 && tx == 8) return ksteps == 4 || ksteps == 3;
    if (ks == 1 && tx == 16) return ksteps == 2 || ksteps == 3 || ksteps == 4 || ksteps == 6;
    return false;
}

// Modelled cycles per (tap, K-step, M-tile) unit, times the number of Cout launches.  The kernel is
// bound by the tensor core's operand reads from shared memory (128 B / cycle): per unit the A tile
// (128 rows x 32 B) is read once per MMA the unit is issued as, the B tile (kd-fused N rows x 32 B) once;
// the MMA itself takes N/2 cycles.  Without ROT a unit splits in two MMAs where the accumulator ring
// wraps (2 of nblk planes).
static double plan_cost(const ConvParams &c, const ConvPlan &p) {
    const double N = (c.k == 3 ? 3.0 : 1.0) * p.n;
    const double units = (double)c.k * c.k * (c.cin / 16);
    double a_reads = 1.0;
    if (c.k == 3) {
        int nblk = 256 / p.n; if (nblk > 8) nblk = 8;
        a_reads = p.rot ? 1.0 + 2.0 / units : 1.0 + 2.0 / nblk + 2.0 / units;
    }
    const double smem_cyc = (a_reads * 4096.0 + N * 32.0) / 128.0, mma_cyc = N / 2.0;
    double cost = (smem_cyc > mma_cyc ? smem_cyc : mma_cyc) * p.n_split;
    if (p.tx == 8) cost *= 1.15;          // wider relative halo per plane load
    if (p.ring == 2) cost *= 1.03;        // shallower prefetch
    if (p.nsub > 1) cost *= 1.02;
    return cost;
}

// K-split plan: k_split launches over input-channel chunks, each with the full Cout (wide N), fp32 partial
// sums round-trip through HBM between launches.  Cost per M-tile plane in cycles: per launch
// max(MMA phase, partial traffic at ~23 B/clk/SM x 1.5 safety).
static double ksplit_cost(const ConvParams &c, const ConvPlan &p) {
    const int cc = c.cin / p.k_split;
    const double N = 3.0 * p.n, units = 9.0 * (cc / 16);
    int nblk = 256 / p.n; if (nblk > 8) nblk = 8;
    const double a_reads = p.rot ? 1.0 + 2.0 / units : 1.0 + 2.0 / nblk + 2.0 / units;
    const double smem_cyc = (a_reads * 4096.0 + N * 32.0) / 128.0, mma_cyc = N / 2.0;
    const double mma = units * (smem_cyc > mma_cyc ? smem_cyc : mma_cyc);
    const double pbytes = 128.0 * p.n * 4.0;            // one fp32 pass over an M-tile
    double total = 0.0;
    for (int g = 0; g < p.k_split; ++g) {
        const double bytes = (g == 0 ? pbytes : g == p.k_split - 1 ? pbytes + 128.0 * p.n * 2.0 : 2.0 * pbytes);
        const double hbm = 1.5 * bytes / 23.0;
        total += mma > hbm ? mma : hbm;
    }
    return total * (p.ring == 2 ? 1.03 : 1.0) / (9.0 * (c.cin / 16)) ;   // per unit of the unsplit problem, like plan_cost
}

static ConvPlan plan_conv(const ConvParams &c) {
    ConvPlan best{0, 0, 0, 0, 0, 0, 0, false, false, 1};
    if (c.cin % 16 || c.cout % 16 || c.cout > 128 || (c.k != 1 && c.k != 3)) return best;
    static const int legacy = getenv("FPL_PLAN_LEGACY") ? atoi(getenv("FPL_PLAN_LEGACY")) : 0;
    int txs[2] = {16, 8};
    if (getenv("FPL_DBG_TX8")) { txs[0] = 8; txs[1] = 16; }
    double best_cost = 0.0;
    for (int ti = 0; ti < 2; ++ti)
        for (int nsub = 1; nsub <= 4; ++nsub) {
            if (c.cin % (16 * nsub)) continue;
            const int ksteps = c.cin / 16 / nsub;
            if (!have_instance(c.k, ksteps, txs[ti])) continue;
            for (int split = 1; split <= 8; split *= 2) {
                if (c.cout % (16 * split)) continue;
                const int n = c.cout / split;
                for (int rot = 1; rot >= 0; --rot) {
                    if (rot && (c.k != 3 || txs[ti] != 16 || legacy || c.no_rot || 3 * n > 256)) continue;
                    if (!rot && ksteps == 1) continue;           // 16-channel sub-planes exist as a ROT instance only
                    for (int ring = 3; ring >= 2; --ring) {
                        if (nsub > 1 && ring > 2) continue;
                        const size_t sm = plan_smem(c.k, c.cin, n, nsub, txs[ti], ring, rot != 0);
                        if (sm > kMaxDynSmem) continue;
                        const ConvPlan cand{split, n, nsub, ksteps, txs[ti], ring, sm, true, rot != 0, 1};
                        // legacy: no split, single plane, wide patch, deep ring (loop order) -- first hit wins
                        if (legacy) return cand;
                        const double cost = plan_cost(c, cand);
                        if (!best.ok || cost < best_cost - 1e-9) { best = cand; best_cost = cost; }
                    }
                }
            }
        }
    // K-split candidates (3x3x3, no Cout split): taken only when clearly better than the best resident plan
    static const int no_ksplit = getenv("FPL_NO_KSPLIT") ? 1 : 0;
    static const double ks_margin = getenv("FPL_KSPLIT_MARGIN") ? atof(getenv("FPL_KSPLIT_MARGIN")) : 0.8;
    if (best.ok && !legacy && !no_ksplit && !c.no_ksplit && !c.no_rot && c.k == 3 && c.cout <= 80) {
        for (int ks_ = 2; ks_ <= 4; ks_ *= 2) {
            if (c.cin % (16 * ks_)) continue;
            const int cc = c.cin / ks_, ksteps = cc / 16;
            if (!have_instance(3, ksteps, 16)) continue;
            for (int rot = 1; rot >= 0; --rot)
                for (int ring = 3; ring >= 2; --ring) {
                    const size_t sm = plan_smem(3, cc, c.cout, 1, 16, ring, rot != 0);
                    if (sm > kMaxDynSmem || (rot && 3 * c.cout > 256) || (!rot && ksteps == 1)) continue;
                    const ConvPlan cand{1, c.cout, 1, ksteps, 16, ring, sm, true, rot != 0, ks_};
                    const double cost = ksplit_cost(c, cand);
                    if (cost < ks_margin * best_cost) { best = cand; best_cost = cost / ks_margin; }
                }
        }
    }
    return best;
}

static bool umma_supported(const ConvParams &c) { return plan_conv(c).ok; }

static int g_force_direct = 0;    // test hook: run every GEMM-shaped conv through the CUDA-core kernel
static int g_no_pool_fusion = 0;  // test hook: keep MaxPooling3D as its own kernel
static int g_no_conv12_fusion = 0; // test hook: run the first two convolutions as separate kernels

// first + second convolution suit conv_fused12_kernel (instantiated for 48/48 and 32/32 channels)
static bool fusable12(const ConvParams &c1, const ConvParams &c2) {
    return c1.cin == 1 && c1.k == 3 && (c1.cout == 48 || c1.cout == 32) && c2.k == 3 && c2.cin == c1.cout &&
           c2.cout == c1.cout;
}

static int pack_weights_hilo(fpl_net *net);

int pack_weights_umma(fpl_net *net) {
    if (net->precision == FPL_PREC_TF32) return pack_weights_hilo(net);
    if (net->ops.size() >= 2 && net->ops[0].kind == OP_CONV && net->ops[1].kind == OP_CONV) {
        ConvParams &c1 = net->convs[net->ops[0].conv_index], &c2 = net->convs[net->ops[1].conv_index];
        c2.no_rot = fusable12(c1, c2) && c1.cout == 48;       // the 32/32 instance of the fused kernel is ROT
    }
    for (ConvParams &c : net->convs) {
        if (c.cin == 1 && c.k == 3 && c.cout % 16 == 0) {   // first layer: K = 27 taps padded to 32
            std::vector<__nv_bfloat16> pk((size_t)2 * 2 * c.cout * 8);
            for (int s = 0; s < 2; ++s)
                for (int h = 0; h < 2; ++h)
                    for (int n = 0; n < c.cout; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int tp = 16 * s + 8 * h + e;
                            float v = tp < 27 ? c.kernel[(size_t)tp * c.cout + n] * c.scale[n] : 0.f;
                            // folded-BN bias rides in two constant-one K slots as bf16 hi + lo parts
                            const float b_hi = __bfloat162float(__float2bfloat16_rn(c.bias[n]));
                            if (tp == 27) v = b_hi;
                            if (tp == 28) v = c.bias[n] - b_hi;
                            pk[(((size_t)s * 2 + h) * c.cout + n) * 8 + e] = __float2bfloat16_rn(v);
                        }
            c.packed_bytes = pk.size() * sizeof(__nv_bfloat16);
            FPL_CUDA_CHECK(cudaMalloc(&c.d_packed, c.packed_bytes));
            FPL_CUDA_CHECK(cudaMemcpy(c.d_packed, pk.data(), c.packed_bytes, cudaMemcpyHostToDevice));
            continue;
        }
        if (c.cin % 16 || c.cout % 16) continue;         // final layer
        // operand-B image per Cout split g: chunk (kh,kw,kstep) = [2 K-halves][k*n rows, kd-major][8 bf16];
        // the k kd taps that share an A tile are adjacent row blocks so that one MMA with N = k*n covers them
        const ConvPlan plan = plan_conv(c);
        if (!plan.ok) continue;
        // ROT plans store the kd row blocks as [kd0 kd1 kd2 kd0 kd1] (all three cyclic orders are windows)
        // one image per launch: n_split Cout splits (all input channels) or k_split input-channel chunks (all Cout)
        const int ks = c.k, n = plan.n, nkb = plan.rot ? 5 : ks;
        const int n_img = plan.n_split * plan.k_split, cc = c.cin / plan.k_split, ksteps = cc / 16;
        std::vector<__nv_bfloat16> pk((size_t)ks * ks * nkb * cc * n * n_img);
        const size_t img_elems = (size_t)ks * ks * nkb * cc * n;
        for (int g = 0; g < n_img; ++g) {
            const int co0 = plan.k_split > 1 ? 0 : g * n, ci0 = plan.k_split > 1 ? g * cc : 0;
            for (int kh = 0; kh < ks; ++kh)
                for (int kw = 0; kw < ks; ++kw)
                    for (int s = 0; s < ksteps; ++s)
                        for (int h = 0; h < 2; ++h)
                            for (int kb = 0; kb < nkb; ++kb)
                                for (int nn = 0; nn < n; ++nn)
                                    for (int e = 0; e < 8; ++e) {
                                        const int kd = kb % ks;
                                        const int ci = ci0 + 16 * s + 8 * h + e, co = co0 + nn;
                                        const int tap = (kd * ks + kh) * ks + kw;
                                        const float v = c.kernel[((size_t)tap * c.cin + ci) * c.cout + co] * c.scale[co];
                                        const size_t chunk = ((size_t)(kh * ks + kw) * ksteps + s);
                                        pk[g * img_elems + (((chunk * 2 + h) * nkb + kb) * n + nn) * 8 + e] = __float2bfloat16_rn(v);
                                    }
        }
        c.packed_bytes = pk.size() * sizeof(__nv_bfloat16);
        FPL_CUDA_CHECK(cudaMalloc(&c.d_packed, c.packed_bytes));
        FPL_CUDA_CHECK(cudaMemcpy(c.d_packed, pk.data(), c.packed_bytes, cudaMemcpyHostToDevice));
    }
    return FPL_OK;
}

void free_packed_umma(fpl_net *net) {
    for (ConvParams &c : net->convs) {
        if (c.d_packed) cudaFree(c.d_packed);
        if (c.d_packed_lo) cudaFree(c.d_packed_lo);
        c.d_packed = nullptr; c.d_packed_lo = nullptr; c.packed_bytes = 0; c.hilo_chunk = 0;
    }
}

// instantiation table of conv_umma_kernel
static int dispatch_umma(const ConvPlan &plan, int ks, int grid, size_t smem, cudaStream_t st, const CUtensorMap &tmap,
                         const ConvArgs &a) {
#define FPL_LAUNCH_UMMA(KS_, KST_, TX_, ROT_)                                                                    \
        do {                                                                                                      \
            FPL_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<KS_, KST_, TX_, ROT_>,                           \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
            conv_umma_kernel<KS_, KST_, TX_, ROT_><<<grid, kThreads, smem, st>>>(tmap, a);                        \
        } while (0)
    const int kst = plan.ksteps;
    if (plan.rot && kst == 1) FPL_LAUNCH_UMMA(3, 1, 16, true);
    else if (plan.rot && kst == 3) FPL_LAUNCH_UMMA(3, 3, 16, true);
    else if (plan.rot && kst == 2) FPL_LAUNCH_UMMA(3, 2, 16, true);
    else if (plan.rot && kst == 4) FPL_LAUNCH_UMMA(3, 4, 16, true);
    else if (plan.rot && kst == 6) FPL_LAUNCH_UMMA(3, 6, 16, true);
    else if (ks == 3 && plan.tx == 16 && kst == 3) FPL_LAUNCH_UMMA(3, 3, 16, false);
    else if (ks == 3 && plan.tx == 16 && kst == 2) FPL_LAUNCH_UMMA(3, 2, 16, false);
    else if (ks == 3 && plan.tx == 16 && kst == 4) FPL_LAUNCH_UMMA(3, 4, 16, false);
    else if (ks == 3 && plan.tx == 16 && kst == 6) FPL_LAUNCH_UMMA(3, 6, 16, false);
    else if (ks == 3 && plan.tx == 8 && kst == 4) FPL_LAUNCH_UMMA(3, 4, 8, false);
    else if (ks == 3 && plan.tx == 8 && kst == 3) FPL_LAUNCH_UMMA(3, 3, 8, false);
    else if (ks == 1 && kst == 2) FPL_LAUNCH_UMMA(1, 2, 16, false);
    else if (ks == 1 && kst == 3) FPL_LAUNCH_UMMA(1, 3, 16, false);
    else if (ks == 1 && kst == 4) FPL_LAUNCH_UMMA(1, 4, 16, false);
    else if (ks == 1 && kst == 6) FPL_LAUNCH_UMMA(1, 6, 16, false);
    else { set_error("conv_umma: no instantiation for k=%d ksteps=%d tx=%d", ks, kst, plan.tx); return FPL_EINVAL; }
#undef FPL_LAUNCH_UMMA
    return FPL_OK;
}

// activation buffer pool: lives in the context (one pool per device, fpl_ctx::act_pool); defined below
static int pool_take(fpl_ctx *ctx, size_t bytes, cudaStream_t st);

static int launch_conv_umma(fpl_ctx *ctx, const ConvParams &c, const __nv_bfloat16 *in, __nv_bfloat16 *out,
                            int n_tiles, int din, int din_z, int relu, int pool, cudaStream_t st) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FPL_ECUDA; }
    const ConvPlan plan = plan_conv(c);
    FPL_REQUIRE(plan.ok, "conv_umma: no plan for k=%d Cin=%d Cout=%d", c.k, c.cin, c.cout);
    FPL_REQUIRE(!pool || plan.n_split == 1, "conv_umma: pooled epilogue needs an unsplit plan");
    const int ks = c.k, dout = din - (ks - 1), dout_z = din_z - (ks - 1);
    const int sx = plan.tx + ks - 1, sy = kTY + ks - 1;
    const int cin_atoms = c.cin / 8;
    CUtensorMap tmap;
    cuuint64_t gdim[4] = {(cuuint64_t)din * 8, (cuuint64_t)din, (cuuint64_t)din_z, (cuuint64_t)n_tiles * cin_atoms};
    cuuint64_t gstride[3] = {(cuuint64_t)din * 16, (cuuint64_t)din * din * 16, (cuuint64_t)din * din * din_z * 16};
    cuuint32_t box[4] = {(cuuint32_t)sx * 8, (cuuint32_t)sy, 1, (cuuint32_t)(2 * plan.ksteps)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return FPL_ECUDA; }
    ConvArgs a;
    a.out = out;
    a.w_bytes = (uint32_t)(c.packed_bytes / (plan.n_split * plan.k_split));
    a.cin_atom_off = 0; a.acc_mode = 0; a.partial = nullptr; a.shared_w = 0; a.sub_stride = 0;
    int partial_buf = -1;
    if (plan.k_split > 1) {
        FPL_REQUIRE(!pool, "conv_umma: pooled epilogue needs an unsplit plan");
        partial_buf = pool_take(ctx, (size_t)n_tiles * dout_z * dout * dout * c.cout * sizeof(float), st);
        if (partial_buf < 0) return FPL_ENOMEM;
        a.partial = (float *)ctx->act_pool[partial_buf].p;
    }
    a.n_tiles = n_tiles; a.din = din; a.dout = dout; a.din_z = din_z; a.dout_z = dout_z;
    a.cin_atoms_total = cin_atoms; a.nsub = plan.nsub; a.cout = plan.n; a.cout_total = c.cout; a.ring = plan.ring;
    a.n_xt = (dout + plan.tx - 1) / plan.tx; a.n_yt = (dout + kTY - 1) / kTY;
    // z chunking: aim for >= 4 work items per SM so the persistent CTAs stay balanced
    const long long base_items = (long long)n_tiles * a.n_xt * a.n_yt;
    int n_zc = (int)((4LL * ctx->sm_count + base_items - 1) / base_items);
    if (n_zc < 1) n_zc = 1;
    int zc_len = (dout_z + n_zc - 1) / n_zc;
    if (zc_len < 8) zc_len = dout_z < 8 ? dout_z : 8;
    if (pool && (zc_len & 1)) ++zc_len;              // z pairs must not straddle work items
    a.zc_len = zc_len; a.n_zc = (dout_z + zc_len - 1) / zc_len;
    a.relu = relu;
    a.pool = pool;
    { const char *e = getenv("FPL_DBG_MAXBLK"); a.max_blk = e ? atoi(e) : 0; }
    a.tmem_cols = 512;
    const size_t smem = plan.smem;
    const long long n_items = base_items * a.n_zc;
    ProfScope prof(ctx, st, ks == 3 ? PROF_CONV3 : PROF_CONV1,
                   2.0 * ks * ks * ks * c.cin * c.cout * (double)n_tiles * dout_z * dout * dout);
    int grid = ctx->sm_count; if (grid > n_items) grid = (int)n_items;
    for (int g = 0; g < plan.n_split * plan.k_split; ++g) {
        a.w_packed = (const __nv_bfloat16 *)((const uint8_t *)c.d_packed + (size_t)g * a.w_bytes);
        if (plan.k_split > 1) {
            a.bias = c.d_bias; a.cout_off = 0;
            a.cin_atom_off = g * (c.cin / plan.k_split / 8);
            a.acc_mode = g == 0 ? 1 : (g == plan.k_split - 1 ? 3 : 2);
        } else {
            a.bias = c.d_bias + g * plan.n;
            a.cout_off = g * plan.n;
        }
        FPL_TRY(dispatch_umma(plan, ks, grid, smem, st, tmap, a));
        FPL_LAUNCH_CHECK(ctx);
    }
    if (partial_buf >= 0) ctx->act_pool[partial_buf].busy = false;     // stream-ordered: later launches may reuse it
    return FPL_OK;
}

static int launch_conv_direct(fpl_ctx *ctx, const ConvParams &c, const __nv_bfloat16 *in, __nv_bfloat16 *out,
                              int n_tiles, int din, int relu, cudaStream_t st) {
    const int dout = din - (c.k - 1);
    const int xb = (dout + 31) / 32;
    const long long blocks = (long long)n_tiles * dout * dout * xb;
    FPL_REQUIRE(blocks < 2147483647LL && c.cout / 8 <= 16, "conv_blocked_direct: launch too large");
    dim3 block(32, c.cout / 8);
    if (c.k == 3)
        conv_blocked_direct_kernel<3><<<(unsigned)blocks, block, 0, st>>>(in, c.d_kernel, c.d_scale, c.d_bias, out,
                                                                         n_tiles, din, c.cin, c.cout, relu);
    else
        conv_blocked_direct_kernel<1><<<(unsigned)blocks, block, 0, st>>>(in, c.d_kernel, c.d_scale, c.d_bias, out,
                                                                         n_tiles, din, c.cin, c.cout, relu);
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

// activation buffer pool (device): exact-size buffers, best fit, grow-only.  It belongs to the context, i.e. to
// one device; a context runs its networks on ONE stream at a time (buffers are recycled in stream order, see
// include/fpl_b200.h "Conventions").
static int pool_take(fpl_ctx *ctx, size_t bytes, cudaStream_t st) {
    std::vector<PoolBuf> &g_bufs = ctx->act_pool;
    int best = -1;
    for (size_t i = 0; i < g_bufs.size(); ++i)
        if (!g_bufs[i].busy && g_bufs[i].cap >= bytes && (best < 0 || g_bufs[i].cap < g_bufs[best].cap)) best = (int)i;
    if (best < 0) {
        cudaStreamSynchronize(st);
        // drop idle buffers that are too small to be useful again before growing
        void *p = nullptr;
        const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (size_t i = 0; i < g_bufs.size();)       // release every idle buffer and retry once
                if (!g_bufs[i].busy) { cudaFree(g_bufs[i].p); g_bufs.erase(g_bufs.begin() + i); } else ++i;
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_error("forward_umma: activation buffer of %zu bytes: %s", want, cudaGetErrorString(e));
                return -1;
            }
        }
        g_bufs.push_back(PoolBuf{p, want, false});
        best = (int)g_bufs.size() - 1;
    }
    g_bufs[best].busy = true;
    return best;
}

// ------------------------------------------------------------------------------------------------
// High-precision tensor-core path (FPL_PREC_TF32 slot of the ABI): "bf16 x 3".
// Every activation tensor is kept as TWO bf16 tensors hi = bf16(v), lo = bf16(v - hi) (channel groups
// [0,C/8) and [C/8,C/4) of one C8-blocked tensor with 2C channels); every weight likewise as w_hi, w_lo.
// A convolution is the sum of three bf16 contractions  hi*w_hi + lo*w_hi + hi*w_lo  (the dropped lo*w_lo
// term is ~2^-18 relative), each run by the same conv_umma_kernel as a K-split launch over the fp32
// partial-sum buffer; the last launch adds the bias, applies ReLU and splits the fp32 result again.
// Products of bf16 pairs are exact in fp32 and the accumulation is fp32, so the result carries ~16 mantissa
// bits end to end -- more than TF32's 10 -- at a third of the bf16 path's MMA rate (vs the CUDA-core fp32
// path: two orders of magnitude faster).  First layer (Cin = 1) and the 1-channel output layer run on CUDA
// cores in fp32.  Cubic staged tiles only (no fused first+second kernel, no direct volume I/O).
// ------------------------------------------------------------------------------------------------
static std::vector<__nv_bfloat16> pack_image(const ConvParams &c, const ConvPlan &plan, int ci0, int cc, bool lo_part) {
    const int ks = c.k, n = c.cout, nkb = plan.rot ? 5 : ks, ksteps = cc / 16;
    std::vector<__nv_bfloat16> pk((size_t)ks * ks * nkb * cc * n);
    for (int kh = 0; kh < ks; ++kh)
        for (int kw = 0; kw < ks; ++kw)
            for (int s = 0; s < ksteps; ++s)
                for (int h = 0; h < 2; ++h)
                    for (int kb = 0; kb < nkb; ++kb)
                        for (int nn = 0; nn < n; ++nn)
                            for (int e = 0; e < 8; ++e) {
                                const int kd = kb % ks, ci = ci0 + 16 * s + 8 * h + e;
                                const int tap = (kd * ks + kh) * ks + kw;
                                const float v = c.kernel[((size_t)tap * c.cin + ci) * c.cout + nn] * c.scale[nn];
                                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                                const size_t chunk = ((size_t)(kh * ks + kw) * ksteps + s);
                                pk[(((chunk * 2 + h) * nkb + kb) * n + nn) * 8 + e] =
                                    lo_part ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
                            }
    return pk;
}

// chunk plan of a hi/lo convolution: resident weights, all Cout in one launch
static ConvPlan hilo_plan(const ConvParams &c, int cc) {
    ConvParams chunk;
    chunk.k = c.k; chunk.cin = cc; chunk.cout = c.cout; chunk.no_ksplit = true;
    ConvPlan p = plan_conv(chunk);
    if (p.ok && p.n_split != 1) p.ok = false;
    return p;
}

static int pack_weights_hilo(fpl_net *net) {
    for (ConvParams &c : net->convs) {
        if (c.cin % 16 || c.cout % 16) continue;            // first layer / output layer: CUDA cores
        int cc = 0;
        const int cands[4] = {48, 32, 16, 64};
        for (int i = 0; i < 4 && !cc; ++i)
            if (c.cin % cands[i] == 0 && hilo_plan(c, cands[i]).ok) cc = cands[i];
        FPL_REQUIRE(cc > 0, "hi/lo path: no chunk plan for k=%d Cin=%d Cout=%d", c.k, c.cin, c.cout);
        const ConvPlan plan = hilo_plan(c, cc);
        const int n_chunks = c.cin / cc;
        std::vector<__nv_bfloat16> hi, lo;
        for (int g = 0; g < n_chunks; ++g) {
            std::vector<__nv_bfloat16> a = pack_image(c, plan, g * cc, cc, false), b = pack_image(c, plan, g * cc, cc, true);
            hi.insert(hi.end(), a.begin(), a.end()); lo.insert(lo.end(), b.begin(), b.end());
        }
        c.hilo_chunk = cc;
        c.packed_bytes = hi.size() * sizeof(__nv_bfloat16);
        FPL_CUDA_CHECK(cudaMalloc(&c.d_packed, c.packed_bytes));
        FPL_CUDA_CHECK(cudaMalloc(&c.d_packed_lo, c.packed_bytes));
        FPL_CUDA_CHECK(cudaMemcpy(c.d_packed, hi.data(), c.packed_bytes, cudaMemcpyHostToDevice));
        FPL_CUDA_CHECK(cudaMemcpy(c.d_packed_lo, lo.data(), c.packed_bytes, cudaMemcpyHostToDevice));
    }
    return FPL_OK;
}

__device__ __forceinline__ void store_hilo8(__nv_bfloat16 *hi_p, __nv_bfloat16 *lo_p, const float (&v)[8]) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[2 * j] - __low2float(hi), v[2 * j + 1] - __high2float(hi));
        ph[j] = *reinterpret_cast<const uint32_t *>(&hi); pl[j] = *reinterpret_cast<const uint32_t *>(&lo);
    }
    *reinterpret_cast<uint4 *>(hi_p) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4 *>(lo_p) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}
__device__ __forceinline__ void load_hilo8(const uint4 *hi_p, const uint4 *lo_p, float (&v)[8]) {
    const uint4 a = __ldg(hi_p), b = __ldg(lo_p);
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162 *>(&ua[j]), l2 = *reinterpret_cast<const __nv_bfloat162 *>(&ub[j]);
        v[2 * j] = __low2float(h2) + __low2float(l2); v[2 * j + 1] = __high2float(h2) + __high2float(l2);
    }
}

// first layer in fp32 on CUDA cores: one thread = one voxel x CG groups of 8 output channels; the folded
// weights sit in shared memory (broadcast float4 reads), threads run along x (coalesced loads and 16-byte stores)
template <int CG>
__global__ void __launch_bounds__(256)
first_hilo_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ scale,
                  const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out, int n_tiles, int din) {
    constexpr int COUT = CG * 8;
    __shared__ __align__(16) float sw[27 * COUT];
    __shared__ float ssc[COUT], sbi[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { ssc[i] = scale[i]; sbi[i] = bias[i]; }
    __syncthreads();
    const int dout = din - 2;
    const long long vox_n = (long long)dout * dout * dout, total = (long long)n_tiles * vox_n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / vox_n);
        const long long o = i - (long long)t * vox_n;
        const int x = (int)(o % dout), y = (int)((o / dout) % dout), z = (int)(o / ((long long)dout * dout));
        const float *ip = in + (size_t)t * din * din * din + ((size_t)z * din + y) * din + x;
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float a = __ldg(ip + ((size_t)kd * din + kh) * din + kw);
                    const float4 *wp = reinterpret_cast<const float4 *>(&sw[((kd * 3 + kh) * 3 + kw) * COUT]);
#pragma unroll
                    for (int j4 = 0; j4 < COUT / 4; ++j4) {
                        const float4 wv = wp[j4];
                        acc[4 * j4] = fmaf(a, wv.x, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(a, wv.y, acc[4 * j4 + 1]);
                        acc[4 * j4 + 2] = fmaf(a, wv.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(a, wv.w, acc[4 * j4 + 3]);
                    }
                }
#pragma unroll
        for (int cg = 0; cg < CG; ++cg) {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fmaxf(fmaf(acc[cg * 8 + j], ssc[cg * 8 + j], sbi[cg * 8 + j]), 0.f);
            store_hilo8(out + (((size_t)t * 2 * CG + cg) * vox_n + o) * 8, out + (((size_t)t * 2 * CG + CG + cg) * vox_n + o) * 8, r);
        }
    }
}

__global__ void __launch_bounds__(256)
pool_hilo_kernel(const uint4 *__restrict__ in, __nv_bfloat16 *__restrict__ out, int n_tiles, int cg_n, int din) {
    const int dout = din / 2;
    const long long vin = (long long)din * din * din, vout = (long long)dout * dout * dout, total = (long long)n_tiles * cg_n * vout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long v = i;
        const int x = (int)(v % dout); v /= dout;
        const int y = (int)(v % dout); v /= dout;
        const int z = (int)(v % dout); v /= dout;
        const int cg = (int)(v % cg_n), t = (int)(v / cg_n);
        const uint4 *hp = in + ((size_t)t * 2 * cg_n + cg) * vin, *lp = in + ((size_t)t * 2 * cg_n + cg_n + cg) * vin;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int dz = 0; dz < 2; ++dz)
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const size_t q = ((size_t)(2 * z + dz) * din + (2 * y + dy)) * din + (2 * x + dx);
                    float f[8];
                    load_hilo8(hp + q, lp + q, f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
                }
        const size_t o = ((size_t)z * dout + y) * dout + x;
        store_hilo8(out + (((size_t)t * 2 * cg_n + cg) * vout + o) * 8, out + (((size_t)t * 2 * cg_n + cg_n + cg) * vout + o) * 8, m);
    }
}

// concatenate([UpSampling3D(2)(a), Cropping3D(crop)(skip)]): hi groups of both first, then lo groups of both
__global__ void __launch_bounds__(256)
upcat_hilo_kernel(const uint4 *__restrict__ a, int da, int cga, const uint4 *__restrict__ skip, int ds, int cgs,
                  int crop, uint4 *__restrict__ out, int n_tiles) {
    const int dout = 2 * da, cgo = cga + cgs;
    const long long vo = (long long)dout * dout * dout, total = (long long)n_tiles * 2 * cgo * vo;
    const long long va = (long long)da * da * da, vs = (long long)ds * ds * ds;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long v = i;
        const int x = (int)(v % dout); v /= dout;
        const int y = (int)(v % dout); v /= dout;
        const int z = (int)(v % dout); v /= dout;
        const int g2 = (int)(v % (2 * cgo)), t = (int)(v / (2 * cgo));
        const int half = g2 / cgo, g = g2 % cgo;                 // half 0: hi, 1: lo
        uint4 r;
        if (g < cga) r = __ldg(a + ((size_t)t * 2 * cga + half * cga + g) * va + ((size_t)(z / 2) * da + y / 2) * da + x / 2);
        else r = __ldg(skip + ((size_t)t * 2 * cgs + half * cgs + (g - cga)) * vs + ((size_t)(z + crop) * ds + y + crop) * ds + x + crop);
        out[i] = r;
    }
}

__global__ void __launch_bounds__(256)
final_hilo_kernel(const uint4 *__restrict__ in, const float *__restrict__ w, float bias, float *__restrict__ out,
                  int n_tiles, int d, int cg_n, int stride) {
    const long long vox = (long long)d * d * d, total = (long long)n_tiles * vox;
    const int dout = d * stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / vox);
        const long long o = i - (long long)t * vox;
        float acc = 0.f;
        for (int cg = 0; cg < cg_n; ++cg) {
            float f[8];
            load_hilo8(in + ((size_t)t * 2 * cg_n + cg) * vox + o, in + ((size_t)t * 2 * cg_n + cg_n + cg) * vox + o, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(f[j], __ldg(w + cg * 8 + j), acc);
        }
        acc += bias;
        const float pr = 1.f / (1.f + expf(-acc));
        const int x = (int)(o % d), y = (int)((o / d) % d), z = (int)(o / ((long long)d * d));
        float *op = out + (size_t)t * dout * dout * dout;
        for (int ez = 0; ez < stride; ++ez)
            for (int ey = 0; ey < stride; ++ey)
                for (int ex = 0; ex < stride; ++ex)
                    op[((size_t)(z * stride + ez) * dout + (y * stride + ey)) * dout + (x * stride + ex)] = pr;
    }
}

// one hi/lo convolution: (Cin/chunk) x 3 K-split launches of conv_umma_kernel over the fp32 partial buffer
static int launch_conv_hilo(fpl_ctx *ctx, const ConvParams &c, const __nv_bfloat16 *in, __nv_bfloat16 *out, int n_tiles,
                            int din, int relu, cudaStream_t st) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FPL_ECUDA; }
    const int cc = c.hilo_chunk, n_chunks = c.cin / cc;
    const ConvPlan plan = hilo_plan(c, cc);
    FPL_REQUIRE(cc > 0 && plan.ok, "hi/lo conv: no plan for k=%d Cin=%d Cout=%d", c.k, c.cin, c.cout);
    const int ks = c.k, dout = din - (ks - 1);
    const int sx = plan.tx + ks - 1, sy = kTY + ks - 1;
    const int atoms_total = 2 * c.cin / 8;                       // hi atoms, then lo atoms
    CUtensorMap tmap;
    cuuint64_t gdim[4] = {(cuuint64_t)din * 8, (cuuint64_t)din, (cuuint64_t)din, (cuuint64_t)n_tiles * atoms_total};
    cuuint64_t gstride[3] = {(cuuint64_t)din * 16, (cuuint64_t)din * din * 16, (cuuint64_t)din * din * din * 16};
    cuuint32_t box[4] = {(cuuint32_t)sx * 8, (cuuint32_t)sy, 1, (cuuint32_t)(2 * plan.ksteps)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return FPL_ECUDA; }
    const int pbuf = pool_take(ctx, (size_t)n_tiles * dout * dout * dout * c.cout * sizeof(float), st);
    if (pbuf < 0) return FPL_ENOMEM;
    ConvArgs a;
    a.out = out; a.bias = c.d_bias; a.partial = (float *)ctx->act_pool[pbuf].p;
    a.w_bytes = (uint32_t)(c.packed_bytes / n_chunks);
    a.n_tiles = n_tiles; a.din = din; a.dout = dout; a.din_z = din; a.dout_z = dout;
    a.cin_atoms_total = atoms_total; a.nsub = plan.nsub; a.cout = c.cout; a.cout_total = c.cout; a.cout_off = 0; a.ring = plan.ring;
    a.n_xt = (dout + plan.tx - 1) / plan.tx; a.n_yt = (dout + kTY - 1) / kTY;
    const long long base_items = (long long)n_tiles * a.n_xt * a.n_yt;
    int n_zc = (int)((4LL * ctx->sm_count + base_items - 1) / base_items);
    if (n_zc < 1) n_zc = 1;
    int zc_len = (dout + n_zc - 1) / n_zc;
    if (zc_len < 8) zc_len = dout < 8 ? dout : 8;
    a.zc_len = zc_len; a.n_zc = (dout + zc_len - 1) / zc_len;
    a.relu = relu; a.pool = 0; a.max_blk = 0; a.tmem_cols = 512; a.shared_w = 0; a.sub_stride = 0;
    const long long n_items = base_items * a.n_zc;
    int grid = ctx->sm_count; if (grid > n_items) grid = (int)n_items;
    ProfScope prof(ctx, st, ks == 3 ? PROF_CONV3 : PROF_CONV1,
                   3.0 * 2.0 * ks * ks * ks * c.cin * c.cout * (double)n_tiles * dout * dout * dout);
    // per chunk two launches: (hi, lo) x w_hi as two sub-planes sharing one weight image (their sum stays in TMEM),
    // then hi x w_lo; the fp32 partial sums pass through HBM between launches
    const bool fuse = plan.nsub == 1 && plan_smem(ks, cc, c.cout, 1, plan.tx, 2, plan.rot) <= kMaxDynSmem;
    const size_t smem_fused = plan_smem(ks, cc, c.cout, 1, plan.tx, 2, plan.rot);
    const int per_chunk = fuse ? 2 : 3, n_launch = per_chunk * n_chunks;
    int li = 0;
    for (int g = 0; g < n_chunks; ++g)
        for (int term = 0; term < per_chunk; ++term, ++li) {
            const bool lo_w = term == per_chunk - 1;              // last term of a chunk: hi x w_lo
            const uint8_t *img = (const uint8_t *)(lo_w ? c.d_packed_lo : c.d_packed) + (size_t)g * a.w_bytes;
            a.w_packed = (const __nv_bfloat16 *)img;
            a.acc_mode = li == 0 ? 1 : (li == n_launch - 1 ? 4 : 2);
            if (fuse && term == 0) {
                a.cin_atom_off = g * (cc / 8); a.nsub = 2; a.ring = 2; a.shared_w = 1; a.sub_stride = c.cin / 8;
                FPL_TRY(dispatch_umma(plan, ks, grid, smem_fused, st, tmap, a));
            } else {
                a.cin_atom_off = g * (cc / 8) + ((!fuse && term == 1) ? c.cin / 8 : 0);
                a.nsub = plan.nsub; a.ring = plan.ring; a.shared_w = 0; a.sub_stride = 0;
                FPL_TRY(dispatch_umma(plan, ks, grid, plan.smem, st, tmap, a));
            }
            FPL_LAUNCH_CHECK(ctx);
        }
    ctx->act_pool[pbuf].busy = false;
    return FPL_OK;
}

static int forward_hilo(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out, cudaStream_t st) {
    fpl_ctx *ctx = net->ctx;
    for (const Op &o : net->ops)
        FPL_REQUIRE(o.kind != OP_ADD && !(o.kind == OP_CONV && (!o.relu || !o.bn || o.src_slot >= 0)),
                    "precision 'tf32' (hi/lo path) does not cover the residual blocks of resnet_like: use 'bf16' or 'fp32'");
    std::vector<PoolBuf> &g_bufs = ctx->act_pool;
    for (PoolBuf &b : g_bufs) b.busy = false;
    auto release = [&](int i) { if (i >= 0) g_bufs[i].busy = false; };
    int cur = -1, skip_buf[4] = {-1, -1, -1, -1}, skip_d[4] = {0, 0, 0, 0}, skip_c[4] = {0, 0, 0, 0};
    bool cur_is_skip = false;
    int d = in_sz, c = 1;
    const int stream_blocks = ctx->sm_count * 8;
    for (size_t oi = 0; oi < net->ops.size(); ++oi) {
        const Op &o = net->ops[oi];
        if (o.kind == OP_CONV) {
            const ConvParams &cp = net->convs[o.conv_index];
            const int dout = d - (o.k - 1);
            const int nb = pool_take(ctx, (size_t)n_tiles * dout * dout * dout * 2 * cp.cout * 2, st);
            if (nb < 0) return FPL_ENOMEM;
            __nv_bfloat16 *dst = (__nv_bfloat16 *)g_bufs[nb].p;
            if (cp.cin == 1) {
                FPL_REQUIRE(cp.k == 3 && (cp.cout == 48 || cp.cout == 32), "hi/lo path: unsupported first layer");
                ProfScope prof(ctx, st, PROF_FIRST, 2.0 * 27 * cp.cout * (double)n_tiles * dout * dout * dout);
                if (cp.cout == 48)
                    first_hilo_kernel<6><<<stream_blocks, 256, 0, st>>>(d_tiles, cp.d_kernel, cp.d_scale, cp.d_bias, dst, n_tiles, d);
                else
                    first_hilo_kernel<4><<<stream_blocks, 256, 0, st>>>(d_tiles, cp.d_kernel, cp.d_scale, cp.d_bias, dst, n_tiles, d);
                FPL_LAUNCH_CHECK(ctx);
            } else {
                FPL_REQUIRE(cp.d_packed && cp.d_packed_lo, "hi/lo path: layer k=%d Cin=%d Cout=%d has no packed weights", cp.k, cp.cin, cp.cout);
                FPL_TRY(launch_conv_hilo(ctx, cp, (const __nv_bfloat16 *)g_bufs[cur].p, dst, n_tiles, d, 1, st));
            }
            if (!cur_is_skip) release(cur);
            cur = nb; cur_is_skip = false; d = dout; c = o.cout;
        } else if (o.kind == OP_POOL) {
            const int nb = pool_take(ctx, (size_t)n_tiles * (d / 2) * (d / 2) * (d / 2) * 2 * c * 2, st);
            if (nb < 0) return FPL_ENOMEM;
            ProfScope prof(ctx, st, PROF_NETAUX, (double)n_tiles * 2 * c * 2.0 * d * d * d * 1.125);
            pool_hilo_kernel<<<stream_blocks, 256, 0, st>>>((const uint4 *)g_bufs[cur].p, (__nv_bfloat16 *)g_bufs[nb].p, n_tiles, c / 8, d);
            FPL_LAUNCH_CHECK(ctx);
            if (!cur_is_skip) release(cur);
            cur = nb; cur_is_skip = false; d /= 2;
        } else if (o.kind == OP_SAVE) {
            skip_buf[o.slot] = cur; skip_d[o.slot] = d; skip_c[o.slot] = c; cur_is_skip = true;
        } else if (o.kind == OP_UPCAT) {
            const int nb = pool_take(ctx, (size_t)n_tiles * 8 * d * d * d * 2 * (c + skip_c[o.slot]) * 2, st);
            if (nb < 0) return FPL_ENOMEM;
            ProfScope prof(ctx, st, PROF_NETAUX, (double)n_tiles * 2 * (c + skip_c[o.slot]) * 2.0 * 8.0 * d * d * d * 2);
            upcat_hilo_kernel<<<stream_blocks, 256, 0, st>>>((const uint4 *)g_bufs[cur].p, d, c / 8, (const uint4 *)g_bufs[skip_buf[o.slot]].p,
                                                            skip_d[o.slot], skip_c[o.slot] / 8, o.crop, (uint4 *)g_bufs[nb].p, n_tiles);
            FPL_LAUNCH_CHECK(ctx);
            if (!cur_is_skip) release(cur);
            release(skip_buf[o.slot]);
            cur = nb; cur_is_skip = false; d *= 2; c += skip_c[o.slot];
        } else if (o.kind == OP_FINAL) {
            const ConvParams &cp = net->convs[o.conv_index];
            ProfScope prof(ctx, st, PROF_NETAUX, (double)n_tiles * d * d * d * (c * 4.0 + 4.0));
            final_hilo_kernel<<<stream_blocks, 256, 0, st>>>((const uint4 *)g_bufs[cur].p, cp.d_kernel, cp.bias[0], d_out, n_tiles, d,
                                                            c / 8, net->info.rf_stride);
            FPL_LAUNCH_CHECK(ctx);
        }
    }
    return FPL_OK;
}

// true when forward_umma will take the fused first+second convolution path, whose builders can read the
// input tile straight from the (uint8 / float32) volume -- no float32 tile staging needed
bool umma_reads_volume(const fpl_net *net) {
    if (g_force_direct || g_no_conv12_fusion || net->precision != FPL_PREC_BF16) return false;
    if (net->ops.size() < 2 || net->ops[0].kind != OP_CONV || net->ops[1].kind != OP_CONV) return false;
    const ConvParams &c1 = net->convs[net->ops[0].conv_index], &c2 = net->convs[net->ops[1].conv_index];
    const ConvPlan p2 = plan_conv(c2);
    return fusable12(c1, c2) && c1.d_packed && p2.ok && p2.n_split == 1 && p2.nsub == 1 && p2.rot == (c1.cout == 32) &&
           p2.k_split == 1;
}

static int forward_hilo(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out, cudaStream_t st);

int forward_umma(fpl_net *net, const float *d_tiles, int n_tiles, int in_sz, float *d_out, cudaStream_t st,
                 const VolumeIO *vio, int in_z) {
    fpl_ctx *ctx = net->ctx;
    if (net->precision == FPL_PREC_TF32) {
        FPL_REQUIRE(vio == nullptr && (in_z <= 0 || in_z == in_sz), "forward_umma: the hi/lo path runs on cubic staged tiles");
        return forward_hilo(net, d_tiles, n_tiles, in_sz, d_out, st);
    }
    if (net->precision != FPL_PREC_BF16) {
        set_error("forward_umma: unknown tensor-core precision %d", net->precision);
        return FPL_ESTATE;
    }
    if (in_z <= 0) in_z = in_sz;
    std::vector<PoolBuf> &g_bufs = ctx->act_pool;
    for (PoolBuf &b : g_bufs) b.busy = false;
    auto release = [&](int i) { if (i >= 0) g_bufs[i].busy = false; };
    int cur = -1;                 // pool buffer holding the current activation (-1: the fp32 input tiles)
    int skip_buf[4] = {-1, -1, -1, -1}, skip_d[4] = {0, 0, 0, 0}, skip_dz[4] = {0, 0, 0, 0}, skip_c[4] = {0, 0, 0, 0};
    bool cur_is_skip = false;
    int d = in_sz, dzv = in_z, c = 1;          // x/y extent, z extent, channels of the current activation
    const int stream_blocks = ctx->sm_count * 8;
    bool skip_next_pool = false;
    for (size_t oi = 0; oi < net->ops.size(); ++oi) {
        const Op &o = net->ops[oi];
        if (o.kind == OP_CONV && oi == 0 && !g_force_direct && !g_no_conv12_fusion && oi + 1 < net->ops.size() &&
            net->ops[1].kind == OP_CONV) {
            // first + second convolution fused (conv_fused12_kernel): the first layer's output stays on chip
            const ConvParams &c1 = net->convs[o.conv_index], &c2 = net->convs[net->ops[1].conv_index];
            const ConvPlan p2 = plan_conv(c2);
            const bool shape_ok = fusable12(c1, c2) && c1.d_packed && p2.ok && p2.n_split == 1 && p2.nsub == 1 &&
                                  p2.rot == (c1.cout == 32) && p2.k_split == 1 && d >= 8 && dzv >= 8;
            if (shape_ok) {
                const int dout = d - 4, dout_z = dzv - 4;
                const bool pool = !g_no_pool_fusion && net->ops.size() > 2 && net->ops[2].kind == OP_POOL &&
                                  dout % 2 == 0 && dout_z % 2 == 0;
                const size_t out_bytes = pool ? (size_t)n_tiles * (dout_z / 2) * (dout / 2) * (dout / 2) * c2.cout * 2
                                              : (size_t)n_tiles * dout_z * dout * dout * c2.cout * 2;
                const int nb = pool_take(ctx, out_bytes, st);
                if (nb < 0) return FPL_ENOMEM;
                FusedArgs fa;
                fa.in = d_tiles; fa.w1_packed = (const __nv_bfloat16 *)c1.d_packed;
                fa.w2_packed = (const __nv_bfloat16 *)c2.d_packed; fa.bias2 = c2.d_bias;
                fa.out = (__nv_bfloat16 *)g_bufs[nb].p; fa.w2_bytes = (uint32_t)c2.packed_bytes;
                fa.n_tiles = n_tiles; fa.din = d; fa.din_z = dzv; fa.dout = dout; fa.dout_z = dout_z;
                fa.n_xt = (dout + kTX - 1) / kTX; fa.n_yt = (dout + kTY - 1) / kTY;
                const long long base_items = (long long)n_tiles * fa.n_xt * fa.n_yt;
                int n_zc = (int)((4LL * ctx->sm_count + base_items - 1) / base_items);
                if (n_zc < 1) n_zc = 1;
                int zc_len = (dout_z + n_zc - 1) / n_zc;
                if (zc_len < 8) zc_len = dout_z < 8 ? dout_z : 8;
                if (pool && (zc_len & 1)) ++zc_len;
                fa.zc_len = zc_len; fa.n_zc = (dout_z + zc_len - 1) / zc_len;
                fa.pool = pool ? 1 : 0;
                if (vio) fa.vio = *vio;
                const long long n_items = base_items * fa.n_zc;
                int grid = ctx->sm_count; if (grid > n_items) grid = (int)n_items;
                const int C = c1.cout;
                const size_t smem = ((c2.packed_bytes + 127) & ~size_t(127)) + 2 * (size_t)(C / 8) * 324 * 16 + 3 * 8192 +
                                    2 * 2 * C * 16 + 4 * 20 * 20 * 4 + 1024 + 512 + 256;
                {
                    ProfScope prof(ctx, st, PROF_CONV3,
                                   2.0 * 27 * C * C * (double)n_tiles * dout_z * dout * dout +
                                   2.0 * 27 * C * (double)n_tiles * (dout_z + 2) * (dout + 2) * (dout + 2));
                    if (C == 48) {
                        FPL_CUDA_CHECK(cudaFuncSetAttribute(conv_fused12_kernel<48, 48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        conv_fused12_kernel<48, 48, false><<<grid, kFusedThreads, smem, st>>>(fa);
                    } else {
                        FPL_CUDA_CHECK(cudaFuncSetAttribute(conv_fused12_kernel<32, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        conv_fused12_kernel<32, 32, true><<<grid, kFusedThreads, smem, st>>>(fa);
                    }
                    FPL_LAUNCH_CHECK(ctx);
                }
                cur = nb; cur_is_skip = false;
                d = dout; dzv = dout_z; c = c2.cout;
                if (pool) { d /= 2; dzv /= 2; }
                oi += pool ? 2 : 1;          // the second convolution (and the pooling) are done
                continue;
            }
        }
        if (o.kind == OP_CONV && o.src_slot >= 0) {
            // resnet_like shortcut: 1x1x1 convolution of a stored tensor, which it replaces
            const ConvParams &cp = net->convs[o.conv_index];
            const int sl = o.src_slot, sd = skip_d[sl], sdz = skip_dz[sl];
            const int nb = pool_take(ctx, (size_t)n_tiles * sdz * sd * sd * cp.cout * 2, st);
            if (nb < 0) return FPL_ENOMEM;
            const __nv_bfloat16 *src = (const __nv_bfloat16 *)g_bufs[skip_buf[sl]].p;
            if (umma_supported(cp) && !g_force_direct)
                FPL_TRY(launch_conv_umma(ctx, cp, src, (__nv_bfloat16 *)g_bufs[nb].p, n_tiles, sd, sdz, o.relu ? 1 : 0, 0, st));
            else {
                FPL_REQUIRE(sdz == sd, "forward_umma: CUDA-core convolution handles cubic tiles only");
                FPL_TRY(launch_conv_direct(ctx, cp, src, (__nv_bfloat16 *)g_bufs[nb].p, n_tiles, sd, o.relu ? 1 : 0, st));
            }
            if (skip_buf[sl] != cur) release(skip_buf[sl]);
            else cur_is_skip = false;                 // the stored tensor was also the current one: it stays alive as cur only
            skip_buf[sl] = nb; skip_c[sl] = cp.cout;
            continue;
        }
        if (o.kind == OP_ADD) {
            const int sl = o.slot;
            FPL_REQUIRE(cur >= 0 && !cur_is_skip && skip_c[sl] == c && skip_d[sl] - 2 * o.crop == d && skip_dz[sl] - 2 * o.crop == dzv,
                        "forward_umma: add: shapes do not match");
            add_blocked_kernel<<<stream_blocks, 256, 0, st>>>((uint4 *)g_bufs[cur].p, d, dzv, (const uint4 *)g_bufs[skip_buf[sl]].p,
                                                             skip_d[sl], skip_dz[sl], o.crop, (long long)n_tiles * (c / 8));
            FPL_LAUNCH_CHECK(ctx);
            release(skip_buf[sl]);
            skip_buf[sl] = -1;
            continue;
        }
        if (o.kind == OP_CONV) {
            const ConvParams &cp = net->convs[o.conv_index];
            const int dout = d - (o.k - 1), dout_z = dzv - (o.k - 1);
            // MaxPooling3D directly after this conv (and the conv output not kept as a skip): fuse it
            const bool fuse_pool = !g_force_direct && !g_no_pool_fusion && cp.cin != 1 && umma_supported(cp) && cp.k == 3 &&
                                   (cp.cout == 32 || cp.cout == 48 || cp.cout == 64) && oi + 1 < net->ops.size() &&
                                   net->ops[oi + 1].kind == OP_POOL && (dout % 2 == 0) && (dout_z % 2 == 0) &&
                                   plan_conv(cp).n_split == 1;
            const size_t out_bytes = fuse_pool ? (size_t)n_tiles * (dout_z / 2) * (dout / 2) * (dout / 2) * cp.cout * 2
                                               : (size_t)n_tiles * dout_z * dout * dout * cp.cout * 2;
            const int nb = pool_take(ctx, out_bytes, st);
            if (nb < 0) return FPL_ENOMEM;
            __nv_bfloat16 *dst = (__nv_bfloat16 *)g_bufs[nb].p;
            if (cp.cin == 1) {
                FPL_REQUIRE(cp.k == 3, "forward_umma: unsupported first layer");
                ProfScope prof(ctx, st, PROF_FIRST, 2.0 * 27 * cp.cout * (double)n_tiles * dout_z * dout * dout);
                if (!g_force_direct && cp.d_packed && cp.cout % 16 == 0 && cp.cout <= 64) {
                    FirstArgs fa;
                    fa.in = d_tiles; fa.w_packed = (const __nv_bfloat16 *)cp.d_packed; fa.bias = cp.d_bias; fa.out = dst;
                    fa.n_tiles = n_tiles; fa.din = d; fa.dout = dout; fa.din_z = dzv; fa.dout_z = dout_z; fa.cout = cp.cout;
                    fa.n_xt = (dout + kTX - 1) / kTX; fa.n_yt = (dout + kTY - 1) / kTY;
                    const long long base_items = (long long)n_tiles * fa.n_xt * fa.n_yt;
                    int n_zc = (int)((4LL * ctx->sm_count + base_items - 1) / base_items);
                    if (n_zc < 1) n_zc = 1;
                    int zc_len = (dout_z + n_zc - 1) / n_zc;
                    if (zc_len < 8) zc_len = dout_z < 8 ? dout_z : 8;
                    fa.zc_len = zc_len; fa.n_zc = (dout_z + zc_len - 1) / zc_len;
                    const int acc_cols = 4 * 2 * cp.cout;       // kSt stages x 2 M-tiles
                    fa.tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
                    if (vio) fa.vio = *vio;
                    { const char *e = getenv("FPL_DBG_FIRST"); fa.dbg = e ? atoi(e) : 0; }
                    const long long n_items = base_items * fa.n_zc;
                    int grid = ctx->sm_count; if (grid > n_items) grid = (int)n_items;
                    const size_t first_smem = 4 * (2 * 4 * 128 * 16);
                    FPL_CUDA_CHECK(cudaFuncSetAttribute(conv_first_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int)first_smem));
                    conv_first_umma_kernel<<<grid, kFirstThreads, first_smem, st>>>(fa);
                } else {
                    const long long blocks = (long long)n_tiles * dout * dout * ((dout + 127) / 128);
                    FPL_REQUIRE(!(vio && vio->img), "forward_umma: direct volume input needs the tensor-core first layer");
                    FPL_REQUIRE(dzv == d && blocks < 2147483647LL, "forward_umma: CUDA-core first layer handles cubic tiles only");
                    if (cp.cout == 48)
                        conv_first_kernel<48><<<(unsigned)blocks, 128, 0, st>>>(d_tiles, cp.d_kernel, cp.d_scale, cp.d_bias, dst, n_tiles, d);
                    else if (cp.cout == 32)
                        conv_first_kernel<32><<<(unsigned)blocks, 128, 0, st>>>(d_tiles, cp.d_kernel, cp.d_scale, cp.d_bias, dst, n_tiles, d);
                    else { set_error("forward_umma: first layer Cout %d unsupported", cp.cout); return FPL_EINVAL; }
                }
                FPL_LAUNCH_CHECK(ctx);
            } else {
                const __nv_bfloat16 *src = (const __nv_bfloat16 *)g_bufs[cur].p;
                if (umma_supported(cp) && !g_force_direct)
                    FPL_TRY(launch_conv_umma(ctx, cp, src, dst, n_tiles, d, dzv, o.relu ? 1 : 0, fuse_pool ? 1 : 0, st));
                else {
                    FPL_REQUIRE(dzv == d, "forward_umma: CUDA-core convolution handles cubic tiles only");
                    FPL_TRY(launch_conv_direct(ctx, cp, src, dst, n_tiles, d, o.relu ? 1 : 0, st));
                }
            }
            if (!cur_is_skip) release(cur);
            cur = nb; cur_is_skip = false;
            d = dout; dzv = dout_z; c = o.cout;
            if (fuse_pool) { d /= 2; dzv /= 2; skip_next_pool = true; }
        } else if (o.kind == OP_POOL) {
            if (skip_next_pool) { skip_next_pool = false; continue; }
            const int nb = pool_take(ctx, (size_t)n_tiles * (dzv / 2) * (d / 2) * (d / 2) * c * 2, st);
            if (nb < 0) return FPL_ENOMEM;
            ProfScope prof(ctx, st, PROF_NETAUX, (double)n_tiles * c * 2.0 * dzv * d * d * 1.125);
            pool_blocked_kernel<<<stream_blocks, 256, 0, st>>>((const uint4 *)g_bufs[cur].p, (uint4 *)g_bufs[nb].p,
                                                              (long long)n_tiles * (c / 8), d, dzv);
            FPL_LAUNCH_CHECK(ctx);
            if (!cur_is_skip) release(cur);
            cur = nb; cur_is_skip = false;
            d /= 2; dzv /= 2;
        } else if (o.kind == OP_SAVE) {
            skip_buf[o.slot] = cur; skip_d[o.slot] = d; skip_dz[o.slot] = dzv; skip_c[o.slot] = c;
            cur_is_skip = true;
        } else if (o.kind == OP_UPCAT) {
            FPL_REQUIRE(dzv == d, "forward_umma: the U-Net runs on cubic tiles");
            const int nb = pool_take(ctx, (size_t)n_tiles * 8 * d * d * d * (c + skip_c[o.slot]) * 2, st);
            if (nb < 0) return FPL_ENOMEM;
            ProfScope prof(ctx, st, PROF_NETAUX, (double)n_tiles * (c + skip_c[o.slot]) * 2.0 * 8.0 * d * d * d * 2);
            upcat_blocked_kernel<<<stream_blocks, 256, 0, st>>>((const uint4 *)g_bufs[cur].p, d, c / 8,
                                                               (const uint4 *)g_bufs[skip_buf[o.slot]].p,
                                                               skip_d[o.slot], skip_c[o.slot] / 8, o.crop,
                                                               (uint4 *)g_bufs[nb].p, n_tiles);
            FPL_LAUNCH_CHECK(ctx);
            if (!cur_is_skip) release(cur);
            release(skip_buf[o.slot]);
            cur = nb; cur_is_skip = false;
            d *= 2; dzv *= 2; c += skip_c[o.slot];
        } else if (o.kind == OP_FINAL) {
            const ConvParams &cp = net->convs[o.conv_index];
            ProfScope prof(ctx, st, PROF_NETAUX, (double)n_tiles * dzv * d * d * (c * 2.0 + 4.0 * net->info.rf_stride * net->info.rf_stride * net->info.rf_stride));
            final_blocked_kernel<<<stream_blocks, 256, 0, st>>>((const uint4 *)g_bufs[cur].p, cp.d_kernel, cp.bias[0],
                                                               d_out, n_tiles, d, c / 8, net->info.rf_stride,
                                                               vio ? *vio : VolumeIO(), dzv);
            FPL_LAUNCH_CHECK(ctx);
        }
    }
    return FPL_OK;
}

}  // namespace net
}  // namespace fpl

extern "C" int fpl_debug_force_direct_conv(int on) { fpl::net::g_force_direct = on; return FPL_OK; }
extern "C" int fpl_debug_no_pool_fusion(int on) { fpl::net::g_no_pool_fusion = on; return FPL_OK; }
extern "C" int fpl_debug_no_conv12_fusion(int on) { fpl::net::g_no_conv12_fusion = on; return FPL_OK; }
