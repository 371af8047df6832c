// train_tc.cuh -- tensor-core (tcgen05 / TMEM) contractions of the training step (BASELINE config 5).
//
// The three GEMMs of a Conv3D layer in training (flypylib/fplnetwork.py:112-128: Keras fit_generator on the
// graphs of flypylib/fplmodels.py:102-172):
//   forward   x[v, co]        = sum_{tap, ci} in[v + tap, ci] * W[tap, ci, co]        M = voxels,   N = Cout, K = k^3*Cin
//   dgrad     din[v', ci]     = sum_{tap, co} dx[v' - tap, co] * W[tap, ci, co]       M = voxels,   N = Cin,  K = k^3*Cout
//             (a valid convolution of the zero-padded dx with the flipped, transposed kernel)
//   wgrad     dW[tap, ci, co] = sum_v in[v + tap, ci] * dx[v, co]                     M = k^3*Cin, N = Cout, K = voxels
// run as tcgen05.mma.cta_group::1.kind::f16 with fp32 accumulators in TMEM.  The training tensors stay float32 in
// HBM ((N,z,y,x,C), the layout of the CUDA-core path); they are converted to bf16 (hi / lo) operand images in shared
// memory in the UMMA no-swizzle canonical layout (core matrix = 8 rows x 16 B; the layout conv_umma.cu validates).
// Two kernel families:
//   * slab kernels (second half of this file; every layer with 48 | Cin): rows are enumerated on the input grid so that
//     a tap is a constant shift -- one contiguous voxel run is staged per kd and the taps are start addresses of the
//     same image (K-major for forward / dgrad, MN-major for wgrad), weights streamed with cp.async.bulk, dedicated MMA
//     issuer warp;
//   * gather kernels (first half; the Cin = 1 first layer, and everything with FPL_TC_GATHER=1): all threads of a CTA
//     gather one K chunk of both operands per tap, one thread issues the MMAs of the chunk while the CTA builds the
//     next chunk into the other stage (two stages, one mbarrier each, armed by tcgen05.commit).
//
// Precision: NS = 3 (default) splits every operand into bf16 hi + bf16 lo (v = hi + lo + O(2^-17 v)) and
// accumulates hi*hi + lo*hi + hi*lo in fp32: products of bf16 pairs are exact in fp32, so the result carries
// ~16 mantissa bits -- the reference trains in float32, and the parity test keeps the fp32 tolerance.  NS = 1
// is the plain bf16 contraction (8 mantissa bits, 3x fewer MMAs).
#pragma once
#include "umma.cuh"

namespace fpl {
namespace train {

using namespace fpl::net;     // PTX helpers of umma.cuh

constexpr int kTcThreads = 256;
constexpr int kTcMaxN = 96;
constexpr uint32_t kTcTmemCols = 128;
// operand images of one stage: A [atom][128 rows][16 B], B [atom][N rows][16 B] (hi image, then lo image when NS = 3)

struct TcArgs {
    const float *act;     // gathered activations (n_img, din^3, ca): conv = the tensor convolved, wgrad = layer input
    const float *mat;     // conv: kernel (tap, cin, cout) float32; wgrad: dx (n_img*dout^3, nn)
    float *out;           // conv: (rows, nn); wgrad: dW (k^3*ca, nn), accumulated with atomics (zeroed by the caller)
    int n_img, din, dout, k, ca, nn;
    int flip;             // conv: 0 = forward (B[n,(tap,c)] = W[tap][c][n]); 1 = dgrad (B[n,(tap,c)] = W[T-1-tap][n][c])
    int K;                // conv: k^3*ca
    int Mtot;             // wgrad: k^3*ca
    long long rows;       // n_img*dout^3: conv = GEMM rows, wgrad = GEMM K extent
    int chunks_per_cta;   // wgrad: K chunks (64 voxels) per CTA
    uint32_t a_bytes, b_bytes;   // bytes of one A / B operand image (atoms per chunk * 128 * 16, atoms * nn * 16)
};

// v -> bf16 hi (and lo = bf16(v - hi)) packed as 8 x bf16 = 16 B
template <int NS>
__device__ __forceinline__ void tc_pack8(const float (&v)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        h[e] = *reinterpret_cast<const uint32_t *>(&hb);
        if (NS == 3) {
            const float2 hf = __bfloat1622float2(hb);
            const __nv_bfloat162 lb = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
            l[e] = *reinterpret_cast<const uint32_t *>(&lb);
        } else {
            l[e] = 0u;
        }
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int NS>
__device__ __forceinline__ void tc_store8(uint8_t *stage_hi, uint32_t lo_offset, uint32_t off, const float (&v)[8]) {
    uint4 hi, lo;
    tc_pack8<NS>(v, hi, lo);
    *reinterpret_cast<uint4 *>(stage_hi + off) = hi;
    if (NS == 3) *reinterpret_cast<uint4 *>(stage_hi + lo_offset + off) = lo;
}

// the MMAs of one chunk: nks K=16 steps, NS products each, then a commit that arms `bar`
template <int NS>
__device__ __forceinline__ void tc_issue(uint32_t tmem_d, uint32_t sa, uint32_t sb, uint32_t a_bytes, uint32_t b_bytes,
                                         int nn, int nks, bool first_chunk, uint32_t idesc, uint64_t *bar) {
    tc_fence_after();
    const uint32_t b_lbo = (uint32_t)nn * 16u;
    for (int ks = 0; ks < nks; ++ks) {
        const uint64_t a_hi = make_desc(sa + (uint32_t)ks * 4096u, 2048u, 128u);
        const uint64_t b_hi = make_desc(sb + (uint32_t)ks * 2u * b_lbo, b_lbo, 128u);
        umma_bf16(tmem_d, a_hi, b_hi, idesc, (first_chunk && ks == 0) ? 0u : 1u);
        if (NS == 3) {
            const uint64_t a_lo = make_desc(sa + a_bytes + (uint32_t)ks * 4096u, 2048u, 128u);
            const uint64_t b_lo = make_desc(sb + b_bytes + (uint32_t)ks * 2u * b_lbo, b_lbo, 128u);
            umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
            umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
        }
    }
    umma_commit(bar);
}

// ------------------------------------------------------------------------------------------------
// valid convolution as a GEMM over (rows = output voxels) x (nn) with K = (tap, channel): forward and dgrad.
// Chunk = `apc` atoms: 6 (48 channels of one tap; ca in {48, 96}) or 4 (the 27 taps of a 1-channel input + 5 zeros).
// ------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv_kernel(const TcArgs a) {
    constexpr int NP = NS == 3 ? 2 : 1;
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t kStage = NP * (a.a_bytes + a.b_bytes);
    const uint32_t idesc = make_idesc_bf16(128, a.nn);
    const bool one_ch = a.ca == 1;
    const int apc = one_ch ? 4 : 6, kc = apc * 8, nks = apc / 2;
    const int n_it = 128 * apc / kTcThreads;             // A items (row, atom) per thread and chunk: 3 or 2
    const int nchunks = (a.K + kc - 1) / kc;
    const int ntaps = a.k * a.k * a.k;
    const long long n_tiles = (a.rows + 127) / 128;
    const int d2 = a.dout * a.dout, d3 = d2 * a.dout;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // this thread's A items: item = tid + j*256 -> (row = item / apc, atom = item % apc), the same for every chunk
        long long base[3]; int arow[3], aatom[3]; bool aok[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int item = tid + j * kTcThreads;
            arow[j] = item / apc; aatom[j] = item - arow[j] * apc;
            const long long m = tile * 128 + arow[j];
            aok[j] = j < n_it && m < a.rows;
            base[j] = 0;
            if (aok[j]) {
                const int img = (int)(m / d3); int rem = (int)(m - (long long)img * d3);
                const int z = rem / d2; rem -= z * d2;
                const int y = rem / a.dout, x = rem - y * a.dout;
                base[j] = ((((long long)img * a.din + z) * a.din + y) * a.din + x) * a.ca;
            }
        }
        for (int c = 0; c < nchunks; ++c, ++it) {
            const uint32_t s = it & 1u, u = it >> 1;
            if (u >= 1) mbar_wait(&bars[s], (u - 1u) & 1u);      // the MMAs that read this stage are done
            uint8_t *sa = tc_smem + s * kStage, *sb = sa + NP * a.a_bytes;
            const int kbase = c * kc;
            // ---- A: gathered activations
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (!aok[j]) continue;                            // rows beyond the tensor: their accumulator rows are dropped
                float v[8];
                const int k0 = kbase + aatom[j] * 8;
                if (!one_ch) {
                    const int tap = k0 / a.ca, ci = k0 - tap * a.ca;
                    const int kd = tap / (a.k * a.k), kh = (tap / a.k) % a.k, kw = tap % a.k;
                    const float4 *p = reinterpret_cast<const float4 *>(
                        a.act + base[j] + ((long long)(kd * a.din + kh) * a.din + kw) * a.ca + ci);
                    const float4 p0 = __ldg(p), p1 = __ldg(p + 1);
                    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int tap = k0 + e;
                        v[e] = 0.f;
                        if (tap < a.K) {
                            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                            v[e] = __ldg(a.act + base[j] + (long long)(kd * a.din + kh) * a.din + kw);
                        }
                    }
                }
                tc_store8<NS>(sa, a.a_bytes, (uint32_t)aatom[j] * 2048u + (uint32_t)arow[j] * 16u, v);
            }
            // ---- B: kernel slice of this chunk
            for (int i = tid; i < a.nn * apc; i += kTcThreads) {
                float v[8];
                int n, at;
                if (a.flip) {                 // dgrad: W[T-1-tap][n][c .. c+7], contiguous in c
                    n = i / apc; at = i - n * apc;
                    const int k0 = kbase + at * 8;
                    const int tap = k0 / a.ca, cc = k0 - tap * a.ca;
                    const float4 *p = reinterpret_cast<const float4 *>(
                        a.mat + ((long long)(ntaps - 1 - tap) * a.nn + n) * a.ca + cc);
                    const float4 p0 = __ldg(p), p1 = __ldg(p + 1);
                    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
                } else {                      // forward: W[tap][c + e][n], contiguous in n
                    at = i / a.nn; n = i - at * a.nn;
                    const int k0 = kbase + at * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int kk = k0 + e;
                        v[e] = kk < a.K ? __ldg(a.mat + (long long)kk * a.nn + n) : 0.f;   // (tap*ca + c)*cout + n
                    }
                }
                tc_store8<NS>(sb, a.b_bytes, (uint32_t)at * (uint32_t)a.nn * 16u + (uint32_t)n * 16u, v);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) tc_issue<NS>(tmem_base, smem_u32(sa), smem_u32(sb), a.a_bytes, a.b_bytes, a.nn, nks, c == 0, idesc, &bars[s]);
        }
        // ---- epilogue: accumulator rows -> out (rows, nn) float32
        {
            const uint32_t last = it - 1u;
            mbar_wait(&bars[last & 1u], (last >> 1) & 1u);
            tc_fence_after();
            const int q = warp & 3, half = warp >> 2;
            const long long m = tile * 128 + q * 32 + lane;
            for (int cb = half; cb < a.nn / 16; cb += 2) {
                uint32_t r[16];
                tmem_ld16(tmem_base + (uint32_t)(cb * 16) + ((uint32_t)(q * 32) << 16), r);
                tmem_ld_wait();
                if (m < a.rows) {
                    float4 *o = reinterpret_cast<float4 *>(a.out + m * a.nn + cb * 16);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        o[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]),
                                           __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
                }
            }
            tc_fence_before();
            __syncthreads();                                      // TMEM is free for the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTcTmemCols);
}

// ------------------------------------------------------------------------------------------------
// weight gradient: rows = (tap, ci), columns = co, K = output voxels (chunks of 64); grid = (row tiles, K splits)
// ------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_kernel(const TcArgs a) {
    constexpr int NP = NS == 3 ? 2 : 1;
    constexpr int KC = 64;
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_src[KC];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t kStage = NP * (a.a_bytes + a.b_bytes);
    const uint32_t idesc = make_idesc_bf16(128, a.nn);
    const int r = tid & 127, ah = tid >> 7;
    const int mrow = blockIdx.x * 128 + r;
    const bool valid = mrow < a.Mtot;
    int a_off = 0;
    if (valid) {
        const int tap = mrow / a.ca, ci = mrow - tap * a.ca;
        const int kd = tap / (a.k * a.k), kh = (tap / a.k) % a.k, kw = tap % a.k;
        a_off = ((kd * a.din + kh) * a.din + kw) * a.ca + ci;
    }
    const long long nchunks_all = (a.rows + KC - 1) / KC;
    const long long c0 = (long long)blockIdx.y * a.chunks_per_cta;
    const long long c1 = c0 + a.chunks_per_cta < nchunks_all ? c0 + a.chunks_per_cta : nchunks_all;
    const int d2 = a.dout * a.dout, d3 = d2 * a.dout;
    uint32_t it = 0;
    for (long long c = c0; c < c1; ++c, ++it) {
        const long long v0 = c * KC;
        if (tid < KC) {                                   // source offset of each voxel of the chunk (tap 0, channel 0)
            const long long v = v0 + tid;
            int src = -1;
            if (v < a.rows) {
                const int img = (int)(v / d3); int rem = (int)(v - (long long)img * d3);
                const int z = rem / d2; rem -= z * d2;
                const int y = rem / a.dout, x = rem - y * a.dout;
                src = (((img * a.din + z) * a.din + y) * a.din + x) * a.ca;
            }
            s_src[tid] = src;
        }
        __syncthreads();
        const uint32_t s = it & 1u, u = it >> 1;
        if (u >= 1) mbar_wait(&bars[s], (u - 1u) & 1u);
        uint8_t *sa = tc_smem + s * kStage, *sb = sa + NP * a.a_bytes;
        // ---- A: in[v + tap, ci] for the 64 voxels (rows beyond Mtot: their accumulator rows are dropped)
        if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int at = ah * 4 + j;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int src = s_src[at * 8 + e];
                    v[e] = src >= 0 ? __ldg(a.act + src + a_off) : 0.f;
                }
                tc_store8<NS>(sa, a.a_bytes, (uint32_t)at * 2048u + (uint32_t)r * 16u, v);
            }
        }
        // ---- B: dx[v, co]
        for (int i = tid; i < a.nn * 8; i += kTcThreads) {
            const int at = i / a.nn, n = i - at * a.nn;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const long long vv = v0 + at * 8 + e;
                v[e] = vv < a.rows ? __ldg(a.mat + vv * a.nn + n) : 0.f;
            }
            tc_store8<NS>(sb, a.b_bytes, (uint32_t)at * (uint32_t)a.nn * 16u + (uint32_t)n * 16u, v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) tc_issue<NS>(tmem_base, smem_u32(sa), smem_u32(sb), a.a_bytes, a.b_bytes, a.nn, 4, it == 0, idesc, &bars[s]);
    }
    if (it > 0) {
        const uint32_t last = it - 1u;
        mbar_wait(&bars[last & 1u], (last >> 1) & 1u);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const int m = blockIdx.x * 128 + q * 32 + lane;
        for (int cb = half; cb < a.nn / 16; cb += 2) {
            uint32_t rr[16];
            tmem_ld16(tmem_base + (uint32_t)(cb * 16) + ((uint32_t)(q * 32) << 16), rr);
            tmem_ld_wait();
            if (m < a.Mtot) {
                float *o = a.out + (size_t)m * a.nn + cb * 16;
#pragma unroll
                for (int e = 0; e < 16; ++e) atomicAdd(o + e, __uint_as_float(rr[e]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTcTmemCols);
}

// zero-padded copy (n, d^3, c) -> (n, (d+2p)^3, c): the dgrad convolves the padded dx
__global__ void __launch_bounds__(256)
tc_pad_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int n, int d, int p, int c4) {
    const int dp = d + 2 * p;
    const long long total = (long long)n * dp * dp * dp * c4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c4); long long v = i / c4;
        const int x = (int)(v % dp) - p; v /= dp;
        const int y = (int)(v % dp) - p; v /= dp;
        const int z = (int)(v % dp) - p; const int t = (int)(v / dp);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0 && x < d && y >= 0 && y < d && z >= 0 && z < d)
            val = in[((((size_t)t * d + z) * d + y) * d + x) * c4 + ch];
        out[i] = val;
    }
}


// ================================================================================================
// Slab kernels (layers with 48 | Cin): no per-tap gathers.
//
// Rows are enumerated on the INPUT grid ("padded-flat": u = (z*din + y)*din + x, all y, x < din), so the source of
// row u under tap (kd,kh,kw) is simply u + (kd*din + kh)*din + kw: a constant shift.  A tile stages, per kd, ONE
// contiguous run of voxels (float32 in HBM -> bf16 hi/lo images [channel atom][voxel][16 B] in shared memory, voxel
// pitch 16 B) and every tap is a different START ADDRESS of the same image:
//   * forward / dgrad: the image is the K-major A operand (rows = voxels: SBO = 128 B, LBO = atom pitch),
//   * wgrad: the same kind of image is an MN-major operand (MN = channels: SBO = atom pitch, K = voxels: LBO = 128 B).
// Rows with y >= dout or x >= dout (about (din/dout)^2 of the work) are computed and dropped / multiplied by zeros.
// The weights of forward and dgrad are packed once per step into bf16 hi/lo operand-B images (tc_pack_w_kernel) and
// streamed through a shared-memory ring with cp.async.bulk (one 48-channel K chunk per slot).
// ================================================================================================
constexpr int kSlabWorkers = 256;                 // warps 0..7: staging + epilogue
constexpr int kSlabThreads = kSlabWorkers + 64;   // warp 8: MMA issuer, warp 9: weight (B) producer
constexpr int kSlabRing = 4;                      // B ring slots

struct SlabConvArgs {
    const float *act;          // (n_img, din^3, ca) float32
    const __nv_bfloat16 *wimg; // packed weights: per K chunk [hi: [6 atoms][nn][8]] [lo: the same]
    float *out;                // (n_img, dout^3, nn) float32
    int n_img, din, dout, k, ca, nn;
    int mt;                    // M tiles (128 rows) per pass: 1 or 2
    int s_pad;                 // voxels per staged slice (pitch of a channel atom, in 16-byte units; == 1 mod 8)
    int pairs_per_img;         // passes per image
    int u_max;                 // rows per image that can hold a valid output
    int flat;                  // k == 1: the batch is one run of n_img*din^3 rows (n_img = 1, out row = u)
    long long img_vox;         // voxels per image (flat: of the whole batch)
};

// fp32 run -> [atom][voxel][16 B] hi/lo images.  item = voxel*apv + atom reads 32 contiguous bytes.
template <int NS>
__device__ __forceinline__ void slab_stage(const float *__restrict__ src, long long src_elems_left, int nvox, int apv,
                                           uint8_t *img_hi, uint32_t lo_off, uint32_t pitch_bytes, int tid, int nthr) {
    const int items = nvox * apv;
    for (int i0 = tid; i0 < items; i0 += 4 * nthr) {
        float4 p[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * nthr;
            p[j][0] = p[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < items && (long long)i * 8 + 8 <= src_elems_left) {
                const float4 *g = reinterpret_cast<const float4 *>(src + (size_t)i * 8);
                p[j][0] = __ldg(g); p[j][1] = __ldg(g + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * nthr;
            if (i < items) {
                const int v = i / apv, at = i - v * apv;
                const float f[8] = {p[j][0].x, p[j][0].y, p[j][0].z, p[j][0].w, p[j][1].x, p[j][1].y, p[j][1].z, p[j][1].w};
                tc_store8<NS>(img_hi, lo_off, (uint32_t)at * pitch_bytes + (uint32_t)v * 16u, f);
            }
        }
    }
}

template <int NS>
__global__ void __launch_bounds__(kSlabThreads, 1)
tc_slab_conv_kernel(const SlabConvArgs a) {
    constexpr int NP = NS == 3 ? 2 : 1;
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ uint64_t bars[7 + 2 * kSlabRing];
    __shared__ uint32_t tmem_slot;
    // per kd slice of the A image: full (workers -> MMA issuer) and free (tcgen05.commit after the slice's last MMA), so
    // the workers re-stage slice kd for the next pass while the MMAs of the other slices still run
    uint64_t *slice_full = &bars[0], *slice_free = &bars[3], *mma_done = &bars[6], *b_full = &bars[7], *b_empty = &bars[7 + kSlabRing];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int apv = a.ca / 8;                                    // channel atoms per voxel
    const uint32_t pitch = (uint32_t)a.s_pad * 16u;              // bytes between channel atoms
    const uint32_t slice_bytes = (uint32_t)apv * pitch;          // one kd slice, one image (hi or lo)
    const uint32_t a_img_bytes = (uint32_t)a.k * slice_bytes;    // hi image; the lo image follows
    const uint32_t b_half = 6u * (uint32_t)a.nn * 16u;           // hi part of one K chunk
    const uint32_t b_slot = NP * b_half;
    uint8_t *sA = tc_smem, *sB = tc_smem + NP * a_img_bytes;
    const uint32_t ncol = (uint32_t)NP * (uint32_t)a.nn;          // accumulator columns per M tile: [hi*hi + lo*hi | hi*lo]
    const uint32_t acc_cols = (uint32_t)a.mt * ncol;
    const uint32_t tmem_cols = 2 * acc_cols <= 256 ? 256u : 512u;
    if (tid == 0) {
        mbar_init(mma_done, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(&slice_full[i], 1); mbar_init(&slice_free[i], 1); }
        for (int i = 0; i < kSlabRing; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int cpt = a.ca / 48;                                   // K chunks per tap
    const int nchunks = a.k * a.k * a.k * cpt;
    const int n_pairs = a.n_img * a.pairs_per_img;
    const int rows_pass = a.mt * 128;
    const int halo = (a.k - 1) * (a.din + 1);
    const long long img_elems = a.img_vox * a.ca;
    const long long tot_elems = (long long)a.n_img * img_elems;

    if (warp < 8) {
        // ======================= workers: stage the A slices, then drain the previous pass =======================
        uint32_t pl = 0;
        int prev_pair = -1;
        auto epilogue = [&](int pair, uint32_t buf) {
            const int img = pair / a.pairs_per_img, u0 = (pair - img * a.pairs_per_img) * rows_pass;
            const int q = warp & 3, half = warp >> 2;
            const int d2 = a.din * a.din;
            for (int t = 0; t < a.mt; ++t) {
                const int u = u0 + t * 128 + q * 32 + lane;
                const int z = u / d2, rem = u - z * d2, y = rem / a.din, x = rem - y * a.din;
                const bool ok = a.flat ? u < a.u_max : (z < a.dout && y < a.dout && x < a.dout);
                float *o = a.out + (a.flat ? (size_t)u : (((size_t)img * a.dout + z) * a.dout + y) * a.dout + x) * a.nn;
                for (int cb = half; cb < a.nn / 16; cb += 2) {
                    uint32_t r[16], r2[16];
                    const uint32_t tcol = tmem_base + buf * acc_cols + (uint32_t)t * ncol + (uint32_t)(cb * 16) + ((uint32_t)(q * 32) << 16);
                    tmem_ld16(tcol, r);
                    if (NS == 3) tmem_ld16(tcol + (uint32_t)a.nn, r2);
                    tmem_ld_wait();
                    if (ok) {
                        float4 *o4 = reinterpret_cast<float4 *>(o + cb * 16);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float4 v = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]),
                                                   __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
                            if (NS == 3) {
                                v.x += __uint_as_float(r2[4 * e]); v.y += __uint_as_float(r2[4 * e + 1]);
                                v.z += __uint_as_float(r2[4 * e + 2]); v.w += __uint_as_float(r2[4 * e + 3]);
                            }
                            o4[e] = v;
                        }
                    }
                }
            }
            tc_fence_before();
        };
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pl) {
            const int img = pair / a.pairs_per_img, u0 = (pair - img * a.pairs_per_img) * rows_pass;
            const int nvox = rows_pass + halo;
            for (int kd = 0; kd < a.k; ++kd) {
                if (pl >= 1) mbar_wait(&slice_free[kd], (pl - 1u) & 1u);          // the MMAs of the previous pass are done with it
                const long long e0 = (long long)img * img_elems + ((long long)u0 + (long long)kd * a.din * a.din) * a.ca;
                slab_stage<NS>(a.act + e0, tot_elems - e0, nvox, apv, sA + kd * slice_bytes, a_img_bytes, pitch, tid,
                               kSlabWorkers);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) mbar_arrive(&slice_full[kd]);
            }
            // the last slice_free wait above implies mma_done of the previous pass: drain its accumulators now, under
            // the MMAs of this pass
            if (prev_pair >= 0) { mbar_wait(mma_done, (pl - 1u) & 1u); tc_fence_after(); epilogue(prev_pair, (pl - 1u) & 1u); }
            prev_pair = pair;
        }
        if (prev_pair >= 0) {
            mbar_wait(mma_done, (pl - 1u) & 1u);
            tc_fence_after();
            epilogue(prev_pair, (pl - 1u) & 1u);
        }
    } else if (warp == 8) {
        // ======================= MMA issuer =======================
        // The whole warp walks the (warp-uniform) loops and one elected lane issues: with `if (lane == 0)` around the loops
        // the compiler cannot keep the descriptors in uniform registers and wraps EVERY tcgen05.mma in an ELECT / R2UR /
        // BRA.U.ANY sequence (~86 clocks per MMA measured, against 24 clocks of tensor-pipe time at N = 48).
        const bool leader = elect_one();
        {
            const uint32_t idesc = make_idesc_bf16(128, a.nn), idesc2 = make_idesc_bf16(128, NP * a.nn);
            const uint32_t b_lbo = (uint32_t)NP * (uint32_t)a.nn * 16u;      // atom pitch of the (stacked) weight image
            const uint32_t hi_w = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
            const uint32_t lo_a0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | ((pitch >> 4) << 16);     // LBO = atom pitch
            const uint32_t lo_b0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | ((b_lbo >> 4) << 16);
            const uint32_t a_ks = (2u * pitch) >> 4, a_lo_img = a_img_bytes >> 4, a_slice = slice_bytes >> 4;
            const uint32_t b_ks = (2u * b_lbo) >> 4, b_slot16 = b_slot >> 4;
            const uint32_t a_half = (6u * pitch) >> 4;
            const bool two = a.mt == 2;
            uint32_t pl = 0, g = 0;
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++pl) {
                const uint32_t d0 = tmem_base + (pl & 1u) * acc_cols, d1 = d0 + ncol;
                uint32_t first = 0u;                                         // 0 on the first chunk: overwrite the accumulator
                for (int kd = 0; kd < a.k; ++kd) {
                    mbar_wait(&slice_full[kd], pl & 1u);
                    tc_fence_after();
                    for (int kh = 0; kh < a.k; ++kh)
                        for (int kw = 0; kw < a.k; ++kw)
                            for (int hf = 0; hf < cpt; ++hf, ++g) {
                                const uint32_t slot = g % kSlabRing, use = g / kSlabRing;
                                mbar_wait(&b_full[slot], use & 1u);
                                tc_fence_after();
                                const uint32_t ac = lo_a0 + (uint32_t)kd * a_slice + (uint32_t)hf * a_half + (uint32_t)(kh * a.din + kw);
                                const uint32_t bc = lo_b0 + slot * b_slot16;
#pragma unroll
                                for (int ks = 0; ks < 3; ++ks) {
                                    const uint32_t al = ac + (uint32_t)ks * a_ks, bl = bc + (uint32_t)ks * b_ks;
                                    const uint32_t acc = ks == 0 ? first : 1u;
                                    if (leader) {
                                        // hi activations x [hi | lo] weights (N = 2*nn), then lo activations x hi weights (N = nn,
                                        // same image: its first nn rows) into the first column block
                                        umma_bf16(d0, desc64(al, hi_w), desc64(bl, hi_w), idesc2, acc);
                                        if (NS == 3) umma_bf16(d0, desc64(al + a_lo_img, hi_w), desc64(bl, hi_w), idesc, 1u);
                                        if (two) {
                                            umma_bf16(d1, desc64(al + 128u, hi_w), desc64(bl, hi_w), idesc2, acc);
                                            if (NS == 3) umma_bf16(d1, desc64(al + 128u + a_lo_img, hi_w), desc64(bl, hi_w), idesc, 1u);
                                        }
                                    }
                                }
                                first = 1u;
                                if (leader) umma_commit(&b_empty[slot]);
                                __syncwarp();
                            }
                    if (leader) umma_commit(&slice_free[kd]);
                    __syncwarp();
                }
                if (leader) umma_commit(mma_done);
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ======================= weight producer: stream the packed K chunks through the ring =======================
        if (lane == 0) {
            uint32_t g = 0;
            for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int c = 0; c < nchunks; ++c, ++g) {
                    const uint32_t slot = g % kSlabRing, use = g / kSlabRing;
                    if (use >= 1) mbar_wait(&b_empty[slot], (use - 1u) & 1u);
                    mbar_expect_tx(&b_full[slot], b_slot);
                    bulk_load_1d(sB + slot * b_slot, reinterpret_cast<const uint8_t *>(a.wimg) + (size_t)c * b_slot, b_slot,
                                 &b_full[slot]);
                }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// packed operand-B images of one layer: chunk c = (tap, 48-channel half); [atom][hi rows | lo rows][8 bf16]
//   flip = 0 (forward): B[n, (tap, c)] = W[tap][c][n]       (W is (tap, ca, nn))
//   flip = 1 (dgrad):   B[n, (tap, c)] = W[T-1-tap][n][c]   (W is (tap, nn, ca))
template <int NS>
__global__ void __launch_bounds__(256)
tc_pack_w_kernel(const float *__restrict__ w, uint8_t *__restrict__ img, int ntaps, int ca, int nn, int flip) {
    const int cpt = ca / 48;
    const int total = ntaps * cpt * 6 * nn;
    // NS = 3: per atom the hi rows are followed by the lo rows ([atom][hi 0..nn-1 | lo nn..2nn-1][16 B]): ONE N = 2*nn MMA
    // of the hi activations then yields hi*hi and hi*lo side by side (the epilogue adds the two column blocks)
    constexpr uint32_t NPK = NS == 3 ? 2u : 1u;
    const uint32_t b_slot = NPK * 6u * (uint32_t)nn * 16u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i % nn; int r = i / nn;
        const int at = r % 6; r /= 6;
        const int hf = r % cpt, tap = r / cpt;
        const int c0 = hf * 48 + at * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
            v[e] = flip ? w[((size_t)(ntaps - 1 - tap) * nn + n) * ca + c0 + e] : w[((size_t)tap * ca + c0 + e) * nn + n];
        uint4 hi, lo;
        tc_pack8<NS>(v, hi, lo);
        uint8_t *dst = img + (size_t)(tap * cpt + hf) * b_slot + (size_t)at * NPK * nn * 16 + (size_t)n * 16;
        *reinterpret_cast<uint4 *>(dst) = hi;
        if (NS == 3) *reinterpret_cast<uint4 *>(dst + (size_t)nn * 16) = lo;
    }
}

// ------------------------------------------------------------------------------------------------
// slab wgrad: dW[tap][ci][co] = sum_u in[u + shift(tap), ci] * dxe[u, co] with dxe = dx on the INPUT grid (zeros where
// there is no output).  Per chunk of 128 voxels: A = dxe^T (MN-major, M = co padded to 128: the rows beyond Cout are
// dropped), B = in^T shifted by the tap (MN-major, N = ci); one accumulator block of Cin columns per tap of the CTA's
// kd group (k^2 taps; k = 1: the single tap).  grid = (k, CTAs per kd group).
// ------------------------------------------------------------------------------------------------
struct SlabWgradArgs {
    const float *act;      // layer input (n_img, din^3, ci)
    const float *dxe;      // dx on the input grid (n_img, din^3, co), zero where x/y/z >= dout
    float *dw;             // (k^3, ci, co), accumulated with atomics
    int n_img, din, dout, k, ci, co;
    int s_pad;             // voxels of a staged input slice (== 1 mod 8)
    int chunks_per_img;    // 128-voxel chunks per image
    int units_per_cta;     // (image, chunk) units per CTA
    uint32_t smem_bytes;   // dynamic shared memory of the launch (two stages + the over-read tail)
    long long img_vox;     // voxels per image (k == 1: of the whole batch, n_img = 1)
};

constexpr int kWgKc = 128;                 // voxels per chunk
constexpr int kWgAPad = 129;               // A (dxe) atom pitch in 16-byte units (== 1 mod 8)

constexpr int kWgThreads = 256 + 32;       // warps 0..7: staging + epilogue, warp 8: MMA issuer

template <int NS>
__global__ void __launch_bounds__(kWgThreads, 1)
tc_slab_wgrad_kernel(const SlabWgradArgs a) {
    constexpr int NP = NS == 3 ? 2 : 1;
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    uint64_t *full = &bars[0], *empty = &bars[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kd = blockIdx.x;
    const int taps = a.k * a.k;                               // taps of this kd group
    const int apa = a.co / 8, apx = a.ci / 8;
    const uint32_t pa = (uint32_t)kWgAPad * 16u, px = (uint32_t)a.s_pad * 16u;
    // stage layout: [A hi (apa atoms)] [A lo] [X hi] [X lo].  The MMA reads M = 128 rows = 16 atom pitches from the A
    // start: the rows beyond Cout read whatever follows (finite bf16 data / the zeroed tail) and are dropped.
    const uint32_t a_img = (uint32_t)apa * pa, x_img = (uint32_t)apx * px;
    const uint32_t stage_bytes = NP * (a_img + x_img);
    const uint32_t tmem_cols = 512;
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    for (uint32_t i = tid; i < a.smem_bytes / 16; i += kWgThreads) reinterpret_cast<uint4 *>(tc_smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_units = a.n_img * a.chunks_per_img;
    const int unit0 = blockIdx.y * a.units_per_cta;
    const int unit1 = unit0 + a.units_per_cta < n_units ? unit0 + a.units_per_cta : n_units;
    const int halo = (a.k - 1) * (a.din + 1);
    const long long img_vox = a.img_vox;
    const long long tot_x = (long long)a.n_img * img_vox * a.ci, tot_d = (long long)a.n_img * img_vox * a.co;
    if (warp < 8) {
        // ======================= workers: stage dxe and the input slice of every unit =======================
        uint32_t it = 0;
        for (int unit = unit0; unit < unit1; ++unit, ++it) {
            const int img = unit / a.chunks_per_img, u0 = (unit - img * a.chunks_per_img) * kWgKc;
            const uint32_t s = it & 1u, use = it >> 1;
            if (use >= 1) mbar_wait(&empty[s], (use - 1u) & 1u);
            uint8_t *sA = tc_smem + s * stage_bytes, *sX = sA + NP * a_img;
            const long long ed = ((long long)img * img_vox + u0) * a.co;
            const long long in_img = (img_vox - u0) * a.co;          // a chunk never takes voxels of the next image
            slab_stage<NS>(a.dxe + ed, in_img < tot_d - ed ? in_img : tot_d - ed, kWgKc, apa, sA, a_img, pa, tid, 256);
            const long long ex = ((long long)img * img_vox + u0 + (long long)kd * a.din * a.din) * a.ci;
            slab_stage<NS>(a.act + ex, tot_x - ex, kWgKc + halo, apx, sX, x_img, px, tid, 256);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) mbar_arrive(&full[s]);
        }
        if (it > 0) {
            // ---- epilogue: rows = co, columns = (tap, ci) -> atomics into dW (tap, ci, co)
            const uint32_t last = it - 1u;
            mbar_wait(&empty[last & 1u], (last >> 1) & 1u);
            tc_fence_after();
            const int q = warp & 3, half = warp >> 2;
            const bool stacked = NS == 3 && 2 * a.co <= 128;        // rows Cout..2*Cout-1 hold the lo*hi products
            int co = q * 32 + lane;
            const bool row_ok = co < a.co || (stacked && co < 2 * a.co);
            if (co >= a.co) co -= a.co;
            const int ncb = taps * a.ci / 16;
            for (int cb = half; cb < ncb; cb += 2) {
                uint32_t r[16];
                tmem_ld16(tmem_base + (uint32_t)(cb * 16) + ((uint32_t)(q * 32) << 16), r);
                tmem_ld_wait();
                if (row_ok) {
                    const int col0 = cb * 16, tl = col0 / a.ci, ci0 = col0 - tl * a.ci;      // 16 | ci: one tap per block
                    float *o = a.dw + ((size_t)(kd * taps + tl) * a.ci + ci0) * a.co + co;
#pragma unroll
                    for (int e = 0; e < 16; ++e) atomicAdd(o + (size_t)e * a.co, __uint_as_float(r[e]));
                }
            }
        }
    } else {
        // ======================= MMA issuer (whole warp in the loops, one elected lane issues) =======================
        const bool leader = elect_one();
        // both operands MN-major, no swizzle: SBO = pitch between 8-channel atoms, LBO = 128 B between 8-voxel groups
        const uint32_t idesc = make_idesc_bf16(128, a.ci) | (1u << 15) | (1u << 16);
        const uint32_t hi_a = (pa >> 4) | (1u << 14), hi_b = (px >> 4) | (1u << 14);
        const uint32_t lbo = (128u >> 4) << 16;
        const uint32_t a_lo_img = a_img >> 4, x_lo_img = x_img >> 4;
        const bool stacked = NS == 3 && 2 * a.co <= 128;
        uint32_t it = 0;
        for (int unit = unit0; unit < unit1; ++unit, ++it) {
            const uint32_t s = it & 1u, use = it >> 1;
            mbar_wait(&full[s], use & 1u);
            tc_fence_after();
            const uint32_t sa = smem_u32(tc_smem + s * stage_bytes);
            const uint32_t al0 = ((sa & 0x3FFFFu) >> 4) | lbo;
            const uint32_t xl0 = (((sa + NP * a_img) & 0x3FFFFu) >> 4) | lbo;
            const uint32_t first = it == 0 ? 0u : 1u;
            for (int kh = 0; kh < a.k; ++kh)
                for (int kw = 0; kw < a.k; ++kw) {
                    const uint32_t d = tmem_base + (uint32_t)((kh * a.k + kw) * a.ci);
                    const uint32_t xt = xl0 + (uint32_t)(kh * a.din + kw);
#pragma unroll
                    for (int ks = 0; ks < kWgKc / 16; ++ks) {
                        const uint32_t al = al0 + (uint32_t)ks * 16u, bl = xt + (uint32_t)ks * 16u;    // 16 voxels = 256 B
                        if (leader) {
                            // NS = 3, Cout = 48: the lo image of dx follows the hi image at the same atom pitch, so the
                            // M = 128 read of the first MMA already covers [hi rows 0..47 | lo rows 48..95]: lo*hi lands
                            // in accumulator rows 48..95 for free (the epilogue adds them) and its own MMA is dropped
                            umma_bf16(d, desc64(al, hi_a), desc64(bl, hi_b), idesc, ks == 0 ? first : 1u);
                            if (NS == 3) {
                                if (!stacked) umma_bf16(d, desc64(al + a_lo_img, hi_a), desc64(bl, hi_b), idesc, 1u);
                                umma_bf16(d, desc64(al, hi_a), desc64(bl + x_lo_img, hi_b), idesc, 1u);
                            }
                        }
                    }
                }
            if (leader) umma_commit(&empty[s]);
            __syncwarp();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// zero-padded copy (n, d^3, c) -> (n, dp^3, c) with the data at offset p in every dimension
__global__ void __launch_bounds__(256)
tc_embed_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int n, int d, int p, int dp, int c4) {
    const long long total = (long long)n * dp * dp * dp * c4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c4); long long v = i / c4;
        const int x = (int)(v % dp) - p; v /= dp;
        const int y = (int)(v % dp) - p; v /= dp;
        const int z = (int)(v % dp) - p; const int t = (int)(v / dp);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0 && x < d && y >= 0 && y < d && z >= 0 && z < d)
            val = in[((((size_t)t * d + z) * d + y) * d + x) * c4 + ch];
        out[i] = val;
    }
}

}  // namespace train
}  // namespace fpl
